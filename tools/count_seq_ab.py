"""count_sequential on C2-shaped device samples: gap-histogram kernel vs the round-1 tile kernel (A/B), and
other max_size values.   python tools/count_seq_ab.py"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import bench
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    X, y, _ = bench.make_data()
    ctx = _lib.Context.get(0)
    design = _lib.Design.upload(X, ctx)
    design.build_gram()
    cfg = LinearSpikeSlab(X, y)._config()
    smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=1000, burn=100, seed=3, mode=_lib.MODE_GRAM)
    bits = smp.bits()
    p, rows = X.shape[1], smp.info.rows
    nbytes = ((p + 31) // 32) * 4 * rows
    for ms_ in (25, 64, 200):
        if os.environ.get("BLIPGPU_COUNT_SEQ_TILES") and ms_ > 64:
            continue
        bits.count_sequential(ms_)
        best = 1e9
        for _ in range(5):
            ctx.timing(reset=True)
            bits.count_sequential(ms_)
            ctx.synchronize()
            best = min(best, ctx.timing()["other_ms"])
        print(f"  max_size {ms_:4d}: {best:.3f} ms = {nbytes / best / 1e6:.0f} GB/s of bit-matrix bytes ({nbytes / 1e6:.0f} MB)")
    sys.exit(0)
for env, name in (({}, "gap histogram (default)"), ({"BLIPGPU_COUNT_SEQ_TILES": "1"}, "round-1 tile kernel")):
    print(name)
    r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
    print(r.stdout + r.stderr[-500:] if r.returncode else r.stdout, end="")
