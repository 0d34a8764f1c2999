"""Stationary-regime launch times of the block sampler on the bench workload (launch log)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--child" in sys.argv:
    import bench
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    mode, bsize, chains = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    X, y, _ = bench.make_data()
    design = _lib.Design.upload(X)
    cfg = LinearSpikeSlab(X, y)._config()
    smp = _lib.sample_spikeslab(design, cfg, y=y, chains=chains, n_keep=64, burn=96, seed=7,
                                mode={"gram": _lib.MODE_GRAM, "residual": _lib.MODE_RESIDUAL}[mode],
                                chunk_sweeps=16, bsize=bsize)
    i = smp.info
    print(f"applied blocks/chain-sweep {i.n_changes / i.n_sweeps_total:.1f} windows {i.n_windows / i.n_sweeps_total:.1f}")
    sys.exit(0)
for mode in ("gram", "residual"):
    for bsize in (3, 5):
        chains = 512
        res = subprocess.run([sys.executable, __file__, "--child", mode, str(bsize), str(chains)],
                             env=dict(os.environ, BLIPGPU_LAUNCH_LOG="1"), capture_output=True, text=True)
        rows = re.findall(r"launch sweeps (\d+)\.\.(\d+) grid (\d+): ([0-9.]+) ms", res.stderr)
        per = sorted(float(t) / (int(b) - int(a)) for a, b, g, t in rows if int(a) >= 96)
        if not per:
            print(res.stdout[-500:], res.stderr[-1500:]); continue
        med = per[len(per) // 2]
        print(f"block {mode} bsize={bsize} chains={chains}: {med:.2f} ms per sweep of all chains (stationary) = "
              f"{chains * 10000 / (med * 1e-3):.3e} coordinate updates/s   {res.stdout.strip()}", flush=True)
