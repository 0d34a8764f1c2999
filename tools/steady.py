"""Steady-state launch time of the sweep kernel on the bench workload: one sampler call,
per-launch kernel times from the library's launch log, median over the launches after burn-in.

    python tools/steady.py --chains 512 --launches 20 [--chunk 32] [--lib path/to/variant.so]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=512)
ap.add_argument("--launches", type=int, default=20)
ap.add_argument("--mode", default="gram")
ap.add_argument("--chunk", type=int, default=32)
ap.add_argument("--lib", default=None, help="variant library copied over the in-tree one first")
ap.add_argument("--kernel", type=int, default=0, help="BLIPGPU_GRAM_KERNEL_* (0 = auto)")
ap.add_argument("--child", action="store_true")
args = ap.parse_args()
if not args.child:
    if args.lib:
        shutil.copy(args.lib, os.path.join(ROOT, "pyblip_b200", "_C", "libblipgpu.so"))
    env = dict(os.environ, BLIPGPU_LAUNCH_LOG="1")
    res = subprocess.run([sys.executable, __file__, "--child", "--chains", str(args.chains), "--launches",
                          str(args.launches), "--mode", args.mode, "--chunk", str(args.chunk), "--kernel", str(args.kernel)],
                         env=env, capture_output=True, text=True)
    ms = [float(m.group(2)) / (int(m.group(1)) / 32.0) for m in
          re.finditer(r"launch sweeps \d+\.\.\d+ grid \d+: ([0-9.]+) ms", res.stderr) for m in [m]] if False else []
    rows = re.findall(r"launch sweeps (\d+)\.\.(\d+) grid (\d+): ([0-9.]+) ms", res.stderr)
    per32 = [float(t) * 32.0 / (int(b) - int(a)) for a, b, g, t in rows if int(a) >= 96]
    per32.sort()
    if not per32:
        print(res.stdout[-2000:], res.stderr[-2000:])
        sys.exit(1)
    med = per32[len(per32) // 2]
    for line in res.stdout.splitlines():
        if "|" in line:
            print(line)
    print(f"{args.lib or 'in-tree'} kernel {args.kernel}: chains {args.chains} chunk {args.chunk} grid {rows[-1][2]}: median {med:.3f} ms per 32 sweeps "
          f"(min {per32[0]:.3f}, max {per32[-1]:.3f}, {len(per32)} launches)")
    sys.exit(0)

import bench  # noqa: E402
from pyblip_b200 import _lib  # noqa: E402
from pyblip_b200.linear import LinearSpikeSlab  # noqa: E402

X, y, _ = bench.make_data()
design = _lib.Design.upload(X)
mode = {"gram": _lib.MODE_GRAM, "residual": _lib.MODE_RESIDUAL}[args.mode]
if mode == _lib.MODE_GRAM:
    design.build_gram()
cfg = LinearSpikeSlab(X, y)._config()
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=args.chains, n_keep=args.chunk * args.launches, burn=96,
                            seed=7, mode=mode, chunk_sweeps=args.chunk, gram_kernel=args.kernel)
smp.close()
