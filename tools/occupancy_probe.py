"""Throughput of the classic gram kernel as a function of CTAs per SM: the AR(1) bench design at another p, so that
q (8p bytes of shared memory per CTA) allows a third resident CTA.  Used with a tuning variant built with
`python tools/build_variant.py b192 -DBLIP_GB_BLOCK=192 -DBLIP_GB_GRAM_REGS=112` (192 threads x 112 registers x 3 CTAs).

    python tools/occupancy_probe.py --p 8800 --chains 888 [--lib pyblip_b200/_C/variants/b192.so]
"""
import argparse, hashlib, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=8800)
ap.add_argument("--chains", type=int, default=888)
ap.add_argument("--launches", type=int, default=8)
ap.add_argument("--lib", default=None)
ap.add_argument("--child", action="store_true")
args = ap.parse_args()
if not args.child:
    lib = os.path.join(ROOT, "pyblip_b200", "_C", "libblipgpu.so")
    keep = lib + ".probe_keep"
    if args.lib:
        shutil.copy(lib, keep)
        shutil.copy(args.lib, lib)
    try:
        res = subprocess.run([sys.executable, __file__, "--child", "--p", str(args.p), "--chains", str(args.chains),
                              "--launches", str(args.launches)], env=dict(os.environ, BLIPGPU_LAUNCH_LOG="1"),
                             capture_output=True, text=True)
    finally:
        if args.lib:
            shutil.move(keep, lib)
    rows = re.findall(r"launch sweeps (\d+)\.\.(\d+) grid (\d+): ([0-9.]+) ms", res.stderr)
    per32 = sorted(float(t) * 32.0 / (int(b) - int(a)) for a, b, g, t in rows if int(a) >= 96)
    if not per32:
        print(res.stdout[-2000:], res.stderr[-3000:])
        sys.exit(1)
    med = per32[len(per32) // 2]
    print(res.stdout.strip())
    print(f"{args.lib or 'in-tree'}: p {args.p} chains {args.chains} grid {rows[-1][2]}: median {med:.3f} ms per 32 sweeps "
          f"= {args.chains * 32 / med / 1e3:.1f} M chain-sweeps/s = {args.chains * 32 * args.p / med / 1e6:.2f} G updates/s")
    sys.exit(0)

import numpy as np
import bench
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab
cfg = dict(bench.CONFIG, p=args.p)
X, y, _ = bench.make_data(cfg)
design = _lib.Design.upload(X)
design.build_gram()
smp = _lib.sample_spikeslab(design, LinearSpikeSlab(X, y)._config(), y=y, chains=args.chains, n_keep=32 * args.launches,
                            burn=96, seed=7, mode=_lib.MODE_GRAM, chunk_sweeps=32, gram_kernel=_lib.GRAM_KERNEL_CLASSIC)
indptr, indices, values = smp.csr()
h = hashlib.sha256()
for a in (indptr, indices, values):
    h.update(np.ascontiguousarray(a).tobytes())
print(f"nnz {int(indptr[-1])}  sha256 of the CSR {h.hexdigest()[:16]}")
smp.close()
