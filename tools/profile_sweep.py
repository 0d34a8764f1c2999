"""Small driver for ncu: one sampler call on the bench workload (or a scaled-down one).

    python tools/profile_sweep.py --mode gram --sweeps 32 --burn 64 --chains 512
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyblip_b200 import _lib  # noqa: E402
from pyblip_b200.linear import LinearSpikeSlab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="gram")
ap.add_argument("--sweeps", type=int, default=32)
ap.add_argument("--burn", type=int, default=64)
ap.add_argument("--chains", type=int, default=512)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--kernel", type=int, default=0, help="BLIPGPU_GRAM_KERNEL_*")
args = ap.parse_args()
X, y, _ = bench.make_data()
design = _lib.Design.upload(X)
mode = {"gram": _lib.MODE_GRAM, "residual": _lib.MODE_RESIDUAL}[args.mode]
if mode == _lib.MODE_GRAM:
    design.build_gram()
cfg = LinearSpikeSlab(X, y)._config()
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=args.chains, n_keep=args.sweeps,
                            burn=args.burn, seed=7, mode=mode, chunk_sweeps=args.chunk, gram_kernel=args.kernel)
i = smp.info
print(f"sweep kernels {i.sweep_ms:.1f} ms in {i.sweep_launches} launches; device total {i.total_ms:.1f} ms; "
      f"changes/chain-sweep {i.n_changes / i.n_sweeps_total:.1f}; windows {i.n_windows / i.n_sweeps_total:.1f}")
