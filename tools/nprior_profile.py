"""Neuronized-prior sampler on one GPU: device time of the sweep kernel next to the wall time of
the API call and of the dense read-back (also the driver for ncu captures of nprior_kernel).

    python tools/nprior_profile.py [chains] [N] [burn]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyblip_b200 import _lib  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
burn = int(sys.argv[3]) if len(sys.argv) > 3 else 100
X, _, _ = bench.make_data()
n, p = X.shape[0], 2000
Xn = np.ascontiguousarray(X[:, :p])
rng = np.random.default_rng(7)
b = np.zeros(p)
b[rng.choice(p, 20, replace=False)] = rng.standard_normal(20)
y = Xn @ b + rng.standard_normal(n)
ctx = _lib.Context.get(0)
d = _lib.Design.upload(Xn, ctx)
cfg = _lib.NPriorConfig(1.0, 1 - 1 / p, 1e-10, 1, 5.0, 1.0, 0, 1, 1, 0)
_lib.sample_nprior(d, cfg, y, chains=chains, n_keep=4, burn=4, seed=1).close()      # Gram build, allocator
t0 = time.perf_counter()
res = _lib.sample_nprior(d, cfg, y, chains=chains, n_keep=N, burn=burn, seed=2)
t1 = time.perf_counter()
betas = res.get("betas")
t2 = time.perf_counter()
res.get("alphas"); res.get("ws")
t3 = time.perf_counter()
act = (betas != 0).sum(1)
print(f"nprior n={n} p={p} {chains} chains x {N + burn} sweeps: call {t1 - t0:.3f} s (sweep kernels "
      f"{res.sweep_ms / 1e3:.3f} s = {1e3 * res.sweep_ms / (N + burn):.1f} us/sweep-step), betas D2H {t2 - t1:.3f} s "
      f"({betas.nbytes / 1e9:.2f} GB), alphas+ws D2H {t3 - t2:.3f} s; kernel-only {chains * p * (N + burn) / (res.sweep_ms * 1e-3):.3e} "
      f"updates/s; active/sample mean {act.mean():.1f} max {act.max()}")
res.close()
d.close()
