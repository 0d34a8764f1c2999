"""Global-q gram kernel (p too large for q in shared memory) on an LD-block design: sweep-kernel
time per chain-sweep and algorithmic bandwidth.   python tools/c5_probe.py [p] [n] [chains] [sweeps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab

p = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 296
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 48
rng = np.random.default_rng(1005)
blk = 100
X = np.empty((n, p))
for b0 in range(0, p, blk):
    rho = min(rng.beta(5, 1), 0.99)
    Z = rng.standard_normal((n, blk))
    X[:, b0] = Z[:, 0]
    for j in range(1, blk):
        X[:, b0 + j] = rho * X[:, b0 + j - 1] + np.sqrt(1 - rho * rho) * Z[:, j]
beta = np.zeros(p)
beta[rng.choice(p, 50, replace=False)] = rng.standard_normal(50)
y = X @ beta + rng.standard_normal(n)
design = _lib.Design.upload(X)
design.build_gram()
cfg = LinearSpikeSlab(X, y)._config()
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=chains, n_keep=sweeps, burn=sweeps, seed=7, mode=_lib.MODE_GRAM)
i = smp.info
ch = i.n_changes / i.n_sweeps_total
gbs = (8.0 * p * i.n_changes + 24.0 * p * i.n_sweeps_total) / (i.sweep_ms * 1e-3) / 1e9
print(f"p={p} n={n} chains={chains}: sweep kernels {i.sweep_ms:.1f} ms for {2 * sweeps} sweeps; changes/chain-sweep {ch:.1f}; "
      f"{chains * p * 2 * sweeps / (i.sweep_ms * 1e-3):.3e} updates/s; algorithmic {gbs:.0f} GB/s (8p per change + 24p per sweep)")
