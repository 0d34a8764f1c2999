"""Probit residual sweep on BASELINE configs[2] shape (n=2000, p=5000, 256 chains): kernel time and the
driver for ncu captures of gibbs_residual_kernel<probit>.   python tools/probit_profile.py [sweeps] [burn]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyblip_b200 import _lib
from pyblip_b200.probit import ProbitSpikeSlab

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(1003)
n, p = 2000, 5000
Z = rng.standard_normal((n, p))
X = np.empty((n, p))
X[:, 0] = Z[:, 0]
for j in range(1, p):
    X[:, j] = 0.5 * X[:, j - 1] + np.sqrt(0.75) * Z[:, j]
b = np.zeros(p)
b[rng.choice(p, 25, replace=False)] = rng.standard_normal(25)
z = (X @ b + rng.standard_normal(n) < 0).astype(np.int64)
d = _lib.Design.upload(X)
cfg = ProbitSpikeSlab(X, z.astype(float))._config()
smp = _lib.sample_spikeslab(d, cfg, z=z, probit=True, chains=256, n_keep=S, burn=B, seed=3, mode=_lib.MODE_RESIDUAL,
                            keep_latents=False, chunk_sweeps=16)
i = smp.info
print(f"probit n={n} p={p} 256 chains x {S + B} sweeps: sweep kernels {i.sweep_ms:.1f} ms in {i.sweep_launches} launches = "
      f"{256 * p * (S + B) / (i.sweep_ms * 1e-3):.3e} updates/s; 8np bytes per chain-sweep: "
      f"{8.0 * n * p * i.n_sweeps_total / (i.sweep_ms * 1e-3) / 1e9:.0f} GB/s; changes/chain-sweep {i.n_changes / i.n_sweeps_total:.1f} "
      f"events {i.n_events / i.n_sweeps_total:.1f}")
