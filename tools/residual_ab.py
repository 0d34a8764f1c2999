"""Residual-form sweeps timed for one build of the library (A/B of kernel variants):
    BLIPGPU_VARIANT=name python tools/residual_ab.py      # pyblip_b200/_C/variants/name.so, else the in-tree library
Linear C2 (n=1000, p=10000, 512 chains) and probit C3 (n=2000, p=5000, 256 chains), stationary launches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyblip_b200 import _lib
v = os.environ.get("BLIPGPU_VARIANT")
if v:
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "variants", v + ".so")
from pyblip_b200.linear import LinearSpikeSlab
from pyblip_b200.probit import ProbitSpikeSlab


def ar1(n, p, rho, rng):
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + np.sqrt(1 - rho * rho) * Z[:, j]
    return X


rng = np.random.default_rng(1003)
n, p = 1000, 10000
X = ar1(n, p, 0.95, rng)
b = np.zeros(p)
b[rng.choice(p, 50, replace=False)] = rng.standard_normal(50)
y = X @ b + rng.standard_normal(n)
d = _lib.Design.upload(X)
cfg = LinearSpikeSlab(X, y)._config()
for rep in range(2):
    smp = _lib.sample_spikeslab(d, cfg, y=y, chains=512, n_keep=32, burn=32, seed=3, mode=_lib.MODE_RESIDUAL)
    i = smp.info
    print(f"[{v or 'in-tree'}] linear C2 residual: {i.sweep_ms:.1f} ms for 64 sweeps; 8np bytes per chain-sweep "
          f"{8.0 * n * p * i.n_sweeps_total / (i.sweep_ms * 1e-3) / 1e9:.0f} GB/s", flush=True)
    chk = float(np.abs(smp.dense()).sum())
    smp.close()
d.close()
print("   checksum", repr(chk))
n, p = 2000, 5000
X = ar1(n, p, 0.5, rng)
b = np.zeros(p)
b[rng.choice(p, 25, replace=False)] = rng.standard_normal(25)
z = (X @ b + rng.standard_normal(n) < 0).astype(np.int64)
d = _lib.Design.upload(X)
cfg = ProbitSpikeSlab(X, z.astype(float))._config()
for rep in range(2):
    smp = _lib.sample_spikeslab(d, cfg, z=z, probit=True, chains=256, n_keep=32, burn=64, seed=3,
                                mode=_lib.MODE_RESIDUAL, keep_latents=False)
    i = smp.info
    print(f"[{v or 'in-tree'}] probit C3 residual: {i.sweep_ms:.1f} ms for 96 sweeps; 8np bytes per chain-sweep "
          f"{8.0 * n * p * i.n_sweeps_total / (i.sweep_ms * 1e-3) / 1e9:.0f} GB/s; events/chain-sweep "
          f"{i.n_events / i.n_sweeps_total:.1f}", flush=True)
    chk = float(np.abs(smp.dense()).sum())
    smp.close()
print("   checksum", repr(chk))
