"""Secondary rows of the hot path (probit, block sampler, neuronized prior) through the
reference-facing API with host buffers on one GPU, next to the reference's own compiled kernel
on the box's host cores (bench.py's cpu-baseline leg, one chain per core, bounded sample).
configs[1] itself is bench.py.  The change-point configuration has no CPU leg: the reference
materialises the dense 20000 x 20000 design (3.2 GB, copied again inside every worker).

    python tools/secondary_vs_cpu.py > profiles/r01_secondary_vs_cpu.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyblip_b200 import LinearSpikeSlab, NPrior, ProbitSpikeSlab  # noqa: E402


def row(name, chains, p, sweeps, wall, cpu):
    dev = chains * p * sweeps / wall
    print(f"{name}\n    device, API with host buffers: {chains} chains x {sweeps} sweeps in {wall:.2f} s = {dev:.3e} "
          f"coordinate updates/s\n    reference on {cpu['cores']} host cores ({cpu['sample']}): {cpu['value']:.3e} "
          f"coordinate updates/s\n    ratio {dev / cpu['value']:.0f}x", flush=True)


which = set(sys.argv[1:]) or {"block", "probit", "nprior"}
X, y, _ = bench.make_data()
n, p = X.shape

# ---- block sampler bsize=3 on the C2 design ------------------------------------------------------
if "block" in which:
    lm = LinearSpikeSlab(X, y)
    lm.sample(N=8, burn=8, chains=512, bsize=3, seed=1)            # warm-up: Gram build, allocator
    lm._release()
    t0 = time.perf_counter()
    lm.sample(N=200, burn=50, chains=512, bsize=3, seed=2)
    wall = time.perf_counter() - t0
    lm._release(); lm._design.close()
    row("block sampler bsize=3, n=1000 p=10000, 512 chains", 512, p, 250, wall,
        bench.reference_throughput(X, y, 200, kind="block", bsize=3))

# ---- probit C3 -----------------------------------------------------------------------------------
if "probit" in which:
    rng = np.random.default_rng(1003)
    n3, p3 = 2000, 5000
    Z = rng.standard_normal((n3, p3))
    X3 = np.empty((n3, p3))
    X3[:, 0] = Z[:, 0]
    for j in range(1, p3):
        X3[:, j] = 0.5 * X3[:, j - 1] + np.sqrt(0.75) * Z[:, j]
    b3 = np.zeros(p3)
    b3[rng.choice(p3, 25, replace=False)] = rng.standard_normal(25)
    z3 = (X3 @ b3 + rng.standard_normal(n3) < 0).astype(float)
    pm = ProbitSpikeSlab(X3, z3)
    pm.sample(N=8, burn=8, chains=256, seed=1)
    pm._release()
    t0 = time.perf_counter()
    pm.sample(N=500, burn=100, chains=256, seed=2)
    wall = time.perf_counter() - t0
    pm._release(); pm._design.close()
    row("probit C3, n=2000 p=5000, 256 chains", 256, p3, 600, wall,
        bench.reference_throughput(X3, z3, 100, kind="probit"))

# ---- neuronized prior ----------------------------------------------------------------------------
if "nprior" in which:
    Xn = np.ascontiguousarray(X[:, :2000])
    rngn = np.random.default_rng(7)
    bn = np.zeros(2000)
    bn[rngn.choice(2000, 20, replace=False)] = rngn.standard_normal(20)
    yn = Xn @ bn + rngn.standard_normal(n)
    nm = NPrior(Xn, yn, tauw2=1.0)
    nm.sample(N=8, burn=8, chains=128, seed=1)
    t0 = time.perf_counter()
    nm.sample(N=500, burn=100, chains=128, seed=2)
    wall = time.perf_counter() - t0
    row("neuronized prior, n=1000 p=2000, 128 chains", 128, 2000, 600, wall,
        bench.reference_throughput(Xn, yn, 300, kind="nprior"))
