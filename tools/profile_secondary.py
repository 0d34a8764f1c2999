"""Device timings of the kernels next to the headline sweep: Gram build, counting kernels,
residual / probit sweep, block sampler, neuronized prior.  CUDA events on the library's stream
(ctx.timing / samples info), inputs resident in HBM.  Prints one line per kernel with the
algorithmic bytes or flops it is judged on.

    python tools/profile_secondary.py > profiles/r01_secondary_kernels.txt
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyblip_b200 import _lib  # noqa: E402
from pyblip_b200.linear import LinearSpikeSlab  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
ctx = _lib.Context.get(0)


def line(name, ms, what, value, unit, frac=None):
    extra = f"  = {frac:.2f} of measured HBM {HBM:.0f} GB/s" if frac is not None else ""
    print(f"{name:34s} {ms:10.3f} ms   {what}: {value:10.1f} {unit}{extra}", flush=True)


# ---- Gram build (C2 design) ------------------------------------------------------------------
X, y, _ = bench.make_data()
n, p = X.shape
design = _lib.Design.upload(X, ctx)
ctx.timing(reset=True)
design.build_gram()
ctx.synchronize()
t = ctx.timing()
line("gram_dmma_kernel n=1000 p=10000", t["other_ms"], "2np^2 flop fp64 (upper triangle computed once)",
     2.0 * n * p * p / (t["other_ms"] * 1e-3) / 1e12, "TFLOP/s-equivalent")

# ---- counting kernels on C2-shaped samples (512 chains x 1000 sweeps) --------------------------
cfg = LinearSpikeSlab(X, y)._config()
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=1000, burn=100, seed=3, mode=_lib.MODE_GRAM)
rows = smp.info.rows
bits = smp.bits()
bits_bytes = ((p + 31) // 32) * 4 * rows
rng = np.random.default_rng(0)
groups = [list(rng.choice(p, size=rng.integers(2, 26), replace=False)) for _ in range(2000)]
group_bytes = rows // 8 * sum(len(g) for g in groups)      # one bit per (row, member) from the column-major copy
for name, fn, nbytes, what in (
        ("count_columns", lambda: bits.count_columns(), bits_bytes + 8 * p, "N*p/8 bytes read once"),
        ("count_sequential(max_size=25)", lambda: bits.count_sequential(25), bits_bytes + 8 * 25 * p,
         "N*p/8 bytes read once"),
        ("count_groups(2000 groups of <=25)", lambda: bits.count_groups(groups), group_bytes,
         "N/8 bytes per group member (column-major copy, built by the first call)")):
    t0 = time.perf_counter()
    fn()
    ctx.synchronize()
    first_ms = 1e3 * (time.perf_counter() - t0)
    ctx.timing(reset=True)
    fn()
    ctx.synchronize()
    t = ctx.timing()
    gbs = nbytes / (t["other_ms"] * 1e-3) / 1e9
    line(f"{name} rows={rows}", t["other_ms"], what, gbs, "GB/s", gbs / HBM)
    if name.startswith("count_groups"):
        print(f"    first call {first_ms:.2f} ms wall (includes the transposition of the {bits_bytes / 1e9:.2f} GB bit matrix and the H2D of the group lists)")
# pair (co-occurrence) counts on the 1024 columns with the largest marginal counts, as all_cand_groups selects them
cc = bits.count_columns()
sub = bits.select_columns(np.sort(np.argsort(-cc)[:1024]))
sub.count_pairs()
ctx.timing(reset=True)
sub.count_pairs()
ctx.synchronize()
t = ctx.timing()
line(f"count_pairs 1024 columns rows={rows}", t["other_ms"], "N*1024/8 bytes read once, 1024^2/2 pair counts",
     rows * 1024 / 8 / (t["other_ms"] * 1e-3) / 1e9, "GB/s")
sub.close()
bits.close()
smp.close()

# ---- residual sweep, linear (C2) --------------------------------------------------------------
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=32, burn=32, seed=3, mode=_lib.MODE_RESIDUAL)
i = smp.info
gbs = 8.0 * n * p * i.n_sweeps_total / (i.sweep_ms * 1e-3) / 1e9
line("gibbs_residual_kernel<linear> C2", i.sweep_ms / i.sweep_launches, "8np bytes per chain-sweep", gbs, "GB/s", gbs / HBM)
print(f"    coordinate updates/s: {512 * p * 64 / (i.total_ms * 1e-3):.3e}")
smp.close()

# ---- block sampler, bsize=3, gram and residual forms (C2) ----------------------------------------
for mode, nm in ((_lib.MODE_GRAM, "gram"), (_lib.MODE_RESIDUAL, "residual")):
    smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=32, burn=32, seed=3, mode=mode, bsize=3)
    i = smp.info
    print(f"gibbs_block_kernel<{nm}> bsize=3 C2        {i.sweep_ms / i.sweep_launches:10.3f} ms/launch   "
          f"coordinate updates/s: {512 * p * 64 / (i.total_ms * 1e-3):.3e}   blocks applied/chain-sweep "
          f"{i.n_changes / i.n_sweeps_total:.1f}", flush=True)
    smp.close()
design.close()

# ---- probit residual sweep (C3: n=2000, p=5000, 256 chains) ------------------------------------
rng = np.random.default_rng(1003)
n3, p3 = 2000, 5000
Z = rng.standard_normal((n3, p3))
X3 = np.empty((n3, p3))
X3[:, 0] = Z[:, 0]
for j in range(1, p3):
    X3[:, j] = 0.5 * X3[:, j - 1] + np.sqrt(0.75) * Z[:, j]
b3 = np.zeros(p3)
b3[rng.choice(p3, 25, replace=False)] = rng.standard_normal(25)
z3 = (X3 @ b3 + rng.standard_normal(n3) < 0).astype(np.int64)
d3 = _lib.Design.upload(X3, ctx)
smp = _lib.sample_spikeslab(d3, cfg, z=z3, probit=True, chains=256, n_keep=32, burn=64, seed=3,
                            mode=_lib.MODE_RESIDUAL, keep_latents=False)
i = smp.info
gbs = 8.0 * n3 * p3 * i.n_sweeps_total / (i.sweep_ms * 1e-3) / 1e9
line("gibbs_residual_kernel<probit> C3", i.sweep_ms / i.sweep_launches, "8np bytes per chain-sweep", gbs, "GB/s", gbs / HBM)
print(f"    coordinate updates/s: {256 * p3 * 96 / (i.total_ms * 1e-3):.3e}   latent redraw events/chain-sweep "
      f"{i.n_events / i.n_sweeps_total:.1f}")
smp.close()
d3.close()

# ---- neuronized prior (n=1000, p=2000, 128 chains) ----------------------------------------------
Xn = X[:, :2000].copy()
rngn = np.random.default_rng(7)
bn = np.zeros(2000)
bn[rngn.choice(2000, 20, replace=False)] = rngn.standard_normal(20)
yn = Xn @ bn + rngn.standard_normal(n)
dn = _lib.Design.upload(Xn, ctx)
ncfg = _lib.NPriorConfig(1.0, 1 - 1 / 2000, 1e-10, 1, 5.0, 1.0, 0, 1, 1, 0)
t0 = time.perf_counter()
res = _lib.sample_nprior(dn, ncfg, yn, chains=128, n_keep=50, burn=10, seed=1)
ctx.synchronize()
wall = time.perf_counter() - t0
act = (res.get("betas") != 0).sum(1)
print(f"nprior sweep n=1000 p=2000 128 chains: wall {wall * 1e3:.1f} ms for 60 sweeps   "
      f"coordinate updates/s: {128 * 2000 * 60 / wall:.3e}   active coordinates per sample: mean {act.mean():.1f} max {act.max()}")
res.close()
dn.close()
