"""count_pairs on C2-shaped device samples (512 chains x 1000 kept sweeps x 10 000 columns): the tcgen05 integer
GEMM (pairs_tc.cuh) against the scalar popc kernel, whole width and the prefiltered widths all_cand_groups uses.
    python tools/pairs_ab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab

X, y, _ = bench.make_data()
ctx = _lib.Context.get(0)
design = _lib.Design.upload(X, ctx)
design.build_gram()
cfg = LinearSpikeSlab(X, y)._config()
smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=1000, burn=100, seed=3, mode=_lib.MODE_GRAM)
bits = smp.bits()
rows, p = smp.info.rows, X.shape[1]
counts = bits.count_columns()
order = np.argsort(-counts)
print(f"{rows} rows x {p} columns, {int((counts > 0).sum())} columns ever active, {counts.sum() / rows:.1f} active per row")
for width in (512, 1024, 2048, 4096, p):
    sel = bits if width == p else bits.select_columns(np.sort(order[:width]).astype(np.int64))
    res = {}
    for name, env in (("tensor cores", "1"), ("scalar popc", "0")):
        if env == "0" and width > 4096:
            continue
        os.environ["BLIPGPU_PAIRS_TC"] = env
        out = sel.count_pairs()                     # first call builds the column-major copy
        best = 1e9
        for _ in range(3):
            ctx.timing(reset=True)
            out = sel.count_pairs()
            best = min(best, ctx.timing()["other_ms"])
        res[name] = (best, out)
    os.environ.pop("BLIPGPU_PAIRS_TC", None)
    line = f"  {width:6d} columns: " + ", ".join(f"{k} {v[0]:.2f} ms" for k, v in res.items())
    tc = res["tensor cores"][0]
    macs = rows * width * width / 2
    line += f"; tensor cores: {2 * macs / tc / 1e9:.0f} Tops/s over the upper triangle"
    if len(res) == 2:
        line += "; equal" if np.array_equal(res["tensor cores"][1], res["scalar popc"][1]) else "; DIFFERENT"
    print(line, flush=True)
    if sel is not bits:
        sel.close()
