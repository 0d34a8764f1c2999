"""The one-off Gram build X'X (fp64) next to the library baseline it has to beat: cuBLAS DGEMM through
torch.matmul on the same shape, and the fp64 peaks probed in the same process.

    python tools/gram_vs_cublas.py            # C2 shape: n=1000, p=10000
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyblip_b200 import _lib  # noqa: E402

X, y, _ = bench.make_data()
n, p = X.shape
ctx = _lib.Context.get()
flop = 2.0 * n * p * p
print(f"shape n={n} p={p}: full product 2np^2 = {flop / 1e9:.0f} GFLOP (the symmetric half is {flop / 2e9:.0f})")
print(f"probed peaks: DFMA {ctx.probe_fp64(0):.1f} TFLOP/s, DMMA m8n8k4 {ctx.probe_fp64(1):.1f} TFLOP/s")
times = []
for rep in range(4):
    d = _lib.Design.upload(X, ctx)
    ctx.timing(reset=True)
    d.build_gram()
    times.append(ctx.timing()["other_ms"])
    if rep == 3:
        G = d.gram(0, 64)
    d.close()
ms = min(times[1:])
print(f"gram_dmma_kernel (upper triangle computed, mirrored): {ms:.2f} ms = {flop / 2 / ms / 1e9:.1f} TFLOP/s executed, "
      f"{flop / ms / 1e9:.1f} TFLOP/s full-product equivalent")

import torch  # noqa: E402
Xt = torch.from_numpy(X).cuda()
XT = Xt.t().contiguous()
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Gt = XT @ Xt
    e1.record()
    torch.cuda.synchronize()
    if rep:
        best = min(best, e0.elapsed_time(e1))
print(f"cuBLAS DGEMM via torch.matmul (X^T X, full product): {best:.2f} ms = {flop / best / 1e9:.1f} TFLOP/s")
err = np.max(np.abs(Gt[:64].cpu().numpy() - G) / np.maximum(np.abs(G), 1e-300))
print(f"max rel difference of the first 64 rows, ours vs cuBLAS: {err:.2e}")
print(f"ratio cuBLAS / ours: {best / ms:.2f}x")
