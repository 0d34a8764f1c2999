"""Where the end-to-end step (bench.py `e2e`) spends its time outside the sweep kernels: host clock around each
stage of LinearSpikeSlab(X, y).sample(...) + samples_csr() at BASELINE configs[1].   python tools/e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab
X, y, _ = bench.make_data()
ctx = _lib.Context.get(0)
for rep in range(4):
    t = [time.perf_counter()]
    lm = LinearSpikeSlab(X, y)
    d = lm._get_design(); ctx.synchronize(); t.append(time.perf_counter())
    d.build_gram(); ctx.synchronize(); t.append(time.perf_counter())
    lm.sample(N=5000, burn=100, chains=512, seed=50 + rep, mode="gram"); t.append(time.perf_counter())
    info = lm._samples.info
    indptr, indices, values = lm.samples_csr(); t.append(time.perf_counter())
    lm._release(); lm._design.close(); t.append(time.perf_counter())
    names = ["upload X", "build gram", "sample() call", "samples_csr()", "release"]
    print(f"rep {rep}: " + "  ".join(f"{n} {1e3 * (b - a):.1f}" for n, a, b in zip(names, t[:-1], t[1:]))
          + f"  | device {info.total_ms:.1f} kernels {info.sweep_ms:.1f}  nnz {info.nnz}  total {1e3 * (t[-1] - t[0]):.1f} ms", flush=True)
