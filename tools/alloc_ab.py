"""Repeated sampler calls in one process: wall time of the call, device time, kernel time, time of close().
    python tools/alloc_ab.py            (BLIPGPU_NO_POOL=1 for the cudaMalloc / cudaFree path)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyblip_b200 import _lib
from pyblip_b200.linear import LinearSpikeSlab
X, y, _ = bench.make_data()
design = _lib.Design.upload(X)
design.build_gram()
cfg = LinearSpikeSlab(X, y)._config()
for rep in range(6):
    t0 = time.perf_counter()
    smp = _lib.sample_spikeslab(design, cfg, y=y, chains=512, n_keep=5000, burn=100, seed=10 + rep, mode=_lib.MODE_GRAM)
    t1 = time.perf_counter()
    i = smp.info
    smp.close()
    t2 = time.perf_counter()
    print(f"rep {rep}: call {1e3 * (t1 - t0):7.1f} ms  device {i.total_ms:7.1f}  kernels {i.sweep_ms:7.1f}  close {1e3 * (t2 - t1):6.1f} ms", flush=True)
