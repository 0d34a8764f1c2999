"""Host-clock breakdown of the end-to-end path bench.py times (LinearSpikeSlab(X, y).sample + samples_csr)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyblip_b200.linear import LinearSpikeSlab  # noqa: E402

X, y, _ = bench.make_data()
for rep in range(3):
    t0 = time.perf_counter()
    lm = LinearSpikeSlab(X, y)
    lm.sample(N=5000, burn=100, chains=512, seed=4000 + rep, mode="gram")
    t1 = time.perf_counter()
    indptr, indices, values = lm.samples_csr()
    t2 = time.perf_counter()
    if rep == 0:                          # integrity of the large pipelined copy: CSR rows == dense rows
        import numpy as np
        rows = lm._samples.info.rows
        for r0 in (0, rows // 2 - 7, rows - 16):
            dense = lm._samples.dense(row0=r0, row1=r0 + 16)
            rec = np.zeros_like(dense)
            for k in range(16):
                a, b = indptr[r0 + k], indptr[r0 + k + 1]
                rec[k, indices[a:b]] = values[a:b]
            assert np.array_equal(rec, dense), r0
        print("CSR rows equal dense rows at the head, middle and tail")
    lm._release()
    lm._design.close()
    t3 = time.perf_counter()
    i = None
    print(f"rep {rep}: construct+sample {1e3 * (t1 - t0):.0f} ms   samples_csr {1e3 * (t2 - t1):.0f} ms "
          f"({(indices.nbytes + values.nbytes + indptr.nbytes) / 1e9:.2f} GB)   release {1e3 * (t3 - t2):.0f} ms", flush=True)
