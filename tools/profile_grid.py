"""grid_peps (continuous-space candidate regions) on the device: wall time of the reference-facing call
on synthetic posterior location samples.   python tools/profile_grid.py [N ...]
With --cpu the reference-equivalent NumPy/Python oracle (oracle/cts_oracle.py) is timed on a bounded N."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_locs(N, n_src=7, seed=0):
    rng = np.random.default_rng(seed)
    true = rng.random((5, 2))
    locs = np.full((N, n_src, 2), np.nan)
    present = rng.random((N, n_src)) < 0.7
    which = rng.integers(5, size=(N, n_src))
    pts = np.clip(true[which] + 0.004 * rng.standard_normal((N, n_src, 2)), 0, 1)
    locs[present] = pts[present]
    return locs


grid = np.around(np.logspace(np.log10(4), 4, 25))
if "--cpu" in sys.argv:
    from oracle import cts_oracle as co
    N = 2000
    locs = make_locs(N)
    t0 = time.perf_counter()
    out = co.grid_peps(locs, grid, max_pep=0.25, shape="circle")
    dt = time.perf_counter() - t0
    print(f"oracle (reference algorithm, Python loops) N={N}: {dt:.2f} s = {N / dt:.0f} samples/s, {len(out)} regions")
    sys.exit(0)
from pyblip_b200 import create_groups_cts as cts  # noqa: E402
cts.grid_peps(make_locs(100), grid, shape="circle")          # context / module load
for N in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1000, 100000, 1000000]:
    locs = make_locs(N)
    for shape in ("square", "circle"):
        t0 = time.perf_counter()
        out = cts.grid_peps(locs, grid, max_pep=0.25, shape=shape)
        dt = time.perf_counter() - t0
        print(f"grid_peps N={N} n_src=7 d=2 25 grid sizes shape={shape}: {dt * 1e3:.1f} ms = {N / dt:.3e} samples/s "
              f"({locs.nbytes / 1e6:.0f} MB of locations, {len(out)} regions kept)", flush=True)
