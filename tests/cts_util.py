"""Loader for the continuous-space golden vectors (tests/golden/cts_*)."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dec(v):
    if isinstance(v, dict):
        return {(k if k == 'pip' else int(k)): float.fromhex(x) for k, x in v.items()}
    return float.fromhex(v)


def load():
    data = np.load(os.path.join(GOLDEN, "cts_cases.npz"))
    with gzip.open(os.path.join(GOLDEN, "cts_expected.json.gz"), "rt") as f:
        expected = json.load(f)
    runs = []
    for name, rs in expected.items():
        for r in rs:
            grid = [10, 100, 1000] if r["grid"] == "small" else data["grid_default"]
            want = {tuple(float.fromhex(x) for x in k): _dec(v) for k, v in r["out"]}
            runs.append(dict(name=name, locs=data["locs_" + name], grid=grid, shape=r["shape"],
                             count_signals=r["count_signals"], max_pep=r["max_pep"],
                             extra=data["extra_" + name] if r["extra"] else None, want=want))
    return runs
