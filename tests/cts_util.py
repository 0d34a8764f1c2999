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


def load_groups():
    """Golden runs of the reference's grid_peps_to_cand_groups (oracle/make_cts_groups_golden.py)."""
    with gzip.open(os.path.join(GOLDEN, "cts_groups.json.gz"), "rt") as f:
        runs = json.load(f)
    for r in runs:
        peps = {}
        for k, v in r["peps"]:
            key = tuple(float.fromhex(x) for x in k)
            if isinstance(v, list):
                peps[key] = {(kk if kk == "pip" else int(kk)): float.fromhex(vv) for kk, vv in v}
            else:
                peps[key] = float.fromhex(v)
        r["peps"] = peps
    return runs


def encode_groups(all_cgs):
    """Same encoding as the golden file: exact (hex) floats, sorted sets."""
    res = []
    for comp in all_cgs:
        lst = []
        for cg in comp:
            data = {}
            for k, v in cg.data.items():
                if k == "center":
                    data[k] = [float(x).hex() for x in v]
                elif k == "nsignals":
                    data[k] = sorted(int(x) for x in v)
                else:
                    data[k] = float(v).hex()
            lst.append(dict(group=sorted(int(x) for x in cg.group), pep=float(cg.pep).hex(), data=data))
        res.append(lst)
    return res


def numpy_overlap_bits(centers, radii, circle, ctx=None):
    """The reference's own float32 expressions (create_groups_cts.py:291-308), bit-packed like the device."""
    centers = np.asarray(centers, dtype=np.float32)
    radii = np.asarray(radii, dtype=np.float32)
    d, G = centers.shape
    deltas = radii.reshape(-1, 1) + radii.reshape(1, -1)
    if circle:
        dists = np.sqrt(np.power(centers.reshape(d, G, 1) - centers.reshape(d, 1, G), 2).sum(axis=0))
        cons = dists < deltas
    else:
        cons = np.ones((G, G), dtype=bool)
        for j in range(d):
            cons = cons & (np.abs(centers[j].reshape(-1, 1) - centers[j].reshape(1, -1)) < deltas)
    pad = (-G) % 32
    packed = np.packbits(np.pad(cons, ((0, 0), (0, pad))), axis=1, bitorder='little')
    return np.ascontiguousarray(packed).view(np.uint32)
