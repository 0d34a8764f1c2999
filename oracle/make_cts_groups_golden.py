"""Golden vectors for grid_peps_to_cand_groups from the REFERENCE's own implementation
(pyblip/create_groups_cts.py:228-384, with networkx and pyblip/ecc.py as installed here).
TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python oracle/make_cts_groups_golden.py     # writes tests/golden/cts_groups.json.gz
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.make_cts_golden import DEFAULT_GRID_SIZES, HAND, random_locs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def hx(v):
    return float(v).hex()


def enc_peps(peps):
    """dict key -> pep (or ordered inner dict), floats as hex, inner insertion order kept."""
    out = []
    for key, val in peps.items():
        k = [hx(x) for x in key]
        if isinstance(val, dict):
            out.append([k, [[("pip" if kk == "pip" else int(kk)), hx(vv)] for kk, vv in val.items()]])
        else:
            out.append([k, hx(val)])
    return out


def enc_groups(all_cgs):
    res = []
    for comp in all_cgs:
        lst = []
        for cg in comp:
            data = {}
            for k, v in cg.data.items():
                if k == "center":
                    data[k] = [hx(x) for x in v]
                elif k == "nsignals":
                    data[k] = sorted(int(x) for x in v)
                else:
                    data[k] = hx(v)
            lst.append(dict(group=sorted(int(x) for x in cg.group), pep=hx(cg.pep), data=data))
        res.append(lst)
    return res


def main():
    cts = ref_loader.load_python("create_groups_cts")
    rng = np.random.default_rng(99)
    cases = {
        "hand": HAND,
        "rand2d": random_locs(rng, 60, 6, 2, 5, 0.01, 0.3),
        "rand2d_dense": random_locs(rng, 30, 6, 2, 12, 0.03, 0.2),
        "rand1d": random_locs(rng, 80, 5, 1, 4, 0.004, 0.2),
        "rand3d": random_locs(rng, 40, 4, 3, 3, 0.01, 0.25),
    }
    runs = []
    for name, locs in cases.items():
        d = locs.shape[2]
        for shape in (["square", "circle"] if d == 2 else ["square"]):
            for count_signals in (False, True):
                for gs in ([10, 100, 1000], DEFAULT_GRID_SIZES[:12]):
                    peps = cts.grid_peps(locs, gs, count_signals=count_signals, max_pep=0.9, shape=shape)
                    if len(peps) == 0:
                        continue
                    for min_blip_size, max_pep, min_pep in ((1000, 0.25, 0.001), (3, 0.5, 0.05)):
                        cgs, comps = cts.grid_peps_to_cand_groups(peps, min_blip_size=min_blip_size, shape=shape,
                                                                  max_pep=max_pep, min_pep=min_pep)
                        runs.append(dict(name=name, shape=shape, count_signals=count_signals,
                                         min_blip_size=min_blip_size, max_pep=max_pep, min_pep=min_pep,
                                         peps=enc_peps(peps), components=[[int(x) for x in c] for c in comps],
                                         groups=enc_groups(cgs)))
        print(f"[golden] cts groups {name}: {sum(1 for r in runs if r['name'] == name)} runs, "
              f"max regions {max(len(r['peps']) for r in runs if r['name'] == name)}")
    with gzip.open(os.path.join(GOLDEN, "cts_groups.json.gz"), "wt") as f:
        json.dump(runs, f)


if __name__ == "__main__":
    main()
