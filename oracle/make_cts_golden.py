"""Golden vectors for grid_peps from the REFERENCE's own implementation.
TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python oracle/make_cts_golden.py     # writes tests/golden/cts_cases.npz + cts_expected.json.gz
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("PYBLIP_REFERENCE", "/root/reference")

HAND = np.array([      # test/test_create_groups.py:424-430
    [[0.13054, 0.410234], [np.nan, np.nan]],
    [[0.12958, 0.406639], [0.46009, 0.459]],
    [[0.46001, 0.45119], [0.1302, 0.4126]],
    [[np.nan, np.nan], [0.95, 0.899]],
    [[np.nan, np.nan], [np.nan, np.nan]],
])
DEFAULT_GRID_SIZES = np.around(np.logspace(np.log10(4), 4, 25))      # blip.py:32


def enc(v):
    if isinstance(v, dict):
        return {str(k): float(x).hex() for k, x in v.items()}
    return float(v).hex()


def dump(peps):
    return [[[float(x).hex() for x in key], enc(val)] for key, val in peps.items()]


def random_locs(rng, N, n_src, d, n_true, jitter, p_missing):
    true = rng.random((n_true, d)) * 0.9 + 0.05
    locs = np.full((N, n_src, d), np.nan)
    for i in range(N):
        for k in range(n_src):
            if rng.random() < p_missing:
                continue
            t = true[rng.integers(n_true)]
            locs[i, k] = np.clip(t + jitter * rng.standard_normal(d), 0, 1)
    return locs


def main():
    cts = ref_loader.load_python("create_groups_cts")
    rng = np.random.default_rng(77)
    post = np.loadtxt(os.path.join(REF, "docs", "source", "data", "post_samples.txt"))
    post = post.reshape(-1, 7, 2)[:50]                                   # docs fixture, 996 x 7 x 2
    post, _, _ = cts.normalize_locs(post)
    cases = {
        "hand": HAND,
        "rand2d": random_locs(rng, 40, 5, 2, 4, 0.004, 0.3),
        "rand1d": random_locs(rng, 80, 5, 1, 3, 0.002, 0.2),
        "rand3d": random_locs(rng, 40, 4, 3, 3, 0.01, 0.25),
        "edges": np.clip(random_locs(rng, 25, 4, 2, 5, 0.2, 0.1), 0, 1),    # many points clipped to 0 / 1
        "post": post,
    }
    extra = {"hand": np.array([[0.13, 0.41], [0.46, 0.455], [0.5, 0.5]]),
             "rand2d": rng.random((5, 2)), "rand1d": rng.random((4, 1))}
    expected = {}
    for name, locs in cases.items():
        d = locs.shape[2]
        runs = []
        for shape in (["square", "circle"] if d == 2 else ["square"]):
            for gs_name, gs in [("small", [10, 100, 1000]), ("default", DEFAULT_GRID_SIZES)]:
                for count_signals in (False, True):
                    for use_extra in ([False, True] if name in extra else [False]):
                        for max_pep in ((1, 0.25) if name == "hand" else (1,)):
                            kw = dict(grid_sizes=gs, count_signals=count_signals, max_pep=max_pep, shape=shape,
                                      extra_centers=extra[name] if use_extra else None)
                            peps = cts.grid_peps(locs, **kw)
                            runs.append(dict(shape=shape, grid=gs_name, count_signals=count_signals,
                                             extra=use_extra, max_pep=max_pep, out=dump(peps)))
        expected[name] = runs
        print(f"[golden] cts {name}: locs {locs.shape}, {len(runs)} runs, "
              f"{sum(len(r['out']) for r in runs)} keys")
    np.savez_compressed(os.path.join(GOLDEN, "cts_cases.npz"),
                        **{"locs_" + k: v for k, v in cases.items()},
                        **{"extra_" + k: v for k, v in extra.items()},
                        grid_default=DEFAULT_GRID_SIZES)
    import gzip
    with gzip.open(os.path.join(GOLDEN, "cts_expected.json.gz"), "wt") as f:
        json.dump(expected, f)


if __name__ == "__main__":
    main()
