"""Golden vectors for the counting path, produced by the REFERENCE's own
create_groups functions.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python oracle/make_counting_golden.py      # writes tests/golden/counting_*.npz
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

HAND = np.array([          # test/test_create_groups.py:159-165
    [0, 1, 0, 0, 1],
    [1, 1, 0, 1, 1],
    [1, 0, 0, 0, 1],
    [0, 1, 0, 1, 1],
    [1, 0, 0, 1, 1],
])


def dump(cgs):
    """list[CandidateGroup] -> JSON list of [sorted group, pep hex] in output order."""
    return json.dumps([[sorted(int(j) for j in cg.group), float(cg.pep).hex()] for cg in cgs])


def sparse_samples(rng, N, p, n_hot, hot_rate=0.7, bg_rate=0.004, smear=0.25):
    """Indicator matrix that looks like sampler output: a few persistent signals,
    mass smeared to neighbours, sparse background."""
    S = rng.random((N, p)) < bg_rate
    hot = rng.choice(p, size=n_hot, replace=False)
    for h in hot:
        on = rng.random(N) < hot_rate
        shift = (rng.random(N) < smear) * rng.integers(-2, 3, size=N)
        cols = np.clip(h + shift, 0, p - 1)
        S[np.arange(N)[on], cols[on]] = True
    return S


def main():
    cg = ref_loader.load_python("create_groups")
    rng = np.random.default_rng(2026)
    cases = {
        "hand": HAND,
        "sparse_a": sparse_samples(rng, 400, 130, 6),
        "sparse_b": sparse_samples(rng, 1000, 77, 4, hot_rate=0.5),
        "wide": sparse_samples(rng, 257, 300, 9, hot_rate=0.9, smear=0.5),
        "float_vals": sparse_samples(rng, 300, 64, 5) * rng.standard_normal((300, 64)),
    }
    for name, samples in cases.items():
        out = dict(samples=samples)
        kw = [
            ("seq_default", dict()),
            ("seq_size4_pep1", dict(max_size=4, max_pep=1)),
            ("seq_prenarrow", dict(prenarrow=True, q=0.1)),
            ("seq_prenarrow_q9", dict(prenarrow=True, q=0.9, max_pep=0.5)),
            ("seq_size40", dict(max_size=40, max_pep=0.9)),
        ]
        for key, k in kw:
            out[key] = dump(cg.sequential_groups(samples, **k))
            out[key + "_kwargs"] = json.dumps(k)
        hk = [
            ("hier_default", dict()),
            ("hier_pep1", dict(filter_sequential=False, max_pep=1)),
            ("hier_filter", dict(filter_sequential=True, max_size=10, max_pep=0.5)),
        ]
        for key, k in hk:
            out[key] = dump(cg.hierarchical_groups(samples, **k))
            out[key + "_kwargs"] = json.dumps(k)
        ak = [
            ("all_default", dict()),
            ("all_noprenarrow", dict(prenarrow=False, max_pep=1)),
            ("all_q", dict(q=0.1, max_pep=0.5, max_size=12)),
        ]
        for key, k in ak:
            out[key] = dump(cg.all_cand_groups(samples, **k))
            out[key + "_kwargs"] = json.dumps(k)
        np.savez_compressed(os.path.join(GOLDEN, f"counting_{name}.npz"), **out)
        print(f"[golden] counting_{name}: N,p={samples.shape} "
              f"seq={len(json.loads(out['seq_default']))} hier={len(json.loads(out['hier_default']))} "
              f"all={len(json.loads(out['all_default']))}")


if __name__ == "__main__":
    main()
