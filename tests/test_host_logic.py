"""Host-side logic of pyblip_b200.create_groups on a CPU box: thresholds, pre-narrowing,
tree -> groups, index remapping.  The device is replaced by tests/fake_device.py (NumPy),
so these tests check everything around the kernels against the reference's golden outputs."""
import numpy as np
import pytest

import blip_testutil as util
import fake_device
from pyblip_b200 import create_groups as cg, dist


@pytest.fixture(autouse=True)
def _numpy_device(monkeypatch):
    fake_device.install(monkeypatch)


@pytest.mark.parametrize("name", util.COUNTING_CASES)
def test_group_functions_match_reference_golden_on_cpu(name):
    case = util.load_counting_case(name)
    samples = case["samples"]
    for key, fn in [("seq", cg.sequential_groups), ("hier", cg.hierarchical_groups),
                    ("all", cg.all_cand_groups)]:
        for k in [k for k in case if k.startswith(key + "_") and not k.endswith("_kwargs")]:
            kwargs = case[k + "_kwargs"]
            got = util.cg_list(fn(samples.copy(), **kwargs))
            want = util.golden_groups(case[k])
            assert sorted(got) == sorted(want), f"{name}/{k}"
            assert got == want, f"{name}/{k}: output order differs from the reference"


def test_reference_known_answers():
    """test/test_create_groups.py:156-212 and :266-300 of the reference."""
    samples = util.load_counting_case("hand")["samples"]
    assert len(cg.sequential_groups(samples, max_size=4, max_pep=1)) == 13
    groups = [x.group for x in cg.sequential_groups(samples, prenarrow=True, q=0.9, max_pep=0.5)]
    assert {0} in groups and {0, 1} not in groups
    groups = [x.group for x in cg.hierarchical_groups(samples, filter_sequential=False, max_pep=1)]
    for expect in [{0}, {1}, {3}, {4}, {0, 1}, {0, 1, 2, 3, 4}]:
        assert expect in groups
    for max_pep in [0.2, 0.4, 0.6, 0.8, 1]:
        out = cg.hierarchical_groups(samples, max_size=4, max_pep=max_pep)
        assert max(x.pep for x in out) < max_pep
        assert len({tuple(sorted(x.group)) for x in out}) == len(out)


def test_all_cand_groups_properties():
    """test/test_create_groups.py:214-264: uniqueness, purity filter, all-zero samples."""
    rng = np.random.default_rng(0)
    n, p = 300, 60
    samples = rng.binomial(1, rng.beta(0.4, 0.4, size=p), size=(n, p))
    X = rng.standard_normal((n, p))
    for j in range(1, p):
        X[:, j] = 0.9 * X[:, j - 1] + np.sqrt(1 - 0.81) * X[:, j]
    cgs = cg.all_cand_groups(samples=samples, X=X, prenarrow=False, max_pep=1)
    keys = [tuple(sorted(c.group)) for c in cgs]
    assert len(keys) == len(set(keys))
    hatC = np.abs(np.corrcoef(X.T))
    cgs = cg.all_cand_groups(samples=samples, X=X, prenarrow=True, max_pep=0.1, min_purity=0.8)
    for c in cgs:
        if len(c.group) > 1:
            g = np.array(list(c.group))
            assert hatC[g][:, g].min() >= 0.8
    assert cg.all_cand_groups(np.zeros((n, p)), X=X, prenarrow=False, max_pep=0.5) == []


def test_susie_branch_and_errors():
    alphas = np.array([[0.9, 0.1, 0.0, 0.0], [0.0, 0.0, 0.5, 0.5]])
    out = cg.sequential_groups(susie_alphas=alphas, max_pep=1)
    peps = {tuple(sorted(c.group)): c.pep for c in out}
    np.testing.assert_allclose(peps[(0,)], 0.1, rtol=1e-12)
    np.testing.assert_allclose(peps[(2, 3)], 1e-15, rtol=1e-6)
    with pytest.raises(ValueError):
        cg.sequential_groups()
    with pytest.raises(ValueError):
        cg.filter_by_purity([cg.CandidateGroup({0, 1}, 0.1)], min_purity=0.5)


def test_candidate_group_to_dict():
    c = cg.CandidateGroup([3, 1], 0.25, data={"blip-group": {7, 8}, "purity": 1})
    d = c.to_dict()
    assert sorted(d["group"]) == [1, 3] and d["pep"] == 0.25 and sorted(d["blip-group"]) == [7, 8]


def test_shard_chains_partition():
    for chains in (1, 7, 10, 512, 513):
        for world in (1, 2, 3, 8):
            parts = [dist.shard_chains(chains, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == chains
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_correlation_from_integer_pair_counts(monkeypatch):
    """`_samples_dist_matrix_device` (exact integer co-occurrence counts) reproduces the reference's
    np.corrcoef-based `_samples_dist_matrix` (create_groups.py:523-535) to rounding, including constant
    columns (the appended all-zero / all-one rows make every column non-constant)."""
    import fake_device
    from pyblip_b200 import create_groups as cg
    fake_device.install(monkeypatch)
    rng = np.random.default_rng(3)
    S = rng.random((400, 30)) < 0.15
    S[:, 4] = True
    S[:, 5] = False
    S[:, 7] = S[:, 6]
    dev = cg._as_device(S)
    d_int = cg._samples_dist_matrix_device(dev)
    d_ref = cg._samples_dist_matrix(S)
    assert np.max(np.abs(d_int - d_ref)) < 1e-12
    assert d_int[6, 7] == 2.0 and np.allclose(np.diag(d_int), 2.0)
    # the size switch: small problems keep the host path (parity with the reference's tie-breaking)
    monkeypatch.setattr(cg, "_DEVICE_CORR_MIN_ENTRIES", 10 ** 9)
    assert np.array_equal(cg._sample_dist_matrix_auto(dev), d_ref)
    monkeypatch.setattr(cg, "_DEVICE_CORR_MIN_ENTRIES", 0)
    assert np.array_equal(cg._sample_dist_matrix_auto(dev), d_int)
