"""NumPy counting oracle vs the reference's golden outputs and known-answer tests."""
import numpy as np
import pytest

import blip_testutil as util
from oracle import counting_oracle as co

HAND = np.array([[0, 1, 0, 0, 1], [1, 1, 0, 1, 1], [1, 0, 0, 0, 1], [0, 1, 0, 1, 1], [1, 0, 0, 1, 1]])


def test_hand_matrix_known_answers():
    """test/test_create_groups.py:156-212 of the reference."""
    groups = co.sequential_groups(HAND, max_size=4, max_pep=1)
    assert len(groups) == 5 + 4 + 3 + 2 - 1
    for max_pep in [0.2, 0.4, 0.6, 0.8, 1]:
        g = co.sequential_groups(HAND, max_size=4, max_pep=max_pep)
        assert max(pep for _, pep in g) < max_pep
        assert len(set(t for t, _ in g)) == len(g)
    g = [t for t, _ in co.sequential_groups(HAND, prenarrow=True, q=0.9, max_pep=0.5)]
    assert (0,) in g and (0, 1) not in g


@pytest.mark.parametrize("name", util.COUNTING_CASES)
def test_sequential_groups_match_reference_golden(name):
    case = util.load_counting_case(name)
    for k in [k for k in case if k.startswith("seq_") and not k.endswith("_kwargs")]:
        got = co.sequential_groups(case["samples"], **case[k + "_kwargs"])
        want = util.golden_groups(case[k])
        assert sorted(got) == sorted(want), (name, k)


@pytest.mark.parametrize("name", util.COUNTING_CASES)
def test_group_counts_reproduce_reference_peps(name):
    """PEP of every golden hierarchical / all_cand group = 1 - any_count / N, bit for bit."""
    case = util.load_counting_case(name)
    S = case["samples"]
    N = S.shape[0]
    for key in ("hier_pep1", "all_default"):
        groups = util.golden_groups(case[key])
        noncontig = [(g, pep) for g, pep in groups if max(g) - min(g) != len(g) - 1]
        counts = co.group_any_counts(S, [g for g, _ in noncontig])
        for (g, pep), cnt in zip(noncontig, counts):
            # all_cand_groups counts in a column-compacted index space, so a group that is
            # not contiguous here may have been a sequential window there (zero_count / N)
            assert pep in (1 - cnt / N, (N - cnt) / N), (name, key, g)


def test_zero_counts_shapes_and_edges():
    S = np.zeros((7, 5))
    z = co.sequential_zero_counts(S, 25)
    assert z.shape == (5, 5)
    assert np.array_equal(z[0], [7] * 5) and np.array_equal(z[4], [7, 0, 0, 0, 0])
    assert np.array_equal(co.column_counts(np.ones((3, 4))), [3] * 4)
