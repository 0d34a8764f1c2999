"""Device PIP counting vs the reference's golden outputs and the NumPy oracle.
Counts and PEPs must be bit-exact (integer work + one float64 divide)."""
import json

import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", util.COUNTING_CASES)
def test_primitive_counts_bit_exact(gpu, name):
    from oracle import counting_oracle as co
    from pyblip_b200 import _lib
    samples = util.load_counting_case(name)["samples"]
    bits = _lib.Bits.from_dense(samples)
    assert np.array_equal(bits.to_dense(), samples != 0)
    assert np.array_equal(bits.count_columns(), co.column_counts(samples))
    for max_size in (1, 4, 25, 33, 64):
        ms = min(max_size, samples.shape[1])
        assert np.array_equal(bits.count_sequential(ms), co.sequential_zero_counts(samples, ms)), max_size
    rng = np.random.default_rng(1)
    p = samples.shape[1]
    groups = [sorted(rng.choice(p, size=int(rng.integers(1, min(p, 25) + 1)), replace=False).tolist())
              for _ in range(200)]
    assert np.array_equal(bits.count_groups(groups), co.group_any_counts(samples, groups))
    cols = np.sort(rng.choice(p, size=max(1, p // 3), replace=False))
    sub = bits.select_columns(cols)
    assert np.array_equal(sub.to_dense(), (samples != 0)[:, cols])
    assert np.array_equal(sub.count_sequential(min(25, len(cols))),
                          co.sequential_zero_counts(samples[:, cols], min(25, len(cols))))


@pytest.mark.parametrize("name", util.COUNTING_CASES)
def test_group_functions_match_reference_golden(gpu, name):
    from pyblip_b200 import create_groups as cg
    case = util.load_counting_case(name)
    samples = case["samples"]
    for key, fn in [("seq", cg.sequential_groups), ("hier", cg.hierarchical_groups),
                    ("all", cg.all_cand_groups)]:
        for k in [k for k in case if k.startswith(key + "_") and not k.endswith("_kwargs")]:
            kwargs = case[k + "_kwargs"]
            got = util.cg_list(fn(samples.copy(), **kwargs))
            want = util.golden_groups(case[k])
            assert sorted(got) == sorted(want), f"{name}/{k}: groups or PEPs differ"
            if key == "seq" and not kwargs.get("prenarrow"):
                assert got == want, f"{name}/{k}: order differs"


def test_large_random_matrix_against_oracle(gpu):
    """Sizes past the fixtures: ragged N (not a multiple of 32), p not a multiple of 32,
    all-zero and all-one columns, > 2 row chunks."""
    from oracle import counting_oracle as co
    from pyblip_b200 import _lib, create_groups as cg
    rng = np.random.default_rng(7)
    N, p = 50021, 203
    S = rng.random((N, p)) < 0.02
    S[:, 17] = True
    S[:, 40:44] = False
    S[rng.random(N) < 0.8, 100] = True
    bits = _lib.Bits.from_dense(S)
    assert np.array_equal(bits.count_columns(), co.column_counts(S))
    assert np.array_equal(bits.count_sequential(25), co.sequential_zero_counts(S, 25))
    got = util.cg_list(cg.sequential_groups(S, max_pep=0.5))
    want = co.sequential_groups(S, max_pep=0.5)
    assert got == want


def test_empty_and_degenerate_inputs(gpu):
    from pyblip_b200 import create_groups as cg
    assert cg.all_cand_groups(np.zeros((100, 40)), prenarrow=False, max_pep=0.5) == []
    one = cg.hierarchical_groups(np.array([[1.0], [0.0], [1.0], [1.0]]))
    assert len(one) == 1 and one[0].group == {0} and one[0].pep == 1 - 3 / 4
    with pytest.raises(ValueError):
        cg.sequential_groups()


def test_sampler_handle_counting_equals_dense_path(gpu):
    """Counting on the device-resident samples == counting on lm.betas."""
    from pyblip_b200 import LinearSpikeSlab, create_groups as cg, changepoint
    case = util.load_sampler_case("linear_c1")
    lm = LinearSpikeSlab(case["X"], case["y"])
    lm.sample(N=60, burn=10, chains=3, seed=3)
    a = util.cg_list(cg.all_cand_groups(lm, max_pep=0.5))
    b = util.cg_list(cg.all_cand_groups(lm.betas, max_pep=0.5))
    assert sorted(a) == sorted(b) and len(a) > 0
    c1 = util.cg_list(changepoint.changepoint_cand_groups(lm, max_pep=0.9))
    dense = lm.betas != 0
    dense[:, 0] = 0
    c2 = util.cg_list(cg.sequential_groups(dense, max_pep=0.9))
    assert c1 == c2


def test_pair_counts_and_device_correlation_path(gpu, monkeypatch):
    """Co-occurrence counts on the device equal S^T S; the correlation matrix built from the exact
    integers equals np.corrcoef of the reference's `_samples_dist_matrix` to rounding; groups proposed
    through the device path carry bit-exact PEPs."""
    from pyblip_b200 import _lib, create_groups as cg
    rng = np.random.default_rng(11)
    for N, p in [(1, 1), (33, 5), (5000, 70), (40013, 131)]:
        S = rng.random((N, p)) < 0.05
        if p > 3:
            S[:, 2] = True
            S[:, 3] = False
        bits = _lib.Bits.from_dense(S)
        pc = bits.count_pairs()
        bits.close()
        Si = S.astype(np.int64)
        assert np.array_equal(pc, Si.T @ Si), (N, p)
    S = rng.random((3000, 40)) < 0.1
    S[:, 5] = S[:, 4]                                 # identical columns: exact tie
    dev = cg._as_device(S)
    d_dev = cg._samples_dist_matrix_device(dev)
    d_host = cg._samples_dist_matrix(S)
    dev.close()
    assert np.max(np.abs(d_dev - d_host)) < 1e-12
    # The trees built from the two matrices need not be the same: np.corrcoef's rounding noise breaks
    # the exact ties of sparse indicator columns one way, exact integers leave them tied (scipy then
    # takes the first pair).  The default therefore keeps the host path wherever the reference itself
    # can run (parity), and the device path above _DEVICE_CORR_MIN_ENTRIES must give VALID groups:
    # PEPs bit-exact for the groups it proposes, size and PEP limits respected.
    monkeypatch.setattr(cg, "_DEVICE_CORR_MIN_ENTRIES", 0)      # force the device path
    for name in util.COUNTING_CASES:
        samples = util.load_counting_case(name)["samples"]
        ind = samples != 0
        groups = cg.hierarchical_groups(samples.copy(), max_pep=0.5, max_size=10)
        assert len(groups) > 0 or ind.shape[1] < 2
        for g in groups:
            cols = sorted(g.group)
            assert len(cols) <= 10 and g.pep < 0.5
            assert g.pep == 1 - ind[:, cols].any(axis=1).sum() / ind.shape[0]
        both = cg.all_cand_groups(samples.copy(), max_pep=0.5, max_size=10)
        for g in both:
            cols = sorted(g.group)
            cnt, N = ind[:, cols].any(axis=1).sum(), ind.shape[0]
            # hierarchical groups: 1 - mean(any) (:459); sequential groups: mean(all zero) (:329-332)
            assert g.pep in (1 - cnt / N, (N - cnt) / N)


def test_selection_fwer_equals_the_reference_row_scan(gpu):
    """blip.py:270-276: false_disc |= all(samples[:, group] == 0, axis=1); fwer = mean -- exact."""
    from pyblip_b200 import create_groups as cg
    rng = np.random.default_rng(4)
    for N, p in [(1, 3), (77, 40), (30011, 333)]:
        S = rng.random((N, p)) < 0.2
        for ng in (0, 1, 5, 40):
            groups = [sorted(rng.choice(p, size=int(rng.integers(1, min(p, 6) + 1)), replace=False).tolist())
                      for _ in range(ng)]
            fd = np.zeros(N, dtype=bool)
            for g in groups:
                fd = fd | np.all(S[:, g] == 0, axis=1)
            assert cg.selection_fwer(S, groups) == fd.mean(), (N, p, ng)
    groups = [cg.CandidateGroup(group={1, 2}, pep=0.1), cg.CandidateGroup(group={5}, pep=0.2)]
    S = rng.random((500, 8)) < 0.5
    want = (np.all(S[:, [1, 2]] == 0, axis=1) | np.all(S[:, [5]] == 0, axis=1)).mean()
    assert cg.selection_fwer(S, groups) == want


def test_column_major_copy_gives_the_same_counts(gpu):
    """count_groups switches to a column-major copy of the bit matrix when the groups hold more than
    p/8 members in total, and count_false_rows uses that copy once it exists: both must equal the
    row-major kernels and NumPy on ragged shapes (N not a multiple of 32 / 128 / 256)."""
    from pyblip_b200 import _lib
    rng = np.random.default_rng(12)
    for N, p in [(1, 5), (33, 70), (4097, 130), (50021, 203)]:
        S = rng.random((N, p)) < 0.05
        S[:, p // 2] = False
        few = [[0], [p - 1, 1]] if p > 2 else [[0]]
        many = [sorted(rng.choice(p, size=int(rng.integers(1, min(p, 9) + 1)), replace=False).tolist())
                for _ in range(max(8, p // 4))]
        sel = many[:7]

        def want_any(groups):
            return np.array([np.any(S[:, g], axis=1).sum() for g in groups], dtype=np.int64)

        def want_false(groups):
            fd = np.zeros(N, dtype=bool)
            for g in groups:
                fd |= np.all(S[:, g] == 0, axis=1)
            return int(fd.sum())

        bits = _lib.Bits.from_dense(S)
        assert np.array_equal(bits.count_groups(few), want_any(few))              # row-major (few members)
        assert bits.count_false_rows(sel) == want_false(sel)                      # row-major scan
        assert np.array_equal(bits.count_groups(many), want_any(many))            # builds the copy
        assert np.array_equal(bits.count_groups(few), want_any(few))              # now served by the copy
        assert bits.count_false_rows(sel) == want_false(sel)                      # column-major scan
        assert bits.count_false_rows([[p // 2]]) == N                             # an all-zero column
        bits.close()


def test_pair_counts_on_the_tensor_cores(gpu, monkeypatch):
    """The tcgen05 integer GEMM (pairs_tc.cuh) gives exactly S^T S: ragged row counts (K padding), column counts
    that end inside a 128 x 256 tile, dense and empty columns, several k-splits; it is the default from 512 columns
    up, can be forced for small sets, and agrees with the scalar kernel entry for entry."""
    from pyblip_b200 import _lib
    rng = np.random.default_rng(23)
    cases = [(1, 1, 0.5), (33, 5, 0.3), (777, 130, 0.2), (5000, 700, 0.03), (40013, 1031, 0.01), (100, 2050, 0.2),
             (300001, 513, 0.02)]
    for N, p, dens in cases:
        S = rng.random((N, p)) < dens
        if p > 3:
            S[:, 2] = True
            S[:, 3] = False
        Si = S.astype(np.float64)                      # exact below 2^53, and BLAS is quick
        want = (Si.T @ Si).astype(np.int64)
        bits = _lib.Bits.from_dense(S)
        monkeypatch.setenv("BLIPGPU_PAIRS_TC", "1")
        pc = bits.count_pairs()
        assert np.array_equal(pc, want), (N, p)
        if p <= 1031:
            monkeypatch.setenv("BLIPGPU_PAIRS_TC", "0")
            assert np.array_equal(bits.count_pairs(), want), (N, p)
        monkeypatch.delenv("BLIPGPU_PAIRS_TC")
        if p >= 512:
            assert np.array_equal(bits.count_pairs(), want), (N, p)
        bits.close()
