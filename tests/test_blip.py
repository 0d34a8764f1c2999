"""BLiP programs on the HiGHS backend (pyblip_b200/blip.py): the reference's own properties
(/root/reference/test/test_blip.py:98-353) -- known answers of the binary program, disjointness,
error control for all four error rates with and without randomisation, backtracking counts and
LP bounds, adversarial FDR examples.  Everything given as candidate groups is host-only; the
cases that start from posterior samples need the device (counting kernels, exact FWER scan)."""
import numpy as np
import pytest

from pyblip_b200 import blip
from pyblip_b200.create_groups import CandidateGroup

TOL = 1e-5


def check_error_control(detections, error, q, samples=None):
    if len(detections) == 0:
        return
    peps = [x.pep for x in detections]
    if error == 'pfer' or (error == 'fwer' and samples is None):
        assert sum(peps) <= q + TOL
    elif error == 'local_fdr':
        assert max(peps) <= q + TOL
    elif error == 'fdr':
        assert np.mean(peps) <= q + TOL
    else:
        false_disc = np.zeros(samples.shape[0], dtype=bool)
        for cg in detections:
            false_disc |= np.all(samples[:, list(cg.group)] == 0, axis=1)
        assert false_disc.mean() <= q + TOL


def check_disjoint(detections):
    seen = set()
    for cg in detections:
        assert not (seen & set(cg.group))
        seen |= set(cg.group)


def test_pfer_known_answer_and_randomised():
    cand_groups = [
        CandidateGroup(group=[0], pep=0.19, data=dict(weight=0.5)),
        CandidateGroup(group=[0, 1], pep=0.0, data=dict(weight=0.25)),
        CandidateGroup(group=[2], pep=0.19, data=dict(weight=1)),
        CandidateGroup(group=[2, 3], pep=0.0, data=dict(weight=0.5)),
    ]
    for cg in cand_groups:
        cg.data['blip-group'] = cg.group
        cg.data['sprob'] = 0.5
    det = blip.binarize_selections(cand_groups, q=0.2, error='pfer', deterministic=True)
    assert set(tuple(x.group) for x in det) == {(2,), (0, 1)}
    for _ in range(10):
        det = blip.binarize_selections(cand_groups, q=0.2, error='pfer', deterministic=False)
        check_disjoint(det)
        check_error_control(det, 'pfer', 0.2)


def test_error_control_on_random_fractional_selections():
    np.random.seed(123)
    cand_groups = []
    inds = np.arange(30)
    for j in range(20):
        sprob = 0.01 + 0.98 * np.random.uniform()
        pep = 0.2 * np.random.uniform()
        group = set(np.random.choice(inds, j + 1, replace=False).tolist())
        cand_groups.append(CandidateGroup(group=group, pep=pep,
                                          data={"weight": 1 / len(group), "sprob": sprob, "blip-group": group}))
    for q in [0.1, 0.2, 0.5, 1.5, 2]:
        for error in ['pfer', 'fdr', 'local_fdr', 'fwer']:
            det = blip.binarize_selections(cand_groups, q=q, error=error, deterministic=True)
            check_disjoint(det)
            check_error_control(det, error, q)
            for _ in range(2):
                det = blip.binarize_selections(cand_groups, q=q, error=error, deterministic=False)
                check_disjoint(det)
                check_error_control(det, error, q)


def test_backtracking_counts_and_lp_bound():
    q = 0.1
    cand_groups = [
        CandidateGroup(group=[0, 1], pep=0.0, data=dict(weight=1)),
        CandidateGroup(group=[1, 2], pep=0.0, data=dict(weight=1)),
        CandidateGroup(group=[2, 0], pep=0.0, data=dict(weight=1)),
        CandidateGroup(group=[3], pep=0.25, data=dict(weight=1)),
    ]
    det, status = blip.BLiP(cand_groups=cand_groups, error='fdr', weight_fn='prespecified', q=q, max_pep=1,
                            deterministic=True, return_problem_status=True, verbose=False)
    assert np.mean([x.pep for x in det]) <= q
    assert status['backtracking_iter'] == 1
    assert abs(status['lp_bound'] - (1.5 + 0.75)) < 1e-3
    cand_groups = [
        CandidateGroup(group=[0, 1], pep=0.0, data=dict(weight=1.1)),
        CandidateGroup(group=[1, 2], pep=0.0, data=dict(weight=1)),
        CandidateGroup(group=[2, 0], pep=0.0, data=dict(weight=1)),
        CandidateGroup(group=[3], pep=0.1, data=dict(weight=1)),
        CandidateGroup(group=[4], pep=0.21, data=dict(weight=1)),
    ]
    det, status = blip.BLiP(cand_groups=cand_groups, error='fdr', weight_fn='prespecified', q=q, max_pep=1,
                            deterministic=True, return_problem_status=True, verbose=False)
    assert set(tuple(x.group) for x in det) == {(0, 1), (3,)}
    assert status['backtracking_iter'] == 1


def test_fdr_adversarial_examples():
    cand_groups = [
        CandidateGroup(group=[0], pep=0.05, data=dict(weight=1)),
        CandidateGroup(group=[1], pep=0.1, data=dict(weight=1)),
        CandidateGroup(group=[2], pep=0.05, data=dict(weight=1 / 100)),
        CandidateGroup(group=[3], pep=0.05, data=dict(weight=1 / 200)),
    ]
    det, _ = blip.BLiP(cand_groups=cand_groups, error='fdr', weight_fn='prespecified', q=0.05,
                       deterministic=True, return_problem_status=True, verbose=False)
    assert set(tuple(x.group) for x in det) == {(0,), (2,), (3,)}
    cand_groups = [
        CandidateGroup(group=[0, 1], pep=0.001, data=dict(weight=0.05)),
        CandidateGroup(group=[1, 2], pep=0.05, data=dict(weight=1)),
        CandidateGroup(group=[2], pep=0.0001, data=dict(weight=0.05)),
    ]
    det = blip.BLiP(cand_groups=cand_groups, error='fdr', weight_fn='prespecified', q=0.05, verbose=False)
    assert set(tuple(x.group) for x in det) == {(1, 2)}


def test_argument_errors_and_empty_problem():
    with pytest.raises(ValueError):
        blip.BLiP(cand_groups=[], error='nope')
    with pytest.raises(ValueError):
        blip.BLiP()
    det, status = blip.BLiP(cand_groups=[CandidateGroup(group=[0], pep=0.9)], q=0.1, return_problem_status=True,
                            verbose=False)
    assert det == [] and status['ngroups'] == 0 and status['nlocs'] == 0


def test_ecc_reduction_keeps_the_overlap_structure():
    from pyblip_b200 import create_groups
    rng = np.random.default_rng(3)
    cgs = [CandidateGroup(group=set(rng.choice(40, rng.integers(1, 6), replace=False).tolist()), pep=0.1)
           for _ in range(30)]
    cgs, nrel = create_groups._elim_redundant_features(cgs)
    assert nrel == len(set(j for cg in cgs for j in cg.group))
    before = [[not a.group.isdisjoint(b.group) for b in cgs] for a in cgs]
    cgs, ncl = create_groups._ecc_reduction(cgs)
    after = [[not a.data['blip-group'].isdisjoint(b.data['blip-group']) for b in cgs] for a in cgs]
    assert before == after and ncl >= 1


@pytest.mark.gpu
def test_blip_from_samples_all_error_rates(gpu):
    """test_blip_regression (reference test_blip.py:176-220): samples from the device sampler and from a
    random indicator matrix; disjointness and error control (exact FWER against the samples)."""
    from pyblip_b200 import LinearSpikeSlab, create_groups
    np.random.seed(110)
    n, p = 100, 200
    X = np.random.randn(n, p)
    for j in range(1, p):
        X[:, j] = 0.8 * X[:, j - 1] + 0.6 * X[:, j]
    beta = np.zeros(p)
    beta[np.random.choice(p, 10, replace=False)] = np.random.randn(10) * 1.5
    y = X @ beta + np.random.randn(n)
    lm = LinearSpikeSlab(X=X, y=y)
    lm.sample(N=500, chains=2, seed=3)
    samples1 = lm.betas != 0
    p2 = 300
    p0s = np.random.beta(a=1, b=3, size=p2)
    samples2 = np.random.binomial(1, p0s, size=(1000, p2))
    cand_groups2 = create_groups.sequential_groups(samples=samples2, max_pep=0.5)
    for kwargs, samples in [(dict(samples=samples1), samples1), (dict(cand_groups=cand_groups2), samples2),
                            (dict(samples=lm), samples1)]:
        for q in [0.05, 0.2]:
            for error, weight_fn in zip(['local_fdr', 'fdr', 'fwer', 'pfer'],
                                        ['log_inverse_size', 'inverse_size', 'inverse_size', 'log_inverse_size']):
                for deterministic in [True, False]:
                    det = blip.BLiP(error=error, q=q, weight_fn=weight_fn, search_method='binary', max_pep=2 * q,
                                    deterministic=deterministic, verbose=False, **kwargs)
                    check_disjoint(det)
                    check_error_control(det, error, q, samples=samples if 'samples' in kwargs else None)


@pytest.mark.gpu
def test_readme_quickstart_end_to_end(gpu):
    """BASELINE configs[0]: LinearSpikeSlab n=100 p=500 AR(1) rho=0.95, N=1000 chains=10, then
    BLiP(q=0.1, error='fdr') -- the whole path of the README on this package."""
    from pyblip_b200 import LinearSpikeSlab
    rng = np.random.default_rng(1001)
    n, p, k, rho = 100, 500, 20, 0.95
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + np.sqrt(1 - rho * rho) * Z[:, j]
    beta = np.zeros(p)
    sig = rng.choice(p, k, replace=False)
    beta[sig] = rng.standard_normal(k)
    y = X @ beta + rng.standard_normal(n)
    lm = LinearSpikeSlab(X=X, y=y)
    lm.sample(N=1000, chains=10, seed=1)
    np.random.seed(0)
    det = blip.BLiP(samples=lm.betas, q=0.1, error='fdr', verbose=False)
    check_disjoint(det)
    check_error_control(det, 'fdr', 0.1)
    assert len(det) >= 1
    # most detections contain a true signal (the realised FDP of one data set is not bounded by q,
    # the Bayesian FDR above is; this only guards against nonsense)
    hits = sum(1 for cg in det if set(cg.group) & set(sig.tolist()))
    assert hits >= 0.7 * len(det)
    # the device handle gives the same detections as the dense array
    np.random.seed(0)
    det2 = blip.BLiP(samples=lm, q=0.1, error='fdr', verbose=False)
    assert sorted(tuple(sorted(c.group)) for c in det) == sorted(tuple(sorted(c.group)) for c in det2)


@pytest.mark.gpu
def test_detect_changepoints_end_to_end(gpu):
    from pyblip_b200 import changepoint
    rng = np.random.default_rng(5)
    T = 300
    jumps = np.zeros(T)
    jumps[[60, 150, 220]] = [3.0, -2.5, 2.0]
    Y = np.cumsum(jumps) + 0.5 * rng.standard_normal(T)
    np.random.seed(1)
    det = changepoint.detect_changepoints(Y, q=0.1, sample_kwargs=dict(N=400, burn=100, chains=4, seed=2),
                                          blip_kwargs=dict(verbose=False))
    check_disjoint(det)
    assert all(0 not in cg.group for cg in det)
    found = [any(abs(j - t) <= 3 for cg in det for j in cg.group) for t in (60, 150, 220)]
    assert all(found), det


@pytest.mark.gpu
def test_blip_cts_disjoint_output(gpu):
    from pyblip_b200 import create_groups_cts
    np.random.seed(7)
    N, max_ndisc = 200, 50
    post = np.random.randn(N, max_ndisc, 2)
    for i in range(N):
        post[i, np.random.choice(np.arange(max_ndisc + 1)):] = np.nan
    post, _, _ = create_groups_cts.normalize_locs(post)
    grid_sizes = np.unique(np.around(np.logspace(np.log10(2), 4, 50)))
    det = blip.BLiP_cts(post, error='fdr', q=0.1, shape='circle', verbose=False, grid_sizes=grid_sizes,
                        deterministic=True)
    check_disjoint(det)
    centers = [x.data['center'] for x in det]
    radii = [x.data['radius'] for x in det]
    for i in range(len(det)):
        for j in range(i):
            assert np.sqrt(np.sum((centers[i] - centers[j]) ** 2)) > radii[i] + radii[j] - 1e-5
