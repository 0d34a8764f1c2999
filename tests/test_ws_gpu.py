"""The warp-specialised few-chains gram kernel (gibbs_gram_ws.cuh) against the classic one-CTA kernel
and against the reference: decision warps + helper warps must reproduce, bit for bit, what one CTA
per chain computes (same fma sequence on every element of q), whatever the interleaving of the two
groups of warps turns out to be."""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu


def _ar1(rng, n, p, rho):
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    s = np.sqrt(1 - rho * rho)
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + s * Z[:, j]
    return X


def _run(design, cfg, kernel, **kw):
    from pyblip_b200 import _lib
    smp = _lib.sample_spikeslab(design, cfg, mode=_lib.MODE_GRAM, gram_kernel=kernel, **kw)
    out = (smp.csr(), np.stack(smp.hyper()), smp.info.n_changes, smp.info.n_windows)
    smp.close()
    return out


def _same(a, b):
    for k in range(3):
        assert np.array_equal(a[0][k], b[0][k])
    assert np.array_equal(a[1], b[1])
    assert a[2] == b[2] and a[3] == b[3]


@pytest.mark.parametrize("kernel", [2, 3, 7, 8])
@pytest.mark.parametrize("name", util.LINEAR_CASES)
def test_specialised_kernel_matches_reference_goldens(gpu, name, kernel):
    """The reference's recorded runs (injected permutations, uniforms, normals, hyper variates)."""
    from pyblip_b200 import _lib
    import test_gibbs_gpu as tg
    case = util.load_sampler_case(name)
    hp = case["hp"]
    cfg = _lib.SpikeSlabConfig(hp["tau2"], int(hp["update_tau2"]), hp["tau2_a0"], hp["tau2_b0"], hp["sigma2"],
                               int(hp["update_sigma2"]), hp["sigma2_a0"], hp["sigma2_b0"], hp["p0"],
                               int(hp["update_p0"]), hp["min_p0"], hp["p0_a0"], hp["p0_b0"])
    design = _lib.Design.upload(case["X"])
    C, N, p = case["chains"], case["N"], case["X"].shape[1]
    smp = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=C, n_keep=N, burn=0, seed=case["tn_seed"],
                                streams=util.stacked_streams(case), mode=_lib.MODE_GRAM,
                                gram_kernel=kernel)
    out = dict(betas=smp.dense().reshape(C, N, p))
    p0s, tau2s, sigma2s = smp.hyper()
    out.update(p0s=p0s.reshape(C, N), tau2s=tau2s.reshape(C, N), sigma2s=sigma2s.reshape(C, N))
    smp.close(); design.close()
    tg._check_against_reference(case, out)


def test_specialised_equals_classic_bit_for_bit(gpu):
    """Keyed stream, burn-in from beta = 0 (hundreds of changes per sweep: the ring throttles), several
    launches (state through HBM), a q rebuild every 7 sweeps, odd p, more chains than one wave would
    need -- identical indicators, coefficients, hyper-parameters and work counters."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(5)
    for (n, p, chains, sweeps, chunk, refresh) in [(120, 777, 5, 40, 0, 0), (200, 1501, 3, 45, 7, 7),
                                                   (64, 130, 9, 60, 16, 0), (50, 33, 2, 30, 0, 5)]:
        X = _ar1(rng, n, p, 0.9)
        beta = np.zeros(p)
        beta[rng.choice(p, max(2, p // 60), replace=False)] = rng.standard_normal(max(2, p // 60))
        y = X @ beta + rng.standard_normal(n)
        cfg = LinearSpikeSlab(X, y)._config()
        design = _lib.Design.upload(X)
        kw = dict(y=y, chains=chains, chain_offset=3, n_keep=sweeps - 5, burn=5, seed=99, chunk_sweeps=chunk,
                  gram_refresh=refresh)
        a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
        for kernel in (_lib.GRAM_KERNEL_SPECIALIZED, _lib.GRAM_KERNEL_CLUSTER, _lib.GRAM_KERNEL_SPECIALIZED2,
                       _lib.GRAM_KERNEL_SPECIALIZED2_CLUSTER):
            b = _run(design, cfg, kernel, **kw)
            _same(a, b)
            b2 = _run(design, cfg, kernel, **kw)       # scheduling differs run to run
            _same(b, b2)
        design.close()


def test_two_position_kernel_equals_classic_on_dense_and_tiny_problems(gpu):
    """gibbs_gram_ws2.cuh: predicted Gram entries and a second tracked position per thread.  Dense signals (more
    active coordinates than the prediction list holds, every change a misprediction), p below one window, p = 1,
    long runs across launches and q rebuilds."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(6)
    for (n, p, chains, sweeps, chunk, refresh, nsig) in [(400, 2000, 2, 30, 0, 0, 150), (300, 3000, 3, 80, 32, 0, 40),
                                                         (600, 1200, 2, 25, 0, 0, 330), (30, 1, 2, 20, 0, 0, 1),
                                                         (80, 300, 4, 150, 16, 11, 8), (60, 257, 3, 40, 0, 0, 5)]:
        X = _ar1(rng, n, p, 0.9)
        beta = np.zeros(p)
        beta[rng.choice(p, nsig, replace=False)] = 2.0 * rng.standard_normal(nsig)
        y = X @ beta + rng.standard_normal(n)
        cfg = LinearSpikeSlab(X, y)._config()
        design = _lib.Design.upload(X)
        kw = dict(y=y, chains=chains, chain_offset=1, n_keep=sweeps - 5, burn=5, seed=17, chunk_sweeps=chunk,
                  gram_refresh=refresh)
        a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
        _same(a, _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED2, **kw))
        _same(a, _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED2_CLUSTER, **kw))
        design.close()


def test_specialised_c2_equals_classic_and_auto_picks_it(gpu):
    """configs[1] shape at the 8-GPU shard size: 64 chains x 40 sweeps from beta = 0."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    import bench
    X, y, _ = bench.make_data()
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, chains=64, chain_offset=128, n_keep=30, burn=10, seed=2024)
    a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
    b = _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED, **kw)
    b3 = _run(design, cfg, _lib.GRAM_KERNEL_CLUSTER, **kw)
    b7 = _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED2, **kw)
    _same(a, _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED2_CLUSTER, **kw))
    c = _run(design, cfg, _lib.GRAM_KERNEL_AUTO, **kw)
    _same(a, b)
    _same(a, b3)
    _same(a, b7)
    _same(a, c)
    # 100 chains: too many for two SMs each
    with pytest.raises(NotImplementedError):
        _run(design, cfg, _lib.GRAM_KERNEL_CLUSTER, y=y, chains=100, n_keep=2, burn=0, seed=1)
    # more chains than SMs: the specialised kernel is refused, auto falls back to the classic one
    with pytest.raises(NotImplementedError):
        _run(design, cfg, _lib.GRAM_KERNEL_SPECIALIZED, y=y, chains=400, n_keep=2, burn=0, seed=1)
    design.close()


def test_specialised_changepoint_equals_classic(gpu):
    from pyblip_b200 import _lib
    from pyblip_b200 import changepoint
    rng = np.random.default_rng(8)
    T = 3000
    jumps = np.zeros(T)
    jumps[rng.choice(T, 12, replace=False)] = rng.normal(0, 2.0, 12)
    Y = np.cumsum(jumps) + rng.standard_normal(T)
    cp = changepoint.ChangepointSpikeSlab(Y)
    cfg = cp._config()
    design = _lib.Design.changepoint(T)
    kw = dict(y=Y, chains=4, n_keep=25, burn=5, seed=4)
    a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
    for kernel in (_lib.GRAM_KERNEL_SPECIALIZED, _lib.GRAM_KERNEL_CLUSTER, _lib.GRAM_KERNEL_SPECIALIZED2,
                       _lib.GRAM_KERNEL_SPECIALIZED2_CLUSTER):
        _same(a, _run(design, cfg, kernel, **kw))
    design.close()
