"""The column-stream few-chains gram kernel (gibbs_gram_stream.cuh) against the classic one-CTA kernel and
against the reference: a resolver warp that decides the tracked redraws from registers and streamers that
apply them to q in checked passes must reproduce, bit for bit, what one CTA per chain computes -- same
decisions, same draws, same fma sequence on every element of q -- however the passes fall, however often a
pass is discarded because a coordinate nobody tracked changes, and whichever kernel ran the launch before."""
import numpy as np
import pytest

import blip_testutil as util
from test_ws_gpu import _ar1, _run

pytestmark = pytest.mark.gpu

STREAM = 6


def _same_results(a, b):
    for k in range(3):
        assert np.array_equal(a[0][k], b[0][k])
    assert np.array_equal(a[1], b[1])
    assert a[2] == b[2]          # number of changes; n_windows counts passes here


@pytest.mark.parametrize("name", util.LINEAR_CASES)
def test_stream_kernel_matches_reference_goldens(gpu, name):
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
                                gram_kernel=STREAM)
    out = dict(betas=smp.dense().reshape(C, N, p))
    p0s, tau2s, sigma2s = smp.hyper()
    out.update(p0s=p0s.reshape(C, N), tau2s=tau2s.reshape(C, N), sigma2s=sigma2s.reshape(C, N))
    smp.close(); design.close()
    tg._check_against_reference(case, out)


def test_stream_equals_classic_bit_for_bit(gpu):
    """Keyed stream, burn-in from beta = 0 (nothing tracked: every change is found by the streamers), dense
    signals (more active coordinates than the resolver tracks, more than the sweep-start list holds), several
    launches, a q rebuild every few sweeps, odd p, p below one pass stride."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(5)
    for (n, p, chains, sweeps, chunk, refresh, nsig) in [(120, 777, 5, 40, 0, 0, 12), (200, 1501, 3, 45, 7, 7, 25),
                                                         (64, 130, 9, 60, 16, 0, 3), (50, 33, 2, 30, 0, 5, 2),
                                                         (400, 2000, 2, 30, 0, 0, 150), (300, 3000, 3, 80, 32, 0, 40),
                                                         (600, 1200, 2, 25, 0, 0, 330), (30, 1, 2, 20, 0, 0, 1)]:
        X = _ar1(rng, n, p, 0.9)
        beta = np.zeros(p)
        beta[rng.choice(p, nsig, replace=False)] = 2.0 * rng.standard_normal(nsig)
        y = X @ beta + rng.standard_normal(n)
        cfg = LinearSpikeSlab(X, y)._config()
        design = _lib.Design.upload(X)
        kw = dict(y=y, chains=chains, chain_offset=3, n_keep=sweeps - 5, burn=5, seed=99, chunk_sweeps=chunk,
                  gram_refresh=refresh)
        a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
        b = _run(design, cfg, STREAM, **kw)
        _same_results(a, b)
        b2 = _run(design, cfg, STREAM, **kw)       # scheduling differs run to run
        _same_results(b, b2)
        design.close()


def test_auto_hands_over_to_the_stream_kernel(gpu, monkeypatch):
    """AUTO with few chains and BLIPGPU_STREAM=1: the role-specialised kernel runs the burn-in launches, the
    column-stream kernel takes over once a launch has seen few changes per sweep; the hand-over changes nothing."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(11)
    n, p = 300, 2500
    X = _ar1(rng, n, p, 0.9)
    beta = np.zeros(p)
    beta[rng.choice(p, 20, replace=False)] = 2.0 * rng.standard_normal(20)
    y = X @ beta + rng.standard_normal(n)
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, chains=6, n_keep=150, burn=10, seed=5, chunk_sweeps=16)
    a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
    c = _run(design, cfg, _lib.GRAM_KERNEL_AUTO, **kw)
    _same_results(a, c)
    monkeypatch.setenv("BLIPGPU_STREAM", "1")
    b = _run(design, cfg, _lib.GRAM_KERNEL_AUTO, **kw)
    _same_results(a, b)
    assert b[3] != c[3]                            # passes were counted: the stream kernel did run
    design.close()


def test_stream_c2_equals_classic(gpu):
    """configs[1] shape at the 8-GPU shard size: 64 chains x 60 sweeps from beta = 0."""
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    import bench
    X, y, _ = bench.make_data()
    cfg = LinearSpikeSlab(X, y)._config()
    design = _lib.Design.upload(X)
    kw = dict(y=y, chains=64, chain_offset=128, n_keep=50, burn=10, seed=2024)
    a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
    _same_results(a, _run(design, cfg, STREAM, **kw))
    with pytest.raises(NotImplementedError):
        _run(design, cfg, STREAM, y=y, chains=400, n_keep=2, burn=0, seed=1)
    design.close()


def test_stream_changepoint_equals_classic(gpu):
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
    kw = dict(y=Y, chains=4, n_keep=45, burn=5, seed=4)
    a = _run(design, cfg, _lib.GRAM_KERNEL_CLASSIC, **kw)
    _same_results(a, _run(design, cfg, STREAM, **kw))
    design.close()
