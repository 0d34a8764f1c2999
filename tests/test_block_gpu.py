"""Block sampler (bsize > 1) on the device vs the reference's `_sample_spikeslab_multi`
(golden fixtures written by oracle/pin_multi.py) and vs the CPU oracle.

Contract (BASELINE.json north_star): driven by an identical injected stream, indicators are
bit-exact and betas / sigma2 / tau2 / p0 agree within 1e-9 relative for the first 50 sweeps.
All calls go through the C ABI (ctypes).
"""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu
RTOL = 1e-9
LINEAR = [c for c in util.MULTI_CASES if "probit" not in c]
PROBIT = [c for c in util.MULTI_CASES if "probit" in c]


def _cfg(hp):
    from pyblip_b200 import _lib
    return _lib.SpikeSlabConfig(hp["tau2"], int(hp["update_tau2"]), hp["tau2_a0"], hp["tau2_b0"],
                                hp["sigma2"], int(hp["update_sigma2"]), hp["sigma2_a0"],
                                hp["sigma2_b0"], hp["p0"], int(hp["update_p0"]), hp["min_p0"],
                                hp["p0_a0"], hp["p0_b0"])


def _run_device(case, mode, chunk_sweeps=0, gram_refresh=0, inject=True):
    from pyblip_b200 import _lib
    design = _lib.Design.upload(case["X"])
    streams = util.stacked_streams(case) if inject else None
    kw = dict(chains=case["chains"], chain_offset=0, n_keep=case["N"], burn=0, seed=case["tn_seed"],
              streams=streams, mode=mode, chunk_sweeps=chunk_sweeps, gram_refresh=gram_refresh,
              bsize=case["bsize"], max_signals_per_block=case["max_signals"])
    if case["probit"]:
        smp = _lib.sample_spikeslab(design, _cfg(case["hp"]), z=case["y"].astype(np.int64), probit=True, **kw)
    else:
        smp = _lib.sample_spikeslab(design, _cfg(case["hp"]), y=case["y"], probit=False, **kw)
    N, p, C = case["N"], case["X"].shape[1], case["chains"]
    out = dict(betas=smp.dense().reshape(C, N, p))
    p0s, tau2s, sigma2s = smp.hyper()
    out.update(p0s=p0s.reshape(C, N), tau2s=tau2s.reshape(C, N), sigma2s=sigma2s.reshape(C, N))
    if case["probit"]:
        out["y_latent"] = smp.latents().reshape(C, N, case["X"].shape[0])
    smp.close()
    design.close()
    return out


def _check(case, out, refs, tol=RTOL):
    for c, ref in enumerate(refs):
        assert np.array_equal(out["betas"][c] != 0, ref["betas"] != 0), \
            f"{case['name']} chain {c}: indicators differ"
        for key in ("betas", "p0s", "tau2s", "sigma2s"):
            err = util.rel_err(out[key][c], ref[key])
            assert err <= tol, f"{case['name']} chain {c}: {key} rel err {err:.3e} > {tol}"
        if case["probit"]:
            err = util.latent_err(out["y_latent"][c], ref["y_latent"])
            assert err <= tol, f"{case['name']} chain {c}: latents err {err:.3e}"


@pytest.mark.parametrize("name", util.MULTI_CASES)
def test_block_residual_form_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_multi_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL)
    _check(case, out, [cd["ref"] for cd in case["chains_data"]])


@pytest.mark.parametrize("name", LINEAR)
def test_block_gram_form_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_multi_case(name)
    out = _run_device(case, _lib.MODE_GRAM)
    _check(case, out, [cd["ref"] for cd in case["chains_data"]])


@pytest.mark.parametrize("name", ["multi_b3", "multi_probit_b3"])
def test_block_small_chunks_and_refresh(gpu, name):
    """State round-trips through HBM every 7 sweeps; gram form also rebuilds q every 5 sweeps."""
    from pyblip_b200 import _lib
    case = util.load_multi_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL, chunk_sweeps=7)
    _check(case, out, [cd["ref"] for cd in case["chains_data"]])
    if not case["probit"]:
        out = _run_device(case, _lib.MODE_GRAM, chunk_sweeps=7, gram_refresh=5)
        _check(case, out, [cd["ref"] for cd in case["chains_data"]])


@pytest.mark.parametrize("name", ["multi_b3", "multi_b5_cap2", "multi_b4_wrap", "multi_probit_b2"])
def test_block_keyed_stream_matches_oracle(gpu, name):
    """No injection: keyed block order / choice uniforms / normals / hyper variates on the device
    against the CPU oracle running the same stream contract."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    case = util.load_multi_case(name)
    probit = case["probit"]
    out = _run_device(case, _lib.MODE_RESIDUAL, inject=False)
    refs = [go.sample_spikeslab_multi(case["N"], case["bsize"], case["X"], y=None if probit else case["y"],
                                      z=case["y"].astype(np.int64) if probit else None, probit=probit,
                                      hp=go.hparams(**case["hp"]), max_signals_per_block=case["max_signals"],
                                      seed=case["tn_seed"], chain=c) for c in range(case["chains"])]
    _check(case, out, refs)
    if not probit:
        out = _run_device(case, _lib.MODE_GRAM, inject=False)
        _check(case, out, refs)


def test_block_sampler_through_the_reference_api(gpu):
    """LinearSpikeSlab(...).sample(bsize=3) / ProbitSpikeSlab(...).sample(bsize=2): shapes and
    attributes of the reference API (linear.py:165-168, probit.py:95-99) and a sane posterior."""
    from pyblip_b200.linear import LinearSpikeSlab
    from pyblip_b200.probit import ProbitSpikeSlab
    rng = np.random.default_rng(5)
    n, p = 120, 50
    X = rng.standard_normal((n, p))
    beta = np.zeros(p)
    beta[[3, 4, 20]] = [2.0, -2.0, 1.5]
    y = X @ beta + rng.standard_normal(n)
    lm = LinearSpikeSlab(X=X, y=y)
    lm.sample(N=300, burn=100, chains=4, bsize=3, seed=11)
    assert lm.betas.shape == (1200, p) and lm.p0s.shape == (1200,)
    pips = (lm.betas != 0).mean(0)
    assert pips[[3, 4, 20]].min() > 0.9 and np.delete(pips, [3, 4, 20]).max() < 0.5
    lm2 = LinearSpikeSlab(X=X, y=y)
    lm2.sample(N=300, burn=100, chains=4, bsize=3, max_signals_per_block=1, seed=11, mode="residual")
    assert (lm2.betas != 0).mean(0)[[3, 20]].min() > 0.9
    z = (X @ beta + rng.standard_normal(n) < 0).astype(float)
    pm = ProbitSpikeSlab(X=X, y=z)
    pm.sample(N=200, burn=100, chains=2, bsize=2, seed=3)
    assert pm.betas.shape == (400, p) and pm.Z.shape == (400, n)
    with pytest.raises(NotImplementedError):
        LinearSpikeSlab(X=X, y=y).sample(N=10, burn=0, chains=1, bsize=9)


def test_block_longrun_pips_within_4_mc_standard_errors(gpu):
    """Long-run posterior summaries of the device block sampler (bsize=3, keyed stream, both forms)
    against the reference's `_sample_spikeslab_multi` run with its native RNG
    (tests/golden/longrun_multi_reference.npz, written by oracle/pin_multi.py --longrun)."""
    import os
    from pyblip_b200 import LinearSpikeSlab
    d = np.load(os.path.join(util.GOLDEN, "longrun_multi_reference.npz"))
    burn, N, bsize = int(d["burn"]), int(d["N"]), int(d["bsize"])
    chains = d["pips"].shape[0]

    def gate(dev, ref, floor):
        se = np.sqrt(dev.var(0, ddof=1) / dev.shape[0] + ref.var(0, ddof=1) / ref.shape[0])
        return np.abs(dev.mean(0) - ref.mean(0)) <= 4 * np.maximum(se, floor)

    for mode in ("gram", "residual"):
        lm = LinearSpikeSlab(d["X"], d["y"])
        lm.sample(N=N, burn=burn, chains=chains, bsize=bsize, seed=79, mode=mode)
        b = lm.betas.reshape(chains, N, -1)
        assert gate((b != 0).mean(1), d["pips"], 2e-3).all(), mode
        assert gate(b.mean(1), d["means"], 2e-3).all(), mode
        assert gate(lm.p0s.reshape(chains, N).mean(1, keepdims=True), d["p0"][:, None], 1e-3).all(), mode
