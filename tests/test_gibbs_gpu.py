"""CUDA sampler vs the reference (golden fixtures) and vs the CPU oracle.

Contract (BASELINE.json north_star): driven by an identical injected stream,
indicators are bit-exact and betas / sigma2 / tau2 / p0 agree within 1e-9 relative
for the first 50 sweeps.  All calls go through the C ABI (ctypes).
"""
import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # the tolerance north_star states for floating point results


def _run_device(case, mode, chunk_sweeps=0, gram_refresh=0, inject=True):
    from pyblip_b200 import _lib
    hp = case["hp"]
    cfg = _lib.SpikeSlabConfig(hp["tau2"], int(hp["update_tau2"]), hp["tau2_a0"], hp["tau2_b0"],
                               hp["sigma2"], int(hp["update_sigma2"]), hp["sigma2_a0"],
                               hp["sigma2_b0"], hp["p0"], int(hp["update_p0"]), hp["min_p0"],
                               hp["p0_a0"], hp["p0_b0"])
    design = _lib.Design.upload(case["X"])
    streams = util.stacked_streams(case) if inject else None
    kw = dict(chains=case["chains"], chain_offset=0, n_keep=case["N"], burn=0,
              seed=case["tn_seed"], streams=streams, mode=mode, chunk_sweeps=chunk_sweeps,
              gram_refresh=gram_refresh)
    if case["probit"]:
        smp = _lib.sample_spikeslab(design, cfg, z=case["y"].astype(np.int64), probit=True, **kw)
    else:
        smp = _lib.sample_spikeslab(design, cfg, y=case["y"], probit=False, **kw)
    N, p, C = case["N"], case["X"].shape[1], case["chains"]
    out = dict(betas=smp.dense().reshape(C, N, p))
    p0s, tau2s, sigma2s = smp.hyper()
    out.update(p0s=p0s.reshape(C, N), tau2s=tau2s.reshape(C, N), sigma2s=sigma2s.reshape(C, N))
    if case["probit"]:
        out["y_latent"] = smp.latents().reshape(C, N, case["X"].shape[0])
        out["n_events"] = smp.info.n_events
    out["info"] = smp.info
    # CSR export must agree with the dense materialisation
    indptr, indices, values = smp.csr()
    dense2 = np.zeros((C * N, p))
    for r in range(C * N):
        dense2[r, indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
    assert np.array_equal(dense2, out["betas"].reshape(C * N, p))
    smp.close()
    design.close()
    return out


def _check_against_reference(case, out, tol=RTOL):
    for c, cd in enumerate(case["chains_data"]):
        ref = cd["ref"]
        assert np.array_equal(out["betas"][c] != 0, ref["betas"] != 0), \
            f"{case['name']} chain {c}: indicators differ from the reference"
        for key in ("betas", "p0s", "tau2s", "sigma2s"):
            err = util.rel_err(out[key][c], ref[key])
            assert err <= tol, f"{case['name']} chain {c}: {key} rel err {err:.3e} > {tol}"
        if case["probit"]:
            err = util.rel_err(out["y_latent"][c], ref["y_latent"])
            assert err <= tol, f"{case['name']} chain {c}: latents rel err {err:.3e}"


@pytest.mark.parametrize("name", util.SAMPLER_CASES)
def test_residual_mode_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", util.LINEAR_CASES)
def test_gram_mode_matches_reference(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_GRAM)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "linear_dense"])
def test_gram_mode_with_refresh_and_small_chunks(gpu, name):
    """Rebuilding q from X'y every 7 sweeps and relaunching every 3 sweeps must
    stay inside the same tolerance (state round-trips through HBM)."""
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_GRAM, chunk_sweeps=3, gram_refresh=7)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "probit_small"])
def test_residual_mode_small_chunks(gpu, name):
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    out = _run_device(case, _lib.MODE_RESIDUAL, chunk_sweeps=7)
    _check_against_reference(case, out)


@pytest.mark.parametrize("name", ["linear_c1", "linear_minp0", "linear_odd", "probit_small"])
def test_keyed_philox_stream_matches_oracle(gpu, name):
    """No injection: device Philox / Fisher-Yates / Marsaglia-Tsang / truncated
    normal against the oracle's independent implementation of the stream contract."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    case = util.load_sampler_case(name)
    modes = [_lib.MODE_RESIDUAL] if case["probit"] else [_lib.MODE_RESIDUAL, _lib.MODE_GRAM]
    hp = go.hparams(**case["hp"])
    N = case["N"]
    for mode in modes:
        out = _run_device(case, mode, inject=False)
        for c in range(case["chains"]):
            if case["probit"]:
                orc = go.sample_spikeslab(N, case["X"], z=case["y"].astype(np.int64), probit=True,
                                          hp=hp, seed=case["tn_seed"], chain=c)
            else:
                orc = go.sample_spikeslab(N, case["X"], y=case["y"], hp=hp,
                                          seed=case["tn_seed"], chain=c)
            assert np.array_equal(out["betas"][c] != 0, orc["betas"] != 0), (name, mode, c)
            for key in ("betas", "p0s", "tau2s", "sigma2s"):
                err = util.rel_err(out[key][c], orc[key])
                assert err <= RTOL, f"{name} mode {mode} chain {c}: {key} rel err {err:.3e}"
            if case["probit"]:
                assert util.rel_err(out["y_latent"][c], orc["y_latent"]) <= RTOL


def test_chain_offset_gives_same_chains_as_one_launch(gpu):
    """Sharding chains over ranks must not change any chain (Philox key = global id)."""
    from pyblip_b200 import _lib
    case = util.load_sampler_case("linear_c1")
    hp = case["hp"]
    cfg = _lib.SpikeSlabConfig(hp["tau2"], 1, hp["tau2_a0"], hp["tau2_b0"], hp["sigma2"], 1,
                               hp["sigma2_a0"], hp["sigma2_b0"], hp["p0"], 1, hp["min_p0"],
                               hp["p0_a0"], hp["p0_b0"])
    design = _lib.Design.upload(case["X"])
    full = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=5, n_keep=20, burn=5, seed=11,
                                 mode=_lib.MODE_GRAM)
    a = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=2, chain_offset=0, n_keep=20,
                              burn=5, seed=11, mode=_lib.MODE_GRAM)
    b = _lib.sample_spikeslab(design, cfg, y=case["y"], chains=3, chain_offset=2, n_keep=20,
                              burn=5, seed=11, mode=_lib.MODE_GRAM)
    assert np.array_equal(full.dense(), np.concatenate([a.dense(), b.dense()]))
    assert np.array_equal(np.stack(full.hyper()),
                          np.concatenate([np.stack(a.hyper()), np.stack(b.hyper())], axis=1))


def test_gram_build_matches_numpy(gpu):
    from pyblip_b200 import _lib
    rng = np.random.default_rng(5)
    for n, p in [(100, 500), (37, 45), (300, 129), (1000, 257)]:
        X = rng.standard_normal((n, p))
        d = _lib.Design.upload(X)
        d.build_gram()
        G = d.gram()
        ref = X.T @ X
        assert np.max(np.abs(G - ref)) <= 1e-10 * np.max(np.abs(ref)), (n, p)
        assert np.array_equal(G, G.T)
        assert util.rel_err(d.xl2(), (X ** 2).sum(0)) <= 1e-13
        d.close()
