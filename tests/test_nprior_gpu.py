"""Neuronized-prior sampler on the device vs the reference (golden fixtures made by
oracle/pin_nprior.py from the reference's compiled kernel) and vs the CPU oracle."""
import os

import numpy as np
import pytest

import blip_testutil as util

pytestmark = pytest.mark.gpu
RTOL = 1e-9
CASES = ["nprior_joint_group", "nprior_plain", "nprior_sigprior", "nprior_fixedp0"]
KW = ["tauw2", "p0_init", "min_p0", "update_p0", "sigma_a0", "sigma_b0", "sigma_prior_type",
      "joint_sample_W", "group_alpha_update"]


def _load(name):
    d = np.load(os.path.join(util.GOLDEN, name + ".npz"))
    kw = {k: float(d["kw_" + k]) for k in KW}
    st = {k[3:]: d[k] for k in d.files if k.startswith("st_")}
    ref = {k[4:]: d[k] for k in d.files if k.startswith("ref_")}
    return dict(X=d["X"], y=d["y"], N=int(d["N"]), seed=int(d["seed"]), kw=kw, st=st, ref=ref)


def _device(case, inject):
    from pyblip_b200 import _lib
    kw = case["kw"]
    cfg = _lib.NPriorConfig(kw["tauw2"], kw["p0_init"], kw["min_p0"], int(kw["update_p0"]),
                            kw["sigma_a0"], kw["sigma_b0"], int(kw["sigma_prior_type"]),
                            int(kw["joint_sample_W"]), int(kw["group_alpha_update"]), 0)
    design = _lib.Design.upload(case["X"])
    streams = None
    if inject:   # one recorded chain -> leading chain axis; sigma_init is one scalar per chain
        streams = {k: (v.reshape(1) if k == "sigma_init" else v[None]) for k, v in case["st"].items()}
    smp = _lib.sample_nprior(design, cfg, case["y"], chains=1, n_keep=case["N"], burn=0,
                             seed=case["seed"], streams=streams)
    out = {k: smp.get(k) for k in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s")}
    bits = smp.bits()
    assert np.array_equal(bits.to_dense(), out["betas"] != 0)
    smp.close()
    design.close()
    return out


@pytest.mark.parametrize("name", CASES)
def test_nprior_matches_reference_under_injected_stream(gpu, name):
    case = _load(name)
    out = _device(case, inject=True)
    ref = case["ref"]
    assert np.array_equal(out["betas"] != 0, ref["betas"] != 0), "indicators differ"
    for k in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s"):
        err = util.rel_err(out[k], ref[k])
        assert err <= RTOL, f"{name}: {k} rel err {err:.3e}"


@pytest.mark.parametrize("name", CASES)
def test_nprior_keyed_stream_matches_oracle(gpu, name):
    from oracle import gibbs_oracle as go
    case = _load(name)
    out = _device(case, inject=False)
    kw = dict(case["kw"])
    for k in ("update_p0", "sigma_prior_type", "joint_sample_W", "group_alpha_update"):
        kw[k] = int(kw[k])
    orc = go.nprior_sample(case["N"], case["X"], case["y"], seed=case["seed"], chain=0, **kw)
    assert np.array_equal(out["betas"] != 0, orc["betas"] != 0)
    for k in ("alphas", "ws", "sigma2s", "alpha0s", "p0s"):
        err = util.rel_err(out[k], orc[k])
        assert err <= RTOL, f"{name}: {k} rel err {err:.3e}"
    # beta_j = w_j * max(alpha_j - alpha0, 0) (_nprior.pyx:204-210) cancels when alpha_j is just
    # above alpha0, so its rounding error is relative to the operands, not to the difference
    scale = np.abs(orc["ws"]) * (np.abs(orc["alphas"]) + np.abs(orc["alpha0s"])[:, None])
    excess = np.abs(out["betas"] - orc["betas"]) - RTOL * np.maximum(scale, np.abs(orc["betas"]))
    assert excess.max() <= 0, f"{name}: betas off by {excess.max():.3e} beyond the operand-relative bound"


def test_nprior_api_and_signal_recovery(gpu):
    """Reference test_mcmc.py:87-134 (nprior leg): min non-null PIP > 0.9, beta estimated."""
    from pyblip_b200 import NPrior, create_groups as cg
    rng = np.random.default_rng(123)
    n, p = 1000, 100
    X = rng.standard_normal((n, p))
    for j in range(1, p):
        X[:, j] = 0.5 * X[:, j - 1] + np.sqrt(0.75) * X[:, j]
    beta = np.zeros(p)
    nn = rng.choice(p, 25, replace=False)
    beta[nn] = rng.uniform(0.5, 1, 25) * rng.choice([-1, 1], 25)
    y = X @ beta + rng.standard_normal(n)
    nlm = NPrior(X=X, y=y, tauw2=1)
    nlm.sample(N=600, chains=2, burn=200, seed=4)
    assert nlm.betas.shape == nlm.alphas.shape == nlm.ws.shape == (1200, p)
    assert nlm.sigma2s.shape == nlm.alpha0s.shape == nlm.p0s.shape == (1200,)
    pips = (nlm.betas != 0).mean(axis=0)
    assert pips[beta != 0].min() > 0.9
    np.testing.assert_allclose(nlm.betas.mean(axis=0), beta, atol=0.15)
    groups = cg.sequential_groups(nlm, max_pep=0.1)
    found = {tuple(g.group)[0] for g in groups if len(g.group) == 1}
    assert set(np.where(beta != 0)[0]) <= found


def test_nprior_workspace_grows_and_replays(gpu):
    """More than 256 active coordinates: the joint-W workspace starts at 256, the launch that outgrows
    it is replayed from the state snapshot with a larger one -- results still equal the oracle's."""
    from oracle import gibbs_oracle as go
    from pyblip_b200 import _lib
    rng = np.random.default_rng(21)
    n, p, N = 120, 420, 4
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal(p) + 0.1 * rng.standard_normal(n)
    kw = dict(tauw2=1.0, p0_init=0.3, min_p0=1e-10, update_p0=1, sigma_a0=5.0, sigma_b0=1.0,
              sigma_prior_type=0, joint_sample_W=1, group_alpha_update=1)
    cfg = _lib.NPriorConfig(kw["tauw2"], kw["p0_init"], kw["min_p0"], kw["update_p0"], kw["sigma_a0"],
                            kw["sigma_b0"], kw["sigma_prior_type"], kw["joint_sample_W"],
                            kw["group_alpha_update"], 0)
    design = _lib.Design.upload(X)
    smp = _lib.sample_nprior(design, cfg, y, chains=2, n_keep=N, burn=0, seed=5)
    out = {k: smp.get(k) for k in ("alphas", "ws", "betas", "sigma2s")}
    smp.close(); design.close()
    assert ((out["betas"] != 0).sum(1) > 256).any(), "the case must outgrow the initial workspace"
    for c in range(2):
        orc = go.nprior_sample(N, X, y, seed=5, chain=c, **kw)
        sl = slice(c * N, (c + 1) * N)
        assert np.array_equal(out["betas"][sl] != 0, orc["betas"] != 0)
        for k in ("alphas", "ws", "sigma2s"):
            assert util.rel_err(out[k][sl], orc[k]) <= 1e-7, k    # 300x300 Cholesky: conditioning, not logic


def test_nprior_waves_when_the_workspace_does_not_fit(gpu, monkeypatch):
    """A workspace budget that holds one chain at a time (BLIPGPU_NPRIOR_WORKSPACE_MB): the chains of a
    launch run in waves, with results identical to the single-launch run."""
    from pyblip_b200 import _lib
    rng = np.random.default_rng(33)
    n, p, N, chains = 100, 300, 6, 5
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal(p) + 0.1 * rng.standard_normal(n)
    cfg = _lib.NPriorConfig(1.0, 0.3, 1e-10, 1, 5.0, 1.0, 0, 1, 1, 0)
    design = _lib.Design.upload(X)

    def run():
        smp = _lib.sample_nprior(design, cfg, y, chains=chains, n_keep=N, burn=2, seed=9)
        out = {k: smp.get(k) for k in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s")}
        smp.close()
        return out

    ref = run()
    assert ((ref["betas"] != 0).sum(1) > 64).any(), "the case must use the global-memory workspace"
    monkeypatch.setenv("BLIPGPU_NPRIOR_WORKSPACE_MB", "1.0")      # 300 x 300 doubles = 0.69 MB per chain
    out = run()
    design.close()
    for k in ref:
        assert np.array_equal(out[k], ref[k]), k
