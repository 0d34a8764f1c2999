"""Pin the CPU oracle against the reference itself and write golden fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/pin_against_reference.py            # check + (re)write tests/golden/*.npz

What it does, per case:
  1. runs the reference's own compiled kernel ``pyblip.linear._linear.
     _sample_spikeslab`` (oracle/_ref) with its NATIVE random sources -- libc
     rand(), NumPy global RandomState, scipy.stats.invgamma -- while recording
     every variate it consumes, indexed by (sweep, position in permutation)
     [linear path]; the probit truncated normals are replaced on the reference
     side by the keyed draw ``orc_truncnorm`` (the module-level name
     ``_linear.sample_truncnorm`` is looked up at call time, _linear.pyx:17);
  2. replays the recording through the C oracle (oracle/gibbs_oracle.c);
  3. requires identical indicators and <=1e-11 relative agreement of betas /
     p0 / sigma2 / tau2 / latents;
  4. stores inputs, the recorded stream and the REFERENCE's outputs in
     tests/golden/<case>.npz, which tests/ load on the GPU box (where
     /root/reference does not exist).
"""
import os
import sys

import numpy as np
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gibbs_oracle as go  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TN_SEED = 20261019


def ar1_design(rng, n, p, rho):
    """AR(1) columns (same law as README.md:47-50, O(np))."""
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    s = np.sqrt(1 - rho * rho)
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + s * Z[:, j]
    return X


def make_data(rng, n, p, k, rho=0.95, probit=False, coef_scale=1.0):
    X = ar1_design(rng, n, p, rho)
    beta = np.zeros(p)
    nn = rng.choice(p, size=k, replace=False)
    beta[nn] = coef_scale * rng.standard_normal(k)
    mu = X @ beta
    if probit:
        y = ((mu + rng.standard_normal(n)) < 0).astype(np.float64)
    else:
        y = mu + rng.standard_normal(n)
    return X, y, beta


class Recorder:
    """Wraps the reference's RNG entry points; records what it consumes."""

    def __init__(self, N, n, p, chain, lin_mod, probit):
        self.N, self.n, self.p, self.chain = N, n, p, chain
        self.perm = np.zeros((N, p), dtype=np.int32)
        self.u = np.full((N, p), np.nan)
        self.z = np.full((N, p), np.nan)
        self.invgamma = None
        self.p0_draw = []
        self.gamma_draw = []
        self.sweep, self.pos = -1, -1
        self.tn_calls = 0
        self.lin = lin_mod
        self.probit = probit

    def __enter__(self):
        self._shuffle, self._randn = np.random.shuffle, np.random.randn
        self._beta, self._gamma = np.random.beta, np.random.gamma
        self._invgamma = scipy.stats.invgamma
        self._unif = self.lin.random_uniform
        self._tn = self.lin.sample_truncnorm
        rec = self

        def shuffle(inds):
            rec._shuffle(inds)
            rec.sweep += 1
            rec.pos = -1
            rec.perm[rec.sweep] = inds

        def uniform():
            v = rec._unif()
            rec.pos += 1
            rec.u[rec.sweep, rec.pos] = v
            return v

        def randn(*a):
            if a:
                return rec._randn(*a)
            v = rec._randn()
            rec.z[rec.sweep, rec.pos] = v
            return v

        def beta(a, b, size=None):
            v = rec._beta(a, b, size=size)
            rec.p0_draw.append(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
            return v

        def gamma(shape, *a, **k):
            v = rec._gamma(shape, *a, **k)
            rec.gamma_draw.append(float(v))
            return v

        class InvGamma:
            def __init__(self, a):
                self.a = a

            def rvs(self, size):
                v = rec._invgamma(self.a).rvs(size)
                rec.invgamma = np.array(v, dtype=np.float64)
                return v

        def truncnorm(mean, var, b, lower_interval):
            event, obs = divmod(rec.tn_calls, rec.n)
            rec.tn_calls += 1
            return go.lib().orc_truncnorm(TN_SEED, rec.chain, event, obs, mean, var, b,
                                          int(lower_interval))

        np.random.shuffle, np.random.randn = shuffle, randn
        np.random.beta, np.random.gamma = beta, gamma
        scipy.stats.invgamma = InvGamma
        self.lin.random_uniform = uniform
        if self.probit:
            self.lin.sample_truncnorm = truncnorm
        return self

    def __exit__(self, *exc):
        np.random.shuffle, np.random.randn = self._shuffle, self._randn
        np.random.beta, np.random.gamma = self._beta, self._gamma
        scipy.stats.invgamma = self._invgamma
        self.lin.random_uniform = self._unif
        self.lin.sample_truncnorm = self._tn

    def streams(self, hp):
        N = self.N
        out = dict(perm=self.perm, u=self.u, zn=self.z)
        if self.invgamma is not None:
            out["invgamma"] = self.invgamma
        if hp["update_p0"]:
            out["p0_draw"] = np.stack(self.p0_draw).reshape(N, -1)
        if hp["update_tau2"]:
            out["gamma_draw"] = np.array(self.gamma_draw)
        return out


HP_DEFAULT = dict(tau2=1.0, update_tau2=True, tau2_a0=2.0, tau2_b0=1.0,
                  sigma2=1.0, update_sigma2=True, sigma2_a0=2.0, sigma2_b0=1.0,
                  p0=0.9, update_p0=True, min_p0=0.0, p0_a0=1.0, p0_b0=1.0)

CASES = [
    # name, n, p, k, N, chains, probit, hyper overrides, data kwargs
    ("linear_c1", 100, 500, 20, 50, 2, False, {}, {}),
    ("linear_minp0", 60, 120, 6, 50, 1, False, dict(min_p0=0.8), {}),
    ("linear_fixed", 50, 40, 4, 50, 1, False,
     dict(update_tau2=False, update_sigma2=False, update_p0=False, p0=0.95, tau2=0.7,
          sigma2=1.3), {}),
    ("linear_dense", 200, 30, 15, 50, 1, False, dict(p0=0.5), dict(rho=0.5)),
    ("linear_odd", 37, 45, 3, 50, 1, False, {}, dict(rho=0.8)),
    ("probit_small", 120, 80, 5, 50, 2, True, {}, dict(rho=0.5, coef_scale=2.0)),
    ("probit_sigma", 75, 33, 4, 50, 1, True, dict(sigma2=1.7, p0=0.8),
     dict(rho=0.3, coef_scale=2.0)),
]


def reference_run(X, y, probit, hp, N, chain, np_seed, lin=None):
    """One chain of the reference's compiled `_sample_spikeslab` (oracle/_ref) with its native RNG,
    every variate recorded.  Returns (reference outputs, recorded streams keyed like the C ABI's
    blipgpu_streams: perm, u, z, invgamma, p0_draw, gamma_draw).  Used by tests/test_configs_gpu.py
    to replay BASELINE's full-size configurations through the CUDA path."""
    if lin is None:
        ref_loader.load()
        from pyblip.linear import _linear as lin
    n, p = X.shape
    np.random.seed(np_seed)
    zlab = np.asarray(y).astype(np.int64) if probit else np.zeros(1, dtype=np.int64)
    ybuf = np.zeros(n) if probit else np.array(y, dtype=np.float64)
    Xc = np.ascontiguousarray(X, dtype=np.float64)
    with Recorder(N, n, p, chain, lin, probit) as rec:
        ref = lin._sample_spikeslab(
            N=N, X=Xc, y=ybuf, z=zlab, probit=int(probit),
            tau2=hp["tau2"], update_tau2=int(hp["update_tau2"]),
            tau2_a0=hp["tau2_a0"], tau2_b0=hp["tau2_b0"],
            sigma2=hp["sigma2"], update_sigma2=int(hp["update_sigma2"]),
            sigma2_a0=hp["sigma2_a0"], sigma2_b0=hp["sigma2_b0"],
            p0=hp["p0"], update_p0=int(hp["update_p0"]), min_p0=hp["min_p0"],
            p0_a0=hp["p0_a0"], p0_b0=hp["p0_b0"])
    st = rec.streams(hp)
    st["z"] = st.pop("zn")
    return ref, st, rec


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def run_case(lin, name, n, p, k, N, chains, probit, hp_over, data_kw, write=True):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else
                                sum(ord(c) for c in name) + 1000)
    X, y, beta_true = make_data(rng, n, p, k, probit=probit, **data_kw)
    hp = dict(HP_DEFAULT)
    hp.update(hp_over)
    out = dict(X=X, y=y, beta_true=beta_true, probit=np.int32(probit), N=np.int32(N),
               chains=np.int32(chains), tn_seed=np.uint64(TN_SEED))
    for k_, v in hp.items():
        out["hp_" + k_] = np.float64(v)
    worst = {}
    for c in range(chains):
        np.random.seed(1234 + c)  # the reference's NumPy global stream
        zlab = y.astype(np.int64) if probit else np.zeros(1, dtype=np.int64)
        ybuf = np.zeros(n) if probit else y.copy()
        with Recorder(N, n, p, c, lin, probit) as rec:
            ref = lin._sample_spikeslab(
                N=N, X=X, y=ybuf, z=zlab, probit=int(probit),
                tau2=hp["tau2"], update_tau2=int(hp["update_tau2"]),
                tau2_a0=hp["tau2_a0"], tau2_b0=hp["tau2_b0"],
                sigma2=hp["sigma2"], update_sigma2=int(hp["update_sigma2"]),
                sigma2_a0=hp["sigma2_a0"], sigma2_b0=hp["sigma2_b0"],
                p0=hp["p0"], update_p0=int(hp["update_p0"]), min_p0=hp["min_p0"],
                p0_a0=hp["p0_a0"], p0_b0=hp["p0_b0"])
        st = rec.streams(hp)
        orc = go.sample_spikeslab(N, X, y=None if probit else y,
                                  z=zlab if probit else None, probit=probit,
                                  hp=go.hparams(**hp), seed=TN_SEED, chain=c, **st)
        ind_equal = np.array_equal(ref["betas"] != 0, orc["betas"] != 0)
        errs = dict(betas=rel_err(orc["betas"], ref["betas"]),
                    p0s=rel_err(orc["p0s"], ref["p0s"]),
                    tau2s=rel_err(orc["tau2s"], ref["tau2s"]),
                    sigma2s=rel_err(orc["sigma2s"], ref["sigma2s"]))
        if probit:
            errs["y_latent"] = rel_err(orc["y_latent"], ref["y_latent"])
            assert rec.tn_calls == orc["n_events"] * n, (rec.tn_calls, orc["n_events"])
        assert ind_equal, f"{name} chain {c}: indicators differ"
        for kk, v in errs.items():
            assert v <= 1e-11, f"{name} chain {c}: {kk} rel err {v}"
            worst[kk] = max(worst.get(kk, 0.0), v)
        pre = f"c{c}_"
        for kk, v in st.items():
            out[pre + "st_" + kk] = v
        for kk in ("betas", "p0s", "tau2s", "sigma2s") + (("y_latent",) if probit else ()):
            out[pre + "ref_" + kk] = ref[kk]
        if probit:
            out[pre + "n_events"] = np.int64(orc["n_events"])
        nact = (ref["betas"] != 0).sum(1)
        worst["active/sweep"] = float(nact.mean())
    if write:
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    return worst


def check_truncnorm_law(lin):
    """KS-test the keyed truncated normal against scipy and the reference's own."""
    from pyblip.cython_utils import _truncnorm as tn
    res = {}
    for mean, lower in [(0.0, 0), (0.0, 1), (1.3, 0), (-2.0, 0), (2.5, 1), (-0.05, 0)]:
        m = 40000
        mine = np.array([go.lib().orc_truncnorm(7, 3, e, 0, mean, 1.5, 0.0, lower)
                         for e in range(m)])
        sd = np.sqrt(1.5)
        if lower:
            dist = scipy.stats.truncnorm(-np.inf, (0 - mean) / sd, loc=mean, scale=sd)
        else:
            dist = scipy.stats.truncnorm((0 - mean) / sd, np.inf, loc=mean, scale=sd)
        pv = scipy.stats.kstest(mine, dist.cdf).pvalue
        theirs = np.array([tn.sample_truncnorm(mean, 1.5, 0.0, lower) for _ in range(m)])
        pv2 = scipy.stats.ks_2samp(mine, theirs).pvalue
        res[(mean, lower)] = (pv, pv2)
        assert pv > 1e-4 and pv2 > 1e-4, (mean, lower, pv, pv2)
    return res


def main():
    ref_loader.load()
    from pyblip.linear import _linear as lin
    for case in CASES:
        worst = run_case(lin, *case)
        print(f"[pin] {case[0]:14s} OK  " +
              "  ".join(f"{k}={v:.2e}" for k, v in worst.items()))
    res = check_truncnorm_law(lin)
    print("[pin] truncnorm law KS p-values (vs scipy, vs reference):",
          {k: (round(a, 3), round(b, 3)) for k, (a, b) in res.items()})


if __name__ == "__main__" and "--longrun" not in sys.argv:
    main()


# ---------------------------------------------------------------------------
# long-run posterior summaries of the reference (native RNG), for the
# "PIPs within 4 Monte-Carlo standard errors" gate
# ---------------------------------------------------------------------------
def longrun_reference():
    ref_loader.load()
    from pyblip.linear import _linear as lin
    out = {}
    # linear: n=100, p=60, 64 chains x (500 burn + 2000)
    rng = np.random.default_rng(4242)
    X, y, beta = make_data(rng, 100, 60, 5, rho=0.9, coef_scale=1.0)
    chains, burn, N = 64, 500, 2000
    pips, means, p0m = [], [], []
    for c in range(chains):
        np.random.seed(900 + c)
        r = lin._sample_spikeslab(N=N + burn, X=X, y=y.copy(), z=np.zeros(1, dtype=np.int64), probit=0,
                                  tau2=1.0, update_tau2=1, tau2_a0=2.0, tau2_b0=1.0, sigma2=1.0,
                                  update_sigma2=1, sigma2_a0=2.0, sigma2_b0=1.0, p0=0.9, update_p0=1,
                                  min_p0=0.0, p0_a0=1.0, p0_b0=1.0)
        b = r["betas"][burn:]
        pips.append((b != 0).mean(0)); means.append(b.mean(0)); p0m.append(r["p0s"][burn:].mean())
    out.update(lin_X=X, lin_y=y, lin_beta=beta, lin_pips=np.array(pips), lin_means=np.array(means),
               lin_p0=np.array(p0m), lin_burn=np.int32(burn), lin_N=np.int32(N))
    # probit: n=150, p=30, 24 chains x (300 burn + 1000), reference's own truncated normal
    rng = np.random.default_rng(4343)
    Xp, yp, betap = make_data(rng, 150, 30, 3, rho=0.5, probit=True, coef_scale=2.0)
    chains, burn, N = 24, 300, 1000
    pips, means = [], []
    for c in range(chains):
        np.random.seed(950 + c)
        r = lin._sample_spikeslab(N=N + burn, X=Xp, y=np.zeros(150), z=yp.astype(np.int64), probit=1,
                                  tau2=1.0, update_tau2=1, tau2_a0=2.0, tau2_b0=1.0, sigma2=1.0,
                                  update_sigma2=1, sigma2_a0=2.0, sigma2_b0=1.0, p0=0.9, update_p0=1,
                                  min_p0=0.0, p0_a0=1.0, p0_b0=1.0)
        b = r["betas"][burn:]
        pips.append((b != 0).mean(0)); means.append(b.mean(0))
    out.update(pro_X=Xp, pro_y=yp, pro_beta=betap, pro_pips=np.array(pips), pro_means=np.array(means),
               pro_burn=np.int32(burn), pro_N=np.int32(N))
    np.savez_compressed(os.path.join(GOLDEN, "longrun_reference.npz"), **out)
    print("[pin] longrun_reference: linear PIP range", out["lin_pips"].mean(0).min(), out["lin_pips"].mean(0).max(),
          " probit PIP range", out["pro_pips"].mean(0).min(), out["pro_pips"].mean(0).max())


if __name__ == "__main__" and "--longrun" in sys.argv:
    longrun_reference()
