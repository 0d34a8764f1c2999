"""Pin the neuronized-prior oracle (oracle/nprior_oracle.c) against the reference's own
compiled kernel and write tests/golden/nprior_*.npz.  TEST INFRASTRUCTURE ONLY.

    python oracle/pin_nprior.py

The reference runs with its native NumPy / libc random sources while every variate is
recorded (by call site and position); only ``sample_truncnorm`` is replaced on the reference
side by the keyed draw, as for probit.  scipy >= 1.18's ``cholesky`` returns a C-ordered
array whose transpose the Cython kernel rejects, so it is wrapped to
return Fortran order -- a runtime shim, no source edit.
"""
import os
import sys

import numpy as np
import scipy.linalg
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gibbs_oracle as go  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.pin_against_reference import make_data, rel_err, GOLDEN  # noqa: E402

SEED = 20261020
NP_TN = 20


class NPRecorder:
    def __init__(self, N, n, p, chain, mod):
        self.N, self.n, self.p, self.chain, self.mod = N, n, p, chain, mod
        self.st = dict(perm=np.zeros((N, p), np.int32), u=np.full((N, p), np.nan),
                       wz=np.full((N, p), np.nan), joint_all=np.full((N, p), np.nan),
                       joint_act=np.full((N, p), np.nan), delta_z=np.full(N, np.nan),
                       p0_draw=np.full(N, np.nan))
        self.big = 0          # number of (N, p) randn calls seen
        self.sweep, self.pos, self.wz_seen = -1, -1, 0
        self.joint_calls = 0  # randn(size) calls since the last shuffle

    def __enter__(self):
        r = self
        self._randn, self._shuffle, self._beta = np.random.randn, np.random.shuffle, np.random.beta
        self._invgamma, self._chol = scipy.stats.invgamma, scipy.linalg.cholesky
        self._unif, self._tn = self.mod.random_uniform, self.mod.sample_truncnorm

        def randn(*a):
            v = r._randn(*a)
            if len(a) == 2:                       # init: alphas then ws
                r.st["alpha_init" if r.big == 0 else "w_fresh"] = (v[0].copy() if r.big == 0 else v.copy())
                r.big += 1
            elif len(a) == 1:                     # inside _sample_w of the *next* sweep
                i = r.sweep + 1
                if r.joint_calls == 0:
                    r.st["joint_all"][i, :] = v
                else:
                    r.st["joint_act"][i, :a[0]] = v
                r.joint_calls += 1
            else:
                if r.big == 2 and r.sweep < 0 and "sigma_init" not in r.st:
                    r.st["sigma_init"] = np.array([v])
                elif r.wz_seen < r.p and r.sweep >= 0 and r.pos >= r.wz_seen:
                    r.st["wz"][r.sweep, r.pos] = v
                    r.wz_seen += 1
                else:
                    r.st["delta_z"][r.sweep] = v
            return v

        def shuffle(inds):
            r._shuffle(inds)
            r.sweep += 1
            r.pos, r.wz_seen, r.joint_calls = -1, 0, 0
            r.st["perm"][r.sweep] = inds

        def uniform():
            v = r._unif()
            r.pos += 1
            r.st["u"][r.sweep, r.pos] = v
            return v

        def truncnorm(mean, var, b, lower_interval):
            f = go.lib().orc_truncnorm_keyed
            f.restype = go.C.c_double
            f.argtypes = [go.C.c_uint64] + [go.C.c_uint32] * 4 + [go.C.c_double] * 3 + [go.C.c_int]
            return f(SEED, r.chain, NP_TN, r.pos, r.sweep, mean, var, b, int(lower_interval))

        def beta(a, b, size=None):
            v = r._beta(a, b, size=size)
            r.st["p0_draw"][r.sweep] = v
            return v

        class InvGamma:
            def __init__(self, a):
                self.a = a

            def rvs(self, size):
                v = r._invgamma(self.a).rvs(size)
                r.st["invgamma"] = np.array(v, dtype=np.float64)
                return v

        def cholesky(a, *args, **kw):
            return np.asfortranarray(r._chol(a, *args, **kw))

        np.random.randn, np.random.shuffle, np.random.beta = randn, shuffle, beta
        scipy.stats.invgamma, scipy.linalg.cholesky = InvGamma, cholesky
        self.mod.random_uniform, self.mod.sample_truncnorm = uniform, truncnorm
        return self

    def __exit__(self, *exc):
        np.random.randn, np.random.shuffle, np.random.beta = self._randn, self._shuffle, self._beta
        scipy.stats.invgamma, scipy.linalg.cholesky = self._invgamma, self._chol
        self.mod.random_uniform, self.mod.sample_truncnorm = self._unif, self._tn


CASES = [
    # name, n, p, k, N, kwargs
    ("nprior_joint_group", 120, 60, 4, 40, dict(tauw2=1.0, joint_sample_W=True, group_alpha_update=True)),
    ("nprior_plain", 80, 50, 3, 40, dict(tauw2=1.0, joint_sample_W=False, group_alpha_update=False)),
    ("nprior_sigprior", 100, 45, 5, 40, dict(tauw2=2.0, joint_sample_W=True, group_alpha_update=False,
                                               sigma_prior_type=1, min_p0=0.5)),
    ("nprior_fixedp0", 60, 33, 3, 40, dict(tauw2=0.5, joint_sample_W=False, group_alpha_update=True,
                                            update_p0=False, p0_init=0.9)),
]


def main():
    ref_loader.load()
    from pyblip.nprior import _nprior as mod
    for name, n, p, k, N, kw in CASES:
        rng = np.random.default_rng(sum(ord(c) for c in name) + 7)
        X, y, beta = make_data(rng, n, p, k, rho=0.6, coef_scale=2.0)
        full = dict(tauw2=1.0, p0_init=1 - min(0.01, 1 / p), min_p0=1e-10, update_p0=True,
                    sigma_a0=5.0, sigma_b0=1.0, sigma_prior_type=0, joint_sample_W=True,
                    group_alpha_update=True)
        full.update(kw)
        np.random.seed(4321)
        with NPRecorder(N, n, p, 0, mod) as rec:
            ref = mod._nprior_sample(N=N, X=X, y=y, tauw2=full["tauw2"], p0_init=full["p0_init"],
                                     min_p0=full["min_p0"], update_p0=int(full["update_p0"]),
                                     sigma_a0=full["sigma_a0"], sigma_b0=full["sigma_b0"],
                                     alpha0_a0=1.0, alpha0_b0=1.0,
                                     sigma_prior_type=int(full["sigma_prior_type"]),
                                     joint_sample_W=int(full["joint_sample_W"]),
                                     group_alpha_update=int(full["group_alpha_update"]),
                                     log_interval=N + 1, time0=0.0)
        st = {k_: v for k_, v in rec.st.items()}
        if not full["joint_sample_W"]:
            st.pop("joint_all"); st.pop("joint_act")
        orc = go.nprior_sample(N, X, y, seed=SEED, chain=0, streams=st, **full)
        assert np.array_equal(ref["betas"] != 0, orc["betas"] != 0), f"{name}: indicators differ"
        errs = {k_: rel_err(orc[k_], ref[k_]) for k_ in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s")}
        for k_, v in errs.items():
            assert v <= 1e-9, f"{name}: {k_} rel err {v}"
        out = dict(X=X, y=y, N=np.int32(N), seed=np.uint64(SEED))
        for k_, v in full.items():
            out["kw_" + k_] = np.float64(v)
        for k_, v in st.items():
            out["st_" + k_] = v
        for k_ in ("alphas", "ws", "betas", "sigma2s", "alpha0s", "p0s"):
            out["ref_" + k_] = ref[k_]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(f"[pin] {name:20s} OK  " + "  ".join(f"{k_}={v:.1e}" for k_, v in errs.items()) +
              f"  active/sweep={np.mean((ref['betas'] != 0).sum(1)):.1f}")


if __name__ == "__main__":
    main()
