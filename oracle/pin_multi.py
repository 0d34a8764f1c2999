"""Pin the block-sampler oracle (oracle/multi_oracle.c) against the reference and write
golden fixtures.  TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python oracle/pin_multi.py            # check + (re)write tests/golden/multi_*.npz

Per case the reference's own compiled ``pyblip.linear._linear_multi._sample_spikeslab_multi``
(oracle/_ref) runs with its NATIVE random sources while every variate it consumes is recorded
-- the block table after the sweep-0 shuffle, the `_weighted_choice` uniform of every block
visit, the normals of every selected model, the hyper-parameter variates; probit truncated
normals are replaced on the reference side by the keyed draw ``orc_truncnorm`` -- and the
recording is replayed through the C oracle.  Required: identical indicators, <=1e-9 relative (the contract's tolerance; the block algebra goes through
LAPACK in the reference and through written-out loops here)
agreement of betas / p0 / sigma2 / tau2 / latents.  Inputs, the recorded stream and the
REFERENCE's outputs go to tests/golden/<case>.npz.
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
from oracle.pin_against_reference import HP_DEFAULT, TN_SEED, make_data, rel_err  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class MultiRecorder:
    def __init__(self, N, n, p, bsize, chain, mod, probit):
        self.N, self.n, self.p, self.bsize, self.chain = N, n, p, bsize, chain
        self.nblock = go.multi_nblock(p, bsize)
        self.blocks = None
        self.u = np.full((N, self.nblock), np.nan)
        self.z = np.zeros((N, self.nblock, bsize))
        self.invgamma = None
        self.p0_draw, self.gamma_draw = [], []
        self.sweep, self.bi, self.k = -1, -1, 0
        self.tn_calls = 0
        self.mod, self.probit = mod, probit

    def __enter__(self):
        self._shuffle, self._randn = np.random.shuffle, np.random.randn
        self._beta, self._gamma = np.random.beta, np.random.gamma
        self._invgamma = scipy.stats.invgamma
        self._unif, self._tn = self.mod.random_uniform, self.mod.sample_truncnorm
        rec = self

        def shuffle(arr):
            rec._shuffle(arr)
            rec.sweep += 1
            rec.bi = -1
            if rec.blocks is None:          # only this one reaches the kernel's buffer
                rec.blocks = np.array(arr, dtype=np.int32)

        def uniform():
            v = rec._unif()
            rec.bi += 1
            rec.k = 0
            rec.u[rec.sweep, rec.bi] = v
            return v

        def randn(*a):
            if a:
                return rec._randn(*a)
            v = rec._randn()
            rec.z[rec.sweep, rec.bi, rec.k] = v
            rec.k += 1
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
        self.mod.random_uniform = uniform
        if self.probit:
            self.mod.sample_truncnorm = truncnorm
        return self

    def __exit__(self, *exc):
        np.random.shuffle, np.random.randn = self._shuffle, self._randn
        np.random.beta, np.random.gamma = self._beta, self._gamma
        scipy.stats.invgamma = self._invgamma
        self.mod.random_uniform = self._unif
        self.mod.sample_truncnorm = self._tn

    def streams(self, hp):
        out = dict(blocks=self.blocks, u=self.u, zn=self.z)
        if self.invgamma is not None:
            out["invgamma"] = self.invgamma
        if hp["update_p0"]:
            out["p0_draw"] = np.stack(self.p0_draw).reshape(self.N, -1)
        if hp["update_tau2"]:
            out["gamma_draw"] = np.array(self.gamma_draw)
        return out


CASES = [
    # name, n, p, k, N, chains, probit, bsize, max_signals, hyper overrides, data kwargs
    ("multi_b3", 80, 60, 6, 50, 2, False, 3, 0, {}, {}),
    ("multi_b5", 100, 103, 8, 50, 1, False, 5, 0, {}, dict(rho=0.9)),
    ("multi_b2_even", 60, 40, 4, 50, 1, False, 2, 0, dict(p0=0.8), dict(rho=0.7)),       # bsize divides p
    ("multi_b4_wrap", 50, 26, 3, 50, 1, False, 4, 0, {}, dict(rho=0.5)),               # last block wraps: coordinates 0, 1 visited twice
    ("multi_b5_cap2", 70, 45, 5, 50, 1, False, 5, 2, dict(min_p0=0.7), dict(rho=0.8)),  # max_signals_per_block
    ("multi_fixed", 40, 30, 3, 50, 1, False, 3, 0,
     dict(update_tau2=False, update_sigma2=False, update_p0=False, p0=0.9, tau2=0.8, sigma2=1.2), {}),
    ("multi_probit_b3", 120, 42, 4, 50, 2, True, 3, 0, {}, dict(rho=0.5, coef_scale=2.0)),
    ("multi_probit_b2", 90, 25, 3, 50, 1, True, 2, 0, dict(sigma2=1.5, p0=0.8), dict(rho=0.3, coef_scale=2.0)),
]


def run_case(mod, name, n, p, k, N, chains, probit, bsize, maxsig, hp_over, data_kw, write=True):
    rng = np.random.default_rng(sum(ord(c) for c in name) + 2000)
    X, y, beta_true = make_data(rng, n, p, k, probit=probit, **data_kw)
    hp = dict(HP_DEFAULT)
    hp.update(hp_over)
    out = dict(X=X, y=y, beta_true=beta_true, probit=np.int32(probit), N=np.int32(N),
               chains=np.int32(chains), tn_seed=np.uint64(TN_SEED), bsize=np.int32(bsize),
               max_signals=np.int32(maxsig))
    for k_, v in hp.items():
        out["hp_" + k_] = np.float64(v)
    worst = {}
    for c in range(chains):
        np.random.seed(4321 + c)
        zlab = y.astype(np.int64) if probit else np.zeros(1, dtype=np.int64)
        ybuf = np.zeros(n) if probit else y.copy()
        with MultiRecorder(N, n, p, bsize, c, mod, probit) as rec:
            ref = mod._sample_spikeslab_multi(
                N=N, bsize=bsize, X=X, y=ybuf, z=zlab, probit=int(probit),
                tau2=hp["tau2"], update_tau2=int(hp["update_tau2"]),
                tau2_a0=hp["tau2_a0"], tau2_b0=hp["tau2_b0"],
                sigma2=hp["sigma2"], update_sigma2=int(hp["update_sigma2"]),
                sigma2_a0=hp["sigma2_a0"], sigma2_b0=hp["sigma2_b0"],
                p0=hp["p0"], update_p0=int(hp["update_p0"]), min_p0=hp["min_p0"],
                p0_a0=hp["p0_a0"], p0_b0=hp["p0_b0"], max_signals_per_block=maxsig)
        st = rec.streams(hp)
        assert st["blocks"].shape == (rec.nblock, bsize), st["blocks"].shape
        orc = go.sample_spikeslab_multi(N, bsize, X, y=None if probit else y,
                                        z=zlab if probit else None, probit=probit,
                                        hp=go.hparams(**hp), max_signals_per_block=maxsig,
                                        seed=TN_SEED, chain=c, **st)
        assert np.array_equal(ref["betas"] != 0, orc["betas"] != 0), f"{name} chain {c}: indicators differ"
        errs = dict(betas=rel_err(orc["betas"], ref["betas"]), p0s=rel_err(orc["p0s"], ref["p0s"]),
                    tau2s=rel_err(orc["tau2s"], ref["tau2s"]), sigma2s=rel_err(orc["sigma2s"], ref["sigma2s"]))
        if probit:
            # latents are mean + N(0,1)-scale noise truncated at 0: values near the truncation point
            # cancel, so their error is measured against the unit scale of the latent
            errs["y_latent"] = float(np.max(np.abs(orc["y_latent"] - ref["y_latent"])
                                            / np.maximum(np.abs(ref["y_latent"]), 1.0)))
            assert rec.tn_calls == orc["n_events"] * n, (rec.tn_calls, orc["n_events"])
        for kk, v in errs.items():
            assert v <= 1e-9, f"{name} chain {c}: {kk} rel err {v}"
            worst[kk] = max(worst.get(kk, 0.0), v)
        pre = f"c{c}_"
        for kk, v in st.items():
            out[pre + "st_" + kk] = v
        for kk in ("betas", "p0s", "tau2s", "sigma2s") + (("y_latent",) if probit else ()):
            out[pre + "ref_" + kk] = ref[kk]
        if probit:
            out[pre + "n_events"] = np.int64(orc["n_events"])
        worst["active/sweep"] = float((ref["betas"] != 0).sum(1).mean())
        worst["nblock"] = rec.nblock
    if write:
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    return worst


def reference_run_multi(X, y, probit, hp, N, bsize, maxsig, chain, np_seed, mod=None):
    """One chain of the reference's compiled `_sample_spikeslab_multi` (oracle/_ref) with its native RNG,
    every variate recorded.  Returns (reference outputs, streams keyed like blipgpu_multi_streams).
    Used by tests/test_configs_gpu.py at BASELINE's full sizes."""
    if mod is None:
        ref_loader.load()
        from pyblip.linear import _linear_multi as mod
    n, p = X.shape
    np.random.seed(np_seed)
    zlab = np.asarray(y).astype(np.int64) if probit else np.zeros(1, dtype=np.int64)
    ybuf = np.zeros(n) if probit else np.array(y, dtype=np.float64)
    Xc = np.ascontiguousarray(X, dtype=np.float64)
    with MultiRecorder(N, n, p, bsize, chain, mod, probit) as rec:
        ref = mod._sample_spikeslab_multi(
            N=N, bsize=bsize, X=Xc, y=ybuf, z=zlab, probit=int(probit),
            tau2=hp["tau2"], update_tau2=int(hp["update_tau2"]),
            tau2_a0=hp["tau2_a0"], tau2_b0=hp["tau2_b0"],
            sigma2=hp["sigma2"], update_sigma2=int(hp["update_sigma2"]),
            sigma2_a0=hp["sigma2_a0"], sigma2_b0=hp["sigma2_b0"],
            p0=hp["p0"], update_p0=int(hp["update_p0"]), min_p0=hp["min_p0"],
            p0_a0=hp["p0_a0"], p0_b0=hp["p0_b0"], max_signals_per_block=maxsig)
    st = rec.streams(hp)
    st["z"] = st.pop("zn")
    return ref, st


def main():
    ref_loader.load()
    from pyblip.linear import _linear_multi as mod
    for case in CASES:
        worst = run_case(mod, *case)
        print(f"[pin-multi] {case[0]:16s} OK  " + "  ".join(f"{k}={v:.2e}" for k, v in worst.items()))


if __name__ == "__main__" and "--longrun" not in sys.argv:
    main()


# ---------------------------------------------------------------------------
# long-run posterior summaries of the reference's block sampler (native RNG), for the
# "PIPs within 4 Monte-Carlo standard errors" gate:  python oracle/pin_multi.py --longrun
# ---------------------------------------------------------------------------
def longrun_reference():
    ref_loader.load()
    from pyblip.linear import _linear_multi as mod
    rng = np.random.default_rng(4242)                       # same data as longrun_reference.npz (linear)
    X, y, beta = make_data(rng, 100, 60, 5, rho=0.9, coef_scale=1.0)
    chains, burn, N, bsize = 48, 300, 1200, 3
    pips, means, p0m = [], [], []
    for c in range(chains):
        np.random.seed(700 + c)
        r = mod._sample_spikeslab_multi(N=N + burn, bsize=bsize, X=X, y=y.copy(), z=np.zeros(1, dtype=np.int64),
                                        probit=0, tau2=1.0, update_tau2=1, tau2_a0=2.0, tau2_b0=1.0, sigma2=1.0,
                                        update_sigma2=1, sigma2_a0=2.0, sigma2_b0=1.0, p0=0.9, update_p0=1,
                                        min_p0=0.0, p0_a0=1.0, p0_b0=1.0, max_signals_per_block=0)
        b = r["betas"][burn:]
        pips.append((b != 0).mean(0)); means.append(b.mean(0)); p0m.append(r["p0s"][burn:].mean())
    out = dict(X=X, y=y, beta=beta, pips=np.array(pips), means=np.array(means), p0=np.array(p0m),
               burn=np.int32(burn), N=np.int32(N), bsize=np.int32(bsize))
    np.savez_compressed(os.path.join(GOLDEN, "longrun_multi_reference.npz"), **out)
    print("[pin-multi] longrun: PIP range", out["pips"].mean(0).min(), out["pips"].mean(0).max())


if __name__ == "__main__" and "--longrun" in sys.argv:
    longrun_reference()
