"""ctypes front-end of the CPU oracle (``oracle/gibbs_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of gibbs_oracle.c.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgibbs_oracle.so")

KIND_SPIKE_U, KIND_SLAB_Z, KIND_PERM, KIND_HYPER, KIND_TRUNCNORM = 0, 1, 2, 3, 4


class HParams(C.Structure):
    """Mirror of the scalar arguments of ``_sample_spikeslab``
    (/root/reference/pyblip/linear/_linear.pyx:34-53)."""
    _fields_ = [
        ("tau2", C.c_double), ("update_tau2", C.c_int32),
        ("tau2_a0", C.c_double), ("tau2_b0", C.c_double),
        ("sigma2", C.c_double), ("update_sigma2", C.c_int32),
        ("sigma2_a0", C.c_double), ("sigma2_b0", C.c_double),
        ("p0", C.c_double), ("update_p0", C.c_int32),
        ("min_p0", C.c_double), ("p0_a0", C.c_double), ("p0_b0", C.c_double),
    ]


class Streams(C.Structure):
    _fields_ = [
        ("perm", C.c_void_p), ("u", C.c_void_p), ("z", C.c_void_p),
        ("invgamma", C.c_void_p), ("p0_draw", C.c_void_p),
        ("gamma_draw", C.c_void_p), ("nprop", C.c_int32),
    ]


def build(force=False):
    src = [os.path.join(HERE, f) for f in
           ("gibbs_oracle.c", "nprior_oracle.c", "multi_oracle.c")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s"],
                          stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_uniform.restype = C.c_double
        _lib.orc_uniform.argtypes = [C.c_uint64] + [C.c_uint32] * 5
        _lib.orc_normal.restype = C.c_double
        _lib.orc_normal.argtypes = [C.c_uint64] + [C.c_uint32] * 5
        _lib.orc_gamma.restype = C.c_double
        _lib.orc_gamma.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        _lib.orc_beta.restype = C.c_double
        _lib.orc_beta.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_double, C.c_double]
        _lib.orc_truncnorm.restype = C.c_double
        _lib.orc_truncnorm.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_double, C.c_double, C.c_double, C.c_int]
        _lib.orc_permutation.restype = None
        _lib.orc_permutation.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_void_p]
        _lib.orc_philox4x32_10.restype = None
        _lib.orc_philox4x32_10.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _lib.orc_sample_spikeslab.restype = C.c_int
    return _lib


def philox(k0, k1, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(k0, k1, c0, c1, c2, c3, out.ctypes.data)
    return out


def permutation(seed, chain, sweep, p):
    out = np.zeros(p, dtype=np.int32)
    lib().orc_permutation(seed, chain, sweep, p, out.ctypes.data)
    return out


def hparams(tau2=1.0, update_tau2=True, tau2_a0=2.0, tau2_b0=1.0,
            sigma2=1.0, update_sigma2=True, sigma2_a0=2.0, sigma2_b0=1.0,
            p0=0.9, update_p0=True, min_p0=0.0, p0_a0=1.0, p0_b0=1.0):
    """Defaults follow LinearSpikeSlab.__init__ (linear/linear.py:63-80)."""
    return HParams(tau2, int(update_tau2), tau2_a0, tau2_b0,
                   sigma2, int(update_sigma2), sigma2_a0, sigma2_b0,
                   p0, int(update_p0), min_p0, p0_a0, p0_b0)


def _ptr(a):
    return None if a is None else a.ctypes.data


def sample_spikeslab(N, X, y=None, z=None, probit=False, hp=None, seed=0, chain=0,
                     perm=None, u=None, zn=None, invgamma=None, p0_draw=None,
                     gamma_draw=None):
    """One chain, like ``_sample_spikeslab``.  Any stream left as None falls
    back to the keyed Philox stream.  Returns a dict like the reference."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, p = X.shape
    hp = hp or hparams()
    y = np.zeros(n) if y is None else np.ascontiguousarray(y, dtype=np.float64)
    zl = np.zeros(n, dtype=np.int64) if z is None else np.ascontiguousarray(z, dtype=np.int64)
    keep = []

    def prep(a, dt, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        assert a.shape == shape, (a.shape, shape)
        keep.append(a)
        return a

    perm = prep(perm, np.int32, (N, p))
    u = prep(u, np.float64, (N, p))
    zn = prep(zn, np.float64, (N, p))
    invgamma = prep(invgamma, np.float64, (N,))
    gamma_draw = prep(gamma_draw, np.float64, (N,))
    nprop = 1
    if p0_draw is not None:
        p0_draw = np.ascontiguousarray(p0_draw, dtype=np.float64)
        if p0_draw.ndim == 1:
            p0_draw = p0_draw.reshape(N, 1)
        nprop = p0_draw.shape[1]
        keep.append(p0_draw)
    st = Streams(_ptr(perm), _ptr(u), _ptr(zn), _ptr(invgamma), _ptr(p0_draw),
                 _ptr(gamma_draw), nprop)
    betas = np.zeros((N, p))
    p0s, tau2s, sigma2s = np.zeros(N), np.zeros(N), np.zeros(N)
    ylat = np.zeros((N, n)) if probit else None
    nev = C.c_int64(0)
    rc = lib().orc_sample_spikeslab(
        C.c_int(N), C.c_int(n), C.c_int(p), C.c_void_p(X.ctypes.data),
        C.c_void_p(y.ctypes.data), C.c_void_p(zl.ctypes.data), C.c_int(int(probit)),
        C.byref(hp), C.c_uint64(seed), C.c_uint32(chain), C.byref(st),
        C.c_void_p(betas.ctypes.data), C.c_void_p(p0s.ctypes.data),
        C.c_void_p(tau2s.ctypes.data), C.c_void_p(sigma2s.ctypes.data),
        C.c_void_p(_ptr(ylat)), C.byref(nev))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    out = {"betas": betas, "p0s": p0s, "tau2s": tau2s, "sigma2s": sigma2s,
           "n_events": nev.value}
    if probit:
        out["y_latent"] = ylat
    return out


# ---------------------------------------------------------------------------
# neuronized prior (oracle/nprior_oracle.c)
# ---------------------------------------------------------------------------
class NPriorStreams(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("perm", "u", "wz", "alpha_init", "w_fresh", "sigma_init", "joint_all",
                 "joint_act", "delta_z", "invgamma", "p0_draw")]


def nprior_sample(N, X, y, tauw2, p0_init=None, min_p0=1e-10, update_p0=True, sigma_a0=5.0,
                  sigma_b0=1.0, sigma_prior_type=0, joint_sample_W=True, group_alpha_update=True,
                  seed=0, chain=0, streams=None):
    """One chain of ``_nprior_sample`` (/root/reference/pyblip/nprior/_nprior.pyx:56).
    `streams`: dict of injected arrays (see NPriorStreams); missing keys use the keyed stream."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = X.shape
    if p0_init is None:
        p0_init = 1 - min(0.01, 1 / p)
    streams = streams or {}
    keep = {}
    shapes = dict(perm=(N, p), u=(N, p), wz=(N, p), alpha_init=(p,), w_fresh=(N, p),
                  sigma_init=(1,), joint_all=(N, p), joint_act=(N, p), delta_z=(N,),
                  invgamma=(N,), p0_draw=(N,))
    st = NPriorStreams()
    for k, shp in shapes.items():
        a = streams.get(k)
        if a is None:
            setattr(st, k, None)
            continue
        a = np.ascontiguousarray(a, dtype=np.int32 if k == "perm" else np.float64)
        assert a.shape == shp, (k, a.shape, shp)
        keep[k] = a
        setattr(st, k, a.ctypes.data)
    alphas, ws, betas = np.zeros((N, p)), np.zeros((N, p)), np.zeros((N, p))
    sigma2s, alpha0s, p0s = np.ones(N), np.zeros(N), np.zeros(N)
    f = lib().orc_nprior_sample
    f.restype = C.c_int
    rc = f(C.c_int(N), C.c_int(n), C.c_int(p), C.c_void_p(X.ctypes.data), C.c_void_p(y.ctypes.data),
           C.c_double(tauw2), C.c_double(p0_init), C.c_double(min_p0), C.c_int(int(update_p0)),
           C.c_double(sigma_a0), C.c_double(sigma_b0), C.c_int(int(sigma_prior_type)),
           C.c_int(int(joint_sample_W)), C.c_int(int(group_alpha_update)),
           C.c_uint64(seed), C.c_uint32(chain), C.byref(st),
           C.c_void_p(alphas.ctypes.data), C.c_void_p(ws.ctypes.data), C.c_void_p(betas.ctypes.data),
           C.c_void_p(sigma2s.ctypes.data), C.c_void_p(alpha0s.ctypes.data), C.c_void_p(p0s.ctypes.data))
    if rc == 2:
        raise ValueError("Cholesky of the active-set precision failed")
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return dict(alphas=alphas, ws=ws, betas=betas, sigma2s=sigma2s, alpha0s=alpha0s, p0s=p0s)


# ---------------------------------------------------------------------------
# block sampler, bsize > 1 (oracle/multi_oracle.c)
# ---------------------------------------------------------------------------
class MultiStreams(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("u", C.c_void_p), ("z", C.c_void_p),
                ("invgamma", C.c_void_p), ("p0_draw", C.c_void_p), ("gamma_draw", C.c_void_p),
                ("nprop", C.c_int32)]


def multi_nblock(p, bsize):
    """`round((p + bsize - 1) / bsize)` with C integer division (_linear_multi.pyx:298)."""
    f = lib().orc_multi_nblock
    f.restype = C.c_int
    return int(f(C.c_int(p), C.c_int(bsize)))


def multi_models(bsize, max_signals_per_block=0):
    """Bit masks of the non-empty models in the reference's enumeration order (:405-409)."""
    f = lib().orc_multi_models
    f.restype = C.c_int
    ms = max_signals_per_block or bsize
    n = f(C.c_int(bsize), C.c_int(ms), None, C.c_int(0))
    out = np.zeros(n, dtype=np.uint32)
    f(C.c_int(bsize), C.c_int(ms), C.c_void_p(out.ctypes.data), C.c_int(n))
    return out


def sample_spikeslab_multi(N, bsize, X, y=None, z=None, probit=False, hp=None,
                           max_signals_per_block=0, seed=0, chain=0, blocks=None, u=None, zn=None,
                           invgamma=None, p0_draw=None, gamma_draw=None):
    """One chain of ``_sample_spikeslab_multi`` (_linear_multi.pyx:271).  Streams left as None
    use the keyed Philox stream."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, p = X.shape
    hp = hp or hparams()
    nblock = multi_nblock(p, bsize)
    y = np.zeros(n) if y is None else np.ascontiguousarray(y, dtype=np.float64)
    zl = np.zeros(n, dtype=np.int64) if z is None else np.ascontiguousarray(z, dtype=np.int64)
    keep = []

    def prep(a, dt, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        assert a.shape == shape, (a.shape, shape)
        keep.append(a)
        return a

    blocks = prep(blocks, np.int32, (nblock, bsize))
    u = prep(u, np.float64, (N, nblock))
    zn = prep(zn, np.float64, (N, nblock, bsize))
    invgamma = prep(invgamma, np.float64, (N,))
    gamma_draw = prep(gamma_draw, np.float64, (N,))
    nprop = 1
    if p0_draw is not None:
        p0_draw = np.ascontiguousarray(p0_draw, dtype=np.float64)
        if p0_draw.ndim == 1:
            p0_draw = p0_draw.reshape(N, 1)
        nprop = p0_draw.shape[1]
        keep.append(p0_draw)
    st = MultiStreams(_ptr(blocks), _ptr(u), _ptr(zn), _ptr(invgamma), _ptr(p0_draw),
                      _ptr(gamma_draw), nprop)
    betas = np.zeros((N, p))
    p0s, tau2s, sigma2s = np.zeros(N), np.zeros(N), np.zeros(N)
    ylat = np.zeros((N, n)) if probit else None
    nev = C.c_int64(0)
    f = lib().orc_sample_spikeslab_multi
    f.restype = C.c_int
    rc = f(C.c_int(N), C.c_int(bsize), C.c_int(n), C.c_int(p), C.c_void_p(X.ctypes.data),
           C.c_void_p(y.ctypes.data), C.c_void_p(zl.ctypes.data), C.c_int(int(probit)),
           C.byref(hp), C.c_int(int(max_signals_per_block)), C.c_uint64(seed), C.c_uint32(chain),
           C.byref(st), C.c_void_p(betas.ctypes.data), C.c_void_p(p0s.ctypes.data),
           C.c_void_p(tau2s.ctypes.data), C.c_void_p(sigma2s.ctypes.data),
           C.c_void_p(_ptr(ylat)), C.byref(nev))
    if rc == 2:
        raise RuntimeError("dpotrf failed: a block model's matrix is not positive definite")
    if rc == 3:
        raise ValueError("oracle supports 1 <= bsize <= 16")
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    out = {"betas": betas, "p0s": p0s, "tau2s": tau2s, "sigma2s": sigma2s, "n_events": nev.value}
    if probit:
        out["y_latent"] = ylat
    return out
