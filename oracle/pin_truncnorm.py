"""Value-level pin of the truncated normal against the reference's own sampler.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs oracle/_ref):

    python oracle/pin_truncnorm.py            # check + (re)write tests/golden/truncnorm_values.npz

The reference's ``sample_truncnorm`` (/root/reference/pyblip/cython_utils/_truncnorm.pyx:88-107)
draws its uniforms from libc ``rand()/RAND_MAX`` (:16-18, :50-53) and its normals from
``np.random.randn()`` (:63).  Both sources are seedable from outside the module:

  1. ``srand(s)`` / ``np.random.seed(s)``, then record K values of ``rand()/RAND_MAX`` and
     ``np.random.randn()`` (legacy RandomState: scalar calls and one vector call consume the
     same gaussian sequence, including the cached second Box-Muller value);
  2. re-seed both, call the REFERENCE's compiled ``sample_truncnorm`` M times in a row over a
     grid of (mean, var, b, lower_interval) that reaches every branch (exponential-proposal
     rejection for a >= 0.15, normal rejection with and without the fold at a >= 0, both tails);
  3. feed the recorded sequences to the oracle's arithmetic (``orc_truncnorm_injected``, the
     same C function that serves the keyed stream) and require the M values to be IDENTICAL,
     bit for bit, and the number of variates consumed to be identical (checked by re-reading
     the generators' next values).

The arguments, the recorded sequences and the reference's values are stored in
``tests/golden/truncnorm_values.npz``; tests/test_oracle.py re-checks the oracle against them on
CPU and tests/test_gibbs_gpu.py runs the DEVICE arithmetic (``blipgpu_debug_truncnorm``) on the
same sequences.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gibbs_oracle as go  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "truncnorm_values.npz")
SEED = 20261020
RAND_MAX = 2147483647


def injected(mean, var, b, lower, u, z):
    """The oracle's arithmetic on recorded variate sequences."""
    lib = go.lib()
    f = lib.orc_truncnorm_injected
    f.restype = C.c_int
    f.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    m = len(mean)
    out = np.zeros(m)
    used = np.zeros(2, dtype=np.int64)
    u_at = np.zeros(m, dtype=np.int64)
    z_at = np.zeros(m, dtype=np.int64)
    mean, var, b = (np.ascontiguousarray(a, dtype=np.float64) for a in (mean, var, b))
    lower = np.ascontiguousarray(lower, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    rc = f(m, mean.ctypes.data, var.ctypes.data, b.ctypes.data, lower.ctypes.data, u.ctypes.data,
           len(u), z.ctypes.data, len(z), out.ctypes.data, used.ctypes.data, u_at.ctypes.data,
           z_at.ctypes.data)
    if rc != 0:
        raise RuntimeError("recorded variate sequence too short")
    return out, used, u_at, z_at


def arguments(rng, m):
    """(mean, var, b, lower) grid: every branch of _truncnorm.pyx:57-106."""
    mean = np.concatenate([
        rng.normal(0.0, 2.0, m // 2),                       # both signs of a, |a| up to ~6
        rng.normal(0.0, 0.05, m // 4),                      # a near 0 and near the 0.15 switch
        np.repeat([0.0, 0.15, -0.15, 0.3, -4.0, 4.0, 7.5, -7.5], m // 32)])
    mean = mean[:m] if len(mean) >= m else np.concatenate([mean, rng.normal(0, 1, m - len(mean))])
    var = np.where(rng.random(m) < 0.5, 1.0, rng.uniform(0.2, 3.0, m))   # probit uses sigma2 = 1 mostly
    b = np.where(rng.random(m) < 0.8, 0.0, rng.normal(0.0, 1.0, m))      # the sampler always passes b = 0
    lower = (rng.random(m) < 0.5).astype(np.int32)
    return mean, var, b, lower


def main():
    ref_loader.load()
    from pyblip.cython_utils import _truncnorm as tn
    libc = C.CDLL(None)
    libc.rand.restype = C.c_int
    rng = np.random.default_rng(SEED)
    m, K = 20000, 400000
    mean, var, b, lower = arguments(rng, m)

    libc.srand(SEED % (2 ** 31))
    u = np.array([libc.rand() for _ in range(K)], dtype=np.float64) / RAND_MAX      # :16-18
    np.random.seed(SEED % (2 ** 31))
    z = np.random.randn(K)

    libc.srand(SEED % (2 ** 31))
    np.random.seed(SEED % (2 ** 31))
    ref = np.array([tn.sample_truncnorm(float(mean[i]), float(var[i]), float(b[i]), int(lower[i]))
                    for i in range(m)])
    next_u = libc.rand() / RAND_MAX              # where the reference's generators stand now
    next_z = np.random.randn()

    out, used, u_at, z_at = injected(mean, var, b, lower, u, z)
    assert np.array_equal(out, ref), f"values differ: max abs {np.max(np.abs(out - ref))}"
    assert u[used[0]] == next_u and z[used[1]] == next_z, "variate consumption differs"
    a = (b - mean) / np.sqrt(var)
    a_eff = np.where(lower == 1, -a, a)
    print(f"[pin] truncnorm: {m} reference values identical; consumed {used[0]} uniforms, {used[1]} normals; "
          f"branches: expo {np.sum(a_eff >= 0.15)}, folded normal {np.sum((a_eff >= 0) & (a_eff < 0.15))}, "
          f"plain normal {np.sum(a_eff < 0)}")
    nu, nz = int(used[0]), int(used[1])
    np.savez_compressed(GOLDEN, mean=mean, var=var, b=b, lower=lower, u=u[:nu], z=z[:nz], ref=ref,
                        u_at=u_at, z_at=z_at)
    print("[pin] wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
