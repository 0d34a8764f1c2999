"""Results in the library's page-locked host blocks (blipgpu_host_alloc): same bytes as through a pageable
destination, blocks recycled when the arrays are collected, views keep a block alive."""
import gc

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _owner(a):
    """The object at the end of an array's base chain (views of views included)."""
    while isinstance(a, np.ndarray) and a.base is not None:
        a = a.base
    return a


def _sample(chains=24, N=40, seed=5):
    from pyblip_b200 import _lib
    from pyblip_b200.linear import LinearSpikeSlab
    rng = np.random.default_rng(17)
    n, p = 300, 700
    X = rng.standard_normal((n, p))
    beta = np.zeros(p)
    beta[rng.choice(p, 12, replace=False)] = rng.standard_normal(12)
    y = X @ beta + rng.standard_normal(n)
    lm = LinearSpikeSlab(X, y)
    design = _lib.Design.upload(X)
    smp = _lib.sample_spikeslab(design, lm._config(), y=y, chains=chains, n_keep=N, burn=5, seed=seed, mode=_lib.MODE_GRAM)
    return design, smp


def test_csr_and_hyper_through_pinned_blocks_equal_the_pageable_path(gpu, monkeypatch):
    from pyblip_b200 import _lib
    design, smp = _sample()
    monkeypatch.setenv("BLIPGPU_NO_PINNED_OUT", "1")
    ref_csr, ref_hyper = smp.csr(), smp.hyper()
    monkeypatch.delenv("BLIPGPU_NO_PINNED_OUT")
    monkeypatch.setattr(_lib, "_HOST_PINNED_MIN", 1)          # the problem is small: pin every result array
    csr, hyper = smp.csr(), smp.hyper()
    assert type(_owner(csr[1])).__name__ == "HostBlock" and type(_owner(hyper[0])).__name__ == "HostBlock"
    assert isinstance(_owner(ref_csr[1]), np.ndarray) and isinstance(_owner(ref_hyper[0]), np.ndarray)
    for a, b in zip(csr, ref_csr):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    for a, b in zip(hyper, ref_hyper):
        assert np.array_equal(a, b)
    assert csr[2].flags.writeable
    # dense rows rebuilt from the pinned CSR equal the dense export
    dense = smp.dense()
    rows, p = dense.shape
    rebuilt = np.zeros_like(dense)
    indptr, indices, values = csr
    for r in range(rows):
        rebuilt[r, indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
    assert np.array_equal(rebuilt, dense)
    smp.close()
    design.close()


def test_blocks_are_recycled_and_views_keep_them_alive(gpu):
    from pyblip_b200 import _lib
    a = _lib.host_empty(40 << 20, np.uint8)
    assert type(_owner(a)).__name__ == "HostBlock"
    addr = a.ctypes.data
    a[:] = 7
    view = a[100:200]
    del a
    gc.collect()
    other = _lib.host_empty(40 << 20, np.uint8)              # the first block is still referenced by `view`
    assert other.ctypes.data != addr
    assert int(view.sum()) == 700
    del view
    gc.collect()
    again = _lib.host_empty(40 << 20, np.uint8)              # ... and now it is idle: handed out again
    assert again.ctypes.data == addr
    small = _lib.host_empty(10, np.float64)                 # below 1 MB: a plain array
    assert small.base is None
    # a request far below an idle block's size does not take it (no 2 GB block for a 40 MB result)
    big = _lib.host_empty(256 << 20, np.uint8)
    big_addr = big.ctypes.data
    del big
    gc.collect()
    modest = _lib.host_empty(40 << 20, np.uint8)
    assert modest.ctypes.data != big_addr
    import ctypes
    assert _lib.load().blipgpu_host_free(ctypes.c_void_p(0)) == 0
    assert _lib.load().blipgpu_host_free(ctypes.c_void_p(12345)) != 0        # not one of ours
