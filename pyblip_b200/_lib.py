"""ctypes binding of libblipgpu.so (the C ABI declared in include/blipgpu.h).

There is no CPU fallback: if the library is missing or no B200 is visible the
calls below raise.  The Python objects here are thin RAII wrappers around the
opaque handles; the reference-facing API lives in ``linear/``, ``probit/``,
``changepoint.py`` and ``create_groups.py``.
"""
import ctypes as C
import os
import sys
import threading
import time
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_C", "libblipgpu.so")

MODE_AUTO, MODE_RESIDUAL, MODE_GRAM = 0, 1, 2
GRAM_KERNEL_AUTO, GRAM_KERNEL_CLASSIC, GRAM_KERNEL_SPECIALIZED, GRAM_KERNEL_CLUSTER = 0, 1, 2, 3
GRAM_KERNEL_BATCHED, GRAM_KERNEL_BATCHED_CLUSTER = 4, 5
GRAM_KERNEL_STREAM = 6
GRAM_KERNEL_SPECIALIZED2 = 7
GRAM_KERNEL_SPECIALIZED2_CLUSTER = 8
DESIGN_DENSE, DESIGN_CHANGEPOINT = 0, 1
DTYPE_F64, DTYPE_U8 = 0, 1
ERR_NUMERIC, ERR_NOMEM, ERR_ARG, ERR_UNSUPPORTED = 4, 3, 2, 5
COUNT_PAIRS_MAX_P = 32768     # blipgpu_count_pairs: columns per call (tensor-core path; the scalar kernel stops at 16384)

# every symbol include/blipgpu.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "blipgpu_abi_version", "blipgpu_sizeof", "blipgpu_last_error", "blipgpu_device_count",
    "blipgpu_ctx_create", "blipgpu_ctx_destroy", "blipgpu_ctx_synchronize", "blipgpu_ctx_timing",
    "blipgpu_ctx_counters", "blipgpu_ctx_event_record", "blipgpu_ctx_event_elapsed_ms",
    "blipgpu_host_alloc", "blipgpu_host_free",
    "blipgpu_design_upload", "blipgpu_design_changepoint", "blipgpu_design_build_gram",
    "blipgpu_design_get_gram", "blipgpu_design_get_xl2", "blipgpu_design_free",
    "blipgpu_sample_spikeslab", "blipgpu_sample_spikeslab_multi",
    "blipgpu_sample_nprior", "blipgpu_nprior_get", "blipgpu_nprior_info", "blipgpu_nprior_bits",
    "blipgpu_nprior_free",
    "blipgpu_samples_info_get", "blipgpu_samples_hyper", "blipgpu_samples_dense",
    "blipgpu_samples_csr", "blipgpu_samples_latents", "blipgpu_samples_free",
    "blipgpu_bits_from_samples", "blipgpu_bits_from_dense", "blipgpu_bits_shape",
    "blipgpu_bits_select_columns", "blipgpu_bits_to_dense", "blipgpu_bits_clear_first_column",
    "blipgpu_count_columns", "blipgpu_count_sequential", "blipgpu_count_groups", "blipgpu_count_pairs", "blipgpu_count_false_rows",
    "blipgpu_bits_free",
    "blipgpu_grid_count", "blipgpu_gridcounts_sizes", "blipgpu_gridcounts_get", "blipgpu_gridcounts_free",
    "blipgpu_overlap_bits",
    "blipgpu_debug_truncnorm",
    "blipgpu_probe_read_bandwidth", "blipgpu_probe_fp64", "blipgpu_probe_latency",
]


class SpikeSlabConfig(C.Structure):
    """``blipgpu_spikeslab_config``: scalars of ``_sample_spikeslab``
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
        ("invgamma", C.c_void_p), ("p0_draw", C.c_void_p), ("gamma_draw", C.c_void_p),
        ("nprop", C.c_int32),
    ]


class MultiStreams(C.Structure):
    """``blipgpu_multi_streams``: injected stream of the block sampler."""
    _fields_ = [
        ("blocks", C.c_void_p), ("u", C.c_void_p), ("z", C.c_void_p),
        ("invgamma", C.c_void_p), ("p0_draw", C.c_void_p), ("gamma_draw", C.c_void_p),
        ("nprop", C.c_int32),
    ]


class SampleOptions(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("chunk_sweeps", C.c_int32), ("keep_latents", C.c_int32),
        ("gram_refresh", C.c_int32), ("gram_kernel", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class SamplesInfo(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("p", C.c_int64), ("n", C.c_int64), ("chains", C.c_int64),
        ("n_keep", C.c_int64), ("nnz", C.c_int64), ("n_events", C.c_int64),
        ("has_latents", C.c_int32), ("mode", C.c_int32),
        ("sweep_ms", C.c_double), ("sweep_launches", C.c_int64), ("bytes_device", C.c_int64),
        ("total_ms", C.c_double), ("n_changes", C.c_int64), ("n_windows", C.c_int64),
        ("n_sweeps_total", C.c_int64),
    ]


class BlipGpuError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load():
    """Load libblipgpu.so; raise loudly if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -m pyblip_b200.build` (needs nvcc). "
                "pyblip_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.blipgpu_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            fn = getattr(lib, name)
            if name not in ("blipgpu_last_error",):
                fn.restype = C.c_int
        _lib = lib
        return lib


def check(rc):
    if rc == 0:
        return
    msg = load().blipgpu_last_error().decode("utf-8", "replace")
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NUMERIC:        # the reference raises RuntimeError (dpotrf) in the same situations
        raise RuntimeError(msg)
    raise BlipGpuError(f"libblipgpu error {rc}: {msg}")


_HOST_PINNED_MIN = 1 << 20


def _host_release(ptr):
    try:
        load().blipgpu_host_free(C.c_void_p(ptr))
    except Exception:      # interpreter shutdown
        pass


def host_empty(count, dtype):
    """1-D result array.  From 1 MB up it lives in a page-locked block of the library
    (``blipgpu_host_alloc``) that goes back to the library's cache when the array and
    every view of it are collected: the device writes such a result with one DMA.
    ``BLIPGPU_NO_PINNED_OUT=1`` (or a failed page-lock) gives a plain ``np.empty``."""
    dtype = np.dtype(dtype)
    nbytes = int(count) * dtype.itemsize
    if nbytes < _HOST_PINNED_MIN or os.environ.get("BLIPGPU_NO_PINNED_OUT"):
        return np.empty(count, dtype=dtype)
    ptr = C.c_void_p()
    if load().blipgpu_host_alloc(C.c_size_t(nbytes), C.byref(ptr)) != 0 or not ptr.value:
        return np.empty(count, dtype=dtype)
    block = type("HostBlock", (C.c_char * nbytes,), {}).from_address(ptr.value)
    weakref.finalize(block, _host_release, ptr.value)
    return np.frombuffer(block, dtype=dtype, count=count)


def device_count():
    n = C.c_int(0)
    check(load().blipgpu_device_count(C.byref(n)))
    return n.value


class Context:
    """One CUDA device (``blipgpu_ctx``)."""
    _cache = {}

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(load().blipgpu_ctx_create(C.c_int(device), C.byref(self._h)))
        self.device = device

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            try:
                ndev = device_count()
            except Exception:
                ndev = 0
            if ndev > 0:
                device %= ndev
        ctx = cls._cache.get(device)
        if ctx is None:
            ctx = cls._cache[device] = Context(device)
        return ctx

    def synchronize(self):
        check(load().blipgpu_ctx_synchronize(self._h))

    def counters(self, reset=False):
        k, h, d = C.c_int64(), C.c_int64(), C.c_int64()
        check(load().blipgpu_ctx_counters(self._h, C.byref(k), C.byref(h), C.byref(d),
                                          C.c_int(int(reset))))
        return dict(kernel_launches=k.value, h2d_bytes=h.value, d2h_bytes=d.value)

    def event_record(self, slot):
        check(load().blipgpu_ctx_event_record(self._h, C.c_int(slot)))

    def event_elapsed_ms(self, slot_begin, slot_end):
        ms = C.c_double()
        check(load().blipgpu_ctx_event_elapsed_ms(self._h, C.c_int(slot_begin), C.c_int(slot_end),
                                                  C.byref(ms)))
        return ms.value

    def probe_read_bandwidth(self, nbytes, passes):
        """GB/s of streaming loads over an `nbytes` buffer (L2-resident if small, HBM if large)."""
        v = C.c_double()
        check(load().blipgpu_probe_read_bandwidth(self._h, C.c_int64(int(nbytes)), C.c_int32(int(passes)), C.byref(v)))
        return v.value

    def probe_fp64(self, kind):
        """TFLOP/s of DFMA (kind 0) or DMMA m8n8k4 (kind 1)."""
        v = C.c_double()
        check(load().blipgpu_probe_fp64(self._h, C.c_int32(int(kind)), C.byref(v)))
        return v.value

    def probe_latency(self, nbytes):
        """Cycles per dependent load over an `nbytes` pointer chase."""
        v = C.c_double()
        check(load().blipgpu_probe_latency(self._h, C.c_int64(int(nbytes)), C.byref(v)))
        return v.value

    def timing(self, reset=False):
        sm, ol = C.c_double(), C.c_double()
        sl, oo = C.c_int64(), C.c_int64()
        check(load().blipgpu_ctx_timing(self._h, C.byref(sm), C.byref(sl), C.byref(ol), C.byref(oo),
                                        C.c_int(int(reset))))
        return dict(sweep_ms=sm.value, sweep_launches=sl.value, other_ms=ol.value,
                    other_launches=oo.value)


class Design:
    """X resident on the device (``blipgpu_design``)."""

    def __init__(self, ctx, handle, n, p, kind):
        self.ctx, self._h, self.n, self.p, self.kind = ctx, handle, n, p, kind

    @classmethod
    def upload(cls, X, ctx=None):
        ctx = ctx or Context.get()
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be a 2-D array")
        h = C.c_void_p()
        check(load().blipgpu_design_upload(ctx._h, C.c_void_p(X.ctypes.data), C.c_int64(X.shape[0]),
                                           C.c_int64(X.shape[1]), C.byref(h)))
        return cls(ctx, h, X.shape[0], X.shape[1], DESIGN_DENSE)

    @classmethod
    def changepoint(cls, T, ctx=None):
        ctx = ctx or Context.get()
        h = C.c_void_p()
        check(load().blipgpu_design_changepoint(ctx._h, C.c_int64(T), C.byref(h)))
        return cls(ctx, h, T, T, DESIGN_CHANGEPOINT)

    def build_gram(self):
        check(load().blipgpu_design_build_gram(self._h))

    def gram(self, r0=0, r1=None):
        r1 = self.p if r1 is None else r1
        out = np.empty((r1 - r0, self.p))
        check(load().blipgpu_design_get_gram(self._h, C.c_int64(r0), C.c_int64(r1),
                                             C.c_void_p(out.ctypes.data)))
        return out

    def xl2(self):
        out = np.empty(self.p)
        check(load().blipgpu_design_get_xl2(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            load().blipgpu_design_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _opt_array(a, dtype, shape):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != tuple(shape):
        raise ValueError(f"injected stream has shape {a.shape}, expected {tuple(shape)}")
    return a


class Samples:
    """Posterior samples kept on the device (``blipgpu_samples``)."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        self._info = None

    @property
    def info(self):
        if self._info is None:
            inf = SamplesInfo()
            check(load().blipgpu_samples_info_get(self._h, C.byref(inf)))
            self._info = inf
        return self._info

    @property
    def shape(self):
        return (self.info.rows, self.info.p)

    def hyper(self):
        rows = self.info.rows
        p0s, tau2s, sigma2s = (host_empty(rows, np.float64) for _ in range(3))
        check(load().blipgpu_samples_hyper(self._h, C.c_void_p(p0s.ctypes.data),
                                           C.c_void_p(tau2s.ctypes.data),
                                           C.c_void_p(sigma2s.ctypes.data)))
        return p0s, tau2s, sigma2s

    def dense(self, row0=0, row1=None):
        row1 = self.info.rows if row1 is None else row1
        out = np.empty((row1 - row0, self.info.p))
        check(load().blipgpu_samples_dense(self._h, C.c_int64(row0), C.c_int64(row1),
                                           C.c_void_p(out.ctypes.data)))
        return out

    def csr(self):
        rows = self.info.rows
        t0 = time.perf_counter()
        indptr = np.empty(rows + 1, dtype=np.int64)
        nnz = self.info.nnz
        indices = host_empty(max(nnz, 1), np.int32)
        values = host_empty(max(nnz, 1), np.float64)
        if os.environ.get("BLIPGPU_LAUNCH_LOG"):
            print(f"[blipgpu] csr(): result arrays in {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
        check(load().blipgpu_samples_csr(self._h, C.c_void_p(indptr.ctypes.data),
                                         C.c_void_p(indices.ctypes.data),
                                         C.c_void_p(values.ctypes.data)))
        return indptr, indices[:nnz], values[:nnz]

    def latents(self, row0=0, row1=None):
        row1 = self.info.rows if row1 is None else row1
        out = np.empty((row1 - row0, self.info.n))
        check(load().blipgpu_samples_latents(self._h, C.c_int64(row0), C.c_int64(row1),
                                             C.c_void_p(out.ctypes.data)))
        return out

    def bits(self, zero_first_column=False):
        h = C.c_void_p()
        check(load().blipgpu_bits_from_samples(self._h, C.c_int32(int(zero_first_column)),
                                               C.byref(h)))
        return Bits(self.ctx, h)

    def close(self):
        if getattr(self, "_h", None):
            load().blipgpu_samples_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_spikeslab(design, cfg, y=None, z=None, probit=False, chains=1, chain_offset=0,
                     n_keep=1, burn=0, seed=0, streams=None, mode=MODE_AUTO, chunk_sweeps=0,
                     keep_latents=True, gram_refresh=0, bsize=1, max_signals_per_block=0,
                     gram_kernel=GRAM_KERNEL_AUTO):
    """All chains of this rank in one call (``blipgpu_sample_spikeslab``, or
    ``blipgpu_sample_spikeslab_multi`` when ``bsize > 1``)."""
    if bsize > 1:
        return _sample_spikeslab_multi(design, cfg, y, z, probit, chains, chain_offset, n_keep, burn,
                                       seed, streams, mode, chunk_sweeps, keep_latents, gram_refresh,
                                       bsize, max_signals_per_block)
    n, p = design.n, design.p
    n_total = n_keep + burn
    keep = []
    if probit:
        z = np.ascontiguousarray(z, dtype=np.int64)
        if z.shape != (n,):
            raise ValueError("z must have shape (n,)")
        yptr, zptr = None, z.ctypes.data
    else:
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.shape != (n,):
            raise ValueError("y must have shape (n,)")
        yptr, zptr = y.ctypes.data, None
    st_ptr = None
    if streams:
        perm = _opt_array(streams.get("perm"), np.int32, (chains, n_total, p))
        u = _opt_array(streams.get("u"), np.float64, (chains, n_total, p))
        zn = _opt_array(streams.get("z"), np.float64, (chains, n_total, p))
        ig = _opt_array(streams.get("invgamma"), np.float64, (chains, n_total))
        gd = _opt_array(streams.get("gamma_draw"), np.float64, (chains, n_total))
        p0d = streams.get("p0_draw")
        nprop = 1
        if p0d is not None:
            p0d = np.ascontiguousarray(p0d, dtype=np.float64)
            if p0d.ndim == 2:
                p0d = p0d[:, :, None]
            if p0d.shape[:2] != (chains, n_total):
                raise ValueError("p0_draw must have shape (chains, sweeps[, nprop])")
            p0d = np.ascontiguousarray(p0d)
            nprop = p0d.shape[2]
        keep += [perm, u, zn, ig, gd, p0d]
        ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        st = Streams(ptr(perm), ptr(u), ptr(zn), ptr(ig), ptr(p0d), ptr(gd), nprop)
        st_ptr = C.byref(st)
    opts = SampleOptions(int(mode), int(chunk_sweeps), int(bool(keep_latents)), int(gram_refresh), int(gram_kernel))
    h = C.c_void_p()
    check(load().blipgpu_sample_spikeslab(
        design.ctx._h, design._h, C.byref(cfg), C.c_void_p(yptr), C.c_void_p(zptr),
        C.c_int32(int(bool(probit))), C.c_int64(chains), C.c_int64(chain_offset),
        C.c_int64(n_keep), C.c_int64(burn), C.c_uint64(seed), st_ptr, C.byref(opts), C.byref(h)))
    del keep
    return Samples(design.ctx, h)


def _sample_spikeslab_multi(design, cfg, y, z, probit, chains, chain_offset, n_keep, burn, seed,
                            streams, mode, chunk_sweeps, keep_latents, gram_refresh, bsize,
                            max_signals_per_block):
    n, p = design.n, design.p
    n_total = n_keep + burn
    nblock = (p + bsize - 1) // bsize
    keep = []
    if probit:
        z = np.ascontiguousarray(z, dtype=np.int64)
        if z.shape != (n,):
            raise ValueError("z must have shape (n,)")
        yptr, zptr = None, z.ctypes.data
    else:
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.shape != (n,):
            raise ValueError("y must have shape (n,)")
        yptr, zptr = y.ctypes.data, None
    st_ptr = None
    if streams:
        blocks = _opt_array(streams.get("blocks"), np.int32, (chains, nblock, bsize))
        u = _opt_array(streams.get("u"), np.float64, (chains, n_total, nblock))
        zn = _opt_array(streams.get("z"), np.float64, (chains, n_total, nblock, bsize))
        ig = _opt_array(streams.get("invgamma"), np.float64, (chains, n_total))
        gd = _opt_array(streams.get("gamma_draw"), np.float64, (chains, n_total))
        p0d = streams.get("p0_draw")
        nprop = 1
        if p0d is not None:
            p0d = np.ascontiguousarray(p0d, dtype=np.float64)
            if p0d.ndim == 2:
                p0d = p0d[:, :, None]
            if p0d.shape[:2] != (chains, n_total):
                raise ValueError("p0_draw must have shape (chains, sweeps[, nprop])")
            p0d = np.ascontiguousarray(p0d)
            nprop = p0d.shape[2]
        keep += [blocks, u, zn, ig, gd, p0d]
        ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        st = MultiStreams(ptr(blocks), ptr(u), ptr(zn), ptr(ig), ptr(p0d), ptr(gd), nprop)
        st_ptr = C.byref(st)
    opts = SampleOptions(int(mode), int(chunk_sweeps), int(bool(keep_latents)), int(gram_refresh))
    h = C.c_void_p()
    check(load().blipgpu_sample_spikeslab_multi(
        design.ctx._h, design._h, C.byref(cfg), C.c_int32(int(bsize)),
        C.c_int32(int(max_signals_per_block or 0)), C.c_void_p(yptr), C.c_void_p(zptr),
        C.c_int32(int(bool(probit))), C.c_int64(chains), C.c_int64(chain_offset),
        C.c_int64(n_keep), C.c_int64(burn), C.c_uint64(seed), st_ptr, C.byref(opts), C.byref(h)))
    del keep
    return Samples(design.ctx, h)


class Bits:
    """Bit-packed ``samples != 0`` on the device (``blipgpu_bits``)."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        N, p = C.c_int64(), C.c_int64()
        check(load().blipgpu_bits_shape(self._h, C.byref(N), C.byref(p)))
        self.N, self.p = N.value, p.value

    @property
    def shape(self):
        return (self.N, self.p)

    @classmethod
    def from_dense(cls, samples, ctx=None):
        ctx = ctx or Context.get()
        samples = np.asarray(samples)
        if samples.ndim != 2:
            raise ValueError("samples must be 2-D")
        if samples.dtype == np.bool_ or samples.dtype == np.uint8:
            arr = np.ascontiguousarray(samples).view(np.uint8)
            dt = DTYPE_U8
        elif samples.dtype == np.float64:
            arr = np.ascontiguousarray(samples)
            dt = DTYPE_F64
        else:
            arr = np.ascontiguousarray(samples != 0).view(np.uint8)
            dt = DTYPE_U8
        h = C.c_void_p()
        check(load().blipgpu_bits_from_dense(ctx._h, C.c_void_p(arr.ctypes.data), C.c_int32(dt),
                                             C.c_int64(arr.shape[0]), C.c_int64(arr.shape[1]),
                                             C.byref(h)))
        return cls(ctx, h)

    def select_columns(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        h = C.c_void_p()
        check(load().blipgpu_bits_select_columns(self._h, C.c_void_p(cols.ctypes.data),
                                                 C.c_int64(cols.shape[0]), C.byref(h)))
        return Bits(self.ctx, h)

    def clear_first_column(self):
        """samples[:, 0] = 0 in place (changepoint.py:7)."""
        check(load().blipgpu_bits_clear_first_column(self._h))

    def to_dense(self):
        out = np.empty((self.N, self.p), dtype=np.uint8)
        check(load().blipgpu_bits_to_dense(self._h, C.c_void_p(out.ctypes.data)))
        return out.view(np.bool_)

    def _out(self, shape, device_out):
        if device_out is not None:
            return device_out, 1
        out = np.zeros(shape, dtype=np.int64)
        return out, 0

    def count_columns(self, device_ptr=None):
        """#rows with a non-zero in each column (int64[p])."""
        if device_ptr is not None:
            check(load().blipgpu_count_columns(self._h, C.c_void_p(device_ptr), C.c_int32(1)))
            return None
        out = np.zeros(self.p, dtype=np.int64)
        check(load().blipgpu_count_columns(self._h, C.c_void_p(out.ctypes.data), C.c_int32(0)))
        return out

    def count_sequential(self, max_size, device_ptr=None):
        """zero_counts[m, j] = #rows with columns j..j+m all zero (int64[max_size, p])."""
        if device_ptr is not None:
            check(load().blipgpu_count_sequential(self._h, C.c_int32(max_size),
                                                  C.c_void_p(device_ptr), C.c_int32(1)))
            return None
        out = np.zeros((max_size, self.p), dtype=np.int64)
        check(load().blipgpu_count_sequential(self._h, C.c_int32(max_size),
                                              C.c_void_p(out.ctypes.data), C.c_int32(0)))
        return out

    def count_groups(self, groups, device_ptr=None):
        """#rows with any non-zero among each group's columns (int64[len(groups)])."""
        G = len(groups)
        offsets = np.zeros(G + 1, dtype=np.int64)
        for i, g in enumerate(groups):
            offsets[i + 1] = offsets[i] + len(g)
        members = np.fromiter((int(j) for g in groups for j in g), dtype=np.int64,
                              count=int(offsets[-1]))
        if members.size == 0:
            members = np.zeros(1, dtype=np.int64)
        if device_ptr is not None:
            check(load().blipgpu_count_groups(self._h, C.c_void_p(offsets.ctypes.data),
                                              C.c_void_p(members.ctypes.data), C.c_int64(G),
                                              C.c_void_p(device_ptr), C.c_int32(1)))
            return None
        out = np.zeros(G, dtype=np.int64)
        if G:
            check(load().blipgpu_count_groups(self._h, C.c_void_p(offsets.ctypes.data),
                                              C.c_void_p(members.ctypes.data), C.c_int64(G),
                                              C.c_void_p(out.ctypes.data), C.c_int32(0)))
        return out

    def count_false_rows(self, groups, device_ptr=None):
        """#rows in which at least one of `groups` has no non-zero column (exact FWER numerator)."""
        G = len(groups)
        offsets = np.zeros(G + 1, dtype=np.int64)
        for i, g in enumerate(groups):
            offsets[i + 1] = offsets[i] + len(g)
        members = np.fromiter((int(j) for g in groups for j in g), dtype=np.int64, count=int(offsets[-1]))
        if members.size == 0:
            members = np.zeros(1, dtype=np.int64)
        if device_ptr is not None:
            check(load().blipgpu_count_false_rows(self._h, C.c_void_p(offsets.ctypes.data),
                                                  C.c_void_p(members.ctypes.data), C.c_int64(G),
                                                  C.c_void_p(device_ptr), C.c_int32(1)))
            return None
        out = np.zeros(1, dtype=np.int64)
        check(load().blipgpu_count_false_rows(self._h, C.c_void_p(offsets.ctypes.data),
                                              C.c_void_p(members.ctypes.data), C.c_int64(G),
                                              C.c_void_p(out.ctypes.data), C.c_int32(0)))
        return int(out[0])

    def count_pairs(self, device_ptr=None):
        """#rows in which both columns are non-zero, for every pair of columns (int64[p, p])."""
        if device_ptr is not None:
            if self.p:
                check(load().blipgpu_count_pairs(self._h, C.c_void_p(device_ptr), C.c_int32(1)))
            return None
        out = np.zeros((self.p, self.p), dtype=np.int64)
        if self.p:
            check(load().blipgpu_count_pairs(self._h, C.c_void_p(out.ctypes.data), C.c_int32(0)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            load().blipgpu_bits_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# neuronized prior
# ---------------------------------------------------------------------------
class NPriorConfig(C.Structure):
    """``blipgpu_nprior_config``: scalars of ``_nprior_sample``
    (/root/reference/pyblip/nprior/_nprior.pyx:56-73)."""
    _fields_ = [
        ("tauw2", C.c_double), ("p0_init", C.c_double), ("min_p0", C.c_double),
        ("update_p0", C.c_int32), ("sigma_a0", C.c_double), ("sigma_b0", C.c_double),
        ("sigma_prior_type", C.c_int32), ("joint_sample_W", C.c_int32),
        ("group_alpha_update", C.c_int32), ("reserved", C.c_int32),
    ]


NPRIOR_STREAM_KEYS = ("perm", "u", "wz", "alpha_init", "w_fresh", "sigma_init", "joint_all",
                      "joint_act", "delta_z", "invgamma", "p0_draw")


class NPriorStreams(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in NPRIOR_STREAM_KEYS]


class NPriorSamples:
    """Dense alphas / ws / betas of all chains on the device (``blipgpu_nprior``)."""
    _WHICH = dict(alphas=0, ws=1, betas=2, sigma2s=3, alpha0s=4, p0s=5)

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        rows, p, ms = C.c_int64(), C.c_int64(), C.c_double()
        check(load().blipgpu_nprior_info(self._h, C.byref(rows), C.byref(p), C.byref(ms)))
        self.rows, self.p, self.sweep_ms = rows.value, p.value, ms.value

    def get(self, name):
        which = self._WHICH[name]
        out = np.empty((self.rows, self.p) if which < 3 else (self.rows,))
        check(load().blipgpu_nprior_get(self._h, C.c_int32(which), C.c_void_p(out.ctypes.data)))
        return out

    def bits(self, zero_first_column=False):
        h = C.c_void_p()
        check(load().blipgpu_nprior_bits(self._h, C.byref(h)))
        bits = Bits(self.ctx, h)
        if zero_first_column:
            bits.clear_first_column()
        return bits

    def close(self):
        if getattr(self, "_h", None):
            load().blipgpu_nprior_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_nprior(design, cfg, y, chains=1, chain_offset=0, n_keep=1, burn=0, seed=0, streams=None):
    """All chains of this rank in one call (``blipgpu_sample_nprior``)."""
    n, p = design.n, design.p
    N = n_keep + burn
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.shape != (n,):
        raise ValueError("y must have shape (n,)")
    st_ptr, keep = None, []
    if streams:
        shapes = dict(perm=(chains, N, p), u=(chains, N, p), wz=(chains, N, p),
                      alpha_init=(chains, p), w_fresh=(chains, N, p), sigma_init=(chains,),
                      joint_all=(chains, N, p), joint_act=(chains, N, p), delta_z=(chains, N),
                      invgamma=(chains, N), p0_draw=(chains, N))
        st = NPriorStreams()
        for k in NPRIOR_STREAM_KEYS:
            a = _opt_array(streams.get(k), np.int32 if k == "perm" else np.float64, shapes[k])
            keep.append(a)
            setattr(st, k, None if a is None else a.ctypes.data)
        st_ptr = C.byref(st)
    h = C.c_void_p()
    check(load().blipgpu_sample_nprior(design.ctx._h, design._h, C.byref(cfg), C.c_void_p(y.ctypes.data),
                                       C.c_int64(chains), C.c_int64(chain_offset), C.c_int64(n_keep),
                                       C.c_int64(burn), C.c_uint64(seed), st_ptr, C.byref(h)))
    del keep
    return NPriorSamples(design.ctx, h)


# ---------------------------------------------------------------------------
# continuous-space region counting
# ---------------------------------------------------------------------------
def grid_count(locs, grid_sizes, circle, extra_centers=None, extra_radii=None, extra_codes=None,
               want_pairs=False, ctx=None):
    """``blipgpu_grid_count``: returns (codes u64, counts u32, pair_codes u64, pair_mult u32)."""
    ctx = ctx or Context.get()
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    N, n_src, d = locs.shape
    grid = np.ascontiguousarray(grid_sizes, dtype=np.float64)
    n_extra = 0
    ecp = erp = ecc = None
    if extra_centers is not None and len(extra_centers) > 0:
        ec = np.ascontiguousarray(extra_centers, dtype=np.float64)
        er = np.ascontiguousarray(extra_radii, dtype=np.float64)
        eco = np.ascontiguousarray(extra_codes, dtype=np.uint64)
        n_extra, ecp, erp, ecc = ec.shape[0], ec.ctypes.data, er.ctypes.data, eco.ctypes.data
    h = C.c_void_p()
    check(load().blipgpu_grid_count(ctx._h, C.c_void_p(locs.ctypes.data), C.c_int64(N), C.c_int64(n_src),
                                    C.c_int32(d), C.c_void_p(grid.ctypes.data), C.c_int32(grid.shape[0]),
                                    C.c_int32(int(bool(circle))), C.c_void_p(ecp), C.c_int64(n_extra),
                                    C.c_void_p(erp), C.c_void_p(ecc), C.c_int32(int(bool(want_pairs))),
                                    C.byref(h)))
    try:
        nk, npair = C.c_int64(), C.c_int64()
        check(load().blipgpu_gridcounts_sizes(h, C.byref(nk), C.byref(npair)))
        codes = np.empty(max(nk.value, 1), dtype=np.uint64)
        counts = np.empty(max(nk.value, 1), dtype=np.uint32)
        pcodes = np.empty(max(npair.value, 1), dtype=np.uint64)
        pmult = np.empty(max(npair.value, 1), dtype=np.uint32)
        check(load().blipgpu_gridcounts_get(h, C.c_void_p(codes.ctypes.data), C.c_void_p(counts.ctypes.data),
                                            C.c_void_p(pcodes.ctypes.data), C.c_void_p(pmult.ctypes.data)))
    finally:
        load().blipgpu_gridcounts_free(h)
    return codes[:nk.value], counts[:nk.value], pcodes[:npair.value], pmult[:npair.value]


def overlap_bits(centers, radii, circle, ctx=None):
    """``blipgpu_overlap_bits``: centers (d, G) float32, radii (G,) float32 -> (G, ceil(G/32)) uint32."""
    ctx = ctx or Context.get()
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    radii = np.ascontiguousarray(radii, dtype=np.float32)
    d, G = centers.shape
    out = np.empty((G, (G + 31) // 32), dtype=np.uint32)
    check(load().blipgpu_overlap_bits(ctx._h, C.c_void_p(centers.ctypes.data), C.c_void_p(radii.ctypes.data),
                                      C.c_int64(G), C.c_int32(d), C.c_int32(int(bool(circle))),
                                      C.c_void_p(out.ctypes.data)))
    return out


def debug_truncnorm(mean, var, b, lower, u, u_at, z, z_at, ctx=None):
    """``blipgpu_debug_truncnorm``: the device arithmetic of ``sample_truncnorm`` on recorded
    uniform / normal sequences (parity hook; the samplers use the keyed stream)."""
    ctx = ctx or Context.get()
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)   # noqa: E731
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)     # noqa: E731
    mean, var, b, u, z = f64(mean), f64(var), f64(b), f64(u), f64(z)
    lower = np.ascontiguousarray(lower, dtype=np.int32)
    u_at, z_at = i64(u_at), i64(z_at)
    m = mean.shape[0]
    out = np.empty(m)
    uu, zu = np.empty(m, dtype=np.int64), np.empty(m, dtype=np.int64)
    ptr = lambda a: C.c_void_p(a.ctypes.data)   # noqa: E731
    check(load().blipgpu_debug_truncnorm(ctx._h, C.c_int64(m), ptr(mean), ptr(var), ptr(b), ptr(lower),
                                         ptr(u), C.c_int64(u.shape[0]), ptr(u_at), ptr(z),
                                         C.c_int64(z.shape[0]), ptr(z_at), ptr(out), ptr(uu), ptr(zu)))
    return out, uu, zu
