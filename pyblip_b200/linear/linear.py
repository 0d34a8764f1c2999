"""Drop-in ``LinearSpikeSlab`` backed by the B200 sampler.

Mirrors /root/reference/pyblip/linear/linear.py:10-168 -- same constructor, same
``.sample(N, burn, chains, num_processes, bsize, max_signals_per_block)``, same
attributes (``betas``, ``p0s``, ``tau2s``, ``sigma2s``) with chains concatenated
chain-major and the burn-in removed (linear.py:165-168).  Instead of mapping one
Cython call per chain over a process pool, ``sample`` makes ONE call into
libblipgpu for all chains of this rank.
"""
import numpy as np

from .. import _lib, dist

_DENSE_LIMIT_BYTES = 64 << 30


class _SpikeSlabBase:
    _probit = False

    def __init__(
        self,
        X,
        y,
        p0=0.9,
        p0_a0=1,
        p0_b0=1,
        update_p0=True,
        min_p0=0,
        sigma2=1,
        update_sigma2=True,
        sigma2_a0=2,
        sigma2_b0=1,
        tau2=1,
        tau2_a0=2,
        tau2_b0=1,
        update_tau2=True,
    ):
        self.X = X
        if not self.X.flags['C_CONTIGUOUS']:      # linear.py:83-84
            self.X = np.ascontiguousarray(self.X)
        self.y = y
        self.sigma2, self.sigma2_a0, self.sigma2_b0 = sigma2, sigma2_a0, sigma2_b0
        self.update_sigma2 = update_sigma2
        self.tau2, self.tau2_a0, self.tau2_b0 = tau2, tau2_a0, tau2_b0
        self.update_tau2 = update_tau2
        self.p0, self.p0_a0, self.p0_b0 = p0, p0_a0, p0_b0
        self.update_p0 = update_p0
        self.min_p0 = min_p0
        self._design = None
        self._samples = None
        self._betas = None
        self._Z = None

    # ---- device plumbing -------------------------------------------------
    def _config(self):
        return _lib.SpikeSlabConfig(
            float(self.tau2), int(bool(self.update_tau2)), float(self.tau2_a0), float(self.tau2_b0),
            float(self.sigma2), int(bool(self.update_sigma2)), float(self.sigma2_a0),
            float(self.sigma2_b0),
            float(self.p0), int(bool(self.update_p0)), float(self.min_p0), float(self.p0_a0),
            float(self.p0_b0))

    def _get_design(self):
        if self._design is None:
            self._design = _lib.Design.upload(self.X)
        return self._design

    def sample(
        self,
        N,
        burn=100,
        chains=1,
        num_processes=1,
        bsize=1,
        max_signals_per_block=None,
        *,
        seed=None,
        mode="auto",
        chunk_sweeps=0,
        gram_refresh=0,
        _streams=None,
    ):
        """
        N : int
            Number of samples per chain
        burn : int
            Number of samples to burn per chain
        chains : int
            Number of chains to run
        num_processes : int
            Accepted for compatibility and ignored: every chain of this rank runs
            in one device launch; under ``torchrun`` chains are sharded over ranks.
        bsize : int
            Maximum block size within gibbs sampling. Default: 1.  The device block
            sampler scores all ``2**bsize`` sub-models of a block at once and supports
            ``bsize <= 8`` (the reference documents 2-5); a larger value raises
            ``NotImplementedError`` instead of falling back to another sampler.
        max_signals_per_block : int
            Maximum number of signals allowed per block (``bsize > 1`` only).
        seed : int, keyword-only
            Philox seed.  ``None`` draws one from ``np.random`` so that
            ``np.random.seed(..)`` makes runs reproducible like the reference.
        mode : {"auto", "gram", "residual"}, keyword-only
            Sweep formulation; probit always uses "residual".
        """
        bsize = min(bsize, self.X.shape[1])        # linear.py:149
        # bsize > 1 runs the block sampler (_linear_multi.pyx); None means "no cap" (linear.py:152-155)
        block_kw = dict(bsize=int(bsize), max_signals_per_block=int(max_signals_per_block or 0)) if bsize > 1 else {}
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62))
        rank, world = dist.world()
        c0, c1 = dist.shard_chains(chains, rank, world)
        self.chain_range = (c0, c1)
        self._release()
        if c1 <= c0:
            p = self.X.shape[1]
            self._betas = np.zeros((0, p))
            self.p0s = self.tau2s = self.sigma2s = np.zeros(0)
            if self._probit:
                self._Z = np.zeros((0, self.X.shape[0]))
            return
        design = self._get_design()
        mode_id = {"auto": _lib.MODE_AUTO, "gram": _lib.MODE_GRAM,
                   "residual": _lib.MODE_RESIDUAL}[mode]
        if self._probit:
            smp = _lib.sample_spikeslab(
                design, self._config(), z=np.asarray(self.y).astype(np.int64), probit=True,
                chains=c1 - c0, chain_offset=c0, n_keep=N, burn=burn, seed=seed,
                streams=_streams, mode=mode_id, chunk_sweeps=chunk_sweeps, **block_kw)
        else:
            smp = _lib.sample_spikeslab(
                design, self._config(), y=np.asarray(self.y, dtype=np.float64), probit=False,
                chains=c1 - c0, chain_offset=c0, n_keep=N, burn=burn, seed=seed,
                streams=_streams, mode=mode_id, chunk_sweeps=chunk_sweeps,
                gram_refresh=gram_refresh, **block_kw)
        self._samples = smp
        self.p0s, self.tau2s, self.sigma2s = smp.hyper()

    def _release(self):
        if self._samples is not None:
            self._samples.close()
        self._samples = None
        self._betas = None
        self._Z = None

    # ---- outputs ---------------------------------------------------------
    @property
    def samples_handle(self):
        """Device-resident samples; accepted by ``create_groups`` functions so the
        indicators never leave the GPU."""
        return self._samples

    @property
    def betas(self):
        """``(chains * N, p)`` float64, materialised from the device on first use."""
        if self._betas is None:
            if self._samples is None:
                raise AttributeError("call .sample() first")
            rows, p = self._samples.shape
            if 8 * rows * p > _DENSE_LIMIT_BYTES:
                raise MemoryError(
                    f"dense betas would need {8 * rows * p / 2**30:.1f} GiB; use "
                    ".samples_handle (device counting) or .samples_csr()")
            self._betas = self._samples.dense()
        return self._betas

    @betas.setter
    def betas(self, value):
        self._betas = value

    def samples_csr(self):
        """(indptr, indices, values) of the sparse coefficient matrix."""
        return self._samples.csr()


class LinearSpikeSlab(_SpikeSlabBase):
    """
    Spike-and-slab model for linear regression (B200 implementation).
    Parameters are those of the reference class (linear.py:10-61).
    """
    _probit = False
