"""Drop-in ``NPrior`` (mirrors /root/reference/pyblip/nprior/nprior.py:10-153)."""
import numpy as np

from .. import _lib, dist


class NPrior():
    """
    Neuronized-prior sampler for spike-and-slab regression on the B200.
    Parameters and attributes follow the reference class: after ``sample`` the object holds
    ``alphas``, ``ws``, ``betas`` (``(chains*N, p)``) and ``sigma2s``, ``alpha0s``, ``p0s``.
    ``alpha0_a0`` / ``alpha0_b0`` are accepted and unused, as in the reference kernel
    (_nprior.pyx:465-468 hard-codes the uniform hyper-prior).
    """

    def __init__(
        self,
        X,
        y,
        tauw2,
        p0=None,
        update_p0=True,
        min_p0=1e-10,
        sigma_prior_type=0,
        sigma_a0=5,
        sigma_b0=1,
        alpha0_a0=1,
        alpha0_b0=1
    ):
        self.n, self.p = X.shape
        self.X = X
        self.y = y
        self.tauw2 = tauw2
        self.sigma_prior_type = sigma_prior_type
        self.sigma_a0, self.sigma_b0 = sigma_a0, sigma_b0
        self.alpha0_a0, self.alpha0_b0 = alpha0_a0, alpha0_b0
        if p0 is None:
            p0 = 1 - min(0.01, 1 / self.p)         # nprior.py:74-75
        self.p0_init = p0
        self.update_p0 = update_p0
        self.min_p0 = min_p0
        self._design = None
        self._samples = None

    def _config(self, joint_sample_W, group_alpha_update):
        return _lib.NPriorConfig(float(self.tauw2), float(self.p0_init), float(self.min_p0),
                                 int(bool(self.update_p0)), float(self.sigma_a0), float(self.sigma_b0),
                                 int(self.sigma_prior_type), int(bool(joint_sample_W)),
                                 int(bool(group_alpha_update)), 0)

    def sample(
        self,
        N=100,
        burn=10,
        chains=1,
        num_processes=1,
        joint_sample_W=True,
        group_alpha_update=True,
        log_interval=None,
        *,
        seed=None,
        _streams=None,
    ):
        """Same parameters as the reference (nprior.py:87-125).  ``num_processes`` and
        ``log_interval`` are accepted and ignored; ``seed`` (keyword-only) fixes the stream."""
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62))
        rank, world = dist.world()
        c0, c1 = dist.shard_chains(chains, rank, world)
        if self._samples is not None:
            self._samples.close()
            self._samples = None
        if c1 <= c0:
            z2, z1 = np.zeros((0, self.p)), np.zeros(0)
            self.alphas, self.ws, self.betas = z2, z2.copy(), z2.copy()
            self.sigma2s, self.alpha0s, self.p0s = z1, z1.copy(), z1.copy()
            return
        if self._design is None:
            self._design = _lib.Design.upload(np.ascontiguousarray(self.X, dtype=np.float64))
        smp = _lib.sample_nprior(self._design, self._config(joint_sample_W, group_alpha_update),
                                 np.asarray(self.y, dtype=np.float64), chains=c1 - c0,
                                 chain_offset=c0, n_keep=N, burn=burn, seed=seed, streams=_streams)
        self._samples = smp
        self.alphas, self.ws, self.betas = smp.get("alphas"), smp.get("ws"), smp.get("betas")
        self.sigma2s, self.alpha0s, self.p0s = smp.get("sigma2s"), smp.get("alpha0s"), smp.get("p0s")

    @property
    def samples_handle(self):
        return self._samples
