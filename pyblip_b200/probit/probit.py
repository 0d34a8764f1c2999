"""Drop-in ``ProbitSpikeSlab`` (mirrors /root/reference/pyblip/probit/probit.py:13-99)."""
import numpy as np

from ..linear.linear import _SpikeSlabBase, _DENSE_LIMIT_BYTES


class ProbitSpikeSlab(_SpikeSlabBase):
    """
    Spike-and-slab model for probit regression. The arguments are identical to
    those for ``LinearSpikeSlab``, except ``y`` should be binary (1 => latent < 0,
    the reference's convention: utilities.py:302, _truncnorm.pyx:95-99).
    After ``sample``, ``Z`` holds the latent variables ``(chains * N, n)``.
    """
    _probit = True

    @property
    def Z(self):
        if self._Z is None:
            if self._samples is None:
                raise AttributeError("call .sample() first")
            rows, n = self._samples.info.rows, self._samples.info.n
            if 8 * rows * n > _DENSE_LIMIT_BYTES:
                raise MemoryError("latents too large to materialise on the host")
            self._Z = self._samples.latents()
        return self._Z

    @Z.setter
    def Z(self, value):
        self._Z = value
