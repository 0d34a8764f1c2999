"""Change-point detection front-end (mirrors /root/reference/pyblip/changepoint.py).

The reference builds the dense T x T design ``X[i, j] = 1{i >= j}`` with a Python
loop (changepoint.py:33-35; 3.2 GB at T = 20000) and runs ``LinearSpikeSlab`` on
it.  Here the design stays implicit: the device sampler runs in gram form with the
closed-form Gram ``(X'X)[a, b] = T - max(a, b)`` and ``X'y`` = suffix sums of ``y``.
"""
import numpy as np

from . import _lib, create_groups
from .linear.linear import LinearSpikeSlab


class _ImplicitStepDesign:
    """Stand-in for the dense lower-triangular ones matrix; only the attributes
    ``LinearSpikeSlab`` touches.  ``np.asarray(model.X)`` materialises it."""

    def __init__(self, T):
        self.shape = (T, T)
        self.flags = {'C_CONTIGUOUS': True}

    def __array__(self, dtype=None, copy=None):
        T = self.shape[0]
        return np.tril(np.ones((T, T), dtype=dtype or np.float64))


class ChangepointSpikeSlab(LinearSpikeSlab):
    """``LinearSpikeSlab`` on the cumulative-sum design of length ``len(Y)``."""

    def __init__(self, Y, **lm_kwargs):
        Y = np.asarray(Y, dtype=np.float64)
        super().__init__(X=_ImplicitStepDesign(Y.shape[0]), y=Y, **lm_kwargs)

    def _get_design(self):
        if self._design is None:
            self._design = _lib.Design.changepoint(self.y.shape[0])
        return self._design

    def sample(self, N, burn=100, chains=1, num_processes=1, bsize=1,
               max_signals_per_block=None, **kw):
        # bsize > 1: the block sampler on the same closed-form Gram (block Grams T - max(a, b))
        kw.setdefault("mode", "gram")
        return super().sample(N, burn=burn, chains=chains, num_processes=num_processes, bsize=bsize,
                              max_signals_per_block=max_signals_per_block, **kw)


def changepoint_cand_groups(model, **kwargs):
    """Sequential candidate groups of a fitted model, never discovering time 0
    (changepoint.py:4-11).  Uses the device-resident samples when available."""
    handle = getattr(model, "samples_handle", None)
    if handle is not None:
        dev = create_groups._as_device(handle, zero_first_column=True)
    else:
        dev = create_groups._as_device(model.betas, zero_first_column=True)
    try:
        return create_groups.sequential_groups(samples=dev, **kwargs)
    finally:
        dev.close()


def detect_changepoints(Y, q=0.1, lm_kwargs={}, sample_kwargs={}, blip_kwargs={}):
    """
    Changepoint detection with BLiP using the LinearSpikeSlab sampler (changepoint.py:13-46):
    sampler and PIP counting on the GPU, then the BLiP programs (HiGHS, ``pyblip_b200.blip``).
    """
    from . import blip
    lm = ChangepointSpikeSlab(Y, **lm_kwargs)
    lm.sample(**sample_kwargs)
    cand_groups = changepoint_cand_groups(lm)
    return blip.BLiP(cand_groups=cand_groups, q=q, **blip_kwargs)
