"""NumPy stand-in for the device sample store, used ONLY by CPU tests of the host logic in
pyblip_b200/create_groups.py (the product never imports this)."""
import numpy as np

from oracle import counting_oracle as co
from pyblip_b200 import dist


class NumpyDeviceSamples:
    def __init__(self, samples):
        self.S = np.asarray(samples) != 0
        self.N_local, self.p = self.S.shape
        self.N = dist.allreduce_scalar(self.N_local) if dist.world()[1] > 1 else self.N_local

    def select(self, cols):
        return NumpyDeviceSamples(self.S[:, cols])

    def column_counts(self):
        return dist.allreduce_counts(co.column_counts(self.S))

    def sequential_counts(self, max_size):
        return dist.allreduce_counts(co.sequential_zero_counts(self.S, max_size))

    def group_counts(self, groups):
        if len(groups) == 0:
            return np.zeros(0, dtype=np.int64)
        return dist.allreduce_counts(co.group_any_counts(self.S, groups))

    def false_rows(self, groups):
        fd = np.zeros(self.S.shape[0], dtype=bool)
        for g in groups:
            fd |= np.all(self.S[:, list(g)] == 0, axis=1)
        return int(dist.allreduce_scalar(int(fd.sum())) if dist.world()[1] > 1 else fd.sum())

    def pair_counts(self):
        Si = self.S.astype(np.int64)
        return dist.allreduce_counts(Si.T @ Si)

    def to_dense(self):
        return self.S

    def close(self):
        pass


def install(monkeypatch_or_module):
    """Route create_groups._as_device to the NumPy stand-in."""
    from pyblip_b200 import create_groups as cg

    def _as_device(samples, zero_first_column=False):
        if isinstance(samples, NumpyDeviceSamples):
            return samples
        arr = np.asarray(samples) != 0
        if zero_first_column:
            arr = arr.copy()
            arr[:, 0] = False
        return NumpyDeviceSamples(arr)

    if hasattr(monkeypatch_or_module, "setattr"):
        monkeypatch_or_module.setattr(cg, "_as_device", _as_device)
    else:
        cg._as_device = _as_device
