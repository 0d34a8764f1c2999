"""NumPy stand-in for the device sample store, used ONLY by CPU tests of the host logic in
pyblip_b200/create_groups.py (the product never imports this)."""
import numpy as np

from oracle import counting_oracle as co
from pyblip_b200 import create_groups as cg
from pyblip_b200 import dist


class NumpyDeviceSamples(cg._SampleSet):
    def __init__(self, samples, N=None):
        self.S = np.asarray(samples) != 0
        self.N_local, self.p = self.S.shape
        self.N = N if N is not None else (self.N_local if dist.world()[1] == 1 else None)
        self.N_per_rank = np.array([self.N_local]) if dist.world()[1] == 1 else None

    def select(self, cols):
        sub = NumpyDeviceSamples(self.S[:, cols], N=self.rows())
        sub.N_per_rank = self.N_per_rank
        return sub

    def local(self, kind, *args):
        if kind == "columns":
            return co.column_counts(self.S)
        if kind == "sequential":
            return co.sequential_zero_counts(self.S, int(args[0]))
        if kind == "groups":
            return co.group_any_counts(self.S, args[0])
        if kind == "false_rows":
            fd = np.zeros(self.S.shape[0], dtype=bool)
            for g in args[0]:
                fd |= np.all(self.S[:, list(g)] == 0, axis=1)
            return np.array([int(fd.sum())], dtype=np.int64)
        if kind == "pairs":
            Si = self.S.astype(np.int64)
            return Si.T @ Si
        raise ValueError(kind)

    def to_dense(self):
        return self.S

    def close(self):
        pass


def install(monkeypatch_or_module):
    """Route create_groups._as_device to the NumPy stand-in."""

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
