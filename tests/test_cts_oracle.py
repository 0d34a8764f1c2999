"""Continuous-space oracle (oracle/cts_oracle.py) against the reference's golden outputs and
the reference's known-answer test (test/test_create_groups.py:483-530)."""
import numpy as np

import cts_util
from oracle import cts_oracle as co


def test_oracle_equals_reference_golden():
    runs = cts_util.load()
    assert len(runs) >= 60
    for r in runs:
        got = co.grid_peps(r["locs"], r["grid"], count_signals=r["count_signals"],
                           extra_centers=r["extra"], max_pep=r["max_pep"], shape=r["shape"])
        assert got == r["want"], (r["name"], r["shape"], r["count_signals"])


def test_reference_known_answers():
    xs = np.array([0.5502, 0.4502, 0.3000001, 0.549, 0.451, 0.3000002, 0.5501, 0.4501, 0.3000003])
    locs = np.stack([xs, 1 - xs], axis=1).reshape(3, 3, 2)[:, :, :]
    locs = np.array([[[0.5502, 0.4502], [0.3000001, 0.2000001]],
                     [[0.549, 0.451], [0.3000002, 0.2000002]],
                     [[0.5501, 0.4501], [0.3000003, 0.2000003]]])
    peps = co.grid_peps(locs, [10, 100, 1000], max_pep=1, shape='square')
    assert peps[(0.55, 0.45, 1 / 20)] == 0
    np.testing.assert_almost_equal(peps[(0.555, 0.455, 1 / 200)], 1 / 3, decimal=5)
    np.testing.assert_almost_equal(peps[(0.3005, 0.2005, 1 / 2000)], 0, decimal=5)


def test_grid_peps_to_cand_groups_host_logic_matches_reference(monkeypatch):
    """Components in networkx's order and the edge clique cover of pyblip/ecc.py, re-derived in
    pyblip_b200/create_groups_cts.py: identical output lists (groups, PEPs, data, components) on
    the reference's golden runs; the overlap matrix comes from the reference's own NumPy
    expressions here (the device kernel is checked against them in the GPU test)."""
    import cts_util
    from pyblip_b200 import _lib, create_groups_cts as cts
    monkeypatch.setattr(_lib, "overlap_bits", cts_util.numpy_overlap_bits)
    runs = cts_util.load_groups()
    assert len(runs) >= 40
    for r in runs:
        cgs, comps = cts.grid_peps_to_cand_groups(r["peps"], min_blip_size=r["min_blip_size"], shape=r["shape"],
                                                  max_pep=r["max_pep"], min_pep=r["min_pep"])
        tag = (r["name"], r["shape"], r["count_signals"], r["min_blip_size"], len(r["peps"]))
        assert [[int(x) for x in c] for c in comps] == r["components"], tag
        assert cts_util.encode_groups(cgs) == r["groups"], tag
    assert cts.grid_peps_to_cand_groups({}) == ([], [])
