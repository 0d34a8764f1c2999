"""Device grid_peps against the reference's golden outputs: keys and PEPs bit-exact."""
import numpy as np
import pytest

import cts_util

pytestmark = pytest.mark.gpu


def test_grid_peps_matches_reference_golden(gpu):
    from pyblip_b200 import create_groups_cts as cts
    for r in cts_util.load():
        got = cts.grid_peps(r["locs"], r["grid"], count_signals=r["count_signals"],
                            extra_centers=r["extra"], max_pep=r["max_pep"], shape=r["shape"])
        assert got == r["want"], (r["name"], r["shape"], r["count_signals"], r["extra"] is not None,
                                  len(got), len(r["want"]))


def test_grid_peps_larger_random_against_oracle(gpu):
    from oracle import cts_oracle as co
    from pyblip_b200 import create_groups_cts as cts
    rng = np.random.default_rng(11)
    N, n_src = 700, 9
    true = rng.random((6, 2))
    locs = np.full((N, n_src, 2), np.nan)
    for i in range(N):
        for k in range(n_src):
            if rng.random() < 0.6:
                locs[i, k] = np.clip(true[rng.integers(6)] + 0.003 * rng.standard_normal(2), 0, 1)
    grid = np.around(np.logspace(np.log10(4), 4, 25))
    extra = rng.random((7, 2))
    for shape in ("square", "circle"):
        for cs in (False, True):
            got = cts.grid_peps(locs, grid, count_signals=cs, extra_centers=extra, max_pep=0.5, shape=shape)
            want = co.grid_peps(locs, grid, count_signals=cs, extra_centers=extra, max_pep=0.5, shape=shape)
            assert got == want and len(got) > 0, (shape, cs)


def test_grid_peps_errors_and_edges(gpu):
    from pyblip_b200 import create_groups_cts as cts
    locs = np.random.default_rng(0).random((5, 3, 2))
    with pytest.raises(ValueError):
        cts.grid_peps(locs * 3, [10])
    with pytest.raises(NotImplementedError):
        cts.grid_peps(np.random.default_rng(0).random((5, 3, 3)), [10], shape='circle')
    with pytest.raises(ValueError):
        cts.grid_peps(locs, [10], shape='hexagon')
    empty = np.full((4, 2, 2), np.nan)
    empty[0, 0] = [0.5, 0.5]
    assert cts.grid_peps(empty, [10], max_pep=0.25) == {}
    out = cts.grid_peps(empty, [10, 10], max_pep=1)            # repeated grid size
    assert out == {(0.55, 0.55, 0.05): 0.75}
    n, s, sc = cts.normalize_locs(np.array([[[2.0, 4.0], [np.nan, np.nan]], [[4.0, 8.0], [3.0, 5.0]]]))
    assert np.nanmin(n) == 0 and np.nanmax(n) == 1 and np.isnan(n[0, 1]).all()


def test_overlap_bits_equal_the_reference_float32_expressions(gpu):
    """Device overlap matrix vs the reference's NumPy float32 expressions (create_groups_cts.py:291-308):
    bit-exact, including centres that touch exactly (|c_a - c_b| == r_a + r_b is NOT an overlap)."""
    from pyblip_b200 import _lib
    rng = np.random.default_rng(5)
    for G, d in [(1, 2), (33, 1), (500, 2), (1031, 3), (4000, 2)]:
        grid = rng.choice([4.0, 10.0, 32.0, 100.0], size=G)
        cells = np.floor(rng.random((d, G)) * grid)
        centers = (cells / grid + 1 / (2 * grid)).astype(np.float32)         # grid centres: exact touching happens
        for circle in ((False, True) if d == 2 else (False,)):
            radii = ((np.sqrt(d) if circle else 1.0) / (2 * grid)).astype(np.float32)
            got = _lib.overlap_bits(centers, radii, circle)
            want = cts_util.numpy_overlap_bits(centers, radii, circle)
            assert np.array_equal(got, want), (G, d, circle)


def test_grid_peps_to_cand_groups_matches_reference_golden(gpu):
    from pyblip_b200 import create_groups_cts as cts
    for r in cts_util.load_groups():
        cgs, comps = cts.grid_peps_to_cand_groups(r["peps"], min_blip_size=r["min_blip_size"], shape=r["shape"],
                                                  max_pep=r["max_pep"], min_pep=r["min_pep"])
        tag = (r["name"], r["shape"], r["count_signals"], r["min_blip_size"], len(r["peps"]))
        assert [[int(x) for x in c] for c in comps] == r["components"], tag
        assert cts_util.encode_groups(cgs) == r["groups"], tag
    # grid_peps -> grid_peps_to_cand_groups end to end on the device
    rng = np.random.default_rng(3)
    locs = np.full((200, 5, 2), np.nan)
    true = rng.random((4, 2)) * 0.8 + 0.1
    for i in range(200):
        for k in range(5):
            if rng.random() < 0.7:
                locs[i, k] = np.clip(true[rng.integers(4)] + 0.01 * rng.standard_normal(2), 0, 1)
    peps = cts.grid_peps(locs, [10, 30, 100], max_pep=0.5, shape="circle")
    cgs, comps = cts.grid_peps_to_cand_groups(peps, shape="circle", max_pep=0.5)
    assert sum(len(c) for c in comps) == len(peps) and sum(len(c) for c in cgs) == len(peps)
