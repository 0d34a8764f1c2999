"""Pure NumPy/Python restatement of the reference's continuous-space PEP computation.

TEST INFRASTRUCTURE ONLY (same rule as oracle/gibbs_oracle.c).

Restates /root/reference/pyblip/create_groups_cts.py:
  grid_peps                    :103-224
  find_centers                 :69-101   (_find_centers)
  additional_circle_centers    :47-66    (_additional_circle_centers)

Pinning status: PINNED by tests/golden/cts_*.json (outputs of the reference's own grid_peps,
written by oracle/make_cts_golden.py) and by the reference's known-answer test
(test/test_create_groups.py:483-530) restated in tests/test_cts_oracle.py.
"""
from collections import Counter

import numpy as np

TOL = 1e-10


def find_centers(samples, gsize, shape):
    """Keys (center..., radius) of the cells / circles that contain each point."""
    m, d = samples.shape
    corners = np.floor(samples * gsize).astype(float) / gsize
    centers = corners + 1 / (2 * gsize)
    out = [tuple(c) for c in centers]
    if shape == 'circle':
        if d != 2:
            raise NotImplementedError("circle needs d == 2")
        radius = np.sqrt(d) / (2 * gsize)
        if m > 0:
            radius2 = np.power(radius, 2)
            for j in range(d):
                for offset in (-1, 1):
                    shifted = centers.copy()
                    shifted[:, j] += offset / gsize
                    inside = np.power(samples - shifted, 2).sum(axis=1) <= radius2
                    inside &= np.max(shifted, axis=1) < 1 + TOL
                    inside &= np.min(shifted, axis=1) > -TOL
                    out.extend(tuple(c) for c in shifted[inside])
    elif shape == 'square':
        radius = 1 / (2 * gsize)
    else:
        raise ValueError(f"Unrecognized shape={shape}")
    return [tuple(np.around(c, 10)) + (radius,) for c in out]


def grid_peps(locs, grid_sizes, count_signals=False, extra_centers=None, max_pep=0.25,
              shape='square'):
    if np.nanmin(locs) < -TOL or np.nanmax(locs) > 1 + TOL:
        raise ValueError("locs are not normalized")
    N, _, d = locs.shape
    pips = {}
    for j in range(N):
        pts = locs[j][~np.any(np.isnan(locs[j]), axis=1)]
        keys = []
        for g in grid_sizes:
            keys.extend(find_centers(pts, g, shape))
        if extra_centers is not None and extra_centers.shape[0] > 0 and pts.shape[0] > 0:
            diff = extra_centers[:, None, :] - pts[None, :, :]
            dist = np.abs(diff).max(axis=2) if shape == 'square' else np.sqrt(np.power(diff, 2).sum(axis=2))
            mind = dist.min(axis=1)
            for g in grid_sizes:
                radius = 1 / (2 * g)
                if shape == 'circle':
                    radius = np.sqrt(d) * radius
                for nc in np.where(mind <= radius)[0]:
                    key = tuple(extra_centers[nc]) + (radius,)
                    reps = int(np.sum(dist[nc] <= radius)) if count_signals else 1
                    keys.extend([key] * reps)
        if not count_signals:
            for key in set(keys):
                pips[key] = pips[key] + 1 / N if key in pips else 1 / N
        else:
            for key, cnt in Counter(keys).items():
                if key not in pips:
                    pips[key] = {cnt: 1 / N, 'pip': 1 / N}
                else:
                    pips[key][cnt] = pips[key][cnt] + 1 / N if cnt in pips[key] else 1 / N
                    pips[key]['pip'] += 1 / N
    out = {}
    for key, val in pips.items():
        pep = 1 - (val['pip'] if count_signals else val)
        if pep <= max_pep:
            out[key] = val if count_signals else pep
    return out
