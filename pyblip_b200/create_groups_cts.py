"""Candidate regions when signals live in a continuous d-dimensional space.

Drop-in for ``grid_peps`` / ``normalize_locs`` of
/root/reference/pyblip/create_groups_cts.py:18-224.  The per-sample loops over grid
sizes and points (:152-206) run on the B200 (``blipgpu_grid_count``): integer cell codes
are de-duplicated per sample and counted over samples on the device; the host only turns
each distinct code back into the reference's float key -- with the reference's own
expression, so keys are bit-identical -- and each count k into the reference's running sum
``1/N + ... + 1/N`` (k terms, :190-194), which is not always ``k/N``.
"""
import numpy as np

from . import _lib, dist

TOL = 1e-10


def normalize_locs(locs):
    """Shift/scale ``locs`` (N, num_disc, d) into [0, 1] per dimension (:18-45).
    Returns (norm_locs, shifts, scales); NaNs stay NaN."""
    min_val = np.nanmin(np.nanmin(locs, axis=0), axis=0)
    max_val = np.nanmax(np.nanmax(locs, axis=0), axis=0)
    shifts, scales = min_val, max_val - min_val
    return (locs - shifts) / scales, shifts, scales


def _merge_over_ranks(codes, counts):
    """(code, count) lists of all ranks -> sorted distinct codes with summed counts.  The lists
    travel as one int64 tensor per rank (codes reinterpreted), one padded all_gather."""
    if dist.world()[1] == 1:
        return codes, counts
    local = np.stack([np.ascontiguousarray(codes, dtype=np.uint64).view(np.int64),
                      np.ascontiguousarray(counts).astype(np.int64)])
    parts = dist.allgather_ragged_int64(local)
    allc = np.concatenate([p[0] for p in parts]).view(np.uint64)
    alln = np.concatenate([p[1] for p in parts])
    uniq, inv = np.unique(allc, return_inverse=True)
    merged = np.zeros(uniq.shape[0], dtype=np.int64)
    np.add.at(merged, inv, alln)
    return uniq, merged


def grid_peps(
    locs,
    grid_sizes,
    count_signals=False,
    extra_centers=None,
    max_pep=0.25,
    log_interval=None,
    time0=None,
    shape='square'
):
    """
    Same parameters and return value as the reference (:103-224): a dict mapping
    ``(center_1, ..., center_d, radius)`` to its PEP (or, with ``count_signals``, to
    ``{n_signals: prob, ..., 'pip': prob}``) for every region with ``pep <= max_pep``.
    ``log_interval`` / ``time0`` are accepted and ignored.  Under ``torchrun`` each rank
    passes ITS samples; counts are merged over ranks.
    """
    locs = np.asarray(locs, dtype=np.float64)
    if np.nanmin(locs) < -TOL or np.nanmax(locs) > 1 + TOL:
        raise ValueError(
            "locs are not normalized: apply create_groups_cts.normalize_locs first."
        )
    if shape not in ('square', 'circle'):
        raise ValueError(f"Unrecognized shape={shape}, must be one of 'square', 'circle'")
    N_local, _, d = locs.shape
    if shape == 'circle' and d != 2:
        raise NotImplementedError(
            f"Shape=circle only implemented for 2-d data, but detected {d} dimensions"
        )
    grid_all = np.asarray(list(grid_sizes), dtype=np.float64)
    # a repeated grid size yields the same keys again: the reference's per-sample set() absorbs
    # it, its Counter (count_signals) counts every repeat -- handled through `repeat` below
    grid, repeat = np.unique(grid_all, return_counts=True)
    world = dist.world()[1]
    N = dist.allreduce_scalar(N_local) if world > 1 else N_local

    # radii exactly as the reference writes them (:86-90 for cells, :175-177 for extra centres)
    cell_radius = [np.sqrt(d) / (2 * g) if shape == 'circle' else 1 / (2 * g) for g in grid]
    extra_radius = []
    for g in grid:
        r = 1 / (2 * g)
        extra_radius.append(np.sqrt(d) * r if shape == 'circle' else r)
    n_extra = 0 if extra_centers is None else extra_centers.shape[0]

    bits = {1: 56, 2: 28, 3: 18}.get(d)
    extra_codes = None
    if n_extra:
        # An extra centre whose key equals the key of the grid cell it lies in (same centre after
        # rounding, same radius) is the SAME dict key in the reference, so both must count as one
        # per sample: such a (centre, grid) pair is counted under the cell's code.
        extra_codes = np.empty((n_extra, grid.shape[0]), dtype=np.uint64)
        for g_idx, g in enumerate(grid):
            cells = np.floor(extra_centers * g)
            own = np.around(cells.astype(float) / g + 1 / (2 * g), 10)
            same = np.all(own == extra_centers, axis=1) & (cell_radius[g_idx] == extra_radius[g_idx])
            head = np.uint64(g_idx) << np.uint64(56)
            cell_code = np.full(n_extra, head, dtype=np.uint64)
            for k in range(d):
                cell_code |= (cells[:, k].astype(np.int64) + 1).astype(np.uint64) << np.uint64(bits * k)
            own_code = (np.uint64(1) << np.uint64(63)) | head | np.arange(n_extra, dtype=np.uint64)
            extra_codes[:, g_idx] = np.where(same, cell_code, own_code)
    codes, counts, pcodes, pmult = _lib.grid_count(
        locs, grid, circle=(shape == 'circle'),
        extra_centers=extra_centers if n_extra else None, extra_radii=extra_radius,
        extra_codes=extra_codes, want_pairs=count_signals)
    codes, counts = _merge_over_ranks(codes, counts.astype(np.int64))
    if count_signals and np.any(repeat > 1):
        pmult = pmult * repeat[((pcodes >> np.uint64(56)) & np.uint64(0x7f)).astype(np.int64)].astype(pmult.dtype)

    # k repeated additions of 1/N (:190-194) -- cumulative sums, not k/N
    running = np.cumsum(np.full(max(N, 1), 1 / N)) if N > 0 else np.zeros(1)

    # decode codes -> the reference's float keys
    is_extra = (codes >> np.uint64(63)).astype(bool)
    gi = ((codes >> np.uint64(56)) & np.uint64(0x7f)).astype(np.int64)
    keys = [None] * codes.shape[0]
    for g_idx in range(grid.shape[0]):
        sel = np.where((gi == g_idx) & ~is_extra)[0]
        if sel.size == 0:
            continue
        g = grid[g_idx]
        cells = np.empty((sel.size, d))
        for k in range(d):
            mask = np.uint64((1 << bits) - 1)
            cells[:, k] = ((codes[sel] >> np.uint64(bits * k)) & mask).astype(np.int64) - 1
        centers = np.around(cells / g + 1 / (2 * g), 10)            # :77-84, :101
        radius = cell_radius[g_idx]
        for row, idx in enumerate(sel):
            keys[idx] = tuple(centers[row]) + (radius,)
    for idx in np.where(is_extra)[0]:
        nc = int(codes[idx] & np.uint64((1 << 56) - 1))
        keys[idx] = tuple(extra_centers[nc]) + (extra_radius[int(gi[idx])],)

    filtered = {}
    if not count_signals:
        for key, k in zip(keys, counts):
            pep = 1 - running[k - 1]
            if pep <= max_pep:
                filtered[key] = pep
        return filtered

    # count_signals: distribution of the per-sample multiplicity of every key (:195-206)
    if world > 1:
        parts = dist.allgather_ragged_int64(np.stack([np.ascontiguousarray(pcodes, dtype=np.uint64).view(np.int64),
                                                      np.ascontiguousarray(pmult).astype(np.int64)]))
        pcodes = np.concatenate([p[0] for p in parts]).view(np.uint64)
        pmult = np.concatenate([p[1] for p in parts]).astype(pmult.dtype)
    index = {int(c): i for i, c in enumerate(codes)}
    hist = {}
    if pcodes.size:
        pairs = np.stack([pcodes, pmult.astype(np.uint64)], axis=1)
        uniq, cnt = np.unique(pairs, axis=0, return_counts=True)
        for (c, m), n_s in zip(uniq, cnt):
            hist.setdefault(int(c), {})[int(m)] = running[n_s - 1]
    for key, c, k in zip(keys, codes, counts):
        total = running[k - 1]
        if 1 - total <= max_pep:
            val = dict(hist.get(int(c), {}))
            val['pip'] = total
            filtered[key] = val
    return filtered


# ---------------------------------------------------------------------------
# grid_peps output -> candidate groups for BLiP (create_groups_cts.py:228-384)
# ---------------------------------------------------------------------------
def _edge_clique_cover(nbrs):
    """Greedy edge clique cover of an undirected graph with self-loops on nodes 0..n-1
    (``nbrs[v]``: ascending neighbour list, v itself included), with the decisions of the
    reference's heuristic (pyblip/ecc.py:8-64): edges are visited in (u, v >= u) order, a clique
    starts from a still uncovered edge and is extended by the common neighbour of highest remaining
    degree (ties: smallest index).  Returns a list of sets."""
    n = len(nbrs)
    nset = [set(x) for x in nbrs]
    rem = [set(x) for x in nbrs]                       # the graph with covered edges deleted (G2)
    degrees = np.array([len(x) + (1 if v in nset[v] else 0) for v, x in enumerate(nbrs)], dtype=np.int64)
    cliques = []
    for v1 in range(n):
        for v2 in nbrs[v1]:
            if v2 < v1:
                continue
            if v2 not in rem[v1]:
                continue
            rem[v1].discard(v2); rem[v2].discard(v1)
            degrees[v1] -= 1; degrees[v2] -= 1
            for v in (v1, v2):
                if v in rem[v]:
                    rem[v].discard(v)
                    degrees[v] -= 1
            c = [v1, v2]
            neighbors = (nset[v1] & nset[v2]) - {v1, v2}
            if neighbors and np.any(degrees[list(neighbors)] > 0):
                while len(neighbors) > 0:
                    ln = sorted(neighbors)
                    vnew = ln[int(np.argmax(degrees[ln]))]
                    c.append(vnew)
                    for vold in c:
                        if vnew in rem[vold]:
                            rem[vold].discard(vnew); rem[vnew].discard(vold)
                            degrees[vold] -= 1
                            if vnew != vold:
                                degrees[vnew] -= 1
                    neighbors = (neighbors & nset[vnew]) - {vnew}
            cliques.append(set(c))
    return cliques


def grid_peps_to_cand_groups(
    filtered_peps,
    time0=None,
    min_blip_size=1000,
    verbose=False,
    shape='square',
    max_pep=0.25,
    min_pep=0.001,
):
    """
    Turns the output of ``grid_peps`` into a list of lists of CandidateGroups, one sub-list per
    (merged) connected component of the overlap graph, plus the components themselves
    (create_groups_cts.py:228-384; same parameters, same return value).  The O(ngroups^2) overlap
    matrix is computed on the device in the reference's float32 arithmetic
    (``blipgpu_overlap_bits``); components are found in the order networkx's BFS yields them and
    the clique cover follows pyblip/ecc.py, so group ids and list orders are the reference's.
    ``time0`` / ``verbose`` are accepted and ignored.
    """
    import copy
    from .create_groups import CandidateGroup
    ngroups = len(filtered_peps)
    if ngroups == 0:
        return [], []
    keys = sorted(filtered_peps.keys())
    if isinstance(filtered_peps[keys[0]], dict):
        count_signals = True
        peps_arr = 1 - np.array([filtered_peps[k]['pip'] for k in keys])
    else:
        count_signals = False
        peps_arr = np.array([filtered_peps[k] for k in keys])
    if shape not in ('square', 'circle'):
        raise ValueError(f"Unrecognized shape={shape}, must be one of 'square', 'circle'")

    # Step 1: overlap matrix on the device
    d = len(keys[0]) - 1
    centers = np.zeros((d, ngroups))
    radii = np.array([float(k[-1]) for k in keys])
    for j in range(d):
        centers[j] = np.array([k[j] for k in keys])
        if centers[j].max() > 1 or centers[j].min() < 0:
            raise ValueError("centers must be between 0 and 1 but this is not true")
    radii = radii.astype(np.float32)
    centers = centers.astype(np.float32)
    bits = _lib.overlap_bits(centers, radii, circle=(shape == 'circle'))
    dense = np.unpackbits(bits.view(np.uint8), axis=1, bitorder='little')[:, :ngroups].astype(bool)
    nbrs = [np.flatnonzero(dense[v]).tolist() for v in range(ngroups)]     # ascending, self included

    # Step 2: connected components, in networkx's order (components by smallest node; inside a
    # component the nodes come out of a Python set filled in BFS order, as in the reference)
    seen_all = np.zeros(ngroups, dtype=bool)
    components = []
    for v in range(ngroups):
        if seen_all[v]:
            continue
        seen = {v}
        nextlevel = [v]
        while nextlevel:
            thislevel, nextlevel = nextlevel, []
            for u in thislevel:
                for w in nbrs[u]:
                    if w not in seen:
                        seen.add(w)
                        nextlevel.append(w)
        for u in seen:
            seen_all[u] = True
        components.append(seen)
    merged_components = [[]]
    for c in components:
        if len(merged_components[-1]) + len(c) > min_blip_size:
            merged_components.append([])
        merged_components[-1].extend(list(c))

    # Step 3: candidate groups, locations = cliques of the (heuristic) edge clique cover
    all_cand_groups = []
    for component in merged_components:
        component_cand_groups = []
        component_groups = [[] for _ in component]
        local = {g: i for i, g in enumerate(component)}
        sub_nbrs = [sorted(local[w] for w in nbrs[g] if w in local) for g in component]
        for cliquenum, clique in enumerate(_edge_clique_cover(sub_nbrs)):
            for j in clique:
                component_groups[j].append(cliquenum)
        for ii, j in enumerate(component):
            group = set(component_groups[ii])
            data_dict = dict(radius=radii[j])
            for k in range(d):
                data_dict[f'dim{k}'] = centers[k, j]
            data_dict['center'] = centers[:, j].tolist()
            if count_signals:
                key = keys[j]
                nsignals = np.array([x for x in filtered_peps[key].keys() if x != 'pip'])
                props = np.array([filtered_peps[key][n] for n in nsignals])
                inds = np.argsort(-1 * props)
                nsignals = nsignals[inds]
                cumprops = np.cumsum(props[inds])
                for ell in range(len(nsignals)):
                    if cumprops[ell] >= 1 - max_pep:
                        data_dict['nsignals'] = set(nsignals[0:(ell + 1)].tolist())
                        component_cand_groups.append(CandidateGroup(
                            group=group, pep=max(0, 1 - cumprops[ell]), data=copy.deepcopy(data_dict)))
                    if cumprops[ell] >= 1 - min_pep:
                        break
            else:
                component_cand_groups.append(CandidateGroup(group=group, pep=peps_arr[j], data=data_dict))
        all_cand_groups.append(component_cand_groups)
    return all_cand_groups, merged_components
