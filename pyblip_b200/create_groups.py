"""Candidate groups from posterior samples, with the PIP counting on the B200.

Drop-in for the counting entry points of /root/reference/pyblip/create_groups.py:

  ``sequential_groups``   (:272-383)  contiguous groups; counting :321-332
  ``hierarchical_groups`` (:386-469)  tree groups;       counting :459
  ``all_cand_groups``     (:56-149)   marginal PIPs :105-106 + the two above on
                                      prefiltered columns :111-143

``samples`` may be what the reference accepts -- an ``(N, p)`` array where a
non-zero marks a signal -- or a device-resident handle: a fitted
``LinearSpikeSlab`` / ``ProbitSpikeSlab`` (its ``samples_handle``), a
``_lib.Samples`` or a ``_lib.Bits``; then the indicators never leave the GPU.
All array math over the sample axis is integer counting in libblipgpu
(``count_columns`` / ``count_sequential`` / ``count_groups``); a PEP is one float64
divide ``count / N`` and therefore bit-identical to the reference's ``np.mean``.
Under ``torchrun`` every rank counts its own chains and one SUM all-reduce merges
the int64 counts.  Tree building, pre-narrowing and purity filtering are host
logic (small, sequential) and keep the reference's ordering conventions.
"""
import numpy as np
import scipy.cluster.hierarchy as hierarchy
import scipy.spatial.distance as ssd

from . import _lib, dist

MIN_PEP = 1e-15  # for numerical stability (susie branch)


class CandidateGroup():
    """
    A single candidate group as an input to BLiP.  Discovering a group asserts
    there is at least one signal in the group (create_groups.py:13-53).

    Parameters
    ----------
    group : list or set
        Set of locations in the group.
    pep : float
        Posterior error probability (1 - PIP) for the group.
    data : dict
        Miscellaneous other attributes to associate with the group.
    """

    def __init__(self, group, pep, data=None):
        self.group = set(group)
        self.pep = pep
        self.data = dict() if data is None else data

    def __str__(self):
        return f"CandidateGroup(group={self.group}, pep={self.pep}, data={self.data})"

    __repr__ = __str__

    def to_dict(self):
        out = {'group': list(self.group), 'pep': self.pep}
        for key, val in self.data.items():
            out[key] = list(val) if isinstance(val, set) else val
        return out


# ---------------------------------------------------------------------------
# device plumbing
# ---------------------------------------------------------------------------
class _SampleSet:
    """Counting interface over the sample rows of this rank, merged over ranks on request.
    ``N`` is the GLOBAL row count (None until the first merge under torchrun: it rides in slot 0
    of that stage's all-reduce).  Subclasses provide ``local(kind, *args)`` and ``p``, ``N_local``."""

    def size_of(self, kind, *args):
        if kind == "columns":
            return (self.p,)
        if kind == "sequential":
            return (int(args[0]), self.p)
        if kind == "groups":
            return (len(args[0]),)
        if kind == "false_rows":
            return (1,)
        if kind == "pairs":
            return (self.p, self.p)
        raise ValueError(kind)

    # ---- one request, merged over ranks (one all-reduce each; stages use dist.CountBatch) ----
    def _one(self, kind, *args):
        batch = dist.CountBatch()
        batch.add(self, kind, *args)
        return batch.run()[0]

    def rows(self):
        """Global number of sample rows."""
        if self.N is None:
            batch = dist.CountBatch()
            batch.want_rows(self)
            batch.run()
        return self.N

    def column_counts(self):
        return self._one("columns")

    def sequential_counts(self, max_size):
        return self._one("sequential", max_size)

    def group_counts(self, groups):
        if len(groups) == 0:
            return np.zeros(0, dtype=np.int64)
        return self._one("groups", groups)

    def false_rows(self, groups):
        return int(self._one("false_rows", groups)[0])

    def pair_counts(self):
        return self._one("pairs")


class _DeviceSamples(_SampleSet):
    """Bit-packed indicators of this rank's chains on the device."""

    def __init__(self, bits, owned, N=None):
        self.bits = bits
        self.owned = owned
        self.N_local = bits.N
        self.N = N if N is not None else (bits.N if dist.world()[1] == 1 else None)
        self.N_per_rank = np.array([bits.N], dtype=np.int64) if dist.world()[1] == 1 else None
        self.p = bits.p
        self.device_index = bits.ctx.device

    def select(self, cols):
        sub = _DeviceSamples(self.bits.select_columns(cols), True, N=self.rows())
        sub.N_per_rank = self.N_per_rank
        return sub

    def local(self, kind, *args):
        """This rank's counts, on the host."""
        if kind == "columns":
            return self.bits.count_columns()
        if kind == "sequential":
            return self.bits.count_sequential(int(args[0]))
        if kind == "groups":
            return self.bits.count_groups(args[0])
        if kind == "false_rows":
            return np.array([self.bits.count_false_rows(args[0])], dtype=np.int64)
        if kind == "pairs":
            return self.bits.count_pairs()
        raise ValueError(kind)

    def local_into(self, kind, device_ptr, *args):
        """The same counts written by the kernels into device memory (``out_is_device``)."""
        if kind == "columns":
            self.bits.count_columns(device_ptr=device_ptr)
        elif kind == "sequential":
            self.bits.count_sequential(int(args[0]), device_ptr=device_ptr)
        elif kind == "groups":
            self.bits.count_groups(args[0], device_ptr=device_ptr)
        elif kind == "false_rows":
            self.bits.count_false_rows(args[0], device_ptr=device_ptr)
        elif kind == "pairs":
            self.bits.count_pairs(device_ptr=device_ptr)
        else:
            raise ValueError(kind)

    def to_dense(self):
        return self.bits.to_dense()

    def close(self):
        if self.owned:
            self.bits.close()


def _as_device(samples, zero_first_column=False):
    """Accepts an array, a fitted sampler, a Samples handle or Bits."""
    if isinstance(samples, _DeviceSamples):
        return samples
    if isinstance(samples, _lib.Bits):
        return _DeviceSamples(samples, False)
    handle = getattr(samples, "samples_handle", None)
    if handle is not None:
        samples = handle
    if isinstance(samples, (_lib.Samples, _lib.NPriorSamples)):
        return _DeviceSamples(samples.bits(zero_first_column=zero_first_column), True)
    arr = np.asarray(samples)
    if arr.ndim != 2:
        raise ValueError("samples must be an (N, p) array")
    if zero_first_column:
        arr = arr != 0
        arr[:, 0] = False
    return _DeviceSamples(_lib.Bits.from_dense(arr), True)


# ---------------------------------------------------------------------------
# purity filter (create_groups.py:543-577) -- host, only active if min_purity > 0
# ---------------------------------------------------------------------------
def filter_by_purity(cand_groups, min_purity=0, corr_matrix=None, X=None):
    """Keeps the groups whose minimum absolute pairwise correlation is at least
    ``min_purity``."""
    if min_purity == 0:
        return cand_groups
    if corr_matrix is None and X is None:
        raise ValueError("To filter by purity, either corr_matrix or X (design matrix) must be provided.")
    if corr_matrix is None:
        corr_matrix = np.corrcoef(X.T)
    absC = np.abs(corr_matrix)
    kept = []
    for cg in cand_groups:
        members = np.array(list(cg.group))
        cg.data['purity'] = 1 if len(members) == 1 else absC[np.ix_(members, members)].min()
        if cg.data['purity'] >= min_purity:
            kept.append(cg)
    return kept


# ---------------------------------------------------------------------------
# sequential groups
# ---------------------------------------------------------------------------
def _sequential_peps_from_counts(zero_counts, N, p, max_size):
    """all_PEPs[m][j] = (#rows with columns j..j+m all zero) / N, j < p - m."""
    return {m: zero_counts[m, :p - m] / N for m in range(max_size)}


def _sequential_peps_susie(susie_alphas, max_size):
    # SuSiE branch (:333-343): L x p alphas, L is tiny -> plain host arithmetic
    L, p = susie_alphas.shape
    cumalphas = np.zeros((L, p + 1))
    cumalphas[:, 1:] = np.cumsum(susie_alphas, axis=1)
    peps = {}
    for m in range(max_size):
        window = 1 - (cumalphas[:, (m + 1):] - cumalphas[:, :p - m])
        window[window < MIN_PEP] = MIN_PEP
        peps[m] = np.exp(np.log(window).sum(axis=0))
    return peps


def _sequential_from_peps(all_PEPs, max_size, q, max_pep, prenarrow):
    """Window PEPs -> candidate groups (create_groups.py:348-378)."""
    # start index of every admissible window of size m+1 (strict <, :350)
    active_inds = {m: np.where(all_PEPs[m] < max_pep)[0] for m in range(max_size)}

    if prenarrow:
        # A window is redundant once a sub-window already has PEP < q/2 (:356-368).
        # Set expressions are kept in the reference's form so the resulting order of
        # the surviving indices (a Python-set order) is the same.
        elim_inds = set(np.where(all_PEPs[0] < q / 2)[0].tolist())
        for m in range(1, max_size):
            elim_inds = elim_inds.union(set([x - 1 for x in elim_inds]))
            update = set(active_inds[m].tolist()) - elim_inds
            active_inds[m] = np.array(list(update))
            elim_inds = elim_inds.union(set(np.where(all_PEPs[m] < q / 2)[0].tolist()))

    cand_groups = []
    for m in range(max_size):
        for ind in active_inds[m]:
            cand_groups.append(CandidateGroup(
                group=set(range(ind, ind + m + 1)), pep=all_PEPs[m][ind]))
    return cand_groups


def sequential_groups(
    samples=None,
    susie_alphas=None,
    q=0,
    max_pep=0.25,
    max_size=25,
    prenarrow=False,
    min_purity=0.0,
    X=None,
    corr_matrix=None,
):
    """
    Calculates all sequential candidate groups below max_size.

    Same parameters and return value as the reference
    (create_groups.py:272-320): a list of ``CandidateGroup`` objects for every
    contiguous window ``{j, ..., j+m}``, ``m < max_size``, with ``pep < max_pep``.
    Under torchrun the window counts (and the row count) of all ranks are merged by
    one all-reduce.
    """
    if samples is not None:
        dev = _as_device(samples)
        try:
            p = dev.p
            max_size = min(max_size, p)
            zero_counts = dev.sequential_counts(max_size)       # one stage: [N | max_size x p]
            all_PEPs = _sequential_peps_from_counts(zero_counts, dev.N, p, max_size)
        finally:
            if dev is not samples:
                dev.close()
    elif susie_alphas is not None:
        p = susie_alphas.shape[1]
        max_size = min(max_size, p)
        all_PEPs = _sequential_peps_susie(susie_alphas, max_size)
    else:
        raise ValueError("Either samples or susie_alphas must be specified.")
    cand_groups = _sequential_from_peps(all_PEPs, max_size, q, max_pep, prenarrow)
    return filter_by_purity(cand_groups, min_purity=min_purity, X=X, corr_matrix=corr_matrix)


# ---------------------------------------------------------------------------
# hierarchical groups
# ---------------------------------------------------------------------------
def _dedup_list_of_lists(x):
    return list(set(tuple(sorted(i)) for i in x))


def _tree_groups(link, p):
    """Leaf sets of every node of a scipy linkage, in the breadth-first order
    (root first, left child before right) in which the reference visits the tree
    (create_groups.py:471-486), built bottom-up from the linkage matrix instead of
    calling ``pre_order()`` on every node."""
    nnode = 2 * p - 1
    leaves = [None] * nnode
    for i in range(p):
        leaves[i] = [i]
    left = np.full(nnode, -1, dtype=np.int64)
    right = np.full(nnode, -1, dtype=np.int64)
    for k in range(p - 1):
        a, b = int(link[k, 0]), int(link[k, 1])
        left[p + k], right[p + k] = a, b
        leaves[p + k] = leaves[a] + leaves[b]
    order, head = [nnode - 1], 0
    while head < len(order):
        node = order[head]
        head += 1
        if left[node] >= 0:
            order.append(int(left[node]))
            order.append(int(right[node]))
    return [leaves[node] for node in order]


def _dist_matrices_to_groups(dist_matrices, cluster_funcs=None):
    """
    Groups from single / average / complete linkage on a distance matrix.
    As in the reference (:491-521) a list of matrices is accepted, each matrix has
    its diagonal zeroed IN PLACE, and only the LAST matrix contributes groups
    (the accumulator is reset inside the loop, :502-514).
    """
    if isinstance(dist_matrices, np.ndarray):
        dist_matrices = [dist_matrices]
    p = dist_matrices[0].shape[0]
    if cluster_funcs is None:
        cluster_funcs = [hierarchy.single, hierarchy.average, hierarchy.complete]
    all_groups = []
    for idx, dist_matrix in enumerate(dist_matrices):
        dist_matrix -= np.diag(np.diag(dist_matrix))     # in place, like the reference
        if idx != len(dist_matrices) - 1:
            continue                                      # overwritten below anyway
        sym = (dist_matrix.T + dist_matrix) / 2
        condensed = ssd.squareform(sym)
        all_groups = []
        for cluster_func in cluster_funcs:
            link = cluster_func(condensed)
            all_groups.extend(_tree_groups(link, p))
            all_groups = _dedup_list_of_lists(all_groups)
    return all_groups


def _samples_dist_matrix(samples):
    """corrcoef of the indicator columns, with one all-zero and one all-one row
    appended so that no column is constant, plus one (:523-535).  Host code: it
    feeds the (host) clustering; ``samples`` is a dense boolean matrix."""
    p = samples.shape[1]
    precorr = np.concatenate([samples, np.zeros((1, p)), np.ones((1, p))], axis=0)
    return np.corrcoef(precorr.T) + 1


# above this many indicator entries the correlation matrix comes from integer pair counts on the
# device instead of a dense host copy (the reference's np.corrcoef needs 8 bytes per entry on the host)
_DEVICE_CORR_MIN_ENTRIES = 50_000_000


def _corr_from_pair_counts(pair_counts, N):
    """The matrix of `_samples_dist_matrix` from exact integer counts: with the all-zero and the
    all-one row appended (:527-533) there are N' = N + 2 rows, column sums c + 1 and co-occurrence
    counts n11 + 1, and the Pearson correlation of two 0/1 columns is
    (N' n11' - c_a' c_b') / sqrt((N' c_a' - c_a'^2)(N' c_b' - c_b'^2)).
    All products stay below 2^53, so every entry is one correctly rounded expression of exact
    integers (np.corrcoef's own result depends on BLAS summation order in the last bits)."""
    n11 = pair_counts.astype(np.float64) + 1.0
    c = np.diag(n11).copy()                         # = column counts + 1
    Np = float(N + 2)
    var = Np * c - c * c
    corr = (Np * n11 - np.outer(c, c)) / np.sqrt(np.outer(var, var))
    np.clip(corr, -1.0, 1.0, out=corr)              # np.corrcoef clips too
    return corr + 1


def _samples_dist_matrix_device(dev):
    pairs = dev.pair_counts()
    return _corr_from_pair_counts(pairs, dev.N)


def _wants_device_corr(dev):
    return dev.rows() * dev.p >= _DEVICE_CORR_MIN_ENTRIES and dev.p <= _lib.COUNT_PAIRS_MAX_P


def _sample_dist_matrix_auto(dev):
    if _wants_device_corr(dev):
        return _samples_dist_matrix_device(dev)
    return _samples_dist_matrix(_gathered_dense(dev))


def _tree_candidates(dist_matrix, max_size, filter_sequential, **kwargs):
    """Tree-node groups that can become candidates (:441-455): at most max_size members and, if
    asked, not contiguous (those are sequential groups already)."""
    kept = []
    for group in _dist_matrices_to_groups(dist_matrix, **kwargs):
        gsize = len(group)
        if gsize > max_size:
            continue
        if filter_sequential and np.max(group) - np.min(group) == gsize - 1:
            continue
        kept.append(group)
    return kept


def _groups_from_counts(kept, counts, N, max_pep):
    cand_groups = []
    for group, cnt in zip(kept, counts):
        pep = 1 - cnt / N                             # :459
        if pep < max_pep:
            cand_groups.append(CandidateGroup(group=set(group), pep=pep))
    return cand_groups


def hierarchical_groups(
    samples,
    dist_matrix=None,
    max_pep=0.25,
    max_size=25,
    filter_sequential=False,
    min_purity=0,
    X=None,
    corr_matrix=None,
    **kwargs
):
    """
    Creates candidate groups by hierarchical clustering (create_groups.py:386-469).
    The per-group count ``#rows with any signal in the group`` (:459) runs on the
    device for all groups at once.
    """
    dev = _as_device(samples)
    try:
        p = dev.p
        if p == 1:                                    # trivial case (:437-439)
            pep = 1 - dev.column_counts()[0] / dev.N
            return [CandidateGroup(group=set([0]), pep=pep)]
        if dist_matrix is None:
            dist_matrix = _sample_dist_matrix_auto(dev)
        kept = _tree_candidates(dist_matrix, max_size, filter_sequential, **kwargs)
        counts = dev.group_counts(kept)
        cand_groups = _groups_from_counts(kept, counts, dev.rows(), max_pep)
    finally:
        if dev is not samples:
            dev.close()
    return filter_by_purity(cand_groups, min_purity=min_purity, X=X, corr_matrix=corr_matrix)


def _pack_rows(dense):
    """bool (N, p) -> uint32 (N, ceil(p/32)) words, 32 columns per word."""
    packed = np.packbits(np.asarray(dense, dtype=bool), axis=1, bitorder="little")
    pad = (-packed.shape[1]) % 4
    if pad:
        packed = np.concatenate([packed, np.zeros((packed.shape[0], pad), dtype=np.uint8)], axis=1)
    return np.ascontiguousarray(packed).view(np.uint32)


def _unpack_rows(words, p):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), axis=1, bitorder="little")[:, :p].astype(bool)


def _gathered_dense(dev):
    """Dense boolean matrix of a (column-compacted) device sample set, rows of all ranks
    concatenated in rank order.  Only feeds the host clustering (small sets: below
    _DEVICE_CORR_MIN_ENTRIES entries); the rows travel bit-packed, one all_gather."""
    local = dev.to_dense()
    if dist.world()[1] == 1:
        return local
    dev.rows()
    words = dist.allgather_rows(_pack_rows(local), dev.N_per_rank, getattr(dev, "device_index", None))
    return _unpack_rows(words, dev.p)


# ---------------------------------------------------------------------------
# all candidate groups
# ---------------------------------------------------------------------------
def all_cand_groups(
    samples,
    X=None,
    q=0,
    max_pep=0.25,
    max_size=25,
    prenarrow=True,
    prefilter_thresholds=[0, 0.01, 0.02, 0.03, 0.05, 0.1, 0.2],
    min_purity=0.0,
    corr_matrix=None,
):
    """
    Creates many candidate groups by prefiltering locations at various thresholds
    and then creating sequential and hierarchical candidate groups
    (create_groups.py:56-149; same parameters, same result).

    The counting is organised in three stages, each merged over ranks by ONE all-reduce of a
    packed int64 buffer: (1) row count + marginal column counts; (2) for every threshold at
    once, the window zero-counts of the prefiltered columns and -- for large sets -- their pair
    counts; (3) for every threshold at once, the any-counts of the tree groups.
    """
    dev = _as_device(samples)
    subs = []
    try:
        marg_pips = dev.column_counts() / dev.N           # stage 1 (:105-106)
        levels = []
        for thresh in prefilter_thresholds:
            rel_features = np.where(marg_pips > thresh)[0]
            if len(rel_features) == 0:
                continue
            sub = dev if len(rel_features) == dev.p else dev.select(rel_features)
            if sub is not dev:
                subs.append(sub)
            levels.append(dict(rel=rel_features, sub=sub, msize=min(max_size, sub.p)))
        # stage 2: sequential counts (:117-123) and pair counts of every level
        batch = dist.CountBatch()
        for lv in levels:
            lv["seq"] = batch.add(lv["sub"], "sequential", lv["msize"])
            lv["pairs"] = batch.add(lv["sub"], "pairs") if _wants_device_corr(lv["sub"]) and lv["sub"].p > 1 else None
        res = batch.run()
        # levels whose tree is built from the dense matrix on the host (small sets): the rows of
        # all ranks are gathered ONCE, for the widest such level, and narrower levels slice it
        need_dense = [lv for lv in levels if lv["pairs"] is None and lv["sub"].p > 1]
        dense_cols, dense_all = None, None
        if need_dense:
            widest = max(need_dense, key=lambda lv: lv["sub"].p)
            dense_cols = {int(c): i for i, c in enumerate(widest["rel"])}
            dense_all = _gathered_dense(widest["sub"])
        for lv in levels:
            sub = lv["sub"]
            peps = _sequential_peps_from_counts(res[lv["seq"]], dev.N, sub.p, lv["msize"])
            lv["cgs"] = _sequential_from_peps(peps, lv["msize"], q, max_pep, prenarrow)
            if sub.p == 1:
                lv["kept"] = None
                continue
            if lv["pairs"] is not None:
                dms = [_corr_from_pair_counts(res[lv["pairs"]], dev.N)]
            else:
                cols = [dense_cols[int(c)] for c in lv["rel"]]          # thresholds nest: rel is a subset
                dms = [_samples_dist_matrix(dense_all[:, cols])]         # :125
            if X is not None:
                dms.append(np.abs(1 - np.corrcoef(X[:, lv["rel"]].T)))
            lv["kept"] = _tree_candidates(dms, max_size, True)
        # stage 3: any-counts of every level's tree groups (:459)
        batch = dist.CountBatch()
        for lv in levels:
            if lv["kept"] is None:
                lv["grp"] = batch.add(lv["sub"], "columns")
            else:
                lv["grp"] = batch.add(lv["sub"], "groups", lv["kept"]) if len(lv["kept"]) else None
        res = batch.run()
        all_groups = set()
        all_cgs = []
        for lv in levels:
            cgs = lv["cgs"]
            if lv["kept"] is None:                                       # p == 1 (:437-439)
                cgs.append(CandidateGroup(group=set([0]), pep=1 - res[lv["grp"]][0] / dev.N))
            elif lv["grp"] is not None:
                cgs.extend(_groups_from_counts(lv["kept"], res[lv["grp"]], dev.N, max_pep))
            # map the compacted indices back and drop duplicates across thresholds
            rel_features = lv["rel"]
            fresh = []
            for cg in cgs:
                group = tuple(sorted(rel_features[list(cg.group)].tolist()))
                if group not in all_groups:
                    cg.group = set(group)
                    all_cgs.append(cg)
                    fresh.append(group)
            all_groups = all_groups.union(fresh)
    finally:
        for sub in subs:
            sub.close()
        if dev is not samples:
            dev.close()
    return filter_by_purity(all_cgs, min_purity=min_purity, X=X, corr_matrix=corr_matrix)


def _prefilter(cand_groups, max_pep):
    """Returns the subset of cand_groups with a pep below max_pep."""
    return [x for x in cand_groups if x.pep < max_pep]


def _elim_redundant_features(cand_groups):
    """Re-indexes the locations that occur in at least one group as 0..nrel-1 and stores every
    group's re-indexed members in ``data['blip-group']`` (create_groups.py:580-613); the BLiP
    programs then have one row per location in use.  Returns (cand_groups, nrel)."""
    new_index = {}
    for cg in cand_groups:
        for j in cg.group:
            new_index.setdefault(j, len(new_index))
    for cg in cand_groups:
        cg.data['blip-group'] = set(new_index[j] for j in cg.group)
    return cand_groups, len(new_index)


def _ecc_reduction(cand_groups, verbose=False):
    """Replaces the locations by the cliques of an edge clique cover of the groups' intersection
    graph (create_groups.py:616-672): two groups share a clique iff they overlap, so the
    disjointness constraints are the same with (usually far) fewer rows.  Returns
    (cand_groups, number of cliques)."""
    from .create_groups_cts import _edge_clique_cover
    for cg in cand_groups:
        cg.data.setdefault('blip-group', cg.group)
    m = len(cand_groups)
    member = {}
    for i, cg in enumerate(cand_groups):
        for loc in cg.data['blip-group']:
            member.setdefault(loc, []).append(i)
    nbrs = [{i} for i in range(m)]
    for groups_here in member.values():
        for i in groups_here:
            nbrs[i].update(groups_here)
    cliques = _edge_clique_cover([sorted(x) for x in nbrs])
    for cg in cand_groups:
        cg.data['blip-group'] = set()
    for k, clique in enumerate(cliques):
        for i in clique:
            cand_groups[i].data['blip-group'].add(k)
    return cand_groups, len(cliques)


def selection_fwer(samples, groups):
    """Exact family-wise error rate of a selection: the fraction of posterior samples in which
    at least one selected group contains no signal -- the quantity the reference's FWER binary
    search recomputes from the full sample matrix in every iteration (blip.py:270-276).
    ``samples``: array, fitted sampler or device handle; ``groups``: iterable of index collections
    (or CandidateGroups)."""
    dev = _as_device(samples)
    try:
        glist = [sorted(getattr(g, "group", g)) for g in groups]
        return dev.false_rows(glist) / dev.N
    finally:
        if dev is not samples:
            dev.close()
