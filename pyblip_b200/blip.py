"""BLiP: resolution-adaptive selection of disjoint candidate groups (host code, HiGHS backend).

Same entry points and semantics as /root/reference/pyblip/blip.py -- ``BLiP`` (:34-313), ``BLiP_cts``
(:315-414) and ``binarize_selections`` (:417-632) -- but without cvxpy: the relaxation
(blip.py:217-252) is handed to ``scipy.optimize.linprog`` and the small binary program that settles
fractional selections (blip.py:556-607) to ``scipy.optimize.milp``, both HiGHS.  The problems are
written once as sparse constraint blocks (:class:`_SelectionProgram`) and reused by every solve.

What touches the posterior samples runs on the device:

* candidate groups come from ``create_groups.sequential_groups`` / ``hierarchical_groups``
  (counting kernels of libblipgpu); ``samples`` may be an array, a fitted sampler or a device
  handle, so the indicators never have to leave the GPU;
* the exact family-wise error of a rounded selection, which the reference recomputes from the
  full ``(N, p)`` matrix in every step of its binary search (blip.py:270-276), is one
  ``count_false_rows`` launch per step.

A vertex solver and an interior-point solver return the same selection whenever the optimum is
unique, which ``perturb=True`` (the default) makes almost sure; tests/test_blip.py checks the
reference's own properties (disjointness, error control, backtracking counts, adversarial
examples of /root/reference/test/test_blip.py:98-353).
"""
import copy
import time
import warnings

import numpy as np
import scipy.optimize as sopt
import scipy.sparse as sparse

from . import create_groups, create_groups_cts, weight_fns

WEIGHT_FNS = {
    'inverse_size': weight_fns.inverse_size,
    'log_inverse_size': weight_fns.log_inverse_size,
}
ERROR_OPTIONS = ['fdr', 'local_fdr', 'fwer', 'pfer']
BINARY_TOL = 1e-3
DEFAULT_GRID_SIZES = np.around(np.logspace(np.log10(4), 4, 25))
DEFAULT_SOLVER = 'HIGHS'


class _SelectionProgram:
    """max sum_g util_g x_g  subject to  sum_{g contains l} x_g <= 1 for every location l,
    0 <= x <= 1 (or x binary), plus optional rows ``a . x <= rhs`` added per solve.

    The location rows are the transpose of the group/location incidence matrix A of
    blip.py:206-209, kept in CSR."""

    def __init__(self, cand_groups, nrel, util):
        rows, cols = [], []
        for g, cg in enumerate(cand_groups):
            for loc in cg.data['blip-group']:
                rows.append(int(loc))
                cols.append(g)
        self.n = len(cand_groups)
        self.loc_rows = sparse.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(nrel, self.n))
        self.util = np.asarray(util, dtype=np.float64)

    def _stack(self, extra):
        if not extra:
            return self.loc_rows, np.ones(self.loc_rows.shape[0])
        a = sparse.vstack([self.loc_rows] + [sparse.csr_matrix(np.asarray(r, dtype=np.float64)[None, :])
                                             for r, _ in extra], format='csr')
        b = np.concatenate([np.ones(self.loc_rows.shape[0]), [float(rhs) for _, rhs in extra]])
        return a, b

    def relax(self, extra=(), verbose=False):
        """LP relaxation.  Returns (x, status) with status in {'optimal', 'infeasible', ...}."""
        a, b = self._stack(list(extra))
        res = sopt.linprog(-self.util, A_ub=a, b_ub=b, bounds=(0.0, 1.0), method='highs',
                           options=dict(disp=bool(verbose)))
        if res.status == 2:
            return None, 'infeasible'
        if res.status != 0 or res.x is None:
            return None, f'failed ({res.message})'
        return np.clip(res.x, 0.0, 1.0), 'optimal'

    def integer(self, extra=()):
        """The same program over x in {0,1}^n."""
        a, b = self._stack(list(extra))
        res = sopt.milp(-self.util, constraints=sopt.LinearConstraint(a, -np.inf, b),
                        integrality=np.ones(self.n), bounds=sopt.Bounds(0.0, 1.0))
        if res.status == 2:
            return None, 'infeasible'
        if res.x is None:
            return None, f'failed ({res.message})'
        return np.around(res.x), 'optimal'


def _error_rows(error, peps, q, level):
    """The error-rate row of the program (blip.py:240-248): PFER / FWER bound the expected number
    of false discoveries by ``level``; FDR asks sum_g x_g (pep_g - q) <= 0; local FDR has none
    (its groups were prefiltered to pep < q)."""
    if error in ('pfer', 'fwer'):
        return [(peps, level)]
    if error == 'fdr':
        return [(peps - q, 0.0)]
    return []


def BLiP(
    samples=None,
    cand_groups=None,
    weight_fn='inverse_size',
    error='fdr',
    q=0.1,
    max_pep=0.25,
    deterministic=True,
    verbose=True,
    solver_verbose=False,
    perturb=True,
    max_iters=100,
    search_method='binary',
    return_problem_status=False,
    **kwargs
):
    """
    Given samples from a posterior or a list of ``CandidateGroup`` objects, performs
    resolution-adaptive signal detection to maximize power while controlling (e.g.) the FDR.
    Arguments and return values are those of the reference (blip.py:34-136).  ``samples`` may
    also be a fitted ``LinearSpikeSlab`` / ``ProbitSpikeSlab`` / ``NPrior`` or a device handle.
    ``solver=`` is accepted for compatibility; the backend is always HiGHS.
    """
    error = str(error).lower()
    if error not in ERROR_OPTIONS:
        raise ValueError(f"error type {error} must be one of {ERROR_OPTIONS}")
    if cand_groups is None and samples is None:
        raise ValueError("At least one of cand_groups and samples must be provided.")
    if error in ['fwer', 'pfer', 'local_fdr']:
        max_pep = min(max_pep, q)            # optimal for all but error == 'fdr' (blip.py:143-144)
    kwargs.pop("solver", None)

    dev = None
    try:
        if samples is not None:
            dev = create_groups._as_device(samples)
        if cand_groups is None:              # blip.py:149-162
            cand_groups = create_groups.sequential_groups(samples=dev, q=q, max_pep=max_pep, prenarrow=True)
            cand_groups.extend(create_groups.hierarchical_groups(samples=dev, max_pep=max_pep,
                                                                 filter_sequential=True))
        cand_groups = create_groups._prefilter(cand_groups, max_pep=max_pep)
        if len(cand_groups) == 0:            # nothing below max_pep
            if return_problem_status:
                return [], dict(nlocs=0, ngroups=0, backtracking_iter=0, ngroups_nonint=0, lp_bound=0,
                                deterministic=deterministic)
            return []
        cand_groups, nrel = create_groups._elim_redundant_features(cand_groups)

        ngroups = len(cand_groups)
        if weight_fn == 'prespecified':
            weights = np.array([x.data['weight'] for x in cand_groups])
        else:
            if 'weight' in cand_groups[0].data:
                warnings.warn("Cand groups have a 'weight' attribute, do you mean to set weight_fn='prespecified'?")
            if isinstance(weight_fn, str):
                weight_fn = weight_fn.lower()
                if weight_fn not in WEIGHT_FNS:
                    raise ValueError(f"Unrecognized weight_fn={weight_fn}, must be one of {list(WEIGHT_FNS.keys())}")
                weight_fn = WEIGHT_FNS[weight_fn]
            weights = np.array([weight_fn(x) for x in cand_groups])
        orig_weights = weights.copy()
        if perturb:                          # one uniform per group, in order (blip.py:198-203)
            weights = np.array([w * (1 + 0.0001 * np.random.uniform()) for w in weights])
        peps = np.array([x.pep for x in cand_groups])
        if verbose:
            print(f"BLiP problem has {ngroups} groups in contention, with {nrel} active features/locations")

        util = (1 - peps) * weights
        prog = _SelectionProgram(cand_groups, nrel, util)
        binary_search = dev is not None and error == 'fwer' and search_method == 'binary'
        if not binary_search:
            selections, status = prog.relax(_error_rows(error, peps, q, q), verbose=solver_verbose)
            if selections is None:
                raise RuntimeError(f"BLiP relaxation could not be solved: {status}")
        else:
            # FWER: bisect the PFER-style level v for the largest one whose ROUNDED selection has
            # exact family-wise error <= q under the posterior samples (blip.py:256-295)
            v_upper, v_lower = q * nrel + 1, 0.0
            for _ in range(max_iters):
                v_current = (v_upper + v_lower) / 2
                x, status = prog.relax(_error_rows(error, peps, q, v_current), verbose=solver_verbose)
                chosen = np.where(np.around(x) > BINARY_TOL)[0] if x is not None else []
                fwer = dev.false_rows([sorted(cand_groups[g].group) for g in chosen]) / dev.N if len(chosen) else 0.0
                if fwer > q:
                    v_upper = v_current
                else:
                    v_lower = v_current
                if v_upper - v_lower < 1e-4:
                    break
            x, status = prog.relax(_error_rows(error, peps, q, v_lower), verbose=solver_verbose)
            selections = np.zeros(ngroups) if x is None else np.around(x)
    finally:
        if dev is not None and dev is not samples:
            dev.close()

    for cand_group, sprob, weight in zip(cand_groups, selections, orig_weights):
        cand_group.data['sprob'] = sprob
        cand_group.data['weight'] = weight
    problem_status = dict(ngroups=ngroups, nlocs=nrel, lp_bound=np.dot(selections, util),
                          backtracking_iter=0, deterministic=deterministic)
    return binarize_selections(cand_groups=cand_groups, q=q, error=error, deterministic=deterministic,
                               problem_status=problem_status, return_problem_status=return_problem_status,
                               verbose=verbose)


def BLiP_cts(
    locs,
    grid_sizes=DEFAULT_GRID_SIZES,
    weight_fn=weight_fns.inverse_radius_weight,
    extra_centers=None,
    max_pep=0.25,
    shape='circle',
    min_blip_size=1500,
    verbose=True,
    **kwargs
):
    """BLiP over a continuous set of locations (blip.py:315-414): ``grid_peps`` on the device,
    overlap components, one BLiP per component."""
    N, num_disc, d = locs.shape
    log_interval = int(max(1, np.around(N / 10))) if verbose else None
    peps = create_groups_cts.grid_peps(locs=locs, grid_sizes=grid_sizes, extra_centers=extra_centers,
                                       max_pep=max_pep, shape=shape, log_interval=log_interval)
    all_cand_groups, _ = create_groups_cts.grid_peps_to_cand_groups(peps, min_blip_size=min_blip_size,
                                                                    shape=shape, max_pep=max_pep)
    all_rej = []
    for cand_groups in all_cand_groups:
        all_rej.extend(BLiP(cand_groups=cand_groups, weight_fn=weight_fn, max_pep=max_pep,
                            verbose=verbose, **kwargs))
    for cand_group in all_rej:               # easier to interpret (blip.py:406-412)
        center = np.zeros(d)
        for k in range(d):
            center[k] = cand_group.data.pop(f'dim{k}')
        cand_group.data['center'] = center
        cand_group.data['radii'] = np.ones(d) * cand_group.data['radius']
    return all_rej


def _trim_to_level(groups, error, q):
    """Drops the highest-PEP discoveries until the estimated FDR (mean PEP) or PFER (sum) is at
    most q (blip.py:529-545)."""
    groups = sorted(groups, key=lambda x: x.pep)
    agg = np.mean if error == 'fdr' else np.sum
    while len(groups) > 0 and agg([x.pep for x in groups]) > q:
        groups = groups[:-1]
    return groups


def _sample_selection(cands, A, sprobs, order):
    """One draw of a disjoint selection whose marginals follow the LP solution (blip.py:489-520):
    features are visited in decreasing marginal selection probability; at each one either nothing
    or one of the still available groups containing it is chosen, and everything it excludes is
    eliminated."""
    ngroups = len(cands)
    eliminated = np.zeros(ngroups, dtype=bool)
    chosen = []
    for feature in order:
        if eliminated.all():
            break
        has = A[:, feature] == 1
        available = has & ~eliminated
        if not available.any():
            continue
        scale = 1 - sprobs[has & eliminated].sum()
        new_probs = sprobs[available] / scale
        if np.random.uniform() <= 1 - new_probs.sum():
            eliminated[has] = True
            continue
        pick = np.where(np.random.multinomial(1, new_probs / new_probs.sum()) != 0)[0][0]
        pick = np.where(available)[0][pick]
        chosen.append(pick)
        eliminated[np.sum(A[:, np.where(A[pick] == 1)[0]], axis=1) != 0] = True
    return chosen


def binarize_selections(
    cand_groups,
    q,
    error,
    deterministic,
    problem_status=None,
    return_problem_status=False,
    tol=1e-3,
    nsample=10,
    verbose=False,
):
    """
    Turns the (possibly fractional) LP selection ``data['sprob']`` into a list of detections
    (blip.py:417-632): integral entries are taken as they are; the fractional remainder is
    settled by a small binary program (``deterministic=True``; FDR may need backtracking) or by
    sampling ``nsample`` disjoint selections and keeping the most powerful one.
    """
    output = []
    if problem_status is None:
        problem_status = dict(backtracking_iter=0)
    if error in ['local_fdr', 'pfer', 'fwer']:
        cand_groups = create_groups._prefilter(cand_groups, max_pep=q)

    nontriv = []
    for cg in cand_groups:
        if cg.data['sprob'] < tol:
            continue
        elif cg.data['sprob'] > 1 - tol:
            output.append(cg)
        else:
            nontriv.append(cg)
    ngroups = len(nontriv)
    problem_status['ngroups_nonint'] = ngroups
    if ngroups <= 1:
        return (output, problem_status) if return_problem_status else output

    nontriv, nrel = create_groups._elim_redundant_features(nontriv)
    nontriv, nrel = create_groups._ecc_reduction(nontriv)
    if verbose:
        print(f"LP had {ngroups} non-integer solutions across {nrel} locations."
              f" Binarizing with deterministic={deterministic}.")
    peps = np.array([cg.pep for cg in nontriv])
    weights = np.array([cg.data['weight'] for cg in nontriv])

    if not deterministic:
        A = np.zeros((ngroups, nrel), dtype=bool)
        for gj, cg in enumerate(nontriv):
            A[gj, list(cg.data['blip-group'])] = True
        sprobs = np.array([cg.data['sprob'] for cg in nontriv])
        marg = np.array([sprobs[A[:, j]].sum() for j in range(nrel)])
        order = np.argsort(-1 * marg)
        best, best_power = None, -np.inf
        for _ in range(nsample):
            chosen = _sample_selection(nontriv, A, sprobs, order)
            out_i = copy.deepcopy(output)
            out_i.extend([nontriv[g] for g in chosen])
            if error != 'local_fdr':
                out_i = _trim_to_level(out_i, error, q)
            power = np.sum([(1 - x.pep) * x.data['weight'] for x in out_i])
            if power > best_power:           # first maximiser, like np.argmax
                best, best_power = out_i, power
        output = best
    else:
        prog = _SelectionProgram(nontriv, nrel, (1 - peps) * weights)
        ndisc_out = len(output)
        v_output = sum(cg.pep for cg in output)
        extra = [(peps, q - v_output)] if error in ['pfer', 'fwer'] else []
        if error in ['pfer', 'fwer', 'local_fdr']:
            x, status = prog.integer(extra)
        else:
            # FDR: the integral part already spent v_output of the budget ndisc_out * q.  If the
            # binary program is infeasible, give back the worst integral discovery and retry
            # (backtracking, blip.py:566-607; extremely rare)
            output = sorted(output, key=lambda cg: cg.pep)
            while True:
                fdr_rhs = ndisc_out * q - v_output
                x, status = prog.integer(extra + [(peps - q, fdr_rhs)])
                if status == 'optimal' and x @ (peps - q) > fdr_rhs + tol:
                    status = 'infeasible'
                if status != 'infeasible':
                    break
                if len(output) == 0:
                    raise RuntimeError("Backtracking for FDR control failed")
                problem_status['backtracking_iter'] += 1
                if verbose:
                    print(f"Starting backtracking iter={problem_status['backtracking_iter']}")
                v_output -= output[-1].pep
                ndisc_out -= 1
                output = output[:-1]
        if x is None:
            if status != 'infeasible':
                raise RuntimeError(f"BLiP binarization could not be solved: {status}")
            x = np.zeros(ngroups)
        for take, cg in zip(x, nontriv):
            if take > 1 - tol:
                output.append(cg)

    return (output, problem_status) if return_problem_status else output
