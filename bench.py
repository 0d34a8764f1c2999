#!/usr/bin/env python
"""Headline benchmark: Gibbs coordinate-updates/sec of the multi-chain
spike-and-slab sampler (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's Cython sampler on the host cores

Workload (BASELINE.json configs[1]): LinearSpikeSlab, n=1000,
p=10000, AR(1) rho=0.95 design, 50 signals, 512 chains IN TOTAL, sharded over the ranks
(strong scaling, north_star's "512 chains sharded over 1/2/4/8 GPUs"; at N > 1 the line also
carries `weak`: 512 chains on every rank), default hyper-priors.  One step = one full sampler pass of `--sweeps` recorded sweeps after
`--burn` burn-in sweeps for every chain -- by default the configuration's full job,
N=5000 after 100 burn-in sweeps.  Rank 0 prints ONE JSON line.

  value     device-resident inputs: X and X'X already in HBM; timed region = the C-ABI
            sampler call (state init, permutations, sweeps, recording into HBM),
            CUDA events on the library's stream, max over ranks.
  e2e       the reference-facing API with HOST buffers: LinearSpikeSlab(X, y).sample(...)
            -> H2D of X, Gram build, sampling, D2H of hyper-parameters + the sparse
            coefficient matrix (everything lm.betas is made of).
  roofline  gram-mode sweep kernel: algorithmic bytes (8p per changed coordinate for the
            Gram column + 24p per chain-sweep) / CUDA-event kernel time vs measured HBM, plus what
            actually binds it: DRAM and L2 fractions from the committed ncu capture against the L2 /
            HBM / fp64 peaks and the L2 latency probed in this process (blipgpu_probe_*).
  e2e_pips  e2e followed by the PIP-counting half of the path on the device-resident samples:
            sequential_groups + all_cand_groups (thresholds >= 0.01), counts merged over ranks by
            one packed all-reduce per stage.
  cpu_baseline  the reference's own compiled sampler (oracle/_ref) on all host cores,
            bounded sample (rank 0, N=1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the duration of the
# run (NCCL prints its banner / debug lines to stdout when NCCL_DEBUG is set in the environment) and
# the result line is written to the original stdout at the end
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)

CONFIG = dict(n=1000, p=10000, rho=0.95, k=50, chains=512, seed=1002)
METRIC = "gibbs_coordinate_updates_per_sec"
UNIT = "updates/s"


def make_data(cfg=CONFIG):
    """AR(1) design, k signals, y = X beta + N(0,1)   (SURVEY.md section 8d)."""
    rng = np.random.default_rng(cfg["seed"])
    n, p, rho = cfg["n"], cfg["p"], cfg["rho"]
    Z = rng.standard_normal((n, p))
    X = np.empty((n, p))
    X[:, 0] = Z[:, 0]
    s = np.sqrt(1 - rho * rho)
    for j in range(1, p):
        X[:, j] = rho * X[:, j - 1] + s * Z[:, j]
    beta = np.zeros(p)
    nn = rng.choice(p, size=cfg["k"], replace=False)
    beta[nn] = rng.standard_normal(cfg["k"])
    y = X @ beta + rng.standard_normal(n)
    return np.ascontiguousarray(X), y, beta


# ---------------------------------------------------------------------------
# clocks (recipe in /opt/skills/guides/B200_PROFILING.md)
# ---------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 200` next to the run.  The process is started BEFORE the warm-up steps: attaching a
    second client to the GPU was seen to stall the device for up to a second (3160 ms for the first timed step
    against 2140 ms for the others); `mark()` at the start of the timed region drops the samples taken before it."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t_mark = 0.0

    def mark(self):
        self.t_mark = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_line, ln in self.lines:
            if t_line < self.t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own Cython kernel on host cores
# ---------------------------------------------------------------------------
_REF_DATA = {}


_HP = dict(tau2=1.0, update_tau2=1, tau2_a0=2.0, tau2_b0=1.0, sigma2=1.0, update_sigma2=1, sigma2_a0=2.0,
           sigma2_b0=1.0, p0=0.9, update_p0=1, min_p0=0.0, p0_a0=1.0, p0_b0=1.0)


def _ref_worker(nsweeps):
    """One chain of the reference's compiled kernel for nsweeps sweeps.  kind "linear" / "probit":
    `_sample_spikeslab` (pyblip/linear/_linear.pyx:34); "block": `_sample_spikeslab_multi`
    (pyblip/linear/_linear_multi.pyx:271); "nprior": `_nprior_sample` (pyblip/nprior/_nprior.pyx:56)."""
    X, y, kind, kw = _REF_DATA["X"], _REF_DATA["y"], _REF_DATA.get("kind", "linear"), _REF_DATA.get("kw", {})
    if kind == "nprior" and "blas" not in _REF_DATA:
        # one chain per core: the nprior kernel's level-3 BLAS / LAPACK calls must stay single-threaded
        # (16 workers x 16 BLAS threads ran 34x slower); the other kernels only make level-1 calls
        try:
            import threadpoolctl
            _REF_DATA["blas"] = threadpoolctl.threadpool_limits(limits=1)
        except Exception:
            _REF_DATA["blas"] = None
    t0 = time.perf_counter()
    if kind == "linear":
        from pyblip.linear._linear import _sample_spikeslab
        _sample_spikeslab(N=nsweeps, X=X, y=y, z=np.zeros(1, dtype=np.int64), probit=0, **_HP)
    elif kind == "probit":
        from pyblip.linear._linear import _sample_spikeslab
        _sample_spikeslab(N=nsweeps, X=X, y=np.zeros(X.shape[0]), z=y.astype(np.int64), probit=1,
                          **dict(_HP, update_sigma2=0))
    elif kind == "block":
        from pyblip.linear._linear_multi import _sample_spikeslab_multi
        _sample_spikeslab_multi(N=nsweeps, bsize=kw.get("bsize", 3), X=X, y=y.copy(), z=np.zeros(1, dtype=np.int64),
                                probit=0, max_signals_per_block=0, **_HP)
    elif kind == "nprior":
        from pyblip.nprior._nprior import _nprior_sample
        import scipy.linalg
        if not getattr(scipy.linalg.cholesky, "_f_ordered", False):
            # scipy >= 1.18 returns a C-ordered factor; the reference (_nprior.pyx:352) takes `.T` of it into
            # a C-contiguous memoryview, which only works with the Fortran-ordered factor of older scipy
            chol = scipy.linalg.cholesky
            scipy.linalg.cholesky = lambda a, *args, **kw: np.asfortranarray(chol(a, *args, **kw))
            scipy.linalg.cholesky._f_ordered = True
        p = X.shape[1]
        _nprior_sample(N=nsweeps, X=X, y=y, tauw2=1.0, p0_init=1 - min(0.01, 1 / p), min_p0=1e-10, update_p0=1,
                       sigma_a0=5.0, sigma_b0=1.0, alpha0_a0=1.0, alpha0_b0=1.0, sigma_prior_type=0,
                       joint_sample_W=1, group_alpha_update=1, log_interval=nsweeps + 1, time0=0.0)
    else:
        raise ValueError(kind)
    return time.perf_counter() - t0


def reference_throughput(X, y, sweeps, workers=None, kind="linear", **kw):
    """Coordinate-updates/s of the reference kernel with one chain per host core.
    Setup cost (XT copy etc.) is removed by differencing a short and a long run."""
    import multiprocessing as mp
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import ref_loader
    ref_loader.load()
    workers = workers or len(os.sched_getaffinity(0))
    _REF_DATA.update(X=X, y=y, kind=kind, kw=kw)
    short = max(2, sweeps // 10)
    ctx = mp.get_context("fork")
    limit = 900                                           # seconds; a stuck worker must not hang the bench
    with ctx.Pool(workers) as pool:
        pool.map_async(_ref_worker, [1] * workers).get(timeout=limit)             # page in, import
        t0 = time.perf_counter()
        pool.map_async(_ref_worker, [short] * workers).get(timeout=limit)
        t_short = time.perf_counter() - t0
        t0 = time.perf_counter()
        pool.map_async(_ref_worker, [short + sweeps] * workers).get(timeout=limit)
        t_long = time.perf_counter() - t0
    dt = max(t_long - t_short, 1e-9)
    p = X.shape[1]
    return dict(value=workers * p * sweeps / dt, seconds=dt, cores=workers, kind="reference",
                sample=f"{workers} chains (one per host core) x {sweeps} sweeps of the "
                       f"n={X.shape[0]}, p={p} workload; setup removed by differencing")


# ---------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return rank, world, local_rank, dist


def reduce_max(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist, ctx=None):
    if ctx is not None:
        ctx.synchronize()
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, _ = make_data()
    vals = []
    for _ in range(args.warmup):
        pass   # the pool's own page-in pass is the warm-up of a CPU run
    res = None
    for _ in range(max(1, args.steps)):
        res = reference_throughput(X, y, args.ref_sweeps)
        vals.append(res["value"])
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, impl="reference"), chains_run=res["cores"],
                       note=f"bounded sample: {res['cores']} chains (one per host core) x {args.ref_sweeps} sweeps of the "
                            f"same n, p, design and hyper-priors; throughput per chain-sweep does not depend on the "
                            f"number of chains or sweeps (BASELINE.md section 3)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, impl="b200", chains=None, chains_per_rank=None):
    c = dict(CONFIG)
    if chains is not None:
        c["chains"] = chains
    return {"workload": "LinearSpikeSlab n=1000 p=10000 AR1 rho=0.95, 512 chains "
                        "(BASELINE.json configs[1])" + (f", {chains} chains in total" if chains not in (None, CONFIG["chains"]) else ""),
            "n": c["n"], "p": c["p"], "chains": c["chains"], "signals": c["k"],
            "chains_per_rank": chains_per_rank,
            "sweeps_per_step": (args.sweeps + args.burn) if impl == "b200" else args.ref_sweeps,
            "burn": args.burn if impl == "b200" else 0,
            "parallelism": f"chains sharded over {args.gpus} GPU(s); sampling has no collective, PIP counts are merged "
                           f"by one packed int64 all-reduce per counting stage (e2e_pips)",
            "mode": args.mode if impl == "b200" else "cython",
            "l2_policy": "no flush: X'X (0.8 GB) + chain state and scratch (0.2 GB) + the sample store exceed "
                         "the 126 MB L2; hot Gram columns legitimately stay L2-resident as in a real run"}


def hardware_probes(ctx):
    """Denominators measured in this process, under this run's clocks (libblipgpu probe kernels)."""
    out = {}
    try:
        out["l2_read_gbs"] = ctx.probe_read_bandwidth(48 << 20, 40)       # 48 MB stays in the 126 MB L2
        out["hbm_read_gbs"] = ctx.probe_read_bandwidth(4 << 30, 2)        # 4 GB does not
        out["fp64_dfma_tflops"] = ctx.probe_fp64(0)
        out["fp64_dmma_tflops"] = ctx.probe_fp64(1)
        out["l2_latency_cycles"] = ctx.probe_latency(32 << 20)
        out["how"] = ("blipgpu_probe_read_bandwidth: 16-byte ld.cg streaming, 8 loads in flight per thread, 8 CTAs "
                      "per SM, 48 MB x 40 passes (L2) / 4 GB x 2 passes (HBM), best of 3; blipgpu_probe_fp64: 8 "
                      "independent DFMA chains per thread / 8 independent mma.sync.m8n8k4.f64 per warp; "
                      "blipgpu_probe_latency: dependent ld.cg pointer chase over 32 MB")
    except Exception as err:       # a probe must not take the result down
        out["error"] = f"{type(err).__name__}: {err}"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=5000, help="recorded sweeps per step (N)")
    ap.add_argument("--burn", type=int, default=100)
    ap.add_argument("--mode", default="gram", choices=["gram", "residual"])
    ap.add_argument("--ref-sweeps", type=int, default=200, dest="ref_sweeps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pips", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): BASELINE's 512 chains in total, sharded over the ranks (64 per GPU at "
                         "N=8); at N > 1 the weak configuration (512 chains on every rank, global chain ids "
                         "rank*512 ..) is run as well and reported under `weak`.  weak: the headline is the weak run")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local_rank, dist = dist_setup(args.gpus)
    from pyblip_b200 import _lib, dist as bdist, create_groups
    from pyblip_b200.linear import LinearSpikeSlab

    X, y, _ = make_data()
    n, p = X.shape
    ctx = _lib.Context.get(local_rank)
    design = _lib.Design.upload(X, ctx)
    mode_id = {"gram": _lib.MODE_GRAM, "residual": _lib.MODE_RESIDUAL}[args.mode]
    if mode_id == _lib.MODE_GRAM:
        design.build_gram()
    cfg = LinearSpikeSlab(X, y)._config()
    S, B = args.sweeps, args.burn

    def timed_run(chains, c0, nloc, want_clocks):
        """args.warmup untimed steps, then args.steps timed ones of `nloc` local chains."""
        def step(seed):
            t0 = time.perf_counter()
            smp = _lib.sample_spikeslab(design, cfg, y=y, chains=nloc, chain_offset=c0, n_keep=S,
                                        burn=B, seed=seed, mode=mode_id)
            t1 = time.perf_counter()
            info = smp.info
            out = dict(total_ms=info.total_ms, sweep_ms=info.sweep_ms, launches=info.sweep_launches,
                       changes=info.n_changes, windows=info.n_windows, nnz=info.nnz)
            smp.close()
            if args.debug and rank == 0:
                print(f"[step] call {1e3 * (t1 - t0):.1f} ms  device {info.total_ms:.1f} ms  sweep kernels "
                      f"{info.sweep_ms:.1f} ms  nnz/row {info.nnz / max(info.rows, 1):.1f}", file=sys.stderr, flush=True)
            return out
        clocks = ClockSampler(local_rank)
        if want_clocks:
            clocks.start()                       # before the warm-up steps: see ClockSampler
        for w in range(args.warmup):
            step(1000 + w)
        barrier(dist, ctx)
        clocks.mark()
        ctx.counters(reset=True)
        ctx.event_record(0)
        stats = [step(2000 + k) for k in range(args.steps)]
        ctx.event_record(1)
        barrier(dist, ctx)
        ms_local = ctx.event_elapsed_ms(0, 1)
        counters = ctx.counters()
        ms_total = reduce_max(dist, ms_local)
        clock_info = clocks.stop() if want_clocks else None
        updates_total = float(chains) * p * (S + B) * args.steps
        return dict(value=updates_total / (ms_total * 1e-3), ms_total=ms_total, ms_local=ms_local, stats=stats,
                    counters=counters, clocks=clock_info, updates_total=updates_total)

    chains = CONFIG["chains"] * (world if args.scaling == "weak" else 1)
    c0, c1 = bdist.shard_chains(chains, rank, world)
    nloc = c1 - c0
    main_run = timed_run(chains, c0, nloc, rank == 0 and not args.no_clocks)
    value, ms_total, stats, counters = main_run["value"], main_run["ms_total"], main_run["stats"], main_run["counters"]
    updates_total = main_run["updates_total"]
    kernel_variant = "classic (one CTA per chain, work queue)" if nloc > 148 else \
        "role-specialised (decision warps + helper warps, one SM per chain)"

    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        wchains = CONFIG["chains"] * world
        w0, w1 = bdist.shard_chains(wchains, rank, world)
        wr = timed_run(wchains, w0, w1 - w0, False)
        weak = {"value": wr["value"], "unit": UNIT, "chains": wchains, "chains_per_rank": w1 - w0,
                "ms_per_step": wr["ms_total"] / args.steps}

    # ---- roofline of the dominant kernel (this rank; ranks are symmetric)
    probes = hardware_probes(ctx) if rank == 0 else {}
    sweep_ms = sum(s["sweep_ms"] for s in stats)
    n_launch = sum(s["launches"] for s in stats)
    chain_sweeps = float(nloc) * (S + B) * args.steps
    changes = float(sum(s["changes"] for s in stats))
    if mode_id == _lib.MODE_GRAM:
        alg_bytes = 8.0 * p * changes + 24.0 * p * chain_sweeps
        kernel = "gibbs_gram_kernel" if nloc > 148 else "gibbs_gram_ws_kernel"
    else:
        alg_bytes = 8.0 * n * p * chain_sweeps
        kernel = "gibbs_residual_kernel"
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
    avg_launch_ms = sweep_ms / max(n_launch, 1)
    # DRAM / L2 traffic of one launch from the committed ncu --set full capture (same launch shape:
    # all 512 chains of one GPU, stationary regime); null for any other shape
    traffic = l2_bytes = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(kernel)
        if tj and nloc == {"gibbs_gram_kernel": 512, "gibbs_gram_ws_kernel": 64}.get(kernel):
            # the capture is a 32-sweep launch of the same chains; a bench launch holds more sweeps (the library's
            # default is 128 per launch in gram form), so the captured bytes are scaled by chain-sweeps per launch
            scale = (chain_sweeps / max(n_launch, 1)) / float(tj.get("chain_sweeps_per_captured_launch", nloc * 32))
            traffic = float(tj["dram_bytes_per_launch"]) * scale
            l2_bytes = float(tj["l2_to_sm_bytes_per_launch"]) * scale if "l2_to_sm_bytes_per_launch" in tj else None
    if mode_id == _lib.MODE_GRAM:
        bound = "latency"
        bound_note = ("the formula of SURVEY 8(d) gives an HBM-EQUIVALENT figure (`achieved`, `frac`); the kernel is not "
                      "HBM-bound: its DRAM traffic is a fraction of the algorithmic bytes (hot Gram columns are L2 hits) "
                      "and L2 / issue slots are far from saturated -- it is bound by the per-chain dependency chain "
                      "(window -> barrier -> first change -> gathered Gram entry at L2 latency -> next window)")
    else:
        bound = "l2"
        bound_note = "X (80 MB) is pinned in L2 and re-read by every chain: the binding resource is L2 -> SM bandwidth"
    l2_peak = probes.get("l2_read_gbs")
    roofline = {"bound": bound, "bound_note": bound_note, "kernel": kernel, "kernel_variant": kernel_variant,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "launches": n_launch, "avg_launch_ms": avg_launch_ms,
                "alg_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "dram_frac": (traffic / (avg_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "l2_peak": l2_peak, "l2_unit": "GB/s",
                "l2_frac": (l2_bytes / (avg_launch_ms * 1e-3) / 1e9 / l2_peak) if (l2_bytes and l2_peak) else None,
                "l2_frac_of_algorithmic_bytes": (achieved / l2_peak) if l2_peak else None,
                "probes": probes,
                "changes_per_chain_sweep": changes / chain_sweeps,
                "windows_per_chain_sweep": float(sum(s["windows"] for s in stats)) / chain_sweeps,
                "kernel_share_of_step": sweep_ms / main_run["ms_local"]}

    # ---- end to end through the reference-facing API, host buffers
    e2e = e2e_pips = None
    if not args.no_e2e:
        def e2e_step(seed, with_pips):
            t_0 = time.perf_counter()
            lm = LinearSpikeSlab(X, y)
            lm.sample(N=S, burn=B, chains=chains, seed=seed, mode=args.mode)
            t_1 = time.perf_counter()
            indptr, indices, values = lm.samples_csr()
            out = dict(nnz=int(indptr[-1]), pip_s=0.0)
            if args.debug:
                print(f"[e2e step rank {rank}] construct + sample() {1e3 * (t_1 - t_0):.1f} ms  samples_csr() "
                      f"{1e3 * (time.perf_counter() - t_1):.1f} ms  nnz {out['nnz']}", file=sys.stderr, flush=True)
            if with_pips:
                t_p = time.perf_counter()
                # the other half of the path, on the device-resident samples of every rank
                seq = create_groups.sequential_groups(lm, q=0.1, max_pep=0.25, max_size=25, prenarrow=True)
                if args.debug:
                    print(f"[e2e step rank {rank}] sequential_groups {1e3 * (time.perf_counter() - t_p):.1f} ms", file=sys.stderr, flush=True)
                allg = create_groups.all_cand_groups(lm, q=0.1, max_pep=0.25, max_size=25,
                                                     prefilter_thresholds=[0.01, 0.02, 0.03, 0.05, 0.1, 0.2])
                out.update(n_seq=len(seq), n_all=len(allg), pip_s=time.perf_counter() - t_p)
                if args.debug:
                    print(f"[e2e step rank {rank}] PIP stage {1e3 * out['pip_s']:.1f} ms", file=sys.stderr, flush=True)
            t_r = time.perf_counter()
            lm._release()
            lm._design.close()
            if args.debug:
                print(f"[e2e step rank {rank}] release {1e3 * (time.perf_counter() - t_r):.1f} ms  whole step "
                      f"{1e3 * (time.perf_counter() - t_0):.1f} ms", file=sys.stderr, flush=True)
            return out

        def timed_e2e(with_pips, ev0, ev1):
            for w in range(2):                           # warm-up: page-locked result blocks, the device pool's steady state
                e2e_step(3000 + w, with_pips)
            barrier(dist, ctx)
            ctx.counters(reset=True)
            c_before = dict(bdist.STATS)
            ctx.event_record(ev0)
            outs = [e2e_step(4000 + k, with_pips) for k in range(args.steps)]
            ctx.event_record(ev1)
            barrier(dist, ctx)
            e_ms = reduce_max(dist, ctx.event_elapsed_ms(ev0, ev1))
            ec = ctx.counters()
            res = {"value": updates_total / (e_ms * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": reduce_sum(dist, ec["h2d_bytes"]) / args.steps,
                   "d2h_bytes_per_step": reduce_sum(dist, ec["d2h_bytes"]) / args.steps,
                   "ms_per_step": e_ms / args.steps}
            return res, outs, {k: (bdist.STATS[k] - c_before[k]) / args.steps for k in c_before}
        e2e, _, _ = timed_e2e(False, 2, 3)
        e2e["api"] = "LinearSpikeSlab(X, y).sample(N, burn, chains) + samples_csr()"
        if not args.no_pips:
            e2e_pips, outs, coll = timed_e2e(True, 4, 5)
            e2e_pips.update(api="... + create_groups.sequential_groups(lm, prenarrow=True) + create_groups.all_cand_groups(lm, "
                                "prefilter_thresholds >= 0.01) on the device-resident samples",
                            pip_stage_ms_per_step=reduce_max(dist, 1e3 * float(np.mean([o["pip_s"] for o in outs]))),
                            sequential_groups=outs[-1]["n_seq"], all_cand_groups=outs[-1]["n_all"],
                            collectives_per_step=coll,
                            sample_rows=int(chains) * S)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_throughput(X, y, args.ref_sweeps)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"]}
        except Exception as err:   # the CPU baseline must not take the GPU number down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": f"failed: {type(err).__name__}: {err}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, chains=chains, chains_per_rank=nloc), "clocks": main_run["clocks"],
            "e2e": e2e, "e2e_pips": e2e_pips, "weak": weak,
            "gpu_launches": int(counters["kernel_launches"]),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    design.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
