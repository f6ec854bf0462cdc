#!/usr/bin/env python
"""bench.py -- batched iLQR iterations/sec on synthetic multi-agent Quadcopter12D scenarios.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--agents 3|5|10|15] [--mode potential|dp] [--scaling weak|strong] [--scenarios B]

Headline (defaults): 4096 x 10-agent Quadcopter12D, Potential-iLQR -- one "step" = one complete batched
ilqrSolver.solve (reference control.py:150-225) of all the scenarios; the metric counts iLQR iterations (one backward
Riccati pass + its line search for one problem) per second.  The other cells of BASELINE.json's sweep
(reference scripts/analysis.py:126-174: agent counts x centralized / distributed) run with --agents / --mode dp,
where a step is one DP-iLQR round (solve_distributed, reference distributed.py:25-103) of every scenario: interaction
graphs, all sub-problems binned by neighbourhood size, stitch, joint cost.

  value         device-timed, inputs resident in HBM when the timed region starts
  e2e           same metric through the C ABI with HOST buffers (dpilqr_solve_batch_host for Potential-iLQR;
                solve_distributed_round with host tensors for DP-iLQR): host->device copies of the inputs and
                device->host copies of the results inside the timed region
  roofline      dominant kernel (backward Riccati, FP64-compute bound) against the FP64 peak measured on this box by
                tools/bin/fp64_peak; second kernel (rollout / line search) alongside
  cpu_baseline  the CPU oracle (oracle/: a restatement of the reference's loop on the reference's own compiled
                dynamics when oracle/_ref is present) on all host cores, on a bounded sample of the same scenarios
  parity_sample the GPU solve of the very scenarios the CPU leg solved, compared per scenario: iteration count,
                accepted step-size trace, final cost and trajectory

Multi-GPU (torchrun, one rank per GPU): scenarios are independent, no data-path collective; time is the max over
ranks.  --scaling weak (default): every rank solves --scenarios of its own; --scaling strong: --scenarios in total,
dealt round-robin over the ranks.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "iterations/s"
S, C, T = 12, 4, 50


def metric_name(a, mode, B=4096):
    if a == 10 and mode == "potential":
        return f"batched iLQR iterations/sec ({B}x10-agent Quad12D)"
    return f"batched iLQR iterations/sec ({B}x{a}-agent Quad12D, {'DP-iLQR round' if mode == 'dp' else 'Potential-iLQR'})"


def backward_flops(a, s=S, c=C, T=T):
    """Structure-aware FLOPs of one backward pass."""
    n, m = a * s, a * c
    step = (4 * n * n * s + 4 * m * n * s + 2 * m * m * s + (2.0 / 3.0) * m ** 3 + 2 * m * m * (n + 1) + 2 * n * m * m
            + 4 * n * n * m + 8 * n * m + 2 * n * (s + c))
    return step * T


def backward_hbm_bytes(a, s=S, c=C, T=T):
    """Algorithmic HBM bytes of one backward pass: stage records in, K and d out."""
    n, m = a * s, a * c
    pairs = a * (a - 1) // 2
    stage = a * s * s + a * s * c + n + m + 9 * a + 9 * pairs
    return 8 * ((T + 1) * stage + T * m * n + T * m)


def rollout_hbm_bytes(a, n_alpha=1, s=S, c=C, T=T):
    """Algorithmic HBM bytes of one line-search launch per problem: K, d, X, U in; n_alpha candidates out."""
    n, m = a * s, a * c
    return 8 * (T * m * n + T * m + (T + 1) * n + T * m + n_alpha * ((T + 1) * n + T * m))


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------
def _cpu_worker(job):
    import numpy as np

    from dpilqr_b200 import scenarios
    from oracle import ilqr_oracle as O

    k, a, mode = job
    x0, xf, U0 = scenarios.quad12_inputs(k, a, T)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a,
                           [100 + i for i in range(a)])
    if mode == "potential":
        solver = O.OracleSolver(prob, T)
        X, U, J = solver.solve(x0, U0)
        return dict(k=k, iters=solver.n_backward, alpha=[r["alpha_index"] for r in solver.trace], J=float(J), X=X)
    Xh, _ = O.OracleSolver(prob, T).rollout(x0, U0)
    count = [0]
    X, U, J, info = O.solve_distributed(prob, Xh, U0, 0.5, [], count=count)
    return dict(k=k, iters=count[0], alpha=[len(v[1]) for v in info.values()], J=float(J), X=X)


def _cpu_sensitivity(job):
    """How far the CPU oracle's OWN result moves (a) when x0 is perturbed by 1e-15 relative (three sign patterns) and
    (b) when its backward pass evaluates the same formulas in another, equally valid floating-point order
    (OracleSolver.arith): the yardstick for scenarios on which iLQR amplifies rounding (the golden fixtures record (a)
    for the unmodified reference)."""
    import numpy as np

    from dpilqr_b200 import scenarios
    from oracle import ilqr_oracle as O

    k, a, mode = job
    x0, xf, U0 = scenarios.quad12_inputs(k, a, T)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a,
                           [100 + i for i in range(a)])

    def run(xp, arith=0):
        if mode == "potential":
            solver = O.OracleSolver(prob, T)
            solver.arith = arith
            X, U, J = solver.solve(xp, U0.copy())
            return [r["alpha_index"] for r in solver.trace], X, float(J)
        Xh, _ = O.OracleSolver(prob, T).rollout(xp, U0)
        count = [0]
        X, U, J, info = O.solve_distributed(prob, Xh, U0, 0.5, [], count=count)
        return [count[0]], X, float(J)

    tr0, X0, J0 = run(x0)
    moved, moved_J, trace_changes = 0.0, 0.0, 0
    trials = [(x0 * (1 + 1e-15 * np.sign(np.random.default_rng(trial).normal(size=x0.shape))), 0) for trial in range(1, 4)]
    if mode == "potential":
        trials += [(x0, 1), (x0, 2)]
    for xp, arith in trials:
        tr, X, J = run(xp, arith)
        trace_changes += tr != tr0
        if X.shape == X0.shape and tr == tr0:
            moved = max(moved, float(np.max(np.abs(X - X0)) / max(np.max(np.abs(X0)), 1e-300)))
            if np.isfinite(J) and np.isfinite(J0) and J0 != 0:
                moved_J = max(moved_J, abs(J - J0) / abs(J0))
    return dict(k=k, trace_changes=int(trace_changes), moved=moved, moved_J=moved_J)


def cpu_reference_run(n_scen, a, mode, first=0, keep=False, pool=None):
    """iterations/s of the CPU oracle over `n_scen` scenarios on all host cores (scenario-level pool, one BLAS thread
    per worker -- the working equivalent of the reference's multiprocessing path)."""
    import multiprocessing as mp

    from oracle import ilqr_oracle as O

    O.build_c_oracle()
    native = O.dynamics_backend("auto").name == "reference-native"
    cores = os.cpu_count() or 1
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    jobs = [(k, a, mode) for k in range(first, first + n_scen)]
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        if own:
            pool.map(_cpu_worker, jobs[:min(cores, n_scen)])  # warm the workers (imports, dlopen)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    finally:
        if own:
            pool.terminate()
    iters = sum(r["iters"] for r in res)
    what = "Potential-iLQR (ilqrSolver.solve)" if mode == "potential" else "one DP-iLQR round (solve_distributed)"
    out = dict(value=iters / dt, unit=UNIT, cores=cores, kind="port", iterations=int(iters), seconds=dt,
               sample=f"{n_scen} of the scenarios (seeds {first}..{first + n_scen - 1}), {a} agents, {what}, "
                      f"multiprocessing.Pool({cores}) over scenarios, 1 BLAS thread/worker; NumPy restatement of the "
                      f"reference loop (oracle/ilqr_oracle.py) calling the "
                      f"{'reference-compiled bbdynamics module (oracle/_ref)' if native else 'plain-C restatement of the dynamics'}")
    if keep:
        out["results"] = res
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_scen = args.cpu_scenarios or max(8 * cores, 64)
    import multiprocessing as mp

    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    vals = []
    with mp.get_context("spawn").Pool(cores) as pool:  # one pool of workers for all steps
        for step in range(args.warmup + args.steps):
            res = cpu_reference_run(n_scen, args.agents, args.mode, pool=pool)
            if step >= args.warmup:
                vals.append(res)
    value = sum(r["iterations"] for r in vals) / sum(r["seconds"] for r in vals)
    base = dict(vals[-1], value=value)
    base.pop("iterations"), base.pop("seconds")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.agents, args.mode), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(r["seconds"] for r in vals) / len(vals),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n_scen}-scenario sample of 4096 x {args.agents}-agent Quadcopter12D "
                               f"{'Potential-iLQR' if args.mode == 'potential' else 'DP-iLQR round'}, N=50, dt=0.1 (CPU)"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum of one backward launch over 4096 ten-agent problems
# (ncu --set full, profiles/r02_backward_ncu_full.txt): extrapolated per problem, not re-measured in the run
NCU_BACKWARD_DRAM_BYTES_PER_PROBLEM_A10 = (4.379636e9 + 7.879973e9) / 4096


def measure_fp64_peak():
    exe = os.path.join(ROOT, "tools", "bin", "fp64_peak")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        vals = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        return max(v["tflops"] for v in vals if str(v.get("kernel", "")).startswith(("dfma", "dmma")))
    except Exception:
        return None


def parity_sample(results, out, a, mode):
    """Per-scenario comparison of the GPU solve with the CPU leg on the same seeds (iteration
    counts per problem must equal the oracle's)."""
    import numpy as np

    rows, n_iter_bad, n_alpha_bad, errs_J, errs_X, suspects = [], 0, 0, [], [], {}
    for j, r in enumerate(results):
        if mode == "potential":
            it = int(out["iters"][j])
            alpha = [int(v) for v in out["trace_alpha"][j, :it]]
            J, X = float(out["J"][j]), out["X"][j]
        else:
            it = int(out["sub_iters"][j].sum())
            adj = out["adjacency"][j]
            alpha = [bin(int(v) & ((1 << a) - 1)).count("1") for v in adj]  # neighbourhood sizes
            J, X = float(out["J_full"][j]), out["X_dec"][j]
        same_it, same_al = (it == r["iters"]), (alpha == list(r["alpha"]))
        n_iter_bad += not same_it
        n_alpha_bad += not same_al
        if same_it and same_al:
            eJ = abs(J - r["J"]) / abs(r["J"]) if np.isfinite(r["J"]) and r["J"] != 0 else (0.0 if (np.isnan(J) and np.isnan(r["J"])) else float("inf"))
            eX = float(np.max(np.abs(X - r["X"])) / max(np.max(np.abs(r["X"])), 1e-300))
            errs_J.append(eJ), errs_X.append(eX)
            if not (eJ <= 1e-9 and eX <= 1e-9):
                suspects[r["k"]] = {"seed": r["k"], "rel_err_X": eX, "rel_err_J": eJ}
        else:
            suspects[r["k"]] = {"seed": r["k"], "iters": [it, r["iters"]], "trace": [alpha, list(r["alpha"])]}
    errs_J, errs_X = np.array(errs_J), np.array(errs_X)
    # every scenario that is not bit-for-bit in its decisions and within 1e-9 is checked against the oracle's OWN
    # sensitivity: explained if a 1e-15 perturbation of x0 or an equally valid evaluation order of the backward pass
    # changes the oracle's own decisions / moves its own result by at least a tenth of the discrepancy
    unexplained = 0
    if suspects:
        import multiprocessing as mp

        with mp.get_context("spawn").Pool(min(len(suspects), os.cpu_count() or 1)) as pool:
            for sres in pool.map(_cpu_sensitivity, [(k, a, mode) for k in suspects]):
                row = suspects[sres["k"]]
                row["oracle_trace_changes_when_perturbed"] = sres["trace_changes"]
                row["oracle_X_moves_when_perturbed"] = sres["moved"]
                row["oracle_J_moves_when_perturbed"] = sres["moved_J"]
                if "trace" in row:
                    row["explained"] = bool(sres["trace_changes"] > 0 or sres["moved"] > 1e-9)
                else:
                    row["explained"] = bool(row["rel_err_X"] <= max(1e-9, 10.0 * sres["moved"])
                                            and row["rel_err_J"] <= max(1e-9, 10.0 * sres["moved_J"]))
                unexplained += not row["explained"]
                rows.append(row)
    return {
        "scenarios": len(results), "what": "GPU vs CPU oracle per scenario: iteration count, "
        + ("accepted step-size index of every iteration" if mode == "potential" else "neighbourhood sizes of the interaction graph")
        + ", final J and X",
        "iteration_count_mismatches": int(n_iter_bad), "trace_mismatches": int(n_alpha_bad),
        "compared_numerically": int(errs_J.size),
        "max_rel_err_J": float(errs_J.max()) if errs_J.size else None, "max_rel_err_X": float(errs_X.max()) if errs_X.size else None,
        "median_rel_err_X": float(np.median(errs_X)) if errs_X.size else None,
        "within_1e-9": int(np.sum((errs_J <= 1e-9) & (errs_X <= 1e-9))),
        "bar": "1e-9 relative where the reference itself is well conditioned; scenarios above it are ill-conditioned "
               "solves that amplify rounding (the golden fixtures record the reference's own sensitivity)",
        "outside_the_bar": sorted(rows, key=lambda r: r["seed"])[:12], "unexplained": int(unexplained),
        "ok": bool(unexplained == 0),
        "ok_means": "every scenario either matches the CPU oracle in iteration count, step-size trace and to 1e-9 in J and X, or is "
                    "one on which the oracle's own decisions change, or its own X / J move by at least a tenth of the discrepancy, when x0 is "
                    "perturbed by 1e-15 or its backward pass is evaluated in another equally valid floating-point order",
    }


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import dpilqr_b200 as dp
    from dpilqr_b200 import _native, scenarios

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    a, mode = args.agents, args.mode
    n, m = a * S, a * C
    if args.scaling == "strong":  # --scenarios in total, dealt round-robin (balances the iteration counts)
        seeds = list(range(rank, args.scenarios, world))
    else:
        seeds = list(range(rank * args.scenarios, (rank + 1) * args.scenarios))
    B = len(seeds)
    # ---- problem construction (excluded from the timed regions, timed on its own)
    t_build = time.perf_counter()
    built = [scenarios.quad12_inputs(k, a, T) for k in seeds]
    specs = [scenarios.quad12_spec(xf, a) for _, xf, _ in built]
    x0_np, U0_np = np.stack([b[0] for b in built]), np.stack([b[2] for b in built])
    t_inputs = time.perf_counter() - t_build
    t_build = time.perf_counter()
    batch = dp.CompiledBatch(specs, T, dev)
    torch.cuda.synchronize(dev)
    t_compile = time.perf_counter() - t_build
    x0_dev, U0_dev = torch.as_tensor(x0_np).to(dev), torch.as_tensor(U0_np).to(dev)
    t_device_build = None
    if args.scaling == "weak":  # the same batch built on the device (scenario kernel + tensor ops): bit-identical inputs
        scenarios.quad12_batch_device(seeds[0], 8, a, T, dev)
        torch.cuda.synchronize(dev)
        t_build = time.perf_counter()
        _, x0_chk, _ = scenarios.quad12_batch_device(seeds[0], B, a, T, dev)
        torch.cuda.synchronize(dev)
        t_device_build = time.perf_counter() - t_build
        assert torch.equal(x0_chk, x0_dev), "device-built scenarios differ from the host-built ones"
    x0_pin, U0_pin = torch.as_tensor(x0_np).pin_memory(), torch.as_tensor(U0_np).pin_memory()
    fp64_peak = measure_fp64_peak() if rank == 0 else None
    lib = _native.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if mode == "potential":
        def step_resident(profile=False):
            return batch.solve(x0_dev, U0_dev, n_lqr_iter=50, tol=1e-3, profile=profile)["total_iters"]

        # end to end: the C ABI entry point a non-Python host binds, with host buffers (pinned) for everything
        out_pin = dict(X=torch.empty((B, T + 1, n), dtype=torch.float64).pin_memory(), U=torch.empty((B, T, m), dtype=torch.float64).pin_memory(),
                       J=torch.empty(B, dtype=torch.float64).pin_memory(), Js=torch.empty(B, dtype=torch.float64).pin_memory(),
                       iters=torch.empty(B, dtype=torch.int32).pin_memory(), status=torch.empty(B, dtype=torch.int32).pin_memory())
        host = dict(model=batch.t_model.cpu(), ndims=batch.t_ndims.cpu(), cidx=batch.t_cidx.cpu(), Q=batch.t_Q.cpu(), R=batch.t_R.cpu(),
                    Qf=batch.t_Qf.cpu(), xf=batch.t_xf.cpu(), radius=batch.t_radius.cpu(), weights=batch.t_weights.cpu(), hasprox=batch.t_hasprox.cpu())
        hb = _native.BatchStruct(B, a, S, C, T, int(batch.t_Q.shape[0]), batch.dt, host["model"].data_ptr(), host["ndims"].data_ptr(),
                                 host["cidx"].data_ptr(), host["Q"].data_ptr(), host["R"].data_ptr(), host["Qf"].data_ptr(),
                                 host["xf"].data_ptr(), host["radius"].data_ptr(), host["weights"].data_ptr(), host["hasprox"].data_ptr(),
                                 batch.model_hint, 0)
        opts = _native.SolveOpts(50, 10, 1e-3, 0.0, 0, 0, int(batch.costs_nonnegative), 0)

        def step_e2e():
            return _native.check(lib.dpilqr_solve_batch_host(
                ctypes.byref(hb), ctypes.byref(opts), x0_pin.data_ptr(), U0_pin.data_ptr(), out_pin["X"].data_ptr(), out_pin["U"].data_ptr(),
                out_pin["J"].data_ptr(), out_pin["Js"].data_ptr(), out_pin["iters"].data_ptr(), out_pin["status"].data_ptr(),
                None, None, None, local))

        h2d = int(x0_pin.numel() * 8 + U0_pin.numel() * 8 + sum(t.numel() * t.element_size() for t in host.values()))
        d2h = int(sum(t.numel() * t.element_size() for t in out_pin.values()))
        e2e_path = "dpilqr_solve_batch_host (C ABI, host buffers, descriptor arrays included)"
    else:
        Xh_dev, _ = batch.rollout(x0_dev, U0_dev)  # the trajectory the interaction graph is built on (hover rollout)
        Xh_dev = Xh_dev.contiguous()
        Xh_pin = Xh_dev.cpu().pin_memory()
        out_pin = dict(X=torch.empty((B, T + 1, n), dtype=torch.float64).pin_memory(), U=torch.empty((B, T, m), dtype=torch.float64).pin_memory(),
                       J=torch.empty(B, dtype=torch.float64).pin_memory())
        last = {}

        def step_resident(profile=False):
            out = dp.solve_distributed_round(batch, Xh_dev, U0_dev, 0.5, profile=profile)
            last.update(bins=out["bins"])
            return out["total_iters"]

        def step_e2e():
            out = dp.solve_distributed_round(batch, Xh_pin.to(dev, non_blocking=True), U0_pin.to(dev, non_blocking=True), 0.5)
            out_pin["X"].copy_(out["X_dec"], non_blocking=True)
            out_pin["U"].copy_(out["U_dec"], non_blocking=True)
            out_pin["J"].copy_(out["J_full"], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out["total_iters"]

        h2d = int(Xh_pin.numel() * 8 + U0_pin.numel() * 8)
        d2h = int(sum(t.numel() * t.element_size() for t in out_pin.values()))
        e2e_path = "solve_distributed_round with pinned host tensors in and out"

    # Steps are issued to a SolvePipeline: `--inflight` solves in flight (worker threads, one stream each): their
    # launches interleave and the straggler tail of a step runs, on a high-priority stream, beside the next steps
    # (csrc/solver.cu).  The timed region spans all K steps, from before the first is issued until the last has finished.
    ws_bytes = int(lib.dpilqr_workspace_bytes(B, a, S, C, T, 10))
    if args.inflight <= 0:  # auto: three steps in flight, four for small (latency-bound) batches, within 100 GB of workspaces
        args.inflight = max(1, min(3 if B >= 1024 else 4, int(100e9 // max(ws_bytes, 1))))
    pipe = dp.SolvePipeline(dev, depth=args.inflight)

    def run_steps(fn, count):
        return sum(pipe.map(lambda _: fn(), range(count)))

    run_steps(step_resident, args.warmup)
    _native.get_profile(reset=True)
    # ---- timed: device-resident
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = run_steps(step_resident, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    # ---- per-kernel times (roofline): the same steps once more, one at a time, every launch bracketed by CUDA events
    # on its stream (with several steps in flight the launches of different steps overlap and their durations do not
    # add up to the step)
    n_prof = min(args.steps, 2)
    _native.get_profile(reset=True)
    for _ in range(n_prof):
        step_resident(profile=True)
    torch.cuda.synchronize(dev)
    prof = _native.get_profile(reset=True)
    torch.cuda.empty_cache()  # the resident arm's cached workspaces make room for the host-buffer arm's arenas
    # ---- timed: end to end with host buffers
    run_steps(step_e2e, args.inflight)  # (every worker thread's arena and streams exist before the timed region)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    iters_e2e = run_steps(step_e2e, args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    pipe.close()
    # ---- reduce over ranks: total work, max time
    stats = torch.tensor([ms, ms_e2e, float(iters), float(iters_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = stats.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, iters, iters_e2e = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3])
    if rank == 0:
        value = iters / (ms * 1e-3)
        e2e_value = iters_e2e / (ms_e2e * 1e-3)
        bms, blaunch, bunits = prof["backward"]
        fms, flaunch, funits = prof["backward_full"]
        lms, llaunch, lunits = prof["linesearch"]
        if mode == "potential":
            bflops = backward_flops(a) * bunits
        else:  # sub-problems of every neighbourhood size: iterations of size k ~ share of the bins (all bins run ~ the same count)
            bins = last.get("bins", {})
            tot = max(sum(bins.values()), 1)
            bflops = sum(backward_flops(k) * bunits * cnt / tot for k, cnt in bins.items())
        achieved = bflops / (bms * 1e-3) * 1e-12 if bms > 0 else None
        peak, peak_src = (fp64_peak, "FP64 peak measured on this box by tools/bin/fp64_peak, max of the DFMA and DMMA m8n8k4 loops "
                                     "(MEASURED_PEAKS.json has no FP64 entry)") \
            if fp64_peak else (37.2, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz (fp64_peak binary missing)")
        hbm_peak = 6453.7
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        total_ms = sum(v[0] for k, v in prof.items() if k != "backward_full")
        kname = f"backward_kernel<12,4,{a}>" if mode == "potential" else "backward_kernel<12,4,k> over the neighbourhood sizes k"
        per_gpu = B
        line = {
            "metric": metric_name(a, mode), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{per_gpu} scenarios/GPU x {a}-agent Quadcopter12D "
                                   + ("Potential-iLQR (ilqrSolver.solve)" if mode == "potential" else "DP-iLQR round (solve_distributed on the hover rollout, radius 0.5)")
                                   + f", N=50, dt=0.1, n_lqr_iter=50, tol=1e-3, hover warm start, random_setup energy={3 * a} (SURVEY 8d)",
                       "iterations_per_step": iters / args.steps / world,
                       "l2": "working set per step (gains, candidate trajectories, stage records: GBs) >> 126 MB L2",
                       "parallelism": f"scenario-sharded x{world} ({args.scaling} scaling), no data-path collective",
                       "inflight": f"{args.inflight} step(s) in flight per GPU (SolvePipeline: the launches of the steps interleave, the straggler "
                                   "tail of a step runs beside the next steps; the timed region spans all steps)",
                       "construction_s": {"scenario_inputs_numpy": t_inputs, "CompiledBatch": t_compile,
                                          "on_the_device_quad12_batch_device": t_device_build,
                                          "note": "outside the timed regions (SURVEY 8d), once per batch"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "path": e2e_path},
            "gpu_launches": int(sum(v[1] for k, v in prof.items() if k != "backward_full") / n_prof * args.steps),
            "clocks": clocks,
            "roofline": {"kernel": kname, "bound": "tensor", "bound_detail": "FP64 pipe: mma.sync.m8n8k4.f64 tiles + DFMA",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "frac_of_the_launches_that_fill_the_machine": (backward_flops(a) * funits / (fms * 1e-3) * 1e-12 / peak)
                         if (mode == "potential" and fms > 0) else None,
                         "launches_that_fill_the_machine": flaunch,
                         "traffic": (NCU_BACKWARD_DRAM_BYTES_PER_PROBLEM_A10 * bunits / max(blaunch, 1)) if (a == 10 and mode == "potential") else None,
                         "traffic_source": "EXTRAPOLATED from one ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum = 12.260 GB "
                                           "for a 4096-problem launch, profiles/r02_backward_ncu_full.txt) to this run's average problems per "
                                           f"launch, not re-measured here; algorithmic bytes/problem = {backward_hbm_bytes(a)}",
                         "peak_source": peak_src,
                         "flops_per_launch_unit": backward_flops(a) if mode == "potential" else None, "launches": blaunch,
                         "avg_launch_ms": bms / max(blaunch, 1), "share_of_step": bms / total_ms if total_ms else None},
            "roofline_linesearch": {"kernel": "rollout_kernel<Quadcopter12D> (staged line search)", "bound": "hbm",
                                    "achieved": rollout_hbm_bytes(a) * lunits / (lms * 1e-3) * 1e-9 if lms > 0 else None,
                                    "peak": hbm_peak, "unit": "GB/s",
                                    "frac": (rollout_hbm_bytes(a) * lunits / (lms * 1e-3) * 1e-9 / hbm_peak) if lms > 0 else None,
                                    "note": "algorithmic bytes = gains read once per problem-iteration + trajectories; the kernel is bound by "
                                            "the latency of the serial RK4 chain, not by either roofline (DESIGN.md)",
                                    "share_of_step": lms / total_ms if total_ms else None},
            "kernel_ms_per_step": dict({k: v[0] / n_prof for k, v in prof.items() if k != "backward_full"},
                                       note=f"{n_prof} step(s) run one at a time after the timed region, every launch between CUDA events"),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = min(args.cpu_scenarios or max(8 * cores, 64), B)
            base = cpu_reference_run(n_cpu, a, mode, first=seeds[0], keep=True)
            results = base.pop("results")
            base.pop("iterations"), base.pop("seconds")
            line["cpu_baseline"] = base
            # the GPU solve of the same seeds (they are the first n_cpu scenarios of this rank's batch)
            if not args.no_parity:
                idx = torch.arange(n_cpu, device=dev)
                sub = batch.select(idx)
                if mode == "potential":
                    o = sub.solve(x0_dev[:n_cpu], U0_dev[:n_cpu], n_lqr_iter=50, tol=1e-3, trace=True)
                else:
                    o = dp.solve_distributed_round(sub, Xh_dev[:n_cpu], U0_dev[:n_cpu], 0.5)
                o = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in o.items()}
                line["parity_sample"] = parity_sample(results, o, a, mode)
                if not line["parity_sample"]["ok"]:
                    print("bench.py: PARITY SAMPLE MISMATCH (see parity_sample in the JSON line)", file=sys.stderr)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_rhc_sharded(args):
    """The one exchange step of the path: a decentralised receding-horizon run (reference
    distributed.py:106-221) of ONE 15-drone scenario (BASELINE config 4) whose agents' sub-problems are sharded over the
    ranks; after every round each rank all-gathers the agents' new trajectories (NCCL over NVLink, one collective per
    round, dpilqr_b200/parallel.py).  A step is one whole run; rank 0 also checks it against the unsharded run."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    a = 15
    x0, xf, U0 = scenarios.quad12_inputs(0, a, T)
    dp._reset_ids()
    ids = [100 + i for i in range(a)]
    dyn = dp.MultiDynamicalModel([dp.QuadcopterDynamics12D(0.1, id_) for id_ in ids])
    costs = [dp.ReferenceCost(xf[12 * i:12 * i + 12], np.eye(12), np.eye(4), 1000 * np.eye(12), id_) for i, id_ in enumerate(ids)]
    prob = dp.ilqrProblem(dyn, dp.GameCost(costs, dp.ProximityCost([12] * a, 0.5, [3] * a)))
    kw = dict(n_d=3, step_size=5, dist_converge=0.2, t_diverge=1.0, U0=U0, n_lqr_iter=8)

    def run(sharded):
        return dp.solve_rhc(prob, x0, T, 0.5, [], centralized=False, sharded=sharded and world > 1, **kw)

    for _ in range(args.warmup):
        run(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Xs, Us, Js = run(True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    Xr, Ur, Jr = run(False)
    same = torch.tensor([1 if (np.array_equal(Xs, Xr) and np.array_equal(Us, Ur) and Js == Jr) else 0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    rounds = (Xs.shape[0] - 1) // kw["step_size"] + 1
    if rank == 0:
        print(json.dumps({
            "metric": "decentralised receding-horizon rounds/s (one 15-drone Quadcopter12D scenario, agents sharded over the GPUs)",
            "value": rounds * args.steps / float(dt), "unit": "rounds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(dt) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE config 4: 15 x Quadcopter12D, solve_rhc(centralized=False, step_size=5, n_lqr_iter=8), seed 0",
                       "parallelism": f"agents' sub-problems sharded x{world}; one NCCL all-gather of the agents' trajectories per round",
                       "rounds_per_run": int(rounds), "identical_to_the_unsharded_run_on_every_rank": bool(same.item()),
                       "note": "latency-bound by construction (one scenario, a handful of sub-problems per rank and round): the record "
                               "of the path's only collective, not a throughput claim"}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios per GPU (weak scaling) or in total (strong scaling)")
    ap.add_argument("--agents", type=int, default=10, choices=[2, 3, 4, 5, 6, 7, 8, 10, 12, 15])
    ap.add_argument("--mode", default="potential", choices=["potential", "dp", "rhc-sharded"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-scenarios", type=int, default=0, help="sample size of the CPU baseline (default 8 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--inflight", type=int, default=0, help="solves in flight per GPU (1 = strictly one after the other; 0 = auto: 3, or 4 for batches under 1024 scenarios, within 100 GB of workspaces)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.mode == "rhc-sharded":
        run_rhc_sharded(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
