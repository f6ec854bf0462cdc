#!/usr/bin/env python
"""bench.py -- batched iLQR iterations/sec on 4096 x 10-agent Quadcopter12D scenarios.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one complete batched Potential-iLQR solve (ilqrSolver.solve semantics, reference
control.py:150-225) of `--scenarios` synthetic scenarios per GPU; the metric counts iLQR
iterations (one backward Riccati pass + its 10-candidate line search for one problem) per second.

  value         device-timed, inputs resident in HBM when the timed region starts
  e2e           same metric through the public API with host (pinned) buffers: H2D of x0/U0 and D2H
                of X/U/J/iters inside the timed region
  roofline      dominant kernel (backward Riccati, FP64-compute bound) against the DFMA peak
                measured on this box by tools/bin/fp64_peak; HBM view alongside
  cpu_baseline  the CPU oracle (oracle/, a restatement of the reference's algorithm using the
                reference's own compiled dynamics when oracle/_ref is present) on all host cores,
                on a bounded sample of the same scenarios

Multi-GPU (torchrun, one rank per GPU): scenarios are independent, each rank solves its own 4096
(weak scaling), no data-path collective; time is the max over ranks.
"""

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched iLQR iterations/sec (4096x10-agent Quad12D)"
UNIT = "iterations/s"
A, S, C, T = 10, 12, 4, 50


def backward_flops(a=A, s=S, c=C, T=T):
    """Structure-aware FLOPs of one backward pass."""
    n, m = a * s, a * c
    step = (4 * n * n * s + 4 * m * n * s + 2 * m * m * s + (2.0 / 3.0) * m ** 3 + 2 * m * m * (n + 1) + 2 * n * m * m
            + 4 * n * n * m + 8 * n * m + 2 * n * (s + c))
    return step * T


def backward_hbm_bytes(a=A, s=S, c=C, T=T):
    """Algorithmic HBM bytes of one backward pass: stage records in, K and d out."""
    n, m = a * s, a * c
    pairs = a * (a - 1) // 2
    stage = a * s * s + a * s * c + n + m + 9 * a + 9 * pairs
    return 8 * ((T + 1) * stage + T * m * n + T * m)


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------
def _cpu_worker(k):
    import numpy as np

    from dpilqr_b200 import scenarios
    from oracle import ilqr_oracle as O

    x0, xf, U0 = scenarios.quad12_inputs(k, A, T)
    prob = O.OracleProblem(["Quadcopter12D"] * A, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * A,
                           [100 + i for i in range(A)])
    solver = O.OracleSolver(prob, T)
    solver.solve(x0, U0)
    return solver.n_backward


def cpu_reference_run(n_scen, first=0):
    """iterations/s of the CPU oracle over `n_scen` scenarios on all host cores (scenario-level pool,
    one BLAS thread per worker -- the working equivalent of the reference's multiprocessing path,
    SURVEY.md section 8d)."""
    import multiprocessing as mp

    from oracle import ilqr_oracle as O

    O.build_c_oracle()
    kind = "reference" if O.dynamics_backend("auto").name == "reference-native" else "port"
    cores = os.cpu_count() or 1
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, range(first, first + min(cores, n_scen)))  # warm the workers (imports, dlopen)
        t0 = time.perf_counter()
        iters = pool.map(_cpu_worker, range(first, first + n_scen), chunksize=1)
        dt = time.perf_counter() - t0
    return dict(value=sum(iters) / dt, unit=UNIT, cores=cores, kind=kind, iterations=int(sum(iters)), seconds=dt,
                sample=f"{n_scen} of the 4096 scenarios (seeds {first}..{first + n_scen - 1}), Potential-iLQR, "
                       f"multiprocessing.Pool({cores}) over scenarios, 1 BLAS thread/worker; python restatement of the "
                       f"reference loop (oracle/ilqr_oracle.py) on the {'reference-compiled' if kind == 'reference' else 'restated C'} dynamics")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_scen = args.cpu_scenarios or max(8 * cores, 64)
    vals = []
    for step in range(args.warmup + args.steps):
        res = cpu_reference_run(n_scen)
        if step >= args.warmup:
            vals.append(res)
    value = sum(r["iterations"] for r in vals) / sum(r["seconds"] for r in vals)
    base = dict(vals[-1], value=value)
    base.pop("iterations"), base.pop("seconds")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(r["seconds"] for r in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n_scen}-scenario sample of 4096 x 10-agent Quadcopter12D Potential-iLQR, N=50, dt=0.1 (CPU)"},
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


# dram__bytes_read.sum + dram__bytes_write.sum of one backward launch over 3908 problems (profiles/r01_backward_ncu_full.txt)
NCU_BACKWARD_DRAM_BYTES_PER_PROBLEM = (4.180168e9 + 7.515227e9) / 3908


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
    B = args.scenarios
    # ---- problem construction (excluded from the timed regions)
    specs, x0_np, U0_np = scenarios.quad12_batch(rank * B, B, A, T)
    batch = dp.CompiledBatch(specs, T, dev)
    x0_dev, U0_dev = torch.as_tensor(x0_np).to(dev), torch.as_tensor(U0_np).to(dev)
    x0_pin, U0_pin = torch.as_tensor(x0_np).pin_memory(), torch.as_tensor(U0_np).pin_memory()
    out_pin = dict(X=torch.empty((B, T + 1, A * S), dtype=torch.float64).pin_memory(),
                   U=torch.empty((B, T, A * C), dtype=torch.float64).pin_memory(),
                   J=torch.empty(B, dtype=torch.float64).pin_memory(), iters=torch.empty(B, dtype=torch.int32).pin_memory())
    fp64_peak = measure_fp64_peak() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident(profile=False):
        return batch.solve(x0_dev, U0_dev, n_lqr_iter=50, tol=1e-3, profile=profile)["total_iters"]

    def step_e2e():
        out = batch.solve(x0_pin, U0_pin, n_lqr_iter=50, tol=1e-3)
        for k, buf in out_pin.items():
            buf.copy_(out[k], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out["total_iters"]

    for _ in range(args.warmup):
        step_resident()
    _native.get_profile(reset=True)
    # ---- timed: device-resident
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 0
    for _ in range(args.steps):
        iters += step_resident(profile=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    prof = _native.get_profile(reset=True)
    # ---- timed: end to end through the public API with host buffers
    step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    iters_e2e = 0
    for _ in range(args.steps):
        iters_e2e += step_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
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
        achieved = backward_flops() * bunits / (bms * 1e-3) * 1e-12 if bms > 0 else None
        peak, peak_src = (fp64_peak, "FP64 peak measured on this box by tools/bin/fp64_peak, max of the DFMA and DMMA m8n8k4 loops "
                                     "(MEASURED_PEAKS.json has no FP64 entry)") \
            if fp64_peak else (37.2, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz (fp64_peak binary missing)")
        hbm_peak = 6453.7
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        total_ms = sum(v[0] for v in prof.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{B} scenarios/GPU x 10-agent Quadcopter12D Potential-iLQR (ilqrSolver.solve), N=50, dt=0.1, "
                                   "n_lqr_iter=50, tol=1e-3, hover warm start, random_setup energy=30 (SURVEY 8d)",
                       "iterations_per_step": iters / args.steps / world,
                       "l2": "working set per step (K 7.9 GB + candidates 5.3 GB + stage 4.4 GB) >> 126 MB L2",
                       "parallelism": f"scenario-sharded x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x0_pin.numel() * 8 + U0_pin.numel() * 8),
                    "d2h_bytes_per_step": int(sum(b.numel() * b.element_size() for b in out_pin.values())),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(sum(v[1] for v in prof.values())),
            "clocks": clocks,
            "roofline": {"kernel": "backward_kernel<12,4,10>", "bound": "tensor", "bound_detail": "FP64 pipe: mma.sync.m8n8k4.f64 tiles + DFMA",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": NCU_BACKWARD_DRAM_BYTES_PER_PROBLEM * bunits / max(blaunch, 1),
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum = 11.695 GB for a 3908-problem launch "
                                           "(profiles/r01_backward_ncu_full.txt), scaled to this run's average problems per launch; "
                                           "algorithmic bytes/problem = %d" % backward_hbm_bytes(),
                         "peak_source": peak_src,
                         "flops_per_launch_unit": backward_flops(), "launches": blaunch, "avg_launch_ms": bms / max(blaunch, 1),
                         "share_of_step": bms / total_ms if total_ms else None},
            "roofline_hbm": {"kernel": "backward_kernel<12,4,10>", "bound": "hbm",
                             "achieved": backward_hbm_bytes() * bunits / (bms * 1e-3) * 1e-9 if bms > 0 else None,
                             "peak": hbm_peak, "unit": "GB/s", "frac": (backward_hbm_bytes() * bunits / (bms * 1e-3) * 1e-9 / hbm_peak) if bms > 0 else None},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            base = cpu_reference_run(args.cpu_scenarios or max(8 * cores, 64))
            base.pop("iterations"), base.pop("seconds")
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--cpu-scenarios", type=int, default=0, help="sample size of the CPU baseline (default 8 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
