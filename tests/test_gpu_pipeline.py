"""Concurrent callers (the bulk token / tail stream of csrc/solver.cu, SolvePipeline), empty and extreme batches."""

import ctypes
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _metric_batch(first, count, a=10):
    import dpilqr_b200 as dp
    from dpilqr_b200 import scenarios

    specs, x0, U0 = scenarios.quad12_batch(first, count, a, 50)
    return dp.CompiledBatch(specs, 50), x0, U0


def test_solve_pipeline_equals_one_solve_at_a_time():
    """Three different batches, two in flight: every result equals the one obtained alone, bit for bit.  One batch is
    above the bulk threshold (600 problems: it takes the bulk token and hands it over for its tail), two are below."""
    import dpilqr_b200 as dp

    jobs = [_metric_batch(0, 640, 3), _metric_batch(700, 48), _metric_batch(900, 32)]
    alone = [b.solve(x0, U0) for b, x0, U0 in jobs]
    with dp.SolvePipeline(depth=2) as pipe:
        piped = list(pipe.map(lambda job: job[0].solve(job[1], job[2]), jobs * 2))
    for k, out in enumerate(piped):
        ref = alone[k % len(jobs)]
        for key in ("X", "U", "J", "iters", "status"):  # (J of a failed line search over NaN costs is NaN in both)
            assert np.array_equal(out[key].cpu().numpy(), ref[key].cpu().numpy(), equal_nan=(key == "J")), (k, key)


def test_host_entry_point_from_two_threads():
    """dpilqr_solve_batch_host from two host threads at once (arena pool, per-thread streams): same answers as one
    after the other."""
    from dpilqr_b200 import _native

    lib = _native.lib()
    results = {}

    inputs = {(0, 40): _metric_batch(0, 40), (100, 24): _metric_batch(100, 24)}  # (the scenario generator uses NumPy's global RNG: not from threads)

    def call(tag, first, count):
        batch, x0, U0 = inputs[(first, count)]
        B, n, m, T = count, batch.n, batch.m, 50
        host = dict(model=batch.t_model.cpu(), ndims=batch.t_ndims.cpu(), cidx=batch.t_cidx.cpu(), Q=batch.t_Q.cpu(), R=batch.t_R.cpu(),
                    Qf=batch.t_Qf.cpu(), xf=batch.t_xf.cpu(), radius=batch.t_radius.cpu(), weights=batch.t_weights.cpu(), hasprox=batch.t_hasprox.cpu())
        hb = _native.BatchStruct(B, batch.a, 12, 4, T, int(batch.t_Q.shape[0]), batch.dt, host["model"].data_ptr(), host["ndims"].data_ptr(),
                                 host["cidx"].data_ptr(), host["Q"].data_ptr(), host["R"].data_ptr(), host["Qf"].data_ptr(), host["xf"].data_ptr(),
                                 host["radius"].data_ptr(), host["weights"].data_ptr(), host["hasprox"].data_ptr(), batch.model_hint, 0)
        opts = _native.SolveOpts(50, 10, 1e-3, 0.0, 0, 0, 1, 0)
        X, U = np.empty((B, T + 1, n)), np.empty((B, T, m))
        J, Js = np.empty(B), np.empty(B)
        iters, status = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
        x0c, U0c = np.ascontiguousarray(x0), np.ascontiguousarray(U0)
        ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        total = lib.dpilqr_solve_batch_host(ctypes.byref(hb), ctypes.byref(opts), ptr(x0c), ptr(U0c), ptr(X), ptr(U), ptr(J), ptr(Js),
                                            ptr(iters), ptr(status), None, None, None, 0)
        assert total >= 0, _native.last_error()
        results[tag] = (X, U, iters.copy())

    call("a0", 0, 40)
    call("b0", 100, 24)
    threads = [threading.Thread(target=call, args=("a1", 0, 40)), threading.Thread(target=call, args=("b1", 100, 24))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for tag in "ab":
        for x, y in zip(results[tag + "0"], results[tag + "1"]):
            assert np.array_equal(x, y)
    lib.dpilqr_release_cache()


def test_empty_batch_and_zero_iterations():
    import dpilqr_b200 as dp

    batch, x0, U0 = _metric_batch(0, 4, 3)
    out = batch.solve(x0, U0, n_lqr_iter=0)  # control.py:176: the loop does not run, the warm start comes back
    Xr, Jr = batch.rollout(x0, U0)
    assert np.array_equal(out["X"].cpu().numpy(), Xr.cpu().numpy()) and np.array_equal(out["U"].cpu().numpy(), U0)
    assert int(out["total_iters"]) == 0 and np.all(out["iters"].cpu().numpy() == 0)
    assert dp.solve_specs([], [], [], 50) == []


def test_largest_and_too_large_teams():
    """16 drones (64 joint controls) is the largest team the backward kernel takes; 17 is refused with an error, not a
    wrong answer."""
    import dpilqr_b200 as dp
    from dpilqr_b200 import _native, scenarios
    from oracle import ilqr_oracle as O

    a, N = 16, 50
    x0, xf, U0 = scenarios.quad12_inputs(1, a, N)
    batch = dp.CompiledBatch([scenarios.quad12_spec(xf, a)], N)
    out = batch.solve(x0[None], U0[None], n_lqr_iter=2, trace=True)
    prob = O.OracleProblem(["Quadcopter12D"] * a, 0.1, xf, np.eye(12), np.eye(4), 1000 * np.eye(12), 0.5, [3] * a, [100 + i for i in range(a)])
    solver = O.OracleSolver(prob, N)
    Xo, Uo, Jo = solver.solve(x0, U0.copy(), n_lqr_iter=2)
    assert [int(v) for v in out["trace_alpha"][0, :int(out["iters"][0])]] == [r["alpha_index"] for r in solver.trace]
    err = np.max(np.abs(out["X"][0].cpu().numpy() - Xo)) / np.max(np.abs(Xo))
    print(f"16 drones, two iterations: X err {err:.1e}")
    assert err < 1e-8
    x0, xf, U0 = scenarios.quad12_inputs(1, 17, N)
    big = dp.CompiledBatch([scenarios.quad12_spec(xf, 17)], N)
    with pytest.raises(_native.NativeError):
        big.solve(x0[None], U0[None], n_lqr_iter=1)
