#!/usr/bin/env python
"""bench.py — headline benchmark of the TT core-contraction hot path on B200.

Workload (BASELINE.json configs[1], "cfg2"): `tt_compress!(ψ, 64; truncerr=0, sweeps=1)` of a random Float64 TT with
d=40, n=2, ranks min(2^k, 2^(40-k), 512) (97.9 MB of cores), i.e. one TT-rounding sweep = 78 two-site bond steps.
One "step" = one full sweep.  Metric: tt_rounding sweeps/s.

  value     device-timed (CUDA events on the library's stream), inputs resident in HBM, a fresh 97.9 MB input copy per step
            (distinct buffers, larger than nothing cached from the previous step);
  e2e       the same sweep through the public host API `ttn_b200.tt_compress_(host_tt, 64)` on pinned host buffers:
            H2D of all cores + sweep + D2H of the rounded cores, wall clock;
  roofline  the kernel family that dominates the step (per-family CUDA-event pass over the same steps);
  extras    local two-site matvec at cfg4 shapes (chi=1024, w=5, n=2): FP64 TFLOP/s and % of the measured FP64 GEMM peak.

cfg2 is a single sequential chain and does not shard: with --gpus N every rank runs an independent
replica ("replicas only", weak scaling, no collective on the data path).

`--impl reference` times the reference's own CPU algorithm for the same sweep (the NumPy restatement in oracle/, *including*
the discarded `orthogonalize` of src/tt_tools.jl:769 that the Julia code executes) on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, NPHYS, RMAX_IN, MAX_BOND = 40, 2, 512, 64
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")


# ---------------------------------------------------------------------------------------------------------------
# synthetic input: rand_tt(dims, 512; normalise=true), seed 1
# ---------------------------------------------------------------------------------------------------------------
def cfg2_ranks(d=D, rmax=RMAX_IN):
    return [min(2 ** k, 2 ** (d - k), rmax) for k in range(d + 1)]


def make_cfg2(seed=1, d=D, rmax=RMAX_IN, pinned=False):
    rks = cfg2_ranks(d, rmax)
    rng = np.random.default_rng(seed)
    cores = []
    keep = []
    for k in range(d):
        c = rng.standard_normal((NPHYS, rks[k], rks[k + 1])) / math.sqrt(NPHYS * rks[k + 1])
        c = np.asfortranarray(c)
        if pinned:
            import torch
            buf = torch.empty(c.size, dtype=torch.float64).pin_memory()
            view = buf.numpy().reshape(c.shape, order="F")
            view[...] = c
            keep.append(buf)
            c = view
        cores.append(c)
    return cores, rks, keep


def step_model(d=D, rmax=RMAX_IN, max_bond=MAX_BOND):
    """algorithmic FLOPs of one sweep per kernel family"""
    rks = cfg2_ranks(d, rmax)
    out = {"gemm_theta": 0.0, "gemm_proj": 0.0, "qr": 0.0, "jacobi": 0.0, "svd_rsvd": 0.0, "steps": []}
    for sweep_dir in (range(0, d - 1), range(d - 2, -1, -1)):
        for k in sweep_dir:
            p, r, q = NPHYS * rks[k], rks[k + 1], NPHYS * rks[k + 2]
            kk, m = min(p, q), max(p, q)
            rn = min(kk, max_bond)
            out["gemm_theta"] += 2.0 * p * r * q
            out["gemm_proj"] += 2.0 * rn * p * q
            out["qr"] += (2.0 * m * kk * kk - 2.0 / 3.0 * kk ** 3) if m > kk else 0.0
            out["jacobi"] += 20.0 * kk ** 3
            out["svd_rsvd"] += 6.0 * m * kk * kk + 20.0 * kk ** 3
            out["steps"].append((p, q, kk))
            rks[k + 1] = rn
    return out


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.time()                      # nvidia-smi needs ~0.1-0.3 s to print its first row: wait for it (bounded)
            while not self.rows and time.time() - t0 < 1.5:
                time.sleep(0.01)
            self.rows.clear()                     # rows printed before the timed region starts are not under load
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only place the oracle is executed: as the timed baseline, never as the product)
# ---------------------------------------------------------------------------------------------------------------
def _orth_flops(rks, i):
    """flop model of orthogonalize(ψ; i) (src/tt_tools.jl:511-543): contraction + geqrf + orgqr per site"""
    d = len(rks) - 1
    f, yl = 0.0, 1
    for j in range(0, i - 1):
        m, n = yl * NPHYS, rks[j + 1]
        k = min(m, n)
        f += 2.0 * yl * rks[j] * NPHYS * n + 2 * (2.0 * m * n * k - 2.0 / 3.0 * k ** 3)
        yl = k
    yr = 1
    for j in range(d - 1, i - 1, -1):
        m, n = yr * NPHYS, rks[j]
        k = min(m, n)
        f += 2.0 * yr * rks[j + 1] * NPHYS * n + 2 * (2.0 * m * n * k - 2.0 / 3.0 * k ** 3)
        yr = k
    return f


def cpu_reference_sweep(budget_s, faithful=True):
    """Runs the reference algorithm (NumPy restatement) for cfg2 until `budget_s` is used up; returns
    (seconds for a full sweep [extrapolated by the flop model if the budget ended first], steps done, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ttn_oracle as o
    cores, rks, _ = make_cfg2()
    x = o.TTvector(D, cores, (NPHYS,) * D, rks, [0] * D)
    order = list(range(1, D)) + list(range(D - 1, 0, -1))
    model = []
    cur = list(rks)
    for k in order:   # per-step flop model on the (deterministic, truncerr = 0) rank profile
        p, r, q = NPHYS * cur[k - 1], cur[k], NPHYS * cur[k + 1]
        kk, m = min(p, q), max(p, q)
        cur[k] = min(kk, MAX_BOND)
        fl = 2.0 * p * r * q + 6.0 * m * kk * kk + 20.0 * kk ** 3
        if faithful:
            fl += _orth_flops(cur, k)
        model.append(fl)
    t0, done = time.perf_counter(), 0
    for idx, k in enumerate(order):
        o.tt_bond_truncate(x, k, max_bond=MAX_BOND, truncerr=0.0, faithful=faithful)
        done = idx + 1
        if time.perf_counter() - t0 > budget_s and done < len(order):
            break
    el = time.perf_counter() - t0
    full = el * sum(model) / sum(model[:done])
    return full, done, len(order)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for N > 1; the CPU legs are meant to use every host core the BLAS can use."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1



# ---------------------------------------------------------------------------------------------------------------
# CPU legs of the extras (rank 0, N = 1 only): the oracle timed beside the GPU numbers, bounded samples
# ---------------------------------------------------------------------------------------------------------------
def cpu_extras(extras, chi=1024, w=5, nn=4):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ttn_oracle as o
    use_all_host_threads()
    cores = host_threads()
    rng = np.random.default_rng(4)
    out = {}
    # cfg4 matvec at full size: K_matfree lowered to three BLAS GEMMs (what @tensoropt does in the reference)
    G = rng.standard_normal((w, chi, chi)); H = rng.standard_normal((w, chi, chi))
    Am = rng.standard_normal((w, nn, nn, w)); V = rng.standard_normal((chi, nn, chi))
    o.dmrg_matvec2_blas(G, Am, V, H)
    t0 = time.perf_counter(); o.dmrg_matvec2_blas(G, Am, V, H); t_mv = time.perf_counter() - t0
    flops = 4.0 * w * nn * chi ** 3 + 2.0 * w * w * nn * nn * chi ** 2
    out["matvec_cfg4"] = {"value": flops / t_mv / 1e12, "unit": "TFLOP/s", "seconds": t_mv, "cores": cores, "kind": "port",
                          "sample": "one full-size application (chi=1024) through three BLAS GEMMs; NumPy restatement, not Julia"}
    # cfg4 DMRG sweep: one bulk bond step at full size = 8 Lanczos matvecs (the reference applies K twice per matvec,
    # dmrg.jl:241) + gesdd of the 2048 x 2048 two-site tensor + one environment update; sweep = 125 bond steps, of which
    # ~105 are at the full bond dimension for L = 64
    chi_svd = 2 * chi
    Th = rng.standard_normal((chi_svd, chi_svd))
    import scipy.linalg as sla
    t0 = time.perf_counter(); sla.svd(Th, full_matrices=False, lapack_driver="gesdd"); t_svd = time.perf_counter() - t0
    bond = 8 * 2 * t_mv + t_svd          # the environment update (43 GF, half a matvec) is left out: lower bound
    out["dmrg_sweep"] = {"value": 105 * bond, "unit": "s", "cores": cores, "kind": "port",
                         "bulk_bond_step_s": bond, "matvec_s": t_mv, "gesdd_s": t_svd, "gesdd_n": chi_svd,
                         "sample": "one bulk bond step measured at full size (8 Lanczos matvecs x 2 applications + gesdd 2048^2; "
                                   "environment update not counted), times the ~105 full-rank bond steps of the L=64 sweep; "
                                   "NumPy restatement, not Julia"}
    # cfg5: one vector through A*x + tt_compress!(y, 64) (algorithm-equivalent: the discarded orthogonalize is not run)
    d, r, W = 30, 64, 4
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
    g = np.random.default_rng(7)
    Aop = o.TToperator(d, [np.asfortranarray((g.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * g.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                             / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
    xv = o.TTvector(d, [np.asfortranarray((g.standard_normal((2, rks[k], rks[k + 1])) + 1j * g.standard_normal((2, rks[k], rks[k + 1])))
                                          / math.sqrt(4.0 * rks[k + 1])) for k in range(d)], (2,) * d, rks, [0] * d)
    t0 = time.perf_counter()
    y = o.apply(Aop, xv)
    for k in list(range(1, d)) + list(range(d - 1, 0, -1)):
        o.tt_bond_truncate(y, k, max_bond=r, truncerr=0.0, faithful=False)
    t_vec = time.perf_counter() - t0
    out["batch_cfg5"] = {"value": 1.0 / t_vec, "unit": "vectors/s", "seconds_per_vector": t_vec, "cores": cores, "kind": "port",
                         "sample": "one of the 4096 vectors (multithreaded BLAS/LAPACK inside the vector); NumPy restatement, not Julia"}
    for k, v in out.items():
        if k in extras:
            extras[k]["cpu_baseline"] = v


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    if rank != 0:
        return
    use_all_host_threads()
    budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    times = []
    for i in range(args.warmup + args.steps):
        full, done, total = cpu_reference_sweep(budget, faithful=True)
        if i >= args.warmup:
            times.append(full)
    t = float(np.mean(times)) if times else float("nan")
    val = 1.0 / t
    cores = host_threads()
    sample = (f"reference-faithful tt_compress! (incl. the discarded orthogonalize, tt_tools.jl:769) on the cfg2 input: "
              f"first {done} of {total} bond steps within a {budget:.0f} s budget per step, extrapolated to the full sweep "
              f"with the per-bond flop model; NumPy restatement of the reference, not Julia")
    line = {"impl": "reference", "metric": "tt_rounding sweeps/s", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2: tt_compress!(rand_tt d=40 n=2 r=512, 64; truncerr=0, sweeps=1)", "d": D, "n": NPHYS,
                       "rank_in": RMAX_IN, "max_bond": MAX_BOND},
            "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import ttn_b200 as t
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        # NCCL may print its version banner on stdout: keep stdout clean for the single JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist_.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        dist = dist_
    K, W = args.steps, args.warmup
    cores, rks, keep = make_cfg2(pinned=True)
    host_tt = lambda: t.TTvector(D, list(cores), (NPHYS,) * D, list(rks))  # noqa: E731
    base = t.DeviceTT.upload(host_tt())
    t.synchronize()
    stream = torch.cuda.ExternalStream(t.stream_handle())

    def barrier():
        t.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(nsteps, copies):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(nsteps):
            t.tt_compress_(copies[i], MAX_BOND)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    # ---- device-resident timing -------------------------------------------------------------------------------
    copies = [base.copy() for _ in range(W)]
    timed(W, copies)
    copies = [base.copy() for _ in range(K)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t.reset_launch_count()
    ms = timed(K, copies)
    launches = t.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    out_rks = copies[0].ttv_rks
    if dist is not None:
        tm = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    ms_per_step = ms / K
    value = world / (ms_per_step * 1e-3)

    # ---- end to end through the host API (pinned host buffers, H2D + sweep + D2H inside the timed region) --------
    for _ in range(min(W, 2)):
        t.tt_compress_(host_tt(), MAX_BOND)
    barrier()
    w0 = time.perf_counter()
    res = None
    for _ in range(K):
        res = t.tt_compress_(host_tt(), MAX_BOND)
    t.synchronize()
    e2e_s = (time.perf_counter() - w0) / K
    if dist is not None:
        tm = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    h2d = int(sum(c.nbytes for c in cores))
    d2h = int(sum(c.nbytes for c in res.ttv_vec))

    # ---- roofline pass: per-kernel-family CUDA events over the same K steps --------------------------------------
    copies = [base.copy() for _ in range(K)]
    t.synchronize()
    t.profile(True)
    ms_prof = timed(K, copies)
    fam = t.profile_read()
    t.profile(False)
    model = step_model()
    peak64 = json.load(open(FP64_PEAK_FILE)) if os.path.exists(FP64_PEAK_FILE) else {"fp64_tflops": 35.45}
    fam_ms = {k: v[0] / K for k, v in fam.items()}
    fam_cnt = {k: v[1] // K for k, v in fam.items()}
    dom = max(fam_ms, key=fam_ms.get)
    alg = {"jacobi": model["jacobi"], "gemm": model["gemm_theta"] + model["gemm_proj"], "qr_panel": model["qr"],
           "qr_apply": model["qr"]}.get(dom, 0.0)
    achieved = alg / (fam_ms[dom] * 1e-3) / 1e12 if fam_ms[dom] > 0 else 0.0
    # DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    # capture, profiles/ncu_cfg2_kernels_r01.txt / ncu_gemm_r01.txt); the bond matrices live in L2, so it is tiny
    traffic = {"jacobi": 161536, "qr_panel": 1132800, "gemm": None}.get(dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak64["fp64_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peak64["fp64_tflops"], "traffic": traffic,
                "traffic_source": "profiles/ncu_cfg2_kernels_r01.txt (per launch; algorithmic bytes of a 128 x 128 Jacobi "
                                  "launch are 262144: one read + one write of the matrix, both served by L2)",
                "peak_source": "measured cuBLAS FP64 GEMM 8192^3 burst on this pool's B200 (profiles/fp64_peak_r01.json; "
                               "MEASURED_PEAKS.json carries no FP64 figure)",
                "algorithmic_flops_per_step": alg, "launches_per_step": fam_cnt[dom], "ms_per_step_in_kernel": fam_ms[dom],
                "family_ms_per_step": fam_ms, "profiled_ms_per_step": ms_prof / K,
                "step_lower_bound_ms": max((model["gemm_theta"] + model["svd_rsvd"]) / (peak64["fp64_tflops"] * 1e12),
                                           (h2d + d2h) / 6547.8e9) * 1e3,
                "note": "cfg2 is a 78-step dependency chain on bond-sized matrices: latency bound, not roofline bound "
                        "(SURVEY.md §8(d)-2)"}

    # ---- extras: the other components of BASELINE.json's metric ---------------------------------------------------
    extras = {}
    if not args.no_extras:
        # cfg5 (sharded over the ranks, no collective): batched ComplexF64 apply + tt_compress!
        bt = bench_batch(t, rank, args.batch_vectors)
        if dist is not None:
            tm = torch.tensor([bt["seconds"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            bt["seconds"] = float(tm.item())
        bt["value"] = world * bt["vectors_per_rank"] / bt["seconds"]
        bt["n_gpus"] = world
        extras["batch_cfg5"] = bt
        if world > 1:
            # cfg4 matvec sharded on the spectator bond, all-gather fused into the last GEMM
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import shard_bench
            extras["matvec_cfg4_sharded"] = shard_bench.run(torch, dist, rank, world, chi=1024, w=5, reps=10, krylovdim=8)
        if rank == 0:
            extras["matvec_cfg4"] = bench_matvec(t, torch, stream, peak64)
            extras["dmrg_sweep"] = bench_dmrg(t, args.dmrg_chi)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import cfg3_bench
            extras["mals_cfg3"] = cfg3_bench.run(bits=20, rmax=128)
            if world == 1 and not args.no_cpu:
                cpu_extras(extras)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            use_all_host_threads()
            full, done, total = cpu_reference_sweep(25.0, faithful=False)
            ffull, fdone, ftotal = cpu_reference_sweep(12.0, faithful=True)
            cpu = {"value": 1.0 / full, "unit": "sweeps/s", "cores": host_threads(), "kind": "port",
                   "sample": f"algorithm-equivalent sweep (dead orthogonalize removed): {done}/{total} bond steps of cfg2 run in full"
                             if done == total else f"algorithm-equivalent sweep: first {done}/{total} bond steps, flop-model extrapolated",
                   "reference_faithful_value": 1.0 / ffull,
                   "reference_faithful_sample": f"first {fdone}/{ftotal} bond steps incl. the discarded orthogonalize "
                                                f"(tt_tools.jl:769), flop-model extrapolated",
                   "note": "NumPy restatement of the reference, not Julia"}
        line = {"metric": "tt_rounding sweeps/s", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "cfg2: tt_compress!(rand_tt d=40 n=2 r=512, 64; truncerr=0, sweeps=1)", "d": D, "n": NPHYS,
                           "rank_in": RMAX_IN, "max_bond": MAX_BOND, "out_max_rank": int(max(out_rks)),
                           "parallelism": "replicas only (cfg2 does not shard)" if world > 1 else "single GPU",
                           "cache": "fresh 97.9 MB device copy of the input per step (distinct buffers)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": world / e2e_s, "unit": "sweeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "api": "ttn_b200.tt_compress_(host TTvector on pinned memory, 64)"},
                "roofline": roofline, "cpu_baseline": cpu, "extras": extras}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_matvec(t, torch, stream, peak64, chi=1024, w=5, nn=4, reps=10):
    """local two-site matvec Y = L·W·V·R at cfg4 shapes; F = 4 w n² chi³ + 2 w² n⁴ chi²"""
    import ctypes as C
    from ttn_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(4)
    G = np.asfortranarray(rng.standard_normal((w, chi, chi)))
    H = np.asfortranarray(rng.standard_normal((w, chi, chi)))
    Am = np.asfortranarray(rng.standard_normal((w, nn, nn, w)))
    V = np.asfortranarray(rng.standard_normal((chi, nn, chi)))
    mv = C.c_void_p()
    _lib.check(lib.ttn_matvec2_create(0, w, w, chi, chi, nn, G.ctypes.data, Am.ctypes.data, H.ctypes.data, C.byref(mv)))
    dV, dY = C.c_void_p(), C.c_void_p()
    _lib.check(lib.ttn_dev_alloc(V.nbytes, C.byref(dV)))
    _lib.check(lib.ttn_dev_alloc(V.nbytes, C.byref(dY)))
    _lib.check(lib.ttn_h2d(dV, V.ctypes.data, V.nbytes))
    for _ in range(3):
        _lib.check(lib.ttn_matvec2_apply(mv, dV, dY))
    t.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        _lib.check(lib.ttn_matvec2_apply(mv, dV, dY))
    e1.record(stream)
    t.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * w * nn * chi ** 3 + 2.0 * w * w * nn * nn * chi ** 2
    tf = flops / (ms * 1e-3) / 1e12
    _lib.check(lib.ttn_matvec2_free(mv))
    _lib.check(lib.ttn_dev_free(dV))
    _lib.check(lib.ttn_dev_free(dY))
    return {"metric": "local-matvec FP64 TFLOP/s", "value": tf, "ms": ms, "gflop": flops / 1e9, "chi": chi, "w": w, "n2": nn,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak64["fp64_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tf / peak64["fp64_tflops_sustained"]},
            "note": "working set 0.5 GB > L2; back-to-back applications"}


def bench_batch(t, rank, nvec, d=30, r=64, W=4):
    """cfg5: `nvec` independent ComplexF64 QTT vectors per rank (d=30, rank 64) through y = A*x (MPO rank W) and
    tt_compress!(y, 64); inputs resident in HBM."""
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
    rng = np.random.default_rng(7)
    A = t.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1])))
                                           / math.sqrt(2.0 * Rk[k + 1])) for k in range(d)], (2,) * d, Rk)
    Ad = t.DeviceTTO.upload(A)
    g = np.random.default_rng(100 + rank)
    cores = []
    for k in range(d):
        shp = (2, rks[k], rks[k + 1], nvec)
        cores.append(np.asfortranarray((g.standard_normal(shp) + 1j * g.standard_normal(shp)) / math.sqrt(4.0 * rks[k + 1])))
    xs = [t.TTvector(d, [c[..., b] for c in cores], (2,) * d, rks) for b in range(nvec)]
    xd = t.DeviceTT.upload(xs)
    t.tt_compress_(t.apply(Ad, xd), r)
    t.synchronize()
    t0 = time.perf_counter()
    y = t.tt_compress_(t.apply(Ad, xd), r)
    t.synchronize()
    el = time.perf_counter() - t0
    return {"metric": "batched apply+round vectors/s", "unit": "vectors/s", "vectors_per_rank": nvec, "seconds": el, "d": d,
            "rank": r, "W": W, "dtype": "c128", "out_max_rank": int(max(y.ttv_rks)),
            "sharding": "vectors split evenly over the ranks, no data-path collective"}


def bench_dmrg(t, chi, L=64, kd=8):
    """cfg4-style DMRG sweep: Heisenberg XYZ chain L=64 (MPO rank 5), one full two-site sweep from a random TT capped at
    bond `chi`, fixed Lanczos budget (krylovdim 8 x 1 restart) as in SURVEY.md §8(d)-4."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dmrg_bench
    H = t.DeviceTTO.upload(dmrg_bench.heisenberg(L))
    x0 = t.DeviceTT.upload(dmrg_bench.rand_tt(L, chi))
    t.synchronize()
    t.reset_launch_count()
    t0 = time.perf_counter()
    E, x, rh = t.dmrg_eigsolve(H, x0, N=2, tol=1e-12, sweep_schedule=[2], rmax_schedule=[chi], linsolv_maxiter=1,
                               linsolv_tol=1e-10, krylovdim=kd)
    t.synchronize()
    el = time.perf_counter() - t0
    return {"metric": "DMRG sweep s", "value": el, "unit": "s", "L": L, "chi": chi, "krylovdim": kd, "bond_steps": len(E),
            "max_rank": int(max(rh)), "E_last": float(E[-1]), "gpu_launches": int(t.launch_count()),
            "note": "cfg4 (BASELINE.json configs[3]): Heisenberg XYZ L=64, MPO rank 5, one full two-site sweep from a random TT capped at chi"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg4 matvec / DMRG sweep / cfg5 batch extras")
    ap.add_argument("--dmrg-chi", type=int, default=1024, help="bond cap of the DMRG sweep extra (cfg4: 1024)")
    ap.add_argument("--batch-vectors", type=int, default=256, help="cfg5 vectors per rank in the batch extra")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
