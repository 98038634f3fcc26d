#!/usr/bin/env python
"""bench.py -- LU GFLOP/s (2n^3/3) for NxN Float64 on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 16384]

One "step" = one LU factorization with partial pivoting of a fresh copy of the same synthetic
U[0,1) matrix (numpy default_rng(12), the distribution of test/runtests.jl:45).

ours:       value = device-resident throughput (matrix already in HBM, CUDA events on the library's
            stream around each factorization; the restore copy between steps is outside the timed
            region); e2e = the same metric through the public host API (`rfb200.lu_` on a pinned host
            matrix: H2D + LU + D2H inside the timed region; finished tiles of the factors travel back
            while the factorization runs, the older row-band scheme is timed beside it);
            roofline = all K4 launches of one LU against max(DMMA microbenchmark, pipe rate at the
            maximum SM clock); checks = residual, pivots against LAPACK dgetrf (also timed).
reference:  the reference is pure Julia and cannot run here (no Julia runtime), so
            this arm times the CPU oracle port of its algorithm (oracle/rf_oracle.c, all host cores)
            at the SAME size as the product arm whenever steps + warmup factorizations fit its time
            budget (--ref-budget-s), otherwise on the largest sample that does (`same_config` says
            which); with a Julia runtime under baseline/_ref it runs the real package instead.

N > 1 (torchrun, one process per GPU): ONE 32768 x 32768 matrix (BASELINE config 4) is factored by all
ranks together -- 1-D block-cyclic block columns, owner-rooted NCCL broadcast of each factored block
column, C++ scheduler behind rfb_mg_* (csrc/rfb_mg.cu; dist_lu.py is a thin ctypes caller); strong
scaling, value = 2n^3/3 / max-over-ranks time; the same matrix is also factored on ONE GPU in the same
run (`strong_n1_value`, pivot equality).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LU GFLOP/s (2n^3/3) for NxN Float64; residual ||PA-LU||_F/||A||_F"


def lu_flops(n):
    return 2.0 * n ** 3 / 3.0


def fill_random(a, seed=12, chunk=1024):
    """U[0,1) entries, generated per column block so the 2 GB matrix never exists twice."""
    n = a.shape[1]
    for j0 in range(0, n, chunk):
        j1 = min(n, j0 + chunk)
        rng = np.random.default_rng([seed, j0 // chunk])
        a[:, j0:j1] = rng.random((a.shape[0], j1 - j0), dtype=a.dtype).reshape(a.shape[0], j1 - j0)


def hutchinson_residual(a0, factors, ipiv, nvec=8, seed=0):
    """Estimate ||P A - L U||_F / ||A||_F with random +-1 probes (O(n^2) per probe, no oracle needed)."""
    rng = np.random.default_rng(seed)
    m, n = a0.shape
    p = np.arange(m)
    for i, ip in enumerate(ipiv):
        ip = int(ip) - 1
        if ip != i:
            p[i], p[ip] = p[ip], p[i]
    mn = min(m, n)
    x = rng.integers(0, 2, size=(n, nvec)).astype(np.float64) * 2 - 1
    ux = np.triu(factors[:mn, :]) @ x
    lux = np.tril(factors[:, :mn], -1) @ ux
    lux[:mn] += ux
    pax = (a0 @ x)[p]                  # P (A x): no permuted copy of the matrix
    return float(np.linalg.norm(pax - lux) / np.sqrt(nvec) / np.linalg.norm(a0))


def near_tie(a0, factors, ipiv, want_ipiv, k):
    """Is the first differing pivot step k a near-tie?  The row the other implementation moved to position k must
    have |L| >= 1 - 20 n eps in OUR factorization (the two candidates were equal to within the factorization's own
    error); everything after a legitimate near-tie diverges by construction and is covered by the residual."""
    m, n = a0.shape

    def perm(ip, upto):
        p = np.arange(m)
        for i in range(upto):
            j = int(ip[i]) - 1
            if j != i:
                p[i], p[j] = p[j], p[i]
        return p
    orig = perm(want_ipiv, k + 1)[k]
    pos = int(np.nonzero(perm(ipiv, len(ipiv)) == orig)[0][0])
    if pos <= k:
        return False
    return abs(float(factors[pos, k])) >= 1.0 - 20 * max(m, n) * float(np.finfo(a0.dtype).eps)


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(smax), power_w_max=max(power),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline(n_sample, threads, reps=1):
    """Oracle port (restatement of the reference algorithm, reference defaults) on the host cores."""
    from oracle import rf_oracle as O
    a0 = np.empty((n_sample, n_sample), dtype=np.float64, order="F")
    fill_random(a0)
    best = None
    for _ in range(reps):
        a = a0.copy(order="F")
        t = time.perf_counter()
        O.lu_c(a, threads=threads)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return lu_flops(n_sample) / best / 1e9, best


def run_julia_reference(args):
    """The real reference (baseline/run_reference.jl -> RecursiveFactorization.lu!) when a Julia runtime with an
    instantiated environment exists under baseline/_ref.  Neither exists in the build image or on the GPU boxes
    (SURVEY.md F2/F3), so this normally returns None and the C port is timed instead."""
    import shutil
    julia = shutil.which("julia")
    proj = os.path.join(ROOT, "baseline", "_ref")
    if not julia or not os.path.exists(os.path.join(proj, "Project.toml")):
        return None
    try:
        res = subprocess.run([julia, f"--project={proj}", "-t", "auto", os.path.join(ROOT, "baseline", "run_reference.jl"),
                              str(args.n), str(args.steps), str(args.warmup)], capture_output=True, text=True,
                             timeout=max(600.0, 2 * args.ref_budget_s))
        d = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    except Exception:
        return None
    cores = int(d.get("threads", os.cpu_count() or 1))
    sample = (f"{args.n}x{args.n} Float64 U[0,1) LU per step (the whole workload), RecursiveFactorization.lu! "
              f"{d.get('version', '')} (the unmodified Julia reference), {cores} Julia threads")
    return {
        "impl": "reference", "metric": METRIC, "value": d["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": d["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.n}x{args.n} Float64 LU with partial pivoting", "sample_n": args.n, "same_config": True},
        "same_config": True,
        "cpu_baseline": {"value": d["value"], "unit": "GFLOP/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the reference algorithm on all host cores.  Same config as the product arm
    (args.n) whenever steps + warmup factorizations at the calibrated rate fit the time budget; otherwise the largest
    sample that does, and then the sample size is part of `config.workload` and `same_config` is false."""
    if rank != 0:
        return
    real = run_julia_reference(args)
    if real is not None:
        print(json.dumps(real))
        return
    from oracle import rf_oracle as O
    cores = os.cpu_count() or 1
    t_start = time.perf_counter()
    cpu_baseline(2048, cores)                         # first call: thread pool start-up, page faults
    gf_cal, _ = cpu_baseline(4096, cores)             # optimistic calibration (only used to skip hopeless sizes)
    total = max(1, args.steps + args.warmup)
    # Largest size whose `steps + warmup` factorizations fit the budget, decided on a MEASURED factorization at that size:
    # the first one (a warm-up when there is one) is timed, and the size is kept only if the projection fits what is left.
    n_s, a0, done = None, None, []
    for cand in [args.n] + [c for c in (24576, 16384, 12288, 8192, 6144, 4096, 3072, 2048) if c < args.n]:
        left = args.ref_budget_s - (time.perf_counter() - t_start)
        if cand > 2048 and lu_flops(cand) / (gf_cal * 1e9) * total > left:
            continue
        a0 = np.empty((cand, cand), dtype=np.float64, order="F")
        fill_random(a0)
        a = a0.copy(order="F")
        t = time.perf_counter()
        O.lu_c(a, threads=cores)
        dt = time.perf_counter() - t
        left = args.ref_budget_s - (time.perf_counter() - t_start)
        if cand <= 2048 or dt * (total - 1) <= left:
            n_s, done = cand, [dt]
            break
    times = []
    for it in range(args.warmup + args.steps):
        if it == 0 and done:
            dt = done[0]                              # the probe above was this first factorization
        else:
            a = a0.copy(order="F")
            t = time.perf_counter()
            O.lu_c(a, threads=cores)
            dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = lu_flops(n_s) / (ms * 1e-3) / 1e9
    same = n_s == args.n
    what = "the whole workload" if same else f"bounded sample of the {args.n}x{args.n} workload"
    sample = (f"{n_s}x{n_s} Float64 U[0,1) LU per step ({what}), "
              f"oracle/rf_oracle.c = C restatement of src/lu.jl with reference defaults (blocksize 8, threshold 48), "
              f"OpenMP {cores} threads; the Julia reference itself cannot run here")
    workload = f"{args.n}x{args.n} Float64 LU with partial pivoting"
    if not same:
        workload += f" -- CPU arm timed on a {n_s}x{n_s} SAMPLE of it (rate, not time, is comparable)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sample_n": n_s, "same_config": same},
        "same_config": same,
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def other_configs(ctx, rfb200, skip_big=False):
    """Device-resident timings (CUDA events, best of 3 after a warm-up, matrix restored by an untimed copy) of the
    other single-GPU BASELINE.json configs and of the rows SURVEY.md section 8f widens into.  Informational: the
    headline `value` / `e2e` above are the 16384 x 16384 Float64 pivoted LU."""
    import ctypes as C
    out = {}

    def timed(fn, pre, reps=3):
        best = None
        for i in range(reps + 1):
            pre()
            ctx.timer_start(); fn(); t = ctx.timer_stop()
            if i:
                best = t if best is None else min(best, t)
        return best

    def lu_case(n, dtype, check=None, residual=True, profile=False, **opt):
        a = np.empty((n, n), dtype=dtype, order="F")
        fill_random(a)
        if opt.get("no_pivot"):
            a[np.arange(n), np.arange(n)] += n / 4       # diagonally dominant: safe without pivoting
        src = rfb200.DeviceMatrix(ctx, n, n, dtype, lda=n); src.upload(a); ctx.sync()
        dst = rfb200.DeviceMatrix(ctx, n, n, dtype, lda=n)
        ms = timed(lambda: dst.lu(**opt), lambda: dst.copy_from(src))
        prof = None
        if profile:                          # separate, untimed pass with events around every launch
            dst.copy_from(src)
            ctx.profile_enable(True)
            dst.lu(**opt)
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            dst.copy_from(src); dst.lu(**opt)
        if not residual:                      # 8.6 GB matrices: no host-side probe (the multi-GPU run reports one at this size)
            info_h = np.zeros(1, dtype=np.int64)
            ctx.d2h(info_h, dst.info_ptr); ctx.sync()
            src.free(); dst.free()
            return {"ms": round(ms, 3), "gflops": round(lu_flops(n) / ms / 1e6, 1), "info": int(info_h[0])}
        f, ipiv, info = dst.download()
        res = hutchinson_residual(a.astype(np.float64, copy=False), f.astype(np.float64, copy=False),
                                  np.arange(1, n + 1) if opt.get("no_pivot") else ipiv, nvec=4)
        src.free(); dst.free()
        out = {"ms": round(ms, 3), "gflops": round(lu_flops(n) / ms / 1e6, 1), "info": info, "residual_fro_rel_est": res,
               "bound_20_n_eps": 20 * n * float(np.finfo(dtype).eps)}
        if prof is not None:
            out["profile"] = prof
        if check and not opt.get("no_pivot"):
            from scipy.linalg import lapack
            getrf = lapack.dgetrf if dtype == np.float64 else lapack.sgetrf
            _, piv, _ = getrf(a.copy(order="F"), overwrite_a=True)
            out["pivots_equal_lapack"] = bool(np.array_equal(ipiv, piv + 1))
            if not out["pivots_equal_lapack"]:
                # Float32: rounding noise of different summation orders reaches the gap between the two largest
                # candidates of a column; report where, and whether that step is a near-tie
                k = int(np.argmax(ipiv != piv + 1))
                out["first_pivot_mismatch"] = k
                out["mismatch_is_near_tie"] = bool(near_tie(a, f, ipiv, piv + 1, k))
            if check == "oracle":
                from oracle import rf_oracle as O
                _, want_p, _ = O.lu_c(a.copy(order="F"), threads=os.cpu_count() or 1)
                out["pivots_equal_oracle"] = bool(np.array_equal(ipiv, want_p))
        return out

    out["4096x4096 Float64 LU with partial pivoting (BASELINE config 2)"] = lu_case(4096, np.float64, check="oracle")
    c5 = lu_case(8192, np.float32, check="lapack", profile=True)
    tf32_peak = max(ctx.tf32_peak_tflops(4000), ctx.tf32_peak_tflops(4000))
    g = c5.pop("profile")["gemm"]
    alg_tf = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    c5["roofline"] = {
        "bound": "tensor", "kernel": "K4' trailing GEMM, tcgen05.mma kind::tf32 (3xTF32 split, FP32 accumulate in TMEM), TMA-fed",
        "achieved": alg_tf, "peak": tf32_peak / 3.0, "unit": "TFLOP/s (FP32-accurate: 3 TF32 MMAs per product)",
        "frac": (3.0 * alg_tf / tf32_peak) if tf32_peak else None, "traffic": None,
        "tf32_dense_peak_measured": tf32_peak,
        "peak_source": "own smem-resident tcgen05.mma kind::tf32 microbenchmark (rfb_bench_tf32_peak), nominal 1.1 PFLOP/s dense",
        "achieved_def": "sum of 2mnk over the GEMM launches of one 8192^2 Float32 LU / sum of their CUDA-event durations",
        "gemm_ms_of_step": round(g["ms"], 3)}
    out["8192x8192 Float32 LU, tcgen05 kind::tf32 3xTF32 trailing update (BASELINE config 5; the default from 4096 columns up)"] = c5
    out["8192x8192 Float32 LU, exact FP32 FFMA trailing update (f32_mode = RFB_F32_FP32)"] = \
        lu_case(8192, np.float32, check="lapack", f32_mode=2)
    out["16384x16384 Float64 LU, pivot = Val(false) (src/lu.jl:27-65)"] = lu_case(16384, np.float64, no_pivot=1)
    if not skip_big:
        out["32768x32768 Float64 LU with partial pivoting on ONE GPU (the N = 1 point of the multi-GPU strong-scaling series, "
            "BASELINE config 4)"] = lu_case(32768, np.float64, residual=False)
    # butterfly transform: algorithmic bytes 2 * 8 * n^2
    n = 16384
    d = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    ctx.memset(d.ptr, 0, d.nbytes)
    uv = rfb200.butterfly_generate_random(n)
    duv = ctx.malloc(uv.nbytes); ctx.h2d(duv, uv); ctx.sync()
    ms = timed(lambda: ctx._check(ctx._lib.rfb_butterfly_mul_f64(ctx.handle, C.c_void_p(d.ptr), n, n, C.c_void_p(duv))), lambda: None, reps=5)
    out["16384x16384 Float64 butterfly transform U'AV (src/butterflylu.jl:93-113)"] = {
        "ms": round(ms, 4), "GBps": round(16.0 * n * n / ms / 1e6, 1), "algorithmic_bytes": 16 * n * n}
    ctx.free(duv); d.free()
    # batched small LU: 16384 matrices of 32 x 32 Float64, one launch
    batch, m = 16384, 32
    a = np.random.default_rng(12).random((batch, m, m))
    d0, d1 = ctx.malloc(a.nbytes), ctx.malloc(a.nbytes)
    ctx.h2d(d0, a); ctx.sync()
    piv, info = ctx.malloc(batch * m * 8), ctx.malloc(batch * 8)
    opts = rfb200._make_opts(rfb200._lib.RFB_MEM_DEVICE)
    ms = timed(lambda: ctx._check(ctx._lib.rfb_lu_batched_f64(ctx.handle, C.c_void_p(d1), m, m, m, m * m, batch, C.c_void_p(piv),
                                                              C.c_void_p(info), C.byref(opts))),
               lambda: ctx.d2d(d1, d0, a.nbytes))
    out["16384 x (32x32 Float64) batched LU with partial pivoting, one launch"] = {
        "ms": round(ms, 4), "matrices_per_s": round(batch / ms * 1e3), "gflops": round(batch * (2.0 * m ** 3 / 3) / ms / 1e6, 1)}
    for p in (d0, d1, piv, info):
        ctx.free(p)
    return out


def run_ours(args, rank, world, local_rank):
    import rfb200

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    ctx = rfb200.Context(local_rank)
    dev = ctx.device_info()
    host = ctx.pinned_empty((n, n), np.float64)
    fill_random(host)
    pristine = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    work = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
    pristine.upload(host)
    ctx.sync()

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)     # started before the warm-up: nvidia-smi needs ~0.5 s to come up
    sampler.start()
    time.sleep(0.5)
    for _ in range(args.warmup):
        work.copy_from(pristine)
        work.lu()
    ctx.sync()
    barrier()
    launches0 = ctx.launch_count()
    times = []
    for _ in range(args.steps):
        work.copy_from(pristine)           # restore (LU is in place); outside the timed region
        ctx.timer_start()
        work.lu()
        times.append(ctx.timer_stop())
    ctx.sync()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    ms = max_over_ranks(sum(times) / len(times))
    value = world * lu_flops(n) / (ms * 1e-3) / 1e9

    # FP64 tensor peak: register-only DMMA microbenchmark, taken right after the timed region (clocks already up;
    # a cold first call has read 20 % low) as the best of two calls of 3 launches each
    dmma_meas = max(ctx.dmma_peak_tflops(40000) for _ in range(3))

    # ---- correctness of what was just timed ---------------------------------------------------
    f, ipiv, info = work.download()
    res = hutchinson_residual(host, f, ipiv)
    lmax = float(np.abs(np.tril(f, -1)).max())
    checks = {"info": info, "residual_fro_rel_est": res, "bound_20_n_eps": 20 * n * float(np.finfo(np.float64).eps),
              "max_abs_L": lmax}
    if args.check_pivots:                       # north_star: "pivot indices bit-exact" -- at the headline size, every run
        from scipy.linalg import lapack
        a_l = np.array(host, order="F", copy=True)
        t_l = time.perf_counter()
        _, piv, _ = lapack.dgetrf(a_l, overwrite_a=True)
        t_l = time.perf_counter() - t_l
        del a_l
        checks["pivots_equal_lapack"] = bool(np.array_equal(ipiv, piv + 1))
        # the stronger CPU sanity baseline, timed on the same matrix on this box's host cores
        checks["lapack_dgetrf_gflops"] = lu_flops(n) / t_l / 1e9
        checks["lapack_dgetrf_cores"] = os.cpu_count() or 1
        if not checks["pivots_equal_lapack"]:
            checks["first_pivot_mismatch"] = int(np.argmax(ipiv != piv + 1))
        del piv
    del f

    # ---- per-kernel-class profile (events around every launch; separate, untimed pass) ---------
    work.copy_from(pristine)
    ctx.profile_enable(True)
    work.lu()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    gemm = prof["gemm"]
    gemm_tf = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # Denominator: the LARGER of the microbenchmark and the pipe's rate at the maximum SM clock (128 DMMA flop per clock
    # per SM).  The register-only microbenchmark draws more power than any real kernel and has read exactly 0.80 of its
    # usual 37.15 TFLOP/s on some boxes of this pool (round-2 runs 1 and 6: 29.7, BELOW the GEMM's own 30.5 -- not a
    # peak); taking the maximum keeps `frac` honest (never above what the hardware can do) on such a box.
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    dmma_nominal = dev["sm_count"] * 128 * sm_max * 1e6 / 1e12
    dmma_peak = max(dmma_meas, dmma_nominal)
    roofline = {
        "bound": "tensor", "kernel": "K4 trailing GEMM (FP64 DMMA mma.sync m8n8k4; tcgen05 has no f64 kind)",
        "achieved": gemm_tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": gemm_tf / dmma_peak if dmma_peak else None,
        "traffic": traffic,
        "peak_measured": dmma_meas, "peak_nominal_at_max_clock": dmma_nominal,
        "peak_source": "max(own register-only DMMA microbenchmark run in this process, sm_count x 128 flop/clk x max SM clock); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 37 TFLOP/s",
        "achieved_def": "sum of 2mnk over all GEMM launches of one LU / sum of their CUDA-event durations",
        "share_of_step_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "launches_by_class": {k: v["launches"] for k, v in prof.items()},
    }

    # ---- end to end through the public host API (pinned host buffers) --------------------------
    e2e = None
    if not args.skip_e2e:
        hwork = ctx.pinned_empty((n, n), np.float64)
        ipiv_h = np.empty(n, dtype=np.int64)

        def e2e_ms(reps):
            et = []
            for it in range(1 + reps):
                np.copyto(hwork, host)         # restore the in-place input (untimed)
                barrier()
                t = time.perf_counter()
                rfb200.lu_(hwork, ipiv_h, ctx=ctx)
                dt = time.perf_counter() - t
                if it > 0:
                    et.append(dt)
            return max_over_ranks(1e3 * sum(et) / len(et))

        # the early-download scheme before round 2's tiles (row bands at the right spine), same buffers, for comparison
        ctx.set_early_download(1)
        bands_ms = e2e_ms(2)
        ctx.set_early_download(2)
        e_ms = e2e_ms(min(args.steps, 3))
        # what was just timed must be a correct factorization too (this path uploads in column chunks, applies
        # the interchanges eagerly and downloads finished rows early: a different schedule from the device path)
        e_res = hutchinson_residual(host, np.asarray(hwork), ipiv_h)
        checks["e2e_residual_fro_rel_est"] = e_res
        checks["e2e_pivots_equal_device_path"] = bool(np.array_equal(ipiv_h, ipiv))
        # raw PCIe rates of the same pinned buffers (explains e2e - device time; not part of any timed region)
        t = time.perf_counter(); ctx.h2d(work.ptr, hwork); ctx.sync(); h2d_s = time.perf_counter() - t
        t = time.perf_counter(); ctx.d2h(hwork, work.ptr); ctx.sync(); d2h_s = time.perf_counter() - t
        e2e = {"value": world * lu_flops(n) / (e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": n * n * 8, "d2h_bytes_per_step": n * n * 8 + n * 8 + 8,
               "pcie_h2d_GBps": n * n * 8 / h2d_s / 1e9, "pcie_d2h_GBps": n * n * 8 / d2h_s / 1e9,
               "early_download": "tiles (rfb_set_early_download 2, the default)", "ms_per_step_row_bands": bands_ms}

    # ---- the other single-GPU BASELINE configs and the widened rows, device resident (not the headline) ----
    others = None
    if world == 1 and not args.skip_others:
        others = other_configs(ctx, rfb200, skip_big=args.skip_big)

    # ---- CPU baseline (rank 0, bounded sample) --------------------------------------------------
    cpu = None
    if rank == 0 and not args.skip_cpu_baseline:
        cores = os.cpu_count() or 1
        n_s = min(n, args.cpu_sample_n)
        gf, secs = cpu_baseline(n_s, cores)
        cpu = {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "seconds": secs,
               "lapack_dgetrf_gflops_same_matrix": checks.get("lapack_dgetrf_gflops"),
               "sample": f"{n_s}x{n_s} Float64 U[0,1) LU, oracle/rf_oracle.c (C restatement of src/lu.jl, reference "
                         f"defaults), OpenMP {cores} threads; the Julia reference cannot run here"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{n}x{n} Float64 LU with partial pivoting, 1 matrix per GPU",
                       "input": "U[0,1) numpy default_rng([12, block]) column-major, lda = n",
                       "parallelism": "replicas" if world > 1 else "1 GPU",
                       "l2": f"input {n * n * 8 / 1e6:.0f} MB > 126 MB L2; matrix restored by an untimed d2d copy between steps",
                       "device": dev},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "checks": checks, "dmma_peak_tflops": dmma_peak, "other_configs": others,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


def run_ours_dist(args, rank, world, local_rank):
    """N > 1: ONE matrix factored by all GPUs (1-D block-cyclic columns + NCCL panel broadcast, C++ driver behind
    rfb_mg_*).  torch.distributed (gloo, CPU tensors) is only the launcher-side plumbing: the 128-byte NCCL id, the
    barriers around the timed region and the max over ranks."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import rfb200
    from rfb200.dist_lu import DistributedLU, block_range

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        # NCCL prints its version banner on stdout at VERSION and WARN; stdout carries ONE JSON line, so (unless the
        # launcher asked for a specific debug level) warnings go to a per-process file instead
        os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/rfb200_nccl_%h_%p.log")
    dist.init_process_group("gloo")
    n = args.n if args.n else 32768
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    # block-column width: measured at 32768^2 (profiles/r02_bench_dist*): 8 GPUs are bound by the chain of per-block-column
    # critical sections (256: 186 ms, 512: 195 ms), 2-4 GPUs by their GEMM share (1024: 307 / 524 ms, 512: 315 / 533 ms)
    nb = args.block if args.block > 0 else (256 if world_env >= 8 else 1024)

    def exchange(mine):
        box = [mine]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    d = DistributedLU(n, np.float64, block=nb, rank=rank, world=world, device=local_rank, exchange_id=exchange)
    ctx = rfb200.Context(local_rank)           # page-locked host buffers, pristine device copies, the 1-GPU comparison run
    dev = ctx.device_info()

    def gen(j):
        c0, w = block_range(j, n, nb)
        return np.asfortranarray(np.random.default_rng([12, j]).random((n, w)))

    # pinned host copies of this rank's block columns (each is contiguous: lda = n) and pristine device copies
    host_in, host_out, pristine = {}, {}, {}
    for j in d.my_blocks:
        c0, w = block_range(j, n, nb)
        host_in[j] = ctx.pinned_empty((n, w), np.float64)
        host_in[j][...] = gen(j)
        host_out[j] = ctx.pinned_empty((n, w), np.float64)
        pristine[j] = ctx.malloc(n * w * 8)
        ctx.h2d(pristine[j], host_in[j])
    ctx.sync()

    def restore():
        for j in d.my_blocks:
            d.restore_block(j, pristine[j])
        d.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.5)
    for _ in range(args.warmup):
        restore()
        dist.barrier()
        d.factor()
        d.synchronize()
    dist.barrier()
    launches0 = d.stats()["launches"]
    times = []
    for _ in range(args.steps):
        restore()                  # outside the timed region (the factorization is in place)
        dist.barrier()
        d.factor()
        times.append(d.synchronize())      # CUDA events on the rank's compute stream around the whole schedule
    dist.barrier()
    clocks = sampler.stop()
    launches = d.stats()["launches"] - launches0
    ms = max_over_ranks(sum(times) / len(times))
    value = lu_flops(n) / (ms * 1e-3) / 1e9
    info = d.info()

    # ---- per-rank breakdown: host-scheduler statistics of the last timed step + one untimed pass with events around every launch ----
    sched = d.sched_stats()
    lib, hctx = d._lib, d.ctx_handle
    restore()
    dist.barrier()
    lib.rfb_profile_enable(hctx, 1)
    d.factor()
    prof_wall = d.synchronize()
    pms, pcnt, pwork = (C.c_double * 8)(), (C.c_int64 * 8)(), (C.c_double * 8)()
    lib.rfb_profile_read(hctx, pms, pcnt, pwork)
    lib.rfb_profile_enable(hctx, 0)
    names = ["panel", "laswp", "trsm_diag", "gemm", "other"]
    mine = {"rank": rank, "sched": sched, "profiled_pass_ms": prof_wall,
            "kernel_ms": {nm: round(pms[i], 2) for i, nm in enumerate(names)},
            "launches": {nm: int(pcnt[i]) for i, nm in enumerate(names)},
            "gemm_tflops": round(pwork[3] / (pms[3] * 1e-3) / 1e12, 2) if pms[3] > 0 else None}
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine)

    # ---- distributed residual probe ||(PA - LU) x|| / ||A|| with +-1 probes (O(n^2) per rank) -----
    ipiv = d.pivots()
    perm = np.arange(n)
    for i, ip in enumerate(ipiv):
        ip = int(ip) - 1
        if ip != i:
            perm[i], perm[ip] = perm[ip], perm[i]
    nvec = 4
    x = np.random.default_rng(0).integers(0, 2, size=(n, nvec)).astype(np.float64) * 2 - 1
    ux = np.zeros((n, nvec)); pax = np.zeros((n, nvec)); nrm2 = 0.0; lmax = 0.0
    fac = {}
    for j in d.my_blocks:
        c0, w = block_range(j, n, nb)
        d.store_block_async(j, host_out[j])
    d.synchronize()
    for j in d.my_blocks:
        c0, w = block_range(j, n, nb)
        f = np.asarray(host_out[j])
        rows = np.arange(n)[:, None]; cols = np.arange(c0, c0 + w)[None, :]
        ux += np.where(rows <= cols, f, 0.0) @ x[c0:c0 + w]
        a0 = np.asarray(host_in[j])
        pax += a0[perm, :] @ x[c0:c0 + w]
        nrm2 += float(np.sum(a0 * a0))
        lmax = max(lmax, float(np.abs(np.where(rows > cols, f, 0.0)).max()))
    ux = sum_over_ranks(ux)
    lux = np.zeros((n, nvec))
    for j in d.my_blocks:
        c0, w = block_range(j, n, nb)
        f = np.asarray(host_out[j])
        rows = np.arange(n)[:, None]; cols = np.arange(c0, c0 + w)[None, :]
        lf = np.where(rows > cols, f, 0.0)
        lf[np.arange(c0, c0 + w), np.arange(w)] = 1.0
        lux += lf @ ux[c0:c0 + w]
    t = sum_over_ranks(np.concatenate([(pax - lux).reshape(-1), [nrm2, 0.0]]))
    res = float(np.linalg.norm(t[:-2]) / np.sqrt(nvec) / np.sqrt(float(t[-2])))
    lmax = max_over_ranks(lmax)

    # ---- end to end: pinned host blocks -> GPUs -> factor -> pinned host blocks (uploads pipelined behind the first blocks) ----
    e2e = None
    if not args.skip_e2e:
        et = []
        for it in range(1 + min(args.steps, 2)):
            d.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for j in d.my_blocks:
                d.set_block(j, host_in[j])             # async H2D on the copy stream; the schedule waits per block column
            d.factor()
            for j in d.my_blocks:
                d.store_block_async(j, host_out[j])
            d.synchronize()
            dist.barrier()
            if it > 0:
                et.append(time.perf_counter() - t0)
        e_ms = max_over_ranks(1e3 * sum(et) / len(et))
        e2e = {"value": lu_flops(n) / (e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": n * n * 8, "d2h_bytes_per_step": n * n * 8}

    # ---- the same matrix on ONE GPU of this box (rank 0): pivot equality + the 1-GPU point of the strong-scaling series ----
    single = None
    if not args.skip_single and rank == 0:
        a_full = np.empty((n, n), dtype=np.float64, order="F")
        for j in range((n + nb - 1) // nb):
            c0, w = block_range(j, n, nb)
            a_full[:, c0:c0 + w] = host_in[j] if j in host_in else gen(j)
        src = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n); src.upload(a_full); ctx.sync()
        del a_full
        work = rfb200.DeviceMatrix(ctx, n, n, np.float64, lda=n)
        st = []
        for it in range(3):
            work.copy_from(src)
            ctx.timer_start(); work.lu(); tms = ctx.timer_stop()
            if it:
                st.append(tms)
        piv1 = np.empty(n, dtype=np.int64)
        ctx.d2h(piv1, work.ipiv_ptr); ctx.sync()
        s_ms = sum(st) / len(st)
        single = {"pivots_equal_single_gpu": bool(np.array_equal(piv1, ipiv)), "strong_n1_ms": s_ms,
                  "strong_n1_value": lu_flops(n) / (s_ms * 1e-3) / 1e9}
        src.free(); work.free()
    dist.barrier()

    if rank == 0:
        checks = {"info": info, "residual_fro_rel_est": res, "bound_20_n_eps": 20 * n * float(np.finfo(np.float64).eps),
                  "max_abs_L": lmax}
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{n}x{n} Float64 LU with partial pivoting, ONE matrix over {world} GPUs",
                       "input": "U[0,1) numpy default_rng([12, block]) per block column",
                       "parallelism": f"1-D block-cyclic columns (block {nb}), owner-rooted ncclBroadcast of each factored block "
                                      f"column + pivots + exchange lists on a dedicated stream, replicated L; C++ schedule (rfb_mg_*)",
                       "l2": f"matrix {n * n * 8 / 1e9:.1f} GB >> L2; restored from a device copy between steps (untimed)",
                       "device": dev},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "K4 trailing GEMM (FP64 DMMA)", "achieved": value / 1e3 / world,
                         "peak": None, "unit": "TFLOP/s per GPU (whole LU, not the kernel alone)", "frac": None, "traffic": None,
                         "note": "per-kernel roofline is reported by the 1-GPU run; here NCCL bytes per rank = "
                                 f"{d.stats()['bcast_bytes_per_rank'] / max(1, args.steps + args.warmup + (0 if args.skip_e2e else 1 + min(args.steps, 2))) / 1e9:.2f} GB per factorization"},
            "cpu_baseline": None, "checks": checks, "per_rank": per_rank,
        }
        if single is not None:
            checks["pivots_equal_single_gpu"] = single["pivots_equal_single_gpu"]
            line["strong_n1_value"] = single["strong_n1_value"]
            line["strong_n1_ms"] = single["strong_n1_ms"]
            line["strong_scaling_efficiency_vs_in_run_n1"] = value / single["strong_n1_value"] / world
        print(json.dumps(line))
    d.close()
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=0, help="matrix size (default 16384 on 1 GPU, 32768 distributed)")
    ap.add_argument("--block", type=int, default=0, help="block-column width of the multi-GPU distribution (0 = by GPU count)")
    ap.add_argument("--cpu-sample-n", type=int, default=8192)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-others", action="store_true", help="skip the informational other_configs block")
    ap.add_argument("--skip-big", action="store_true", help="1 GPU: skip the 32768^2 single-GPU entry of other_configs")
    ap.add_argument("--skip-single", action="store_true", help="multi-GPU: skip the in-run 1-GPU factorization of the same matrix")
    ap.add_argument("--check-pivots", dest="check_pivots", action="store_true", default=True,
                    help="compare the pivot vector of the timed factorization with LAPACK dgetrf (default: on)")
    ap.add_argument("--no-check-pivots", dest="check_pivots", action="store_false")
    ap.add_argument("--ref-budget-s", type=float, default=1500.0,
                    help="reference arm: wall-clock budget for all its factorizations; the arm runs the same size as the "
                         "product arm when that fits")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.n = args.n or (16384 if world == 1 else 32768)
        run_reference(args, rank, world)
    elif world > 1:
        run_ours_dist(args, rank, world, local_rank)
    else:
        args.n = args.n or 16384
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
