#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: Gauss-Seidel throughput (Gnnz/s) and fraction of the HBM roofline
on the 4096^2 x 3-channel Poisson system (configs[2]), 1/2/4/8 B200.

A "step" is one pass of the hot path over one batch of synthetic input: `--sweeps` Gauss-Seidel sweeps
(reference cadence: the stop rule is evaluated every sweep) over the reference-faithful full-grid
matrix with the three colour channels as three right-hand sides sharing one pass over the CSR.

  value     whole-job Gnnz/s = nnz * sweeps * channels * steps / time, inputs resident in HBM
  e2e       same metric through the reference-facing API with HOST buffers:
            initializeFromEigenRowMajor(host CSR) + gaussSeidel(host b) -> host x, copies timed
  roofline  algorithmic bytes (12*nnz + 4*n + 24*k*n per sweep) / device time of the sweep
            loop (CUDA events on the library stream, gsb_gs_stats.solve_ms) vs MEASURED_PEAKS.json
  cpu_baseline  oracle/_ref (the unmodified reference header) on the host cores, bounded sample

`--impl reference` times the reference's own CPU gaussSeidel on the same system (bounded sample).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--size", type=int, default=4096, help="image width = height")
    p.add_argument("--channels", type=int, default=3)
    p.add_argument("--sweeps", type=int, default=100, help="GS sweeps per step")
    p.add_argument("--check-every", type=int, default=1)
    p.add_argument("--kernel", type=int, default=0)
    p.add_argument("--e2e-steps", type=int, default=10)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--strips", action="store_true", help="N = 1 through the strip solver (measurement aid)")
    p.add_argument("--no-time-to-tol", action="store_true", help="skip the time-to-tolerance leg (masked blend)")
    p.add_argument("--time-to-tol-only", action="store_true", help=argparse.SUPPRESS)
    p.add_argument("--no-default-eps", action="store_true", help="time-to-tol: skip the epsilon = 1e-6 run")
    p.add_argument("--no-other-configs", action="store_true", help="skip the legs for BASELINE configs[0], [1], [4]")
    p.add_argument("--other-config-only", default=None, help=argparse.SUPPRESS)
    p.add_argument("--no-c4", action="store_true", help="N >= 2: skip the 16384^2 strip leg (BASELINE configs[3])")
    p.add_argument("--c4-size", type=int, default=16384)
    p.add_argument("--c5-n", type=int, default=10_000_000, help="rows of the random SPD system (configs[4])")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def e2e_warm_done(warm_ms, warm_min=3, warm_max=12, tol=0.1):
    """The e2e leg's untimed warm-up is over: at least `warm_min` steps, then as soon as two consecutive steps agree to
    `tol` (or after `warm_max` steps).  On the pool's shared boxes the first steps after the GB-sized page-locked
    allocations are erratic; the timed region starts once the step time has settled."""
    if len(warm_ms) < warm_min:
        return False
    return len(warm_ms) >= warm_max or abs(warm_ms[-1] - warm_ms[-2]) <= tol * min(warm_ms[-2:])


def pinned_like(pkg, a):
    """Copy of numpy array `a` in page-locked host memory (gsb_host_alloc): the e2e leg copies from / to it."""
    p = C.c_void_p()
    pkg._lib.check(pkg.load().gsb_host_alloc(C.byref(p), max(a.nbytes, 1)), "gsb_host_alloc")
    buf = (C.c_char * max(a.nbytes, 1)).from_address(p.value)
    out = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
    out[...] = a
    return out


def src_sha16(files):
    """Hash of the source files that define a kernel (names relative to coursecomputationalphotography_b200/csrc):
    what an ncu capture recorded in profiles/traffic.json is valid for.  (The .so itself is no use as a key: nvcc /
    the linker do not produce the same bytes twice from the same sources.)"""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "coursecomputationalphotography_b200", "csrc")
    try:
        for f in sorted(files):
            h.update(f.encode())
            h.update(open(os.path.join(csrc, f), "rb").read())
    except OSError:
        return None
    return h.hexdigest()[:16]


def recorded_traffic(kernel_used, W, H, ch, check_every):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full capture
    recorded in profiles/traffic.json (one entry per kernel / shape).  It is a recorded number, not a measurement of
    this run: `traffic_source` names the capture and `traffic_stale` says when the kernel's source files have changed since."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        db = json.load(open(path))
    except Exception:
        return {}
    key = "k%d_%dx%d_rhs%d_ce%d" % (kernel_used, W, H, ch, check_every)
    ent = db.get("captures", {}).get(key)
    if not ent:
        return {"traffic_source": "no ncu capture recorded for %s" % key}
    out = {"traffic": ent["dram_bytes_per_launch"], "traffic_source": ent.get("source")}
    if ent.get("src_files"):
        sha = src_sha16(ent["src_files"])
        if ent.get("src_sha16") and sha and ent["src_sha16"] != sha:
            out["traffic_stale"] = "captured when %s hashed to %s, now %s" % (ent["src_files"], ent["src_sha16"], sha)
    return out


def workload_config(W, H, ch, n, nnz):
    """The workload-defining keys, identical for the GPU arm, the strip arm and the reference arm (what each arm did
    with the workload -- sweeps per step, ordering, cadence -- goes into "run")."""
    return {"workload": "poisson_%dx%d_x%dch_full_grid (BASELINE configs[2])" % (W, H, ch), "n": int(n), "nnz": int(nnz),
            "channels": int(ch),
            "l2": "working set %.2f GB per sweep >> 126 MB L2 (no flush needed)" % (algorithmic_bytes_per_sweep(nnz, n, ch) / 1e9)}


def algorithmic_bytes_per_sweep(nnz, n, k):
    return 12.0 * nnz + 4.0 * n + 24.0 * k * n


def synth_rhs(pkg, wl, W, H, ch):
    """Gradients of a synthetic two-exposure image -> A^T b on the device path, host result (ch, n)."""
    img = wl.synth_image(W, H, ch, seed=7)
    gx, gy = wl.seamless_gradients(img)
    b = pkg.poisson_rhs(W, H, gx, gy, img[:, 0, 0].astype(np.float64))
    return np.ascontiguousarray(b.reshape(ch, W * H))


# --------------------------------------------------------------------------------------------------
# CPU reference arm (also the cpu_baseline leg of the GPU arm)
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(W, H, ch, sweeps, steps, warmup, b_host=None):
    """The reference's own gaussSeidel (oracle/_ref, else the oracle port) on the W x H Poisson system:
    `ch` channels run concurrently on `ch` host threads (each sweep is serial by construction,
    v2 :359-374).  Returns (Gnnz/s, ms_per_step, cores, kind, nnz)."""
    from oracle import pyoracle
    pyoracle.build()
    from coursecomputationalphotography_b200 import workloads as wl
    ro, ci, va = pyoracle.poisson_csr(W, H)
    n, nnz = W * H, int(ro[-1])
    kind = "reference" if pyoracle.ref_available() else "port"
    if b_host is None:
        img = wl.synth_image(W, H, ch, seed=7)
        gx, gy = wl.seamless_gradients(img)
        b_host = np.stack([pyoracle.poisson_rhs(W, H, gx[c], gy[c], float(img[c, 0, 0])) for c in range(ch)])
    if kind == "reference":
        mats = [pyoracle.Ref(2, "f64").import_csr(va, ro[:-1], ci, n) for _ in range(1)]
        solve = lambda c: mats[0].gauss_seidel(b_host[c], 0.0, sweeps)
    else:
        mats = [pyoracle.Oracle().import_csr(va, ro[:-1], ci, n)]
        solve = lambda c: mats[0].gauss_seidel(b_host[c], 0.0, sweeps)
    cores = min(ch, os.cpu_count() or 1)

    def step():
        if cores <= 1:
            for c in range(ch):
                solve(c)
            return
        ths = [threading.Thread(target=solve, args=(c,)) for c in range(ch)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return nnz * sweeps * ch * steps / dt / 1e9, dt / steps * 1e3, cores, kind, nnz


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = H = args.size
    sweeps = 2  # bounded sample: 2 sweeps x channels per step (~0.3 s per sweep per channel at 4096^2)
    warmup = max(args.warmup, 3)  # the same clamp as the GPU arm
    val, ms, cores, kind, nnz = cpu_reference_run(W, H, args.channels, sweeps, args.steps, warmup)
    line = {
        "impl": "reference", "metric": "gauss_seidel_throughput", "value": val, "unit": "Gnnz/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(W, H, args.channels, W * H, nnz),
        "run": {"sweeps_per_step": sweeps, "ordering": "natural (the reference's lexicographic sweep)",
                "check_every": 1, "sweeps_per_s": sweeps * args.steps / (ms * args.steps * 1e-3)},
        "cpu_baseline": {"value": val, "unit": "Gnnz/s", "cores": cores, "kind": kind,
                         "sample": "%d sweeps x %d channels per step (one channel per thread), %d steps, %dx%d" %
                                   (sweeps, args.channels, args.steps, W, H)},
        "e2e": {"value": val, "unit": "Gnnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import coursecomputationalphotography_b200 as pkg
    from coursecomputationalphotography_b200 import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    L = pkg.load()
    pkg._lib.check(L.gsb_set_device(local), "gsb_set_device")
    if world > 1 or args.strips:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29555")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        import datetime
        # a short collective timeout: a rank that dies must not hold a multi-GPU box for ten minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        from coursecomputationalphotography_b200 import dist_bench
        return dist_bench.run(args, pkg, wl, dist, rank, world, local)

    W = H = args.size
    ch = args.channels
    n = W * H
    stream = torch.cuda.ExternalStream(L.gsb_stream())
    sp = pkg.SparseMatrix(np.float64)
    sp.poisson(W, H)
    nnz = sp._nnz
    b_host = synth_rhs(pkg, wl, W, H, ch)
    b_dev = torch.from_numpy(b_host).cuda()
    x_dev = torch.empty_like(b_dev)
    opts = pkg.SparseMatrix.options(check_every=args.check_every, kernel=args.kernel)
    info = sp.analyze()
    torch.cuda.synchronize()

    def step():
        return sp.gaussSeidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), ch, 0.0, args.sweeps, opts)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, solve_ms, sweeps_done = 0, 0.0, 0
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        st = step()
        launches += st.kernel_launches
        solve_ms += st.solve_ms
        sweeps_done += st.sweeps
    e1.record(stream)
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    assert sweeps_done == args.sweeps * args.steps, "stop rule fired early: %d sweeps" % sweeps_done
    value = nnz * ch * sweeps_done / (total_ms * 1e-3) / 1e9
    # residual after the last step (sanity: the sweeps did real work)
    x_host = x_dev.cpu().numpy()
    resid = [sp.residual(b_host[c], x_host[c]) for c in range(ch)]

    peak, peak_src = peaks()
    abytes = algorithmic_bytes_per_sweep(nnz, n, ch)
    # dominant kernel: kernel 5 runs a whole sweep (both colours) per launch, kernels 1-4 one colour phase per launch
    kernel_used = int(st.kernel_used)
    launches_per_sweep = 1 if kernel_used == 5 else info["n_colors"]
    kname = ("gs_sweep_fused<%d,%s,2> (one sweep, both colours, %d RHS)" % (ch, "true" if args.check_every == 1 else "false", ch)
             if kernel_used == 5 else "gs_phase (one colour phase, %d RHS, kernel %d)" % (ch, kernel_used))
    per_launch_ms = solve_ms / (sweeps_done * launches_per_sweep)
    achieved = (abytes / launches_per_sweep) / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kname, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes / launches_per_sweep, "avg_launch_ms": per_launch_ms,
                "avg_launch_ms_includes": "the sweep loop's gs_end_sweep launches (library CUDA events / launches)",
                "frac_of_8TBps_nominal": achieved / 8000.0}
    roofline.update(recorded_traffic(kernel_used, W, H, ch, args.check_every))

    # ---- e2e: reference-facing API, host buffers, copies inside the timed region -------------------
    e2e = None
    if not args.no_e2e:
        # host CSR as Eigen would hand it over (valuePtr / outerIndexPtr / innerIndexPtr): taken from the
        # device-built matrix, untimed (it stands for the reference's Eigen A^T*A step)
        va, ci, _, rn, _ = sp.layout()
        ro_in = np.zeros(n, np.int32)
        ro_in[1:] = np.cumsum(rn[:-1])
        va, ci, ro_in, b_pin = pinned_like(pkg, va), pinned_like(pkg, ci), pinned_like(pkg, ro_in), pinned_like(pkg, b_host)
        x_pin = pinned_like(pkg, np.zeros_like(b_host))
        spe = pkg.SparseMatrix(np.float64)
        t_e2e, t_imp, t_setup, steps_e2e = 0.0, 0.0, 0.0, max(1, args.e2e_steps)
        per_step, per_step_parts = [], []
        st_e = pkg.GsStats()
        # untimed warm-up: at least 3 steps (the first imports into a fresh handle allocate GB-sized buffers), then --
        # on a shared box the first steps after the pinned allocations are erratic -- until two consecutive steps
        # agree to 10 %, at most 12
        warm_min, warm_ms = max(args.warmup, 3), []
        def dev_allocs():
            a, f = C.c_int64(0), C.c_int64(0)
            L.gsb_alloc_counters(C.byref(a), C.byref(f))
            return a.value + f.value
        allocs_timed = 0
        while True:
            timed = e2e_warm_done(warm_ms, warm_min)
            if timed and len(per_step) >= steps_e2e:
                break
            torch.cuda.synchronize()
            a0 = dev_allocs()
            t0 = time.perf_counter()
            spe.initializeFromEigenRowMajor(va, len(va), ro_in, n, ci, n)
            t1 = time.perf_counter()
            pkg._lib.check(L.gsb_gauss_seidel(spe._h, pkg._lib.ptr(b_pin), ch, 0.0, args.sweeps, C.byref(opts),
                                              pkg._lib.ptr(x_pin), C.byref(st_e)), "gsb_gauss_seidel")
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if not timed and not per_step:
                warm_ms.append((t2 - t0) * 1e3)
            else:
                allocs_timed += dev_allocs() - a0
                t_e2e += t2 - t0
                t_imp += t1 - t0
                t_setup += st_e.setup_ms
                per_step.append((t2 - t0) * 1e3)
                per_step_parts.append([round((t1 - t0) * 1e3, 2), round(float(st_e.setup_ms), 2), round(float(st_e.solve_ms), 2)])
        h2d = va.nbytes + ci.nbytes + ro_in.nbytes + b_pin.nbytes
        d2h = x_pin.nbytes
        e2e = {"value": nnz * ch * args.sweeps * steps_e2e / t_e2e / 1e9, "unit": "Gnnz/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e / steps_e2e * 1e3,
               "steps": steps_e2e, "warmup": len(warm_ms), "warmup_ms": [round(v, 2) for v in warm_ms], "ms_per_step_min": min(per_step), "ms_per_step_median": float(np.median(per_step)),
               "per_step_ms": [round(v, 2) for v in per_step], "per_step_import_analysis_sweeps_ms": per_step_parts,
               "includes": "CSR import (H2D) + ordering analysis + b H2D + sweeps + x D2H",
               "import_ms": t_imp / steps_e2e * 1e3, "analysis_ms": t_setup / steps_e2e,
               "host_memory": "pinned (gsb_host_alloc)",
               "device_allocs_and_frees_per_step": allocs_timed / steps_e2e}
        assert np.array_equal(x_pin, x_host), "e2e result differs from the resident-input result"

    cpu = None
    if not args.no_cpu_baseline:
        sw_cpu = 3
        v, ms, cores, kind, _ = cpu_reference_run(W, H, ch, sw_cpu, 1, 0, b_host)
        cpu = {"value": v, "unit": "Gnnz/s", "cores": cores, "kind": kind,
               "sample": "%d sweeps x %d channels (one channel per thread), %dx%d, %.1f s" % (sw_cpu, ch, W, H, ms / 1e3)}

    line = {
        "metric": "gauss_seidel_throughput", "value": value, "unit": "Gnnz/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(W, H, ch, n, nnz),
        "run": {"sweeps_per_step": args.sweeps, "ordering": "red-black", "n_colors": info["n_colors"],
                "check_every": args.check_every, "kernel": kernel_used,
                "sweeps_per_s": sweeps_done / (total_ms * 1e-3), "residual_l2": resid},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_time_to_tol:
        try:  # nothing in this leg may cost the main line
            # release this process' device memory first: the child builds its own matrices
            del sp, b_dev, x_dev
            torch.cuda.empty_cache()
            line["time_to_tol"] = time_to_tol_child(args)
        except Exception as e:
            line["time_to_tol"] = {"error": repr(e)[:300]}
    if not args.no_other_configs:
        line["other_configs"] = {"c1_lab3_n1e4": other_config_child(args, "c1", 120),
                                 "c2_poisson_1024": other_config_child(args, "c2", 120),
                                 "c5_random_spd": other_config_child(args, "c5", 420),
                                 "cg_reference_call": other_config_child(args, "cg", 180)}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# time-to-tolerance leg (BASELINE metric: "...; time-to-tol"): SURVEY 8d C3 -- the Dirichlet-masked 5-point
# blend with an irregular mask (random blobs, inscribed thickness <= 48 px, 30 % of the frame), three channels,
# solved from x0 = 1 until the reference's stop rule fires (L1 norm of a sweep's update <= epsilon, v2 :350-376).
# On the full-grid Neumann + pin system plain GS does not converge in any useful time (SURVEY finding 3), so this
# is the system time-to-tolerance is defined on.  Checked against an independent
# solver on the device (conjugate gradients on the same system): max-abs difference vs 1e-4 * 255.
# Runs in a child process so that nothing it does can take the main bench line down.
# --------------------------------------------------------------------------------------------------
def run_time_to_tol(args):
    import coursecomputationalphotography_b200 as pkg
    from coursecomputationalphotography_b200 import workloads as wl
    W = H = args.size
    ch = args.channels
    t0 = time.perf_counter()
    ro, ci, va, b, pix, colors = wl.c3_masked_system(W, ch)
    n, nnz = len(pix), len(va)
    t_gen = time.perf_counter() - t0
    # golden: the REFERENCE's own gaussSeidel (oracle/_ref) run to its stop rule on this very system, generated once
    # by tests/golden/make_golden_c3.py (tens of CPU-minutes); strided sample of x, sweeps, residual, CPU seconds
    gold, gmeta = None, None
    gpath = os.path.join(ROOT, "tests", "golden", "c3_masked_%d" % W)
    if ch == 3 and os.path.exists(gpath + ".npz") and os.path.exists(gpath + ".json"):
        gmeta = json.load(open(gpath + ".json"))
        if gmeta["n"] == n and gmeta["nnz"] == nnz:
            gold = np.load(gpath + ".npz")
    sm = pkg.SparseMatrix(np.float64)
    sm.initializeFromEigenRowMajor(va, len(va), ro[:-1], n, ci, n)
    sm.analyze(pkg._lib.ORDER_USER, colors)
    opts = pkg.SparseMatrix.options(check_every=args.check_every, kernel=args.kernel)
    sm.gaussSeidel(b, epsilon=0.0, max_iteration=20, options=opts)  # warm-up: plan, workspaces
    tol = 1e-4 * 255.0

    def one(eps, cap):
        t1 = time.perf_counter()
        x = sm.gaussSeidel(b, epsilon=eps, max_iteration=cap, options=opts)
        wall_ms = (time.perf_counter() - t1) * 1e3
        st = sm.last_stats
        sweeps, solve_ms = int(st.sweeps), float(st.solve_ms)
        out = {"epsilon": eps, "max_iteration": cap, "stopped": bool(sweeps < cap), "sweeps": sweeps, "ms": solve_ms,
               "wall_ms_with_copies": wall_ms,
               "Gnnz_per_s": nnz * ch * sweeps / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else None,
               "last_eps": [float(v) for v in list(st.last_eps)[:ch]],
               "residual_l2": [float(sm.residual(b[c], x[c])) for c in range(ch)],
               "kernel": int(st.kernel_used), "n_colors": int(st.n_colors)}
        if gold is not None:
            runs = [r for r in gmeta["runs"] if r["epsilon"] == eps]
            if runs:
                idx = gold["index"]
                diffs, u8_equal = [], True
                for c in range(ch):
                    ref = gold["x_eps%g_ch%d" % (eps, c)]
                    mine = x[c][idx]
                    diffs.append(float(np.abs(mine - ref).max()))
                    # write-back as the reference does it: uchar(clamp(x, 0, 255)), truncation (PhotoMontage.cpp:617-626)
                    u8 = lambda v: np.clip(v, 0.0, 255.0).astype(np.uint8)
                    u8_equal = u8_equal and bool(np.array_equal(u8(mine), u8(ref)))
                out.update({"max_abs_vs_reference": max(diffs), "tolerance": tol,
                            "within_tolerance": bool(max(diffs) <= tol), "u8_equal_on_sample": u8_equal,
                            "reference_sample_points": int(len(idx)) * ch,
                            "reference": {"what": gmeta["source"], "sweeps": [r["sweeps"] for r in runs],
                                          "stopped": [r["stopped"] for r in runs],
                                          "residual_l2": [r["residual_l2"] for r in runs],
                                          "cpu_s": max(r["cpu_s"] for r in runs), "cpu_threads": gmeta["host_threads"],
                                          "oracle_bitwise": all(r["oracle_bitwise"] for r in runs)}})
                out["speedup_vs_reference_cpu"] = out["reference"]["cpu_s"] / (wall_ms * 1e-3)
        return out, x

    # epsilon: the reference's rule is an absolute L1 norm (v2 :376) whose rounding floor grows with n; the default
    # 1e-6 sits near that floor at 5 M unknowns, so the headline leg uses 1e-5 (an average update of 2e-12 per unknown)
    # and the default is reported beside it
    cap = int(gmeta["cap"]) if gmeta else 50000
    main, x = one(1e-5, cap)
    if gold is None:
        # no golden for this size: independent check against conjugate gradients on the same (SPD, Dirichlet) system
        diffs, iters = [], []
        for c in range(ch):
            xc = sm.conjugateGradient(b[c], 1e-6, 3000)
            iters.append(int(sm.last_iters))
            diffs.append(float(np.abs(xc - x[c]).max()))
        main.update({"max_abs_vs_cg": max(diffs), "tolerance": tol, "within_tolerance": bool(max(diffs) <= tol),
                     "cg_iterations": iters, "note": "no reference golden for this size (tests/golden/make_golden_c3.py)"})
    out = {"workload": "dirichlet_masked_blend_%dx%d_x%dch, blob mask 30 %%, thickness <= 48 px (SURVEY 8d C3)" % (W, H, ch),
           "n": int(n), "nnz": int(nnz), "x0": 1.0, "host_generation_s": t_gen}
    out.update(main)
    if gold is not None and not args.no_default_eps:
        out["default_epsilon_1e-6"], _ = one(1e-6, cap)
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------
# The other BASELINE configurations, as extra keys on the N = 1 line (each in a child process, bounded):
#   configs[0] C1  lab3 diagonally dominant system, n = 1e4: microseconds per solve (reference defaults) next to the
#                  compiled reference on this host, no roofline claim (launch-latency bound)
#   configs[1] C2  1024^2 single channel: throughput; the 92 MB working set is L2-resident, so the bytes/s figure is
#                  an EFFECTIVE one (stated), not an HBM claim
#   configs[4] C5  random sparse SPD, n = 1e7, ~27 entries per row: unsorted triplets -> device radix sort
#                  (initializeFromTriplets) -> Jones-Plassmann multicolour -> sweeps; Gnnz/s, algorithmic GB/s
#   (configs[3] C4, 16384^2 row strips, is reported by the N >= 2 arm: dist_bench.py)
# --------------------------------------------------------------------------------------------------
def run_other_config(args):
    import coursecomputationalphotography_b200 as pkg
    from coursecomputationalphotography_b200 import workloads as wl
    which = args.other_config_only
    peak, _ = peaks()
    if which == "c1":
        r, c, v, b, xstar = wl.diag_dominant_system(10_000, 4, seed=42)
        sp = pkg.SparseMatrix(np.float64)
        sp.initializeFromVector(r, c, v)
        sp.gaussSeidel(b)  # analysis, plan, graph capture
        reps, t = 20, []
        for _ in range(reps):
            t0 = time.perf_counter()
            x = sp.gaussSeidel(b)
            t.append((time.perf_counter() - t0) * 1e6)
        st = sp.last_stats
        out = {"workload": "lab3 diag-dominant, n = 10000, 4 off-diagonals per row (BASELINE configs[0])",
               "sweeps": int(st.sweeps), "n_colors": int(st.n_colors), "last_eps": float(st.last_eps[0]),
               "us_per_solve_wall_median": float(np.median(t)), "us_per_solve_wall_min": float(min(t)),
               "us_device_sweep_loop": float(st.solve_ms) * 1e3, "includes": "b H2D, the solve in one persistent launch (kernel 6), x D2H",
               "max_abs_vs_xstar": float(np.abs(x - xstar).max()), "roofline": "none claimed: latency bound (a chain of dependent colour steps)"}
        try:
            from oracle import pyoracle
            if pyoracle.ref_available():
                ref = pyoracle.Ref(2, "f64").init_from_vector(r, c, v)
                tr = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    xr = ref.gauss_seidel(b)
                    tr.append((time.perf_counter() - t0) * 1e6)
                out["reference_cpu_us_per_solve"] = float(np.median(tr))
                out["max_abs_vs_reference"] = float(np.abs(x - xr).max())
        except Exception as e:  # the CPU arm is a courtesy here
            out["reference_cpu_error"] = repr(e)[:200]
    elif which == "c2":
        W = H = 1024
        sp = pkg.SparseMatrix(np.float64)
        sp.poisson(W, H)
        b = synth_rhs(pkg, wl, W, H, 1)[0]
        sweeps = 2000
        sp.gaussSeidel(b, epsilon=0.0, max_iteration=50)
        sp.gaussSeidel(b, epsilon=0.0, max_iteration=sweeps)
        st = sp.last_stats
        ab = algorithmic_bytes_per_sweep(sp._nnz, W * H, 1)
        gbs = ab * st.sweeps / (st.solve_ms * 1e-3) / 1e9
        out = {"workload": "poisson_1024x1024_x1ch_full_grid (BASELINE configs[1])", "n": W * H, "nnz": int(sp._nnz),
               "sweeps": int(st.sweeps), "ms": float(st.solve_ms), "kernel": int(st.kernel_used),
               "Gnnz_per_s": sp._nnz * st.sweeps / (st.solve_ms * 1e-3) / 1e9, "sweeps_per_s": st.sweeps / (st.solve_ms * 1e-3),
               "effective_GBps": gbs, "effective_frac_of_hbm_peak": gbs / peak,
               "note": "working set %.0f MB per sweep is L2-resident (126 MB): effective bytes/s, not an HBM figure" % (ab / 1e6)}
    elif which == "c5":
        n = args.c5_n
        t0 = time.perf_counter()
        rows, cols, vals = wl.random_spd_coo(n, 13, seed=5)
        t_gen = time.perf_counter() - t0
        sp = pkg.SparseMatrix(np.float64)
        sp.initialize(n, n)
        t0 = time.perf_counter()
        sp.initializeFromTriplets(rows, cols, vals)
        t_asm = time.perf_counter() - t0
        del rows, cols, vals
        nnz = int(sp._nnz)
        rng = np.random.default_rng(1)
        xstar = rng.uniform(-1.0, 1.0, n)
        b = sp.applyToVector(xstar)
        t0 = time.perf_counter()
        info = sp.analyze()
        t_ana = time.perf_counter() - t0
        sweeps = 20
        sp.gaussSeidel(b, epsilon=0.0, max_iteration=3)
        x = sp.gaussSeidel(b, epsilon=0.0, max_iteration=sweeps)
        st = sp.last_stats
        ab = algorithmic_bytes_per_sweep(nnz, n, 1)
        gbs = ab * st.sweeps / (st.solve_ms * 1e-3) / 1e9
        out = {"workload": "random sparse SPD, n = %d, ~27 entries per row, unsorted triplets (BASELINE configs[4])" % n,
               "n": n, "nnz": nnz, "nnz_per_row": nnz / n, "n_colors": int(info["n_colors"]), "kernel": int(st.kernel_used),
               "sweeps": int(st.sweeps), "ms": float(st.solve_ms), "Gnnz_per_s": nnz * st.sweeps / (st.solve_ms * 1e-3) / 1e9,
               "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak,
               "note": "x gathers hit random 32-byte sectors: DRAM traffic exceeds the algorithmic bytes (profiles/)",
               "max_abs_vs_xstar_after_%d_sweeps" % sweeps: float(np.abs(x - xstar).max()),
               "residual_l2": float(sp.residual(b, x)), "residual_l2_x0": float(sp.residual(b, np.ones(n))),
               "host_generation_s": t_gen, "assembly_s_triplets_to_slack_csr": t_asm, "analysis_s_colouring_and_permute": t_ana}
    elif which == "cg":
        # what lab8 / the project actually call (hw8_pa.cc:972, PhotoMontage.cpp:613): conjugateGradient(A^T b, 1e-10,
        # 50, composite) per channel on the full-grid Neumann + pin system; here the three channels in one call
        out = {"workload": "conjugateGradient(A^T b, 1e-10, 50, init = the composite), 3 channels, full-grid Poisson system "
                           "(the reference's own solver call, hw8_pa.cc:972 / PhotoMontage.cpp:613)"}
        for (W, H) in ((566, 752), (4096, 4096)):
            img = wl.synth_image(W, H, 3, seed=7)
            gx, gy = wl.seamless_gradients(img)
            b = np.asarray(pkg.poisson_rhs(W, H, gx, gy, img[:, 0, 0].astype(np.float64))).reshape(3, W * H)
            init = img.reshape(3, -1).astype(np.float64)
            sp = pkg.SparseMatrix(np.float64)
            sp.poisson(W, H)
            sp.conjugateGradientMulti(b, 1e-10, 50, init)  # workspaces, module load
            t = []
            for _ in range(3):
                t0 = time.perf_counter()
                x = sp.conjugateGradientMulti(b, 1e-10, 50, init)
                t.append((time.perf_counter() - t0) * 1e3)
            # the same call with the caller's vectors in page-locked memory (as the Gauss-Seidel e2e leg has them): the
            # three 8 n k-byte copies then run at PCIe speed instead of the driver's pageable staging rate
            bp, ip, xp = pinned_like(pkg, b), pinned_like(pkg, init), pinned_like(pkg, np.zeros_like(b))
            tp = []
            for _ in range(4):
                t0 = time.perf_counter()
                sp.conjugateGradientMulti(bp, 1e-10, 50, ip, out=xp)
                tp.append((time.perf_counter() - t0) * 1e3)
            assert np.array_equal(xp, x), "pinned-buffer call differs"
            ent = {"n": W * H, "nnz": int(sp._nnz), "iterations": list(sp.last_iters), "ms_wall_median": float(np.median(tp[1:])),
                   "ms_wall_median_pageable_vectors": float(np.median(t)),
                   "includes": "b, init H2D (page-locked host vectors); 50 iterations x 3 channels on the device (one SpMV "
                               "pass per iteration for all three); x D2H",
                   "Gnnz_per_s_spmv_equivalent": sp._nnz * 3 * 50 / (np.median(tp[1:]) * 1e-3) / 1e9,
                   "residual_l2": [float(sp.residual(b[c], x[c])) for c in range(3)],
                   "residual_l2_init": [float(sp.residual(b[c], init[c])) for c in range(3)]}
            if (W, H) == (566, 752):
                try:
                    from oracle import pyoracle
                    if pyoracle.ref_available():
                        va, ci, _, rn, _ = sp.layout()
                        ro = np.zeros(W * H, np.int32)
                        ro[1:] = np.cumsum(rn[:-1])
                        ref = pyoracle.Ref(2, "f64").import_csr(va, ro, ci, W * H)
                        xr = [None] * 3

                        def one(c):
                            xr[c] = ref.cg(b[c], 1e-10, 50, init[c])
                        t0 = time.perf_counter()
                        ths = [threading.Thread(target=one, args=(c,)) for c in range(3)]
                        [th.start() for th in ths]
                        [th.join() for th in ths]
                        ent["reference_cpu_ms"] = (time.perf_counter() - t0) * 1e3
                        ent["reference_cpu_threads"] = 3
                        ent["max_abs_vs_reference"] = float(max(np.abs(x[c] - xr[c]).max() for c in range(3)))
                except Exception as e:
                    ent["reference_cpu_error"] = repr(e)[:200]
            out["%dx%d" % (W, H)] = ent
            del sp
    else:
        raise SystemExit("unknown config " + str(which))
    print(json.dumps(out))


def other_config_child(args, which, timeout):
    cmd = [sys.executable, os.path.abspath(__file__), "--other-config-only", which, "--c5-n", str(args.c5_n)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}
    except Exception as e:
        return {"error": repr(e)[:400]}


def time_to_tol_child(args):
    """Run the leg in a child process; returns its dict or {"error": ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--time-to-tol-only", "--size", str(args.size), "--channels",
           str(args.channels), "--check-every", str(args.check_every), "--kernel", str(args.kernel)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}
    except Exception as e:  # timeout, spawn failure, bad JSON
        return {"error": repr(e)[:400]}


def main():
    args = parse()
    if args.time_to_tol_only:
        return run_time_to_tol(args)
    if args.other_config_only:
        return run_other_config(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
