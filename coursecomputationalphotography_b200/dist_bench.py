"""The N > 1 leg of bench.py: 4096^2 x 3-channel Poisson system split into row strips, one rank per GPU."""
import json
import os
import time

import numpy as np


def run(args, pkg, wl, dist, rank, world, local):
    import torch
    from . import strips
    from bench import ClockSampler, algorithmic_bytes_per_sweep, peaks, workload_config

    L = pkg.load()
    dev = torch.device("cuda", local)
    uid = strips.broadcast_unique_id(dist, rank, dev)
    W = H = args.size
    ch = args.channels
    y0, y1 = wl.strip_bounds(H, world)[rank]
    solver = strips.StripSolver(uid, rank, world, local)
    solver.poisson_strip(W, H, y0, y1)
    n_local = W * (y1 - y0)
    n = W * H
    nnz = wl.poisson_nnz(W, H)
    b_host = strips.strip_rhs(W, H, ch, y0, y1)
    b_dev = torch.from_numpy(b_host).to(dev)
    x_dev = torch.empty_like(b_dev)
    opts = pkg.SparseMatrix.options(check_every=args.check_every, kernel=args.kernel)
    stream = torch.cuda.ExternalStream(L.gsb_stream())

    def step():
        return solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), ch, 0.0, args.sweeps, opts)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, solve_ms, sweeps_done = 0, 0.0, 0
    e0.record(stream)
    for _ in range(args.steps):
        st = step()
        launches += st.kernel_launches
        solve_ms += st.solve_ms
        sweeps_done += st.sweeps
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1), solve_ms, float(launches)], device=dev, dtype=torch.float64)
    mx = ms.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = ms.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    total_ms, solve_ms_max = float(mx[0]), float(mx[1])
    clocks = sampler.stop() if rank == 0 else None
    resid = [solver.residual_dev(b_dev[c].data_ptr(), x_dev[c].data_ptr()) for c in range(ch)]

    # e2e: the SAME workload as the N = 1 line -- the host CSR as Eigen would hand it over, imported every step
    # (each rank uploads its rows: gsb_dist_matrix_rows), + b H2D + sweeps + x D2H; pinned host buffers
    e2e = None
    if not args.no_e2e:
        from bench import pinned_like
        # everything that can fail locally comes BEFORE the leg's first collective; the ranks then agree (min) on
        # whether to run it, so that a failure on one rank cannot leave the others waiting in a barrier
        prep, err = None, None
        try:
            full = pkg.SparseMatrix(np.float64)   # untimed: stands for the reference's Eigen A^T*A step
            full.poisson(W, H)
            va, ci, rb, rn, _ = full.layout()
            r0, r1 = y0 * W, y1 * W
            ro_s = np.zeros(n_local + 1, np.int32)
            ro_s[1:] = np.cumsum(rn[r0:r1])
            k0 = int(rb[r0])
            k1 = k0 + int(ro_s[-1])  # compressed layout: the strip's entries are one contiguous span
            prep = (pinned_like(pkg, va[k0:k1]), pinned_like(pkg, ci[k0:k1]), pinned_like(pkg, ro_s), r0)
            del full, va, ci, rb, rn
            torch.cuda.empty_cache()
        except Exception as e:
            err = repr(e)[:300]
        ok = torch.tensor([1 if prep is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0:
            e2e = {"error": err or "another rank could not prepare the host CSR"}
        else:
            va_s, ci_s, ro_s, r0 = prep
            bh = torch.from_numpy(b_host).pin_memory()
            xh = torch.empty_like(bh).pin_memory()
            steps_e2e = max(1, args.e2e_steps)
            per_step = []
            warm_e2e = max(args.warmup, 3)  # untimed, as at N = 1
            for it in range(steps_e2e + warm_e2e):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                solver.matrix_rows(va_s, ro_s, ci_s, r0, n, W)
                bd = bh.to(dev, non_blocking=True)
                torch.cuda.synchronize()
                xd = torch.empty_like(bd)
                solver.gauss_seidel_dev(bd.data_ptr(), xd.data_ptr(), ch, 0.0, args.sweeps, opts)
                xh.copy_(xd)
                torch.cuda.synchronize()
                dist.barrier()
                if it >= warm_e2e:
                    per_step.append(time.perf_counter() - t0)
            tt = torch.tensor(per_step, device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # per step: the slowest rank
            per = [float(v) for v in tt]
            h2d = torch.tensor([float(va_s.nbytes + ci_s.nbytes + ro_s.nbytes + b_host.nbytes)], device=dev, dtype=torch.float64)
            dist.all_reduce(h2d)
            same = torch.tensor([1 if torch.equal(xd, x_dev) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            e2e = {"value": nnz * ch * args.sweeps * steps_e2e / sum(per) / 1e9, "unit": "Gnnz/s",
                   "h2d_bytes_per_step": int(h2d[0]), "d2h_bytes_per_step": int(b_host.nbytes) * world,
                   "ms_per_step": sum(per) / steps_e2e * 1e3, "ms_per_step_min": min(per) * 1e3,
                   "ms_per_step_median": float(np.median(per)) * 1e3, "steps": steps_e2e, "warmup": warm_e2e,
                   "per_step_ms": [round(v * 1e3, 2) for v in per],
                   "includes": "host CSR import of every rank's rows (H2D) + ordering + halo setup + b H2D + sweeps + x D2H "
                               "(all ranks, slowest rank per step)", "host_memory": "pinned",
                   "equals_resident_input_result": bool(int(same[0]))}

    parity = parity_vs_one_gpu(args, pkg, wl, dist, torch, solver, b_dev, rank, world, dev, W, H, ch, n_local, opts)

    ttt = None
    if not args.no_time_to_tol:
        try:
            ttt = time_to_tol_strips(args, pkg, wl, dist, torch, solver, rank, world, dev)
        except Exception as e:  # nothing in this leg may cost the main line
            ttt = {"error": repr(e)[:300]}

    c4 = None
    if not getattr(args, "no_c4", False):
        try:
            c4 = config4_strips(args, pkg, wl, dist, torch, solver, rank, world, dev)
        except Exception as e:  # nothing in this leg may cost the main line
            c4 = {"error": repr(e)[:300]}

    if rank == 0:
        assert sweeps_done == args.sweeps * args.steps
        value = nnz * ch * sweeps_done / (total_ms * 1e-3) / 1e9
        peak, peak_src = peaks()
        abytes = algorithmic_bytes_per_sweep(nnz, n, ch) / world  # per GPU
        per_launch_ms = solve_ms_max / (sweeps_done * 2)
        achieved = (abytes / 2) / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "gauss_seidel_throughput", "value": value, "unit": "Gnnz/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(W, H, ch, n, nnz),
            "run": {"partition": "%d row strips, one rank per GPU" % world,
                       "sweeps_per_step": args.sweeps, "ordering": "red-black (global parity)",
                       "check_every": args.check_every,
                       "halo": "1 image row per neighbour per colour phase, " +
                               ("stored into the neighbour's ghost slots by the phase kernel (NVLink peer memory)"
                                if st.kernel_used >= 10 else "packed ncclSend/ncclRecv"),
                       "stop_rule_allreduce": ("fused into the end-of-sweep kernel (peer memory)" if st.kernel_used >= 30
                                               else "ncclAllReduce"),
                       "per_gpu_working_set_gb_per_sweep": abytes / 1e9,
                       "sweeps_per_s": sweeps_done / (total_ms * 1e-3), "residual_l2": resid},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "gs_phase per GPU incl. halo exchange gaps", "peak_source": peak_src,
                         "avg_launch_ms": per_launch_ms},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(sm[2]), "clocks": clocks,
        }
        line.update(parity)
        if c4 is not None:
            line["config4_16384_strips"] = c4
        if ttt is not None:
            line["time_to_tol"] = ttt
        print(json.dumps(line), flush=True)
    solver.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and parity.get("parity_bitwise_vs_1gpu") is False:
        raise SystemExit("dist_bench: the %d-strip solution differs from the 1-GPU solution" % world)
    if rank == 0 and c4 and c4.get("parity_bitwise_vs_1gpu") is False:
        raise SystemExit("dist_bench: the %d-strip 16384^2 solution differs from the 1-GPU solution" % world)


def time_to_tol_strips(args, pkg, wl, dist, torch, solver, rank, world, dev):
    """BASELINE configs[2] "with irregular mask, 1 and 8 B200": the 4096^2 x 3-channel Dirichlet-masked blend (compact
    unknowns, parity colouring handed over as bytes) split into `world` row blocks, solved from x0 = 1 to the
    reference's stop rule (eps = 1e-5), gathered on rank 0 and compared with the REFERENCE's own solution (the golden
    of tests/golden/c3_masked_4096.*).  Every rank generates the (deterministic) host system and uploads its rows."""
    import os
    size, ch = args.size, args.channels
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gpath = os.path.join(root, "tests", "golden", "c3_masked_%d" % size)
    prep, err = None, None
    try:  # local work first, then the ranks agree on whether the leg runs (see the e2e leg)
        t0 = time.perf_counter()
        ro, ci, va, b, pix, colors = wl.c3_masked_system(size, ch)
        n = len(pix)
        r0, r1 = [(n * q) // world for q in (rank, rank + 1)]
        k0, k1 = int(ro[r0]), int(ro[r1])
        prep = (np.ascontiguousarray(va[k0:k1]), (ro[r0:r1 + 1] - ro[r0]).astype(np.int32), np.ascontiguousarray(ci[k0:k1]),
                r0, r1, n, len(va), colors.astype(np.uint8), np.ascontiguousarray(b[:, r0:r1]), time.perf_counter() - t0)
        del ro, ci, va, b
    except Exception as e:
        err = repr(e)[:300]
    ok = torch.tensor([1 if prep is not None else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok[0]) == 0:
        return {"error": err or "another rank could not generate the system"}
    va_s, ro_s, ci_s, r0, r1, n, nnz, col8, b_s, t_gen = prep
    solver.set_colors(col8)
    solver.matrix_rows(va_s, ro_s, ci_s, r0, n, 1)
    b_dev = torch.from_numpy(b_s).to(dev)
    x_dev = torch.empty_like(b_dev)
    opts = pkg.SparseMatrix.options(check_every=args.check_every)
    cap, eps = 40000, 1e-5
    solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), ch, 0.0, 20, opts)  # warm-up: plan, peer mappings
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), ch, eps, cap, opts)
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t1, float(st.solve_ms)], device=dev, dtype=torch.float64)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    resid = [solver.residual_dev(b_dev[c].data_ptr(), x_dev[c].data_ptr()) for c in range(ch)]
    sizes = [(n * (q + 1)) // world - (n * q) // world for q in range(world)]
    nmax = max(sizes)
    pad = torch.zeros(ch, nmax, device=dev, dtype=torch.float64)
    pad[:, :r1 - r0] = x_dev
    got = [torch.empty(ch, nmax, device=dev, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, got, dst=0)
    solver.set_colors(np.zeros(0, np.uint8))  # back to pixel parity for whatever the caller builds next
    if rank != 0:
        return None
    x = torch.cat([got[q][:, :sizes[q]] for q in range(world)], dim=1).cpu().numpy()
    out = {"workload": "dirichlet_masked_blend_%dx%d_x%dch, blob mask 30 %%, thickness <= 48 px (SURVEY 8d C3), %d row blocks"
                       % (size, size, ch, world), "n": int(n), "nnz": int(nnz), "epsilon": eps, "max_iteration": cap,
           "x0": 1.0, "sweeps": int(st.sweeps), "stopped": bool(st.sweeps < cap), "ms": float(wall[1]),
           "wall_ms": float(wall[0]) * 1e3, "Gnnz_per_s": nnz * ch * st.sweeps / (float(wall[1]) * 1e-3) / 1e9,
           "last_eps": [float(v) for v in list(st.last_eps)[:ch]], "residual_l2": resid, "kernel": int(st.kernel_used),
           "host_generation_s": t_gen}
    if ch == 3 and os.path.exists(gpath + ".npz"):
        import json as _json
        meta, gold = _json.load(open(gpath + ".json")), np.load(gpath + ".npz")
        if meta["n"] == n and meta["nnz"] == nnz:
            idx = gold["index"]
            diffs = [float(np.abs(x[c][idx] - gold["x_eps%g_ch%d" % (eps, c)]).max()) for c in range(ch)]
            u8 = lambda v: np.clip(v, 0.0, 255.0).astype(np.uint8)
            runs = [r for r in meta["runs"] if r["epsilon"] == eps]
            out.update({"max_abs_vs_reference": max(diffs), "tolerance": 1e-4 * 255.0,
                        "within_tolerance": bool(max(diffs) <= 1e-4 * 255.0),
                        "u8_equal_on_sample": bool(all(np.array_equal(u8(x[c][idx]), u8(gold["x_eps%g_ch%d" % (eps, c)]))
                                                       for c in range(ch))),
                        "reference": {"sweeps": [r["sweeps"] for r in runs], "cpu_s": max(r["cpu_s"] for r in runs),
                                      "what": meta["source"]},
                        "speedup_vs_reference_cpu": max(r["cpu_s"] for r in runs) / float(wall[0])})
    return out


def _c4_rhs(torch, dev, r0, r1):
    """A deterministic right-hand side as a function of the GLOBAL row index, generated on the device (a 16384^2
    image never exists on one host): values in [-0.5, 0.5)."""
    idx = torch.arange(r0, r1, device=dev, dtype=torch.int64)
    return (((idx * 2654435761) % 1000003).to(torch.float64) / 1000003.0 - 0.5).reshape(1, -1).contiguous()


def config4_strips(args, pkg, wl, dist, torch, solver, rank, world, dev):
    """BASELINE configs[3]: 16384 x 16384 full-grid Poisson system (n = 268 M, nnz = 1.342 G), one channel, row strips
    generated on the devices, 200 sweeps with the reference's stop-rule cadence; then 10 sweeps from x0 = 1 gathered
    on rank 0 and compared BIT FOR BIT with the single-GPU solver on the same system (it fits one B200)."""
    import hashlib
    from bench import algorithmic_bytes_per_sweep, peaks
    W = H = int(getattr(args, "c4_size", 16384))
    sweeps = 200
    y0, y1 = wl.strip_bounds(H, world)[rank]
    n_local = W * (y1 - y0)
    solver.poisson_strip(W, H, y0, y1)
    b_dev = _c4_rhs(torch, dev, y0 * W, y1 * W)
    x_dev = torch.empty_like(b_dev)
    opts = pkg.SparseMatrix.options(check_every=args.check_every)
    solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, 20, opts)  # warm-up: plan, peer mappings
    dist.barrier()
    torch.cuda.synchronize()
    st = solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, sweeps, opts)
    ms = torch.tensor([float(st.solve_ms)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # parity: PARITY_SWEEPS sweeps, strips vs one GPU
    stp = solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), 1, 0.0, PARITY_SWEEPS, opts)
    bounds = [b1 - a1 for a1, b1 in wl.strip_bounds(H, world)]
    nmax = max(bounds) * W
    pad = torch.zeros(1, nmax, device=dev, dtype=torch.float64)
    pad[:, :n_local] = x_dev
    got = [torch.empty(1, nmax, device=dev, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, got, dst=0)
    del pad
    if rank != 0:
        return None
    x_full = torch.cat([got[r][:, :bounds[r] * W] for r in range(world)], dim=1).contiguous()
    del got
    nnz = wl.poisson_nnz(W, H)
    n = W * H
    peak, _ = peaks()
    ab = algorithmic_bytes_per_sweep(nnz, n, 1) / world
    per_sweep_ms = float(ms[0]) / sweeps
    out = {"workload": "poisson_%dx%d_x1ch_full_grid, %d row strips generated on the devices (BASELINE configs[3])" % (W, H, world),
           "n": n, "nnz": int(nnz), "sweeps": sweeps, "ms": float(ms[0]), "ms_per_sweep": per_sweep_ms,
           "Gnnz_per_s": nnz * sweeps / (float(ms[0]) * 1e-3) / 1e9, "kernel": int(st.kernel_used),
           "per_gpu_algorithmic_GBps": ab / (per_sweep_ms * 1e-3) / 1e9,
           "per_gpu_frac_of_hbm_peak": ab / (per_sweep_ms * 1e-3) / 1e9 / peak}
    try:
        sp = pkg.SparseMatrix(np.float64)
        sp.poisson(W, H)
        b_full = _c4_rhs(torch, dev, 0, n)
        x_one = torch.empty_like(b_full)
        st1 = sp.gaussSeidel_dev(b_full.data_ptr(), x_one.data_ptr(), 1, 0.0, PARITY_SWEEPS,
                                 pkg.SparseMatrix.options(check_every=args.check_every))
        torch.cuda.synchronize()
        out["parity_bitwise_vs_1gpu"] = bool(torch.equal(x_full, x_one))
        out["parity"] = {"what": "%d sweeps from x0 = 1, %d strips vs one GPU" % (PARITY_SWEEPS, world),
                         "max_abs_diff": float((x_full - x_one).abs().max()), "kernel_1gpu": int(st1.kernel_used),
                         "x_sha256": hashlib.sha256(x_full.cpu().numpy().tobytes()).hexdigest()}
        del sp, b_full, x_one
    except Exception as e:
        out["parity_error"] = repr(e)[:300]
    torch.cuda.empty_cache()
    return out


PARITY_SWEEPS = 10


def parity_vs_one_gpu(args, pkg, wl, dist, torch, solver, b_dev, rank, world, dev, W, H, ch, n_local, opts):
    """A fixed PARITY_SWEEPS-sweep solve from x0 = 1 on the strips, gathered on rank 0 and compared BIT FOR BIT with
    the single-GPU solver (SparseMatrix::gaussSeidel on the whole system, which fits one B200) on the same b.
    Returns the keys rank 0 adds to its line; a mismatch makes the bench exit non-zero."""
    import hashlib
    x_dev = torch.empty_like(b_dev)
    st = solver.gauss_seidel_dev(b_dev.data_ptr(), x_dev.data_ptr(), ch, 0.0, PARITY_SWEEPS, opts)
    assert st.sweeps == PARITY_SWEEPS
    bounds = [y1 - y0 for y0, y1 in wl.strip_bounds(H, world)]
    nmax = max(bounds) * W

    def padded(t):
        out = torch.zeros(ch, nmax, device=dev, dtype=torch.float64)
        out[:, :n_local] = t
        return out

    got_x = [torch.empty(ch, nmax, device=dev, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    got_b = [torch.empty(ch, nmax, device=dev, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(padded(x_dev), got_x, dst=0)
    dist.gather(padded(b_dev), got_b, dst=0)
    if rank != 0:
        return {}
    x_full = torch.cat([got_x[r][:, :bounds[r] * W] for r in range(world)], dim=1).contiguous()
    b_full = torch.cat([got_b[r][:, :bounds[r] * W] for r in range(world)], dim=1).contiguous()
    del got_x, got_b
    sp = pkg.SparseMatrix(np.float64)
    sp.poisson(W, H)
    x_one = torch.empty_like(x_full)
    st1 = sp.gaussSeidel_dev(b_full.data_ptr(), x_one.data_ptr(), ch, 0.0, PARITY_SWEEPS,
                             pkg.SparseMatrix.options(check_every=args.check_every))
    torch.cuda.synchronize()
    same = bool(torch.equal(x_full, x_one))
    out = {"parity_bitwise_vs_1gpu": same, "parity": {
        "what": "%d sweeps from x0 = 1: %d strips gathered on rank 0 vs SparseMatrix::gaussSeidel on one GPU, same b"
                % (PARITY_SWEEPS, world),
        "x_sha256": hashlib.sha256(x_full.cpu().numpy().tobytes()).hexdigest(),
        "x_sha256_1gpu": hashlib.sha256(x_one.cpu().numpy().tobytes()).hexdigest(),
        "max_abs_diff": float((x_full - x_one).abs().max()), "kernel_1gpu": int(st1.kernel_used),
        "kernel_strips": int(st.kernel_used), "last_eps_strips": [float(v) for v in list(st.last_eps)[:ch]],
        "last_eps_1gpu": [float(v) for v in list(st1.last_eps)[:ch]]}}
    del sp
    return out
