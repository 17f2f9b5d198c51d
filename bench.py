#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Cahn-Hilliard stepper.

Metric (BASELINE.json): CH steps/sec at N=512 per B200 (aggregated over an ensemble of
independent simulations; whole-job value over all GPUs).  One bench "step" = one CH time
step (reference solver.py:165-249) of EVERY member of a batch of `--batch` N=512
simulations per GPU -- the experiment.py A0/A1 ensemble with the fixed-length (`full_sim`)
variant of SURVEY.md 8d config 3 so that no member drops out during the timed region.
`value` = sim-steps/s with state resident in HBM; `e2e` = the same through the public
BatchStepper API from HOST buffers (upload of the initial field, K steps, download of the
TimeData rows and final fields inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
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

N_GRID = 512
ALGO_BYTES_PER_SIM_STEP = 32 * N_GRID * N_GRID      # read+write U-equivalent and hat_U, fp64
# kappa_tilde at the (fac_A0, fac_A1) corners and centre, from the reference (tests/golden/n512_corner_*.npz)
_KAPPA = {(0.995, 0.995): 2.8527824637628903e-4, (0.995, 1.005): 4.088183715201557e-4,
          (1.005, 0.995): 2.053833366570028e-4, (1.005, 1.005): 3.127937637279922e-4}


def member_scalars(runs, seed=85972):
    """A0/A1 factors exactly as experiment.py:157-160; kappa_tilde by bilinear interpolation of
    the reference's corner values (synthetic stand-in for the 0.5 s/member sympy solve)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fac = rng.uniform(0.995, 1.005, size=(runs, 2))
    u = (fac[:, 0] - 0.995) / 0.01
    v = (fac[:, 1] - 0.995) / 0.01
    k = ((1 - u) * (1 - v) * _KAPPA[(0.995, 0.995)] + (1 - u) * v * _KAPPA[(0.995, 1.005)]
         + u * (1 - v) * _KAPPA[(1.005, 0.995)] + u * v * _KAPPA[(1.005, 1.005)])
    return fac, k


def workload_text(members):
    return (f"A0/A1 ensemble (BASELINE configs[2] shape): {members} independent N={N_GRID} simulations per GPU, "
            f"each the configs[1] stepper with full_sim (fixed length); one bench step = one CH time step of "
            f"every member")


def physical_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


# ------------------------------------------------------------------------------- CPU legs
def _oracle_member(args):
    """One ensemble member on one host core with the oracle port; returns seconds for `steps`."""
    fac0, fac1, kappa, warm, steps = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ch_oracle as orc
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        k = orc.Consts.from_params(N=N_GRID, A0=orc.redlich_kister_A0(923.15) * fac0,
                                   A1=orc.redlich_kister_A1(923.15) * fac1, kappa_tilde=kappa)
        U0, draw = orc.initial_field(N_GRID, 0.875, "uniform", 2023)
        s = orc.OracleSolver(k, U0, full_sim=True, create_rand=draw)
        s.prepare()
        s.run(1 + warm)
        t = time.perf_counter()
        s.run(steps)
        return time.perf_counter() - t


def cpu_baseline(sample_steps=300):
    """Oracle port, one core, default config (BASELINE config 1 shape), bounded sample."""
    dt = _oracle_member((1.0, 1.0, 2.989112919661156e-4, 20, sample_steps))
    return {"value": round(sample_steps / dt, 2), "unit": "sim-steps/s", "cores": 1, "kind": "port",
            "port_over_reference": port_over_reference(),
            "sample": f"oracle/ch_oracle.py (numpy+scipy.fftpack restatement of solver.py), 1 sim N={N_GRID}, "
                      f"{sample_steps} steps after 20 warm-up, single thread"}


def port_over_reference():
    """Speed of the oracle port relative to the unmodified reference, measured in the build container where
    both run (tools/port_vs_reference.py -> profiles/port_over_reference.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "port_over_reference.json")))["port_over_reference"]
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is pure
    Python and cannot travel to the GPU box) on all physical host cores, experiment.py-style
    process pool (experiment.py:197-211), TWO members per core, `sample` CH steps per member
    per bench step."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = physical_cores()
    members = 2 * cores
    sample = 1                                   # CH steps per member per bench step
    fac, kap = member_scalars(members)
    jobs = [(fac[i, 0], fac[i, 1], kap[i], args.warmup * sample, args.steps * sample) for i in range(members)]
    t0 = time.perf_counter()
    with mp.get_context("forkserver").Pool(cores) as pool:
        tw = time.perf_counter()
        secs = pool.map(_oracle_member, jobs, chunksize=1)
        wall_pool = time.perf_counter() - tw
    # `cores` workers run concurrently, each its members back to back: the job's stepping time is the workers'
    # busy time (the pool's wall time also holds every member's set-up and warm-up, which the GPU arm does not
    # time either -- it is reported as pool_wall_s)
    wall = sum(secs) / cores
    value = members * args.steps * sample / wall
    por = port_over_reference()
    line = {"impl": "reference", "metric": "ch_steps_per_sec_n512", "value": round(value, 2), "unit": "sim-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(wall / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(members) + f" [CPU arm: two members per physical host core, "
                                   f"{sample} CH step per member per bench step]", "members_per_gpu": members, "N": N_GRID},
            "cpu_baseline": {"value": round(value, 2), "unit": "sim-steps/s", "cores": cores, "kind": "port",
                             "port_over_reference": por,
                             "sample": f"{members} members x {args.steps * sample} steps, mp.Pool({cores}), "
                                       f"BLAS/FFT single-threaded per process as in the reference; "
                                       f"port_over_reference = oracle port speed / unmodified reference speed on one "
                                       f"core (profiles/port_over_reference.json, build container)",
                             "pool_wall_s": round(wall_pool, 2)},
            "e2e": {"value": round(value, 2), "unit": "sim-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "setup_s": round(time.perf_counter() - t0 - wall_pool, 2)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= t <= t1 + 0.05):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


# ------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import chsimpy_b200 as ch
    from chsimpy_b200.solver import BatchStepper, make_params_struct

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 stepper has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: send fd 1 to stderr
        # for that moment so that stdout carries nothing but the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B, K, W = args.batch, args.steps, args.warmup

    # ---- members of this rank (weak scaling: B per GPU, run ids rank*B .. rank*B+B-1)
    fac_all, kap_all = member_scalars(B * world)
    sl = slice(rank * B, (rank + 1) * B)
    base = ch.Parameters()
    base.no_gui, base.full_sim, base.kappa_tilde = True, True, 1.0
    structs = []
    for f, kap in zip(fac_all[sl], kap_all[sl]):
        p = base.deepcopy()
        p.kappa_tilde = float(kap)
        p.func_A0 = (lambda f0: (lambda T: ch.utils.A0(T) * f0))(float(f[0]))
        p.func_A1 = (lambda f1: (lambda T: ch.utils.A1(T) * f1))(float(f[1]))
        structs.append(make_params_struct(p, ch.Solution(p)))
    rng = np.random.Generator(np.random.PCG64(2023))
    U0 = 0.875 + 0.875 * 0.01 * (rng.random((N_GRID, N_GRID)) - 0.5)          # solver.py:78-82, seed 2023

    st = BatchStepper(N_GRID, structs, rows_cap=max(K + 8, W + 8))
    st.set_U(U0)
    st.prepare()
    st.begin()
    st.steps(W)
    st.poll(); st.take_rows()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: K steps, state resident in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = st.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    st.steps(K)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = st.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    running, stop, cs, rw = st.poll()
    assert running == B and int(rw.min()) == K, "a member dropped out of the timed region"
    st.take_rows()
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * B * K / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events around every launch, same stream), second pass
    st.set_timing(True)
    st.steps(K)
    tk, nk = st.get_timing()
    st.set_timing(False)
    st.poll(); st.take_rows()
    st.end()
    col_us, row_us = tk["col"] / nk * 1e3, tk["row"] / nk * 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    step_bytes = ALGO_BYTES_PER_SIM_STEP * B
    achieved = step_bytes / ((col_us + row_us) * 1e-6) / 1e9
    traffic = None
    fp64_instr = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("dram_bytes_per_step_per_sim")
        traffic = traffic * B if traffic else None
        fp64_instr = tj.get("fp64_warp_instr_per_step_per_sim")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_col<512,STEP> + k_row<512,STEP> (the two launches of one step)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_launch_pair": step_bytes,
                "k_col_us": round(col_us, 2), "k_row_us": round(row_us, 2),
                "fp64_issue_floor": fp64_floor(fp64_instr, clocks, value / world),
                "k_col_gbs_own_32N2": round(32 * N_GRID ** 2 * B / (col_us * 1e-6) / 1e9, 1),
                "k_row_gbs_own_16N2": round(16 * N_GRID ** 2 * B / (row_us * 1e-6) / 1e9, 1)}

    # ---- e2e: public API from host buffers, copies inside the timed region.  The members run as G groups:
    # while group g+1 steps, the results of group g drain to pinned host memory on a copy stream (double
    # buffering in time).  Two variants are timed:
    #   e2e.value              host U_init -> set_U (H2D) -> prepare -> K steps -> the steps' results (the TimeData
    #                          rows: E, E2, SA, ... of every member and step) in pinned host memory (D2H)
    #   e2e.with_final_fields  the same plus every member's final N x N field (what Solution.U / the per-run
    #                          exports of experiment.py need once per simulation, i.e. once per ~1700 steps; at
    #                          K = 20 steps per run it is 99.9 % of the bytes and bound by the host's D2H rate)
    del st
    torch.cuda.empty_cache()
    G = 4 if B % 4 == 0 and B >= 64 else 1
    Bg = B // G
    numa = bind_to_gpu_numa(local)
    rows_host = torch.empty((B, K, 9), dtype=torch.float64).pin_memory()
    U_host = torch.empty((B, N_GRID, N_GRID), dtype=torch.float64).pin_memory()
    groups = [BatchStepper(N_GRID, structs[g * Bg:(g + 1) * Bg], rows_cap=K + 8) for g in range(G)]
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()

    def e2e_run(with_fields):
        cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(G)]
        done = [torch.cuda.Event() for _ in range(G)]
        barrier()
        t0_ = time.perf_counter()
        for sg in groups:
            sg.set_U(U0)                                    # H2D: the members share one initial field (quirk Q11)
            sg.prepare()                                    # row 0 (reads it back: the only host syncs of the run)
        for g, sg in enumerate(groups):
            sg.begin()
            sg.steps(K, last=True)
            if with_fields:
                sg.end()                                    # queues U = idctn(hat_U)
            done[g].record(main_stream)
            copy_stream.wait_event(done[g])
            with torch.cuda.stream(copy_stream):
                cev[g][0].record()
                rows_host[g * Bg:(g + 1) * Bg].copy_(sg.rows[:, :K, :], non_blocking=True)     # D2H TimeData
                if with_fields:
                    U_host[g * Bg:(g + 1) * Bg].copy_(sg.U, non_blocking=True)                   # D2H final fields
                cev[g][1].record()
        copy_stream.synchronize()
        barrier()
        dt = time.perf_counter() - t0_
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        for sg in groups:
            sg.poll()                                       # (outside the timed region) every member ran K steps
            assert int(sg._rw.min()) == K
        assert np.isfinite(rows_host.numpy()).all()
        return dt, sum(a.elapsed_time(b) for a, b in cev)
    e2e_run(False)                                          # warm-up of the group handles
    te, _ = e2e_run(False)
    tf, cms = e2e_run(True)
    assert abs(float(U_host[B - 1].mean()) - float(U0.mean())) < 1e-9
    row_bytes, fld_bytes = rows_host.numel() * 8, U_host.numel() * 8
    e2e = {"value": round(world * B * K / te, 1), "unit": "sim-steps/s",
           "h2d_bytes_per_step": int((G * U0.size * 8 + B * 136) / K),
           "d2h_bytes_per_step": int(row_bytes / K),
           "what": f"BatchStepper.set_U (host U_init, H2D) -> prepare -> K steps -> every step's TimeData row of every member "
                   f"in pinned host memory (D2H on a copy stream); {G} member groups",
           "groups": G, "pinned_numa_node": numa,
           "with_final_fields": {"value": round(world * B * K / tf, 1), "unit": "sim-steps/s",
                                 "d2h_bytes_per_step": int((row_bytes + fld_bytes) / K),
                                 "d2h_gbs_this_rank": round((row_bytes + fld_bytes) / 1e6 / cms, 1),
                                 "what": "the same plus the final N x N field of every member (Solution.U): the D2H of a group "
                                         "overlaps the steps of the next; bound by the host's D2H rate when all GPUs of the box copy"}}
    del groups, rows_host, U_host
    torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: one large domain on the row-slab path over ALL ranks of this run (collective)
    large = None
    if not args.no_cpu:
        large = large_domain_suite(ch, rank, world)

    if rank == 0:
        line = {"metric": "ch_steps_per_sec_n512", "value": round(value, 1), "unit": "sim-steps/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms / K, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_text(B), "members_per_gpu": B, "N": N_GRID, "l2": "working set 6 MiB/member >> 126 MB L2 "
                           "(inputs larger than L2, no flush needed)" if B * 6 > 252 else "L2-resident batch"},
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "sims_per_s_at_1674_steps": round(value / 1673.0, 2)}
        if large is not None:
            line["large_domain"] = large
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
            line["single_sim"] = single_sim_probe(ch)
            line["ensemble_to_stop"] = ensemble_to_stop_probe(ch)
            line["jitter_adaptive"] = jitter_adaptive_probe(ch)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def fp64_floor(instr_per_sim_step, clocks, rate_per_gpu):
    """The other ceiling of this path: the step needs ~139 FP64 instructions per grid point (ncu), and
    B200 issues one FP64 warp instruction per 2 cycles per SM sub-partition -- at N=512 that floor
    (~2 us per sim-step) is ABOVE the HBM time of the 32*N^2 algorithmic bytes (1.3 us)."""
    if not instr_per_sim_step:
        return None
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    floor_rate = sms * 4 * 0.5 * mhz * 1e6 / instr_per_sim_step
    return {"fp64_warp_instr_per_sim_step": instr_per_sim_step, "sim_steps_per_s_at_full_fp64_issue": round(floor_rate, 1),
            "frac": round(rate_per_gpu / floor_rate, 4), "source": "profiles/traffic.json (ncu sm__inst_executed_pipe_fp64)"}


def single_sim_probe(ch):
    """BASELINE configs[1]: one N=512 simulation run to the energy stop (latency-bound)."""
    import torch
    p = ch.Parameters()
    p.no_gui = True
    p.kappa_tilde = 2.989112919661156e-4
    s = ch.Solver(p)
    s.prepare()
    torch.cuda.synchronize()
    t = time.perf_counter()
    sol = s.solve_or_resume(p.ntmax)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return {"workload": "configs[1]: N=512 to the energy stop, device-side stop flag, host poll every 128 steps",
            "computed_steps": sol.computed_steps, "stop_reason": sol.stop_reason,
            "steps_per_s": round((sol.computed_steps - 1) / dt, 1), "wall_s": round(dt, 4)}


def ensemble_to_stop_probe(ch, members=1024):
    """BASELINE configs[2] as the reference runs it: every member to ITS energy stop (device-side
    flags, grid compaction at each poll); solve only (kappa_tilde interpolated, no exports)."""
    import torch
    from chsimpy_b200.solver import BatchStepper, make_params_struct
    fac, kap = member_scalars(members)
    structs = []
    for f, k in zip(fac, kap):
        p = ch.Parameters()
        p.no_gui, p.kappa_tilde = True, float(k)
        p.func_A0 = (lambda f0: (lambda T: ch.utils.A0(T) * f0))(float(f[0]))
        p.func_A1 = (lambda f1: (lambda T: ch.utils.A1(T) * f1))(float(f[1]))
        structs.append(make_params_struct(p, ch.Solution(p)))
    U0 = 0.875 + 0.875 * 0.01 * (np.random.Generator(np.random.PCG64(2023)).random((N_GRID, N_GRID)) - 0.5)
    st = BatchStepper(N_GRID, structs, rows_cap=128)
    st.set_U(U0)
    st.prepare()
    torch.cuda.synchronize()
    t = time.perf_counter()
    rows, done = st.run(10 ** 6, poll_every=128)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    stops = done + 1
    return {"workload": f"{members} members, A factors PCG64(85972), each to its own energy stop, host poll every 128 steps",
            "sims_per_s": round(members / dt, 2), "sim_steps_per_s": round(float(done.sum()) / dt, 1), "wall_s": round(dt, 3),
            "stop_step_min_mean_max": [int(stops.min()), round(float(stops.mean()), 1), int(stops.max())]}


def jitter_adaptive_probe(ch):
    """BASELINE configs[3]: N=512, --jitter 0.01 --adaptive-time (delt_max=2e-10, the stable variant),
    the reference's PCG64 noise stream regenerated bit-exactly on the device per 64-step chunk."""
    import torch
    p = ch.Parameters()
    p.no_gui, p.full_sim, p.jitter, p.adaptive_time, p.delt_max, p.ntmax = True, True, 0.01, True, 2e-10, 700
    p.kappa_tilde = 2.989112919661156e-4
    s = ch.Solver(p)
    s.prepare()
    torch.cuda.synchronize()
    t = time.perf_counter()
    sol = s.solve_or_resume(p.ntmax)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return {"workload": "configs[3]: N=512 jitter 0.01 + adaptive dt (delt_max 2e-10), 700 steps", "steps_per_s": round(699 / dt, 1),
            "wall_s": round(dt, 3), "delt_last": float(sol.delt[-1]), "note": "noise = numpy PCG64 stream reproduced bit-exactly on the device (k_pcg64_fill); diagnostics via k_diag"}


def bind_to_gpu_numa(index):
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers it
    allocates next are node-local (8 ranks sharing one host otherwise halve each other's D2H rate).
    Returns the node, or a short reason why no binding was made."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bus = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(path):
            return f"no {path}"
        node = int(open(path).read())
        if node < 0:
            return "numa_node=-1 (single node or not exposed)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: none of its CPUs is in this process's affinity mask"
        os.sched_setaffinity(0, cpus)
        return node
    except Exception as e:                               # noqa: BLE001 -- best effort
        return f"{type(e).__name__}: {e}"


def large_domain_suite(ch, rank, world):
    """BASELINE configs[4] on the ranks of this run: the n8192_k4 fixture of the unmodified reference is
    replayed on this world size first (parity), then N=8192 and N=16384 are timed on the slab path."""
    import torch
    W = (rank, world) if world > 1 else None
    out = {"n_gpus": world, "path": "row slabs + peer-memory transposes over NVLink (symmetric memory), one exchange launch "
                                    "per pass, sums gathered peer-to-peer" if world > 1 else "row slabs on one GPU"}
    z = np.load(os.path.join(ROOT, "tests", "golden", "n8192_k4.npz"))
    m = json.loads(str(z["meta"]))
    p = ch.Parameters()
    p.no_gui = True
    for k, v in m["params"].items():
        setattr(p, k, v)
    s = ch.Solver(p, _world=W)
    s.prepare()
    sol = s.solve_or_resume(p.ntmax)
    rows, ref = sol.timedata.data(), z["rows"]
    rel = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-300)
    rel[ref == 0] = np.abs(rows[ref == 0])
    st_ = p.N // 64
    du = float(np.abs(sol.U[::st_, ::st_] - z["U_sample"]).max())
    assert rel.max() < 1e-9 and du < 1e-11, ("large-domain parity", float(rel.max()), du)
    out["parity_n8192_k4"] = {"rows_max_rel": float(rel.max()), "U_max_abs": du, "fixture": "tests/golden/n8192_k4.npz (unmodified reference)"}
    del s, sol
    torch.cuda.empty_cache()
    for N, steps in ((8192, 30), (16384, 12)):
        out[f"N{N}"] = large_domain_probe(ch, N, steps, rank, world)
    return out


def large_domain_probe(ch, N, steps, rank=0, world=1):
    """One N x N domain, `steps` CH steps on the slab path over `world` ranks; device-timed (CUDA events, max
    over ranks); the exchange share is measured with events around the transposes in a second run."""
    import torch
    import torch.distributed as dist
    p = ch.Parameters()
    p.no_gui, p.full_sim, p.N = True, True, N
    p.kappa_tilde = 2.989112919661156e-4
    s = ch.Solver(p, _world=(rank, world) if world > 1 else None)
    s.prepare()
    eng = s._stepper
    eng.run(3)

    def timed(profile):
        eng._prof = [] if profile else None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.begin()
        e0.record()
        for i in range(steps):
            eng._step(last=(i == steps - 1))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        x = sum(a.elapsed_time(b) for a, b in eng._prof) if profile else 0.0
        eng._prof = None
        if world > 1:
            tt = torch.tensor([ms, x], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, x = float(tt[0]), float(tt[1])
        return ms, x
    l0 = eng.launch_count()
    ms, _ = timed(False)
    launches = eng.launch_count() - l0
    ms2, xms = timed(True)
    st = eng.get_state(0)
    out = {"workload": f"configs[4]: N={N} single domain, {steps} steps, {world} GPU(s)",
           "ms_per_step": round(ms / steps, 4), "steps_per_s": round(steps / (ms * 1e-3), 1),
           "algorithmic_gbs_32N2": round(32.0 * N * N * steps / (ms * 1e-3) / 1e9, 1),
           "exchange_share": round(xms / ms2, 3), "exchange_ms_per_step": round(xms / steps, 4),
           "nvlink_bytes_per_step_per_gpu": int(2 * 8 * N * N // world * (world - 1) // world),
           "launches_per_step": round(launches / steps, 1), "computed_steps": int(st.computed_steps)}
    del s, eng
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1024, help="ensemble members per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / single-sim probes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
