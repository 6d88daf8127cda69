#!/usr/bin/env python
"""bench.py -- zone-updates/s of the FP64 CTU-PPM hydro step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 512] [--solver 2]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A "step" is one full coarse time step of the single-level 3-D Sedov problem (512^3 zones per GPU,
gamma = 1.4, HLLC, outflow, cfl 0.5): clean_state(S_old) -> halo exchange (N > 1) -> physical-BC fill +
clean_state(Sborder) -> construct_ctu_hydro_source (+ fused check_for_negative_density, clean_state x2, estdt x2)
-> dt/status min-all-reduce -> host dt control; exactly what the parity tests check against the reference.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     : the whole fused hydro pipeline of one box against the HBM roofline, algorithmic bytes =
                 16*NUM_STATE B per zone-update (read U once, write U once), timed live with CUDA events
                 on the launching stream (the C ABI's per-stage event timers, summed); "kernels" gives each kernel
                 family's share so that the ncu launch list under profiles/ can be cross-checked
  cpu_baseline : the UNMODIFIED reference executable (oracle/_ref, OpenMP build) on this box's host cores, bounded sample
  e2e          : the same step driven through the C ABI with HOST buffers: pinned-host -> device upload of the state,
                 the step, device -> pinned-host download of the new state, all inside the timed region
`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "zone-updates/sec (FP64 CTU-PPM, 3D Sedov 512^3 per GPU)"
UNIT = "zone-updates/s"
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sedov_omp.ex")
REF_INPUTS = os.path.join(ROOT, "bench_inputs", "inputs.sedov3d")
HBM_FALLBACK_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ---- reference CPU arm ------------------------------------------------------------------------------------------
def run_reference_cpu(n_sample, steps, warmup, solver, threads=None, plot_dir=None):
    """Runs the unmodified reference executable on a bounded sample (n_sample^3 zones) for warmup+steps coarse steps and
    returns (zones/s over the timed steps, mean s/step, threads, kind, description).  With plot_dir the run also writes its final
    state as plot_dir/ref_pltNNNNN (after the last timed step; Amr::coarseTimeStep prints the step time before it writes), which
    the GPU arm compares its own run of the same inputs against (parity_check).  Falls back to the CPU port (liboracle, 1 thread)
    when the reference executable is not present."""
    threads = threads or (os.cpu_count() or 1)
    if os.path.exists(REF_EXE):
        wd = plot_dir or tempfile.mkdtemp(prefix="refbench_")
        mgs = 32 if n_sample >= 64 else 16
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        args = [REF_EXE, REF_INPUTS, "amr.n_cell=%d %d %d" % ((n_sample,) * 3), "amr.max_grid_size=%d" % mgs,
                "max_step=%d" % (warmup + steps), "castro.riemann_solver=%d" % solver]
        if plot_dir:
            args += ["amr.plot_files_output=1", "amr.plot_int=%d" % (warmup + steps), "amr.plot_file=%s" % os.path.join(plot_dir, "ref_plt"),
                     "amr.derive_plot_vars=NONE"]
        out = subprocess.run(args, cwd=wd, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("reference executable failed: " + out.stdout[-2000:] + out.stderr[-2000:])
        ts = [float(x) for x in re.findall(r"\[STEP \d+\] Coarse TimeStep time: ([0-9.eE+-]+)", out.stdout)]
        ts = ts[warmup:warmup + steps]
        assert len(ts) == steps, "could not parse per-step times of the reference run"
        tot = sum(ts)
        desc = "%d^3 Sedov (max_grid_size %d), %d timed coarse steps after %d warm-up, OMP_NUM_THREADS=%d" % (
            n_sample, mgs, steps, warmup, threads)
        return n_sample ** 3 * steps / tot, tot / steps, threads, "reference", desc
    # CPU port (single thread)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from cases import sedov_params
    from castro_b200 import abi
    n = min(n_sample, 32)
    par = sedov_params(solver)
    O.params_finalize(par)
    geom = abi.make_geom((n, n, n))
    U = O.sedov_init(par, geom, (n, n, n))
    st, t, dt, _ = O.run(par, geom, U, warmup)
    t0 = time.perf_counter()
    st, t, dt, _ = O.run(par, geom, U, steps, time=t, dt=dt)
    el = time.perf_counter() - t0
    return n ** 3 * steps / el, el / steps, 1, "port", "%d^3 Sedov, %d steps, oracle/liboracle.so single thread" % (n, steps)


REF_GPU_EXE = os.path.join(ROOT, "oracle", "_ref", "Castro3d_sedov_cuda.ex")


def run_reference_gpu(n, steps, warmup, solver, mgs=128):
    """The UNMODIFIED reference built with its own GPU back end (oracle/build_ref.sh gpu: USE_CUDA=TRUE CUDA_ARCH=100) on this
    box's GPU 0: Castro's stock unfused ParallelFor pipeline, the same inputs file, n^3 zones.  -> dict or None."""
    if not os.path.exists(REF_GPU_EXE):
        return None
    wd = tempfile.mkdtemp(prefix="refgpu_")
    a = [REF_GPU_EXE, REF_INPUTS, "amr.n_cell=%d %d %d" % ((n,) * 3), "amr.max_grid_size=%d" % mgs, "amr.blocking_factor=32",
         "max_step=%d" % (warmup + steps), "castro.riemann_solver=%d" % solver]
    try:
        out = subprocess.run(a, cwd=wd, env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]),
                             capture_output=True, text=True, timeout=600)
    except Exception as e:      # noqa: BLE001
        return {"unavailable": "run failed: %r" % (e,)}
    if out.returncode != 0:
        return {"unavailable": "exit %d: %s" % (out.returncode, (out.stdout[-300:] + out.stderr[-300:]).replace("\n", " | "))}
    ts = [float(x) for x in re.findall(r"\[STEP \d+\] Coarse TimeStep time: ([0-9.eE+-]+)", out.stdout)][warmup:warmup + steps]
    if len(ts) != steps:
        return {"unavailable": "could not parse step times"}
    tot = sum(ts)
    return {"value": n ** 3 * steps / tot, "unit": UNIT, "ms_per_step": 1e3 * tot / steps,
            "kind": "reference, its own CUDA back end (USE_CUDA=TRUE CUDA_ARCH=100, stock code path, --use_fast_math as its build system sets)",
            "sample": "%d^3 Sedov, max_grid_size %d, %d timed coarse steps after %d warm-up, whole coarse step (as [STEP] lines report), "
                      "one GPU" % (n, mgs, steps, warmup)}


def reference_arm(args, rank):
    if rank != 0:
        return
    val, s_per_step, threads, kind, desc = run_reference_cpu(args.cpu_n, args.steps, args.warmup, args.solver)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus, cpu_sample=desc),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, cpu_sample=None):
    from castro_b200.dist import proc_grid
    dims = proc_grid(world)
    prob = getattr(args, "problem", "sedov")
    if getattr(args, "scaling", "weak") == "strong":
        w = ("Exec/hydro_tests/%s 3-D, global %d^3 zones cut over %d GPU(s) as %s, gamma-law EOS, CTU-PPM + %s, single level" % (
            {"sedov": "Sedov", "kh": "KH (NUMADV=1 + %d species)" % args.nspec, "rt": "RT"}[prob], args.global_n, world,
            "x".join(str(d) for d in dims), "HLLC" if args.solver == 2 else "CGF"))
    elif prob != "sedov":
        w = ("Exec/hydro_tests/%s 3-D, %d^3 zones per GPU, gamma-law EOS, CTU-PPM + %s, single level (developer option, not the "
             "benchmark configuration)" % (prob.upper(), args.n, "HLLC" if args.solver == 2 else "CGF"))
    else:
        w = ("Exec/hydro_tests/Sedov 3-D, %d^3 zones per GPU (global %s), gamma-law EOS gamma=1.4, CTU-PPM + %s, "
             "single level, outflow BCs, cfl 0.5" % (args.n, "x".join(str(args.n * d) for d in dims), "HLLC" if args.solver == 2 else "CGF"))
    if cpu_sample:      # the reference CPU arm times a bounded sample of that workload (its zone rate does not depend on the box size)
        w += "; TIMED HERE ON THE HOST CORES AS A BOUNDED SAMPLE: " + cpu_sample
    return {"workload": w,
            "zones_per_gpu": (args.global_n ** 3 // world) if getattr(args, "scaling", "weak") == "strong" else args.n ** 3, "num_state": 7 + getattr(args, "nspec", 1) + (1 if getattr(args, "problem", "sedov") == "kh" else 0),
            "problem": getattr(args, "problem", "sedov"), "proc_grid": list(dims),
            "l2": "inputs larger than L2 (state %.1f GB per GPU); no explicit flush" % (
                8 * (7 + getattr(args, "nspec", 1) + (1 if getattr(args, "problem", "sedov") == "kh" else 0)) * (args.n + 8) ** 3 / 1e9)}


# ---- clocks -------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---- checks that travel with the bench line ----------------------------------------------------------------------------
def parity_check(par_factory, n, steps, solver, ref_plot_dir):
    """The GPU path on the inputs the reference CPU arm just ran (n^3 Sedov, `steps` coarse steps), judged by the reference's own
    comparator: fcompare --rel_tol 1e-12 between the reference's plotfile and ours.  -> dict for the JSON line."""
    from castro_b200 import lib
    fcmp = os.path.join(ROOT, "oracle", "_ref", "fcompare.ex")
    ref_pf = os.path.join(ref_plot_dir, "ref_plt%05d" % steps)
    if not (os.path.exists(fcmp) and os.path.isdir(ref_pf)):
        return {"n": n, "steps": steps, "max_rel": None, "agree": None, "note": "reference plotfile or fcompare missing"}
    over = "amr.n_cell=%d %d %d castro.riemann_solver=%d" % (n, n, n, solver)
    par, geom, nc, _, _, prob = lib.parse_inputs(REF_INPUTS, over)
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in nc))
    sim.init_problem("Sedov", {k: (int(v) if k == "nsub" else float(v)) for k, v in (kv.split("=") for kv in prob.split())})
    rc, t, dt, hist = sim.advance(steps)
    my_pf = os.path.join(ref_plot_dir, "b200_plt%05d" % steps)
    sim.write_plotfile(my_pf, t, steps)
    sim.close()
    res = subprocess.run([fcmp, "--rel_tol", "1e-12", "-a", ref_pf, my_pf], capture_output=True, text=True)
    rows = re.findall(r"^\s*(\w+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s*$", res.stdout, flags=re.M)
    return {"n": n, "steps": steps, "status": rc, "fields": len(rows),
            "max_rel": max(float(r[2]) for r in rows) if rows else None, "max_abs": max(float(r[1]) for r in rows) if rows else None,
            "agree": res.returncode == 0 and "PLOTFILE AGREE" in res.stdout and len(rows) == 8,
            "tool": "oracle/_ref/fcompare.ex --rel_tol 1e-12 (reference executable's plotfile vs this library's, same inputs file)"}


def decomp_check(torch, dist, rank, world, dims, args, n_global=128, steps=5):
    """N > 1: a 128^3 global box split over the N ranks (halo exchange over NCCL) against the same box run whole on every rank:
    each rank compares its own part bit for bit; the verdicts are AND-reduced."""
    from castro_b200 import lib, abi
    from castro_b200.dist import box_of, rank_coords
    par = lib.default_params()
    par.riemann_solver = args.solver
    par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
    par.eos_gamma, par.eos_assume_neutral = 1.4, 1
    lib.params_finalize(par)
    prob = dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10)
    n_cell = (n_global,) * 3
    geom = abi.make_geom(n_cell)
    whole = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    whole.init_problem("Sedov", prob)
    rc0, t0, dt0, hist0 = whole.advance(steps)
    U = whole.download()
    whole.close()
    ds = lib.Dist(par, geom, dims, rank=rank, world=world, comm=dist)
    ds.init_problem("Sedov", prob)
    st, hist = ds.advance(steps)
    mine = ds.sim.download()
    lo, hi = box_of(rank_coords(rank, dims), dims, n_cell)
    ref = U[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
    ok = (st == 0 and rc0 == 0 and list(hist) == list(hist0) and mine.shape == ref.shape and bool((mine == ref).all()))
    ds.close()
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"verdict": "bitwise" if float(t.item()) == 1.0 else "MISMATCH", "global_box": "%d^3" % n_global, "steps": steps,
            "proc_grid": list(dims), "halo_exchange": "ctu_dist_* (C++): fused pack, one ncclGroup of send/recv, fused unpack; %d ranks" % world}


# ---- GPU arm ------------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "--gpus %d but WORLD_SIZE %d" % (args.gpus, world)

    from castro_b200 import lib, abi
    from castro_b200.dist import proc_grid
    sys.path.insert(0, os.path.join(ROOT, "tests"))

    dims = proc_grid(world)
    n = args.n
    strong = args.scaling == "strong"
    if strong:       # fixed global box (BASELINE config[3]: KH 1024^3 over the GPUs of the node), cut into equal boxes
        assert all(args.global_n % d == 0 for d in dims)
        n_cell = (args.global_n,) * 3
    else:
        n_cell = tuple(n * d for d in dims)
    box = tuple(n_cell[a] // dims[a] for a in range(3))          # zones of one rank's box (x, y, z)
    par = lib.default_params()
    par.riemann_solver = args.solver
    par.cfl, par.init_shrink, par.change_max = 0.5, 0.01, 1.1
    par.eos_gamma, par.eos_assume_neutral = 1.4, 1
    if args.nspec != 1:          # developer option (NOT the benchmark configuration): a multi-species network
        par.nspec = args.nspec
        for m in range(args.nspec):
            par.aion[m], par.zion[m] = 1.0 + 3.0 * m, 1.0 + m
    problem = "Sedov"
    geom_kw = {}
    prob = dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10)
    if args.problem == "kh":      # developer option: bench_inputs/inputs.kh3d at n^3 (NUM_STATE = 9)
        problem = "KH"
        par.cfl, par.init_shrink, par.change_max = 0.5, 0.1, 1.1
        par.small_dens, par.small_temp = 1.e-10, 1.e-10
        par.eos_gamma = 5.0 / 3.0
        par.nadv, par.nspec = 1, args.nspec
        if args.nspec == 4:       # networks/general_null/triple_alpha_plus_o.net: He4 C12 O16 Fe56 (BASELINE config[3])
            for m, (a_, z_) in enumerate(((4.0, 2.0), (12.0, 6.0), (16.0, 8.0), (56.0, 26.0))):
                par.aion[m], par.zion[m] = a_, z_
        geom_kw = dict(periodic=(1, 1, 1), lo_bc=(0, 0, 0), hi_bc=(0, 0, 0))
        prob = dict(problem=3, rho1=1.0, rho2=2.0, pressure=2.5, bulk_velocity=0.0)
    elif args.problem == "rt":    # developer option: bench_inputs/inputs.rt3d on a cube of n^3 zones, constant gravity, grav_source_type 4
        problem = "RT"
        par.cfl, par.init_shrink, par.change_max = 0.8, 0.3, 1.1
        par.eos_gamma = 5.0 / 3.0
        par.do_grav, par.grav_source_type, par.const_grav = 1, 4, -1.0
        geom_kw = dict(periodic=(1, 1, 0), lo_bc=(0, 0, 4), hi_bc=(0, 0, 4))
        prob = dict(frac=0.5, rho_1=1.0, rho_2=2.0, p0_base=5.0)
    for kv in args.par:          # developer option: any castro.* field of ctu_params_t, e.g. --par ppm_type=0 --par use_pslope=1
        k, v = kv.split("=", 1)
        setattr(par, k, type(getattr(par, k))(float(v)))
    lib.params_finalize(par)
    # weak scaling: dx = 1/n whatever the process grid; strong scaling: the unit cube
    geom = abi.make_geom(n_cell, probhi=(1.0, 1.0, 1.0) if strong else tuple(float(d) for d in dims), **geom_kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the decomposed run must reproduce the single-box run bit for bit (before anything is timed)
    dcheck = decomp_check(torch, dist, rank, world, dims, args) if (dist is not None and not args.no_checks) else None

    # the C++ multi-box driver (ctu_dist_*): one box per rank, halo exchange + dt/status all-reduce over NCCL inside the library
    sim = lib.Dist(par, geom, dims, rank=rank, world=world, comm=dist, workspace_bytes=args.workspace_gb << 30)
    sim.init_problem(problem, prob)
    eng = sim.sim
    ws = eng.ws_handle
    zones_local = box[0] * box[1] * box[2]
    zones_total = zones_local * world

    # ---- warm-up -------------------------------------------------------------------------------------------------
    st, _ = sim.advance(args.warmup)
    assert st == 0, "warm-up step failed with status %d" % st
    barrier()

    # ---- timed region: K steps, device-resident state --------------------------------------------------------------
    lib.workspace_set_profiling(ws, True)
    lib.workspace_stage_times(ws, reset=True)
    l0 = eng.launches
    r0 = lib.workspace_replays(ws)
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    st, hist = sim.advance(args.steps)
    e1.record()
    barrier()
    clk = clocks.stop() if clocks else None
    assert st == 0 and len(hist) == args.steps, "timed steps failed with status %d" % st
    ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    replays = lib.workspace_replays(ws) - r0
    stages = lib.workspace_stage_times(ws, reset=True)
    lib.workspace_set_profiling(ws, False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = zones_total * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + step + D2H per step ------------------------------------------
    # ctu_dist_step_host streams this rank's box through the device in z-chunks (copy-in, kernels, copy-out overlap); at N > 1 the
    # halo exchange runs inside the pipeline (boundary planes first, then every chunk's x/y halos as soon as it is on the device).
    from castro_b200.dist import new_dt
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    nbytes = 8 * par.num_state * zones_local
    do_e2e = args.e2e_steps > 0 and 2 * nbytes <= (args.e2e_max_gb << 30)     # two pinned copies of the box per rank must fit the host
    if not do_e2e:
        e2e_steps = 0
    host = torch.empty((2, par.num_state, box[2], box[1], box[0]) if do_e2e else (2, 1, 1, 1, 1 << 27), dtype=torch.float64, pin_memory=True)
    hin, hout = host[0].numpy(), host[1].numpy()
    if do_e2e:
        eng.download(out=hin)
    streamed = not par.do_grav                      # the host-streamed step covers pure hydro
    e2e_mode = "streamed z-chunks (ctu_dist_step_host%s)" % (", halo exchange inside the pipeline" if world > 1 else "") if streamed \
        else "upload + step + download"

    def e2e_step():
        nonlocal hin, hout
        if streamed:
            st_, est = sim.step_host(hin, hout, (sim.time + sim.dt) - sim.time)
            assert st_ == 0
            sim.time += sim.dt
            sim.dt = new_dt(par, est[1], sim.dt, sim.time, -1.0)
            hin, hout = hout, hin
        else:
            eng.upload(hin)                 # pinned host -> device (valid zones of this rank's box)
            st_, _ = sim.advance(1)
            assert st_ == 0
            eng.download(out=hin)           # device -> pinned host
    if do_e2e:
        e2e_step()                          # one untimed step (allocates the staging buffers)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    t = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = zones_total * e2e_steps / (e2e_ms * 1e-3) if do_e2e else None
    # what the host side can deliver when all N ranks copy at once: 1 GiB up and 1 GiB down per rank, concurrently (explains e2e)
    pc_n = min(1 << 27, host.numel() // 2)
    pc_h = host.view(-1)[:2 * pc_n]
    pc_d = torch.empty(2 * pc_n, dtype=torch.float64, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    s_up.wait_stream(torch.cuda.current_stream()); s_dn.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_up):
        pc_d[:pc_n].copy_(pc_h[:pc_n], non_blocking=True)
    with torch.cuda.stream(s_dn):
        pc_h[pc_n:].copy_(pc_d[pc_n:], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
    g1.record()
    barrier()
    t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcie_gbs = 8.0 * pc_n / (float(t.item()) * 1e-3) / 1e9          # per rank, each direction, all ranks copying both ways at once
    del pc_d

    if rank == 0:
        peak, peak_src = measured_peaks()
        hydro_ms = sum(v[0] for v in stages.values())
        hydro_launches = sum(v[1] for v in stages.values())
        alg_bytes = 16 * par.num_state          # B per zone-update: read U once + write U once
        achieved = zones_local * args.steps * alg_bytes / (hydro_ms * 1e-3) / 1e9 if hydro_ms > 0 else 0.0
        top = max(stages, key=lambda k: stages[k][0])
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src,
            "kernel": "ctu_hydro_advance pipeline (all kernels of construct_ctu_hydro_source for one box)",
            "algorithmic_bytes_per_zone": alg_bytes,
            "pipeline_ms_per_step": hydro_ms / args.steps, "pipeline_launches_per_step": hydro_launches / args.steps,
            "top_kernel": top,
            "kernels": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                            "share": (v[0] / hydro_ms if hydro_ms else 0.0)} for k, v in stages.items() if v[1]},
            "note": "FP64-pipe bound, not HBM bound, once fused (SURVEY §8d); see DESIGN.md and profiles/",
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                # ncu-measured DRAM bytes and FP64-pipe busy time of ONE coarse step of this workload (profiles/r2_ncu_raw_512.csv ->
                # tools/ncu_traffic.py -> profiles/traffic.json, with the commit it was captured at), beside the algorithmic bytes
                tj = json.load(open(traffic_file))
                tpz = tj.get("pipeline_dram_bytes_per_zone")
                roofline["traffic"] = tpz * zones_local
                roofline["traffic_bytes_per_zone"] = tpz
                roofline["traffic_source"] = {"file": "profiles/traffic.json", "raw": tj.get("source"), "captured_at_commit": tj.get("commit"),
                                              "workload": tj.get("workload")}
                roofline["algorithmic_bytes"] = alg_bytes * zones_local
                fam = {"k_ctoprim": "k_ctoprim", "k_flat_div": "k_flat_div", "k_trace": "k_trace_riemann_tma", "k_trans_riemann": "k_trans2_tma",
                       "k_final": "k_final_tma", "k_update": "k_update_tma"}
                for k, v in roofline["kernels"].items():
                    t = tj.get("kernels", {}).get(fam.get(k, ""), None)
                    if t:
                        v["bytes_per_zone"] = t["bytes_per_zone"]
                        v["dram_GBps_live"] = t["bytes_per_zone"] * zones_local / (v["ms_per_step"] * 1e-3) / 1e9 if v["ms_per_step"] else None
                        v["hbm_frac_live"] = v["dram_GBps_live"] / peak if v["dram_GBps_live"] else None
                        v["fp64_pipe_pct_ncu"] = t["fp64_pipe_pct"]
                        v["l2_hit_pct_ncu"] = t["l2_hit_pct"]
                if "fp64_pipe_busy_ns_per_zone" in tj and hydro_ms > 0:
                    busy_ms = tj["fp64_pipe_busy_ns_per_zone"] * zones_local * 1e-6
                    roofline["fp64_pipe"] = {"busy_ms_per_step": busy_ms, "frac": busy_ms / (hydro_ms / args.steps),
                                             "note": "fraction of the pipeline time the FP64 pipe would need at 100 % utilisation (ncu)"}
                    roofline["binding"] = ("k_final: HBM (its ncu bytes over its live time: see kernels.k_final.hbm_frac_live); k_trace / k_trans2: "
                                           "mixed FP64 + integer/select instruction issue at the 16-27 warps/SM the register file allows; "
                                           "k_ctoprim / k_update: HBM")
            except Exception:
                pass
        cpu = None
        pcheck = None
        if world == 1 and not args.no_cpu:
            # the reference executable on the host cores (bounded sample); its final plotfile then judges the GPU path on the same inputs
            pdir = tempfile.mkdtemp(prefix="parity_")
            v, spp, threads, kind, desc = run_reference_cpu(args.cpu_n, args.cpu_steps, 1, args.solver, plot_dir=pdir)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc}
            if kind == "reference" and not args.no_checks and args.problem == "sedov" and args.nspec == 1:
                pcheck = parity_check(None, args.cpu_n, args.cpu_steps + 1, args.solver, pdir)
            import shutil
            shutil.rmtree(pdir, ignore_errors=True)
        line = {
            "metric": METRIC if (args.problem == "sedov" and not strong) else
            "zone-updates/sec (FP64 CTU-PPM, 3D %s, %s)" % (args.problem.upper(), "global %d^3 strong scaling" % args.global_n if strong
                                                             else "%d^3 per GPU" % n),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps if do_e2e else None, "mode": e2e_mode,
                    "host_copy_gbs_per_rank_each_way": pcie_gbs,
                    "copy_floor_ms_per_step": nbytes / (pcie_gbs * 1e9) * 1e3,
                    "note": "host_copy_gbs = measured here with all N ranks copying 1 GiB up and 1 GiB down at once from pinned memory; "
                            "copy_floor = this step's bytes at that rate"},
            "gpu_launches": launches, "clocks": clk,
            "halo_bytes_per_step_per_gpu": sim.halo_bytes_per_step,
            "exact_replay_tiles_per_step": replays / args.steps,
            "parity_check": pcheck, "decomp_check": dcheck,
        }
    sim.close()
    if rank == 0:
        if world == 1 and not args.no_cpu and not args.no_ref_gpu:
            # Castro's own GPU build on the same B200 (after this arm's device memory has been released)
            torch.cuda.empty_cache()
            line["reference_gpu"] = run_reference_gpu(args.ref_gpu_n, args.steps, 3, args.solver)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="zones per GPU per dimension")
    ap.add_argument("--solver", type=int, default=2, help="castro.riemann_solver: 2 = HLLC, 0 = CGF")
    ap.add_argument("--workspace-gb", type=int, default=48, help="device scratch bound for the z-slab pipeline")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-max-gb", type=int, default=24, help="skip the e2e leg when two pinned host copies of a rank's box exceed this")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --n^3 zones per GPU (the benchmark); strong: a fixed --global-n^3 box cut over the GPUs (config[3])")
    ap.add_argument("--global-n", type=int, default=1024)
    ap.add_argument("--cpu-n", type=int, default=128, help="edge of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the run of the reference's own CUDA build (reference_gpu)")
    ap.add_argument("--ref-gpu-n", type=int, default=512, help="edge of the box the reference's CUDA build is timed on")
    ap.add_argument("--no-checks", action="store_true", help="skip parity_check (N = 1) / decomp_check (N > 1)")
    ap.add_argument("--problem", default="sedov", choices=["sedov", "kh", "rt"],
                    help="developer option: kh = Kelvin-Helmholtz as shipped (advected scalar + species, periodic), rt = Rayleigh-Taylor with "
                         "constant gravity (cube of n^3 zones); the benchmark configuration is sedov")
    ap.add_argument("--par", action="append", default=[], help="developer option: override a ctu_params_t field (name=value), repeatable")
    ap.add_argument("--nspec", type=int, default=1, help="developer option: species in the network (default 1 = gammalaw.net, the benchmark)")
    ap.add_argument("--profile-mode", action="store_true", help="allow warmup < 3 (ncu runs only; never a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and not args.profile_mode:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args, int(os.environ.get("RANK", "0")))
        return
    gpu_arm(args)


if __name__ == "__main__":
    main()
