#!/usr/bin/env python
"""Developer measurement (not the bench): per-kernel-family times of a 3-D step with the Helmholtz EOS -- the Sod_stellar tube along x
on an n^3 box (rho 1e7 | 1e6 g/cc, T 1e8 | 1e6 K, ignition network: 3 species), `--gamma` runs the same box with the gamma law on
the same kernel family (CTU_NO_TMA=1) or on the tile-staged family for comparison.  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--solver", type=int, default=1)
    ap.add_argument("--gamma", action="store_true")
    ap.add_argument("--grav", action="store_true")
    args = ap.parse_args()
    import torch
    from castro_b200 import lib, abi
    from cases import IGNITION_NET, HELM_TABLE
    par = abi.default_params()
    par.riemann_solver = args.solver
    par.cfl, par.init_shrink, par.change_max = 0.8, 0.1, 1.1
    par.nspec = 3
    for m, (a, z) in enumerate(IGNITION_NET):
        par.aion[m], par.zion[m] = a, z
    if args.gamma:
        par.eos_gamma = 5.0 / 3.0
        prob = dict(p_l=1.e24, u_l=0.0, rho_l=1.0e7, p_r=1.e22, u_r=0.0, rho_r=1.e6, idir=1, frac=0.5)
    else:
        lib.helm_load(HELM_TABLE)
        par.eos_kind = 1
        prob = dict(T_l=1.e8, u_l=0.0, rho_l=1.0e7, T_r=1.e6, u_r=0.0, rho_r=1.e6, use_Tinit=1, idir=1, frac=0.5)
    if args.grav:
        par.do_grav, par.const_grav = 1, -1.e10
    n = args.n
    geom = abi.make_geom((n, n, n), probhi=(1.e6, 1.e6, 1.e6), lo_bc=(2, 3, 3), hi_bc=(2, 3, 3))
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), (n - 1, n - 1, n - 1))
    sim.init_problem("Sod", prob)
    rc, t, dt, hist = sim.advance(args.warmup)
    assert rc == 0
    ws = sim.ws_handle
    lib.workspace_set_profiling(ws, True)
    lib.workspace_stage_times(ws, reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc, t, dt, hist = sim.advance(args.steps)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0
    ms = e0.elapsed_time(e1) / args.steps
    st = lib.workspace_stage_times(ws, reset=True)
    import hashlib
    sha = hashlib.sha256(sim.download().tobytes()).hexdigest()[:16]
    print(json.dumps({"n": n, "state_sha256": sha, "eos": "gamma" if args.gamma else "helmholtz", "solver": args.solver, "grav": args.grav,
                      "tile_staged": bool(sim.tile_staged), "no_tma_env": os.environ.get("CTU_NO_TMA", ""),
                      "ms_per_step": ms, "zones_per_s": n ** 3 / (ms * 1e-3),
                      "kernels_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in st.items() if v[1]}}))
    sim.close()


if __name__ == "__main__":
    main()
