"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libcastro_ctu.so; the CPU oracle (oracle/ctu_oracle.cpp, pinned bit-for-bit to the reference executable by
tests/test_oracle_golden.py) and the committed golden fixtures are the checkers.

Tolerance: the north star asks for fcompare rel-Linf <= 1e-12 per conserved field.  The kernels are compiled
with -fmad=false and restate the reference's operation order, so the tests demand BIT-EXACT agreement and
report the rel-Linf they would have had otherwise."""
import numpy as np
import pytest

from cases import CASES, load_case

pytestmark = pytest.mark.gpu

NG = 4
REL_TOL = 1.0e-12   # north-star tolerance (fcompare: max|a-b| / max|a| per component)


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU test needs a CUDA device"
    return torch


def rel_linf(a, b, vector_momentum=False):
    """fcompare's relative L-inf per field: max|a - b| / max|b|.  vector_momentum: the three momentum components are measured against
    the largest of them (used where a transverse component is a small difference of pressures that themselves agree to ~1e-13: the
    Helmholtz EOS with gravity across the tube; everywhere else every field is measured against itself)."""
    scale = [max(np.abs(b[n]).max(), 1e-300) for n in range(a.shape[0])]
    if vector_momentum:
        scale[1] = scale[2] = scale[3] = max(scale[1:4])
    return [float(np.abs(a[n] - b[n]).max() / scale[n]) for n in range(a.shape[0])]


def assert_bitwise(a, b, what):
    nbad = [int((a[n] != b[n]).sum()) for n in range(a.shape[0])]
    rl = rel_linf(a, b)
    assert max(rl) <= REL_TOL, "%s: rel-Linf %s exceeds %g" % (what, rl, REL_TOL)
    assert sum(nbad) == 0, "%s: not bit-exact: zones differing per comp %s, rel-Linf %s" % (what, nbad, rl)


@pytest.mark.parametrize("name", sorted(CASES))
def test_sim_matches_reference_golden(name):
    """full time loop (IC, ghost fill, clean_state x4, hydro, dt control) vs the reference executable's states"""
    from castro_b200 import lib
    _torch()
    par, geom, prob, g, snaps = load_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("Sedov", prob)
    t, dt, step = 0.0, -1.0, 0
    times = [0.0]
    for s in snaps:
        if s > step:
            rc, t, dt, hist = sim.advance(s - step, time=t, dt=dt)
            assert rc == 0
            tt = times[-1]
            for h in hist:
                tt += h
                times.append(tt)
            step = s
        assert_bitwise(sim.download(), g["U%d" % s], "%s step %d" % (name, s))
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    assert sim.launches > 0
    sim.close()


@pytest.mark.parametrize("solver,n_cell,nslab_bytes", [(2, (16, 16, 16), 0), (0, (16, 16, 16), 0), (2, (24, 16, 20), 6 << 20)])
def test_hydro_advance_vs_oracle(solver, n_cell, nslab_bytes):
    """one call of ctu_hydro_advance on an evolved state: S_new, flux registers, mass fluxes, dt estimate, status"""
    torch = _torch()
    import oracle_lib as O
    from castro_b200 import lib, abi
    from cases import sedov_params
    par = sedov_params(solver)
    O.params_finalize(par)
    geom = abi.make_geom(n_cell, probhi=tuple(n / 16.0 for n in n_cell))
    U = O.sedov_init(par, geom, n_cell, r_init=0.08, nsub=4)
    st, t, dt, _ = O.run(par, geom, U, 6)          # evolve on the CPU so that velocities are non-trivial
    assert st == 0
    hi = tuple(n - 1 for n in n_cell)
    glo = (-NG, -NG, -NG)
    O.clean_state(par, U, glo, (0, 0, 0), hi)
    O.fill_boundary(par, geom, U, (0, 0, 0), hi)
    O.clean_state(par, U, glo, glo, tuple(h + NG for h in hi))
    ns = par.num_state
    Sn_ref = np.ascontiguousarray(U[:, NG:-NG, NG:-NG, NG:-NG])

    def flux_arrays(nc, init):
        return [np.full((nc, n_cell[2] + (d == 2), n_cell[1] + (d == 1), n_cell[0] + (d == 0)), init) for d in range(3)]
    F_ref = flux_arrays(ns, 0.0)
    st_ref, dt_ref, loc_ref, nbad_ref = O.hydro_advance(par, geom, (0, 0, 0), hi, U, glo, Sn_ref, (0, 0, 0), dt, do_clean=1,
                                                        fluxes=F_ref, flux_los=[(0, 0, 0)] * 3)
    dSb = torch.from_numpy(U).cuda()
    dSn = torch.from_numpy(np.ascontiguousarray(U[:, NG:-NG, NG:-NG, NG:-NG])).cuda()
    dF = [torch.from_numpy(a).cuda() for a in flux_arrays(ns, 0.25)]     # the registers accumulate (+=)
    dM = [torch.from_numpy(a).cuda() for a in flux_arrays(1, 9.0)]       # the mass fluxes are overwritten (=)
    ws = lib.Workspace(nslab_bytes)
    rc, est, loc, nbad = lib.hydro_advance(ws, par, geom, (0, 0, 0), hi, dSb, glo, dSn, (0, 0, 0), dt, n_clean=1,
                                           fluxes=dF, flux_los=[(0, 0, 0)] * 3, mass_fluxes=dM, mass_los=[(0, 0, 0)] * 3,
                                           want_loc=True)
    torch.cuda.synchronize()
    if nslab_bytes:
        assert ws.last_slabs > 1, "test is meant to exercise the multi-slab path"
    assert rc == st_ref and nbad == nbad_ref
    assert_bitwise(dSn.cpu().numpy(), Sn_ref, "S_new")
    for d in range(3):
        assert_bitwise(dF[d].cpu().numpy(), 0.25 + F_ref[d], "flux register %d" % d)
        np.testing.assert_array_equal(dM[d].cpu().numpy()[0], F_ref[d][0])
    assert est[0] == dt_ref and est[1] == dt_ref
    assert loc == loc_ref
    # same answer when S_new's input content is ignored
    dSn2 = torch.zeros_like(dSn)
    lib.hydro_advance(ws, par, geom, (0, 0, 0), hi, dSb, glo, dSn2, (0, 0, 0), dt, n_clean=1, from_sborder=True)
    assert torch.equal(dSn2, dSn)


def test_pointwise_entry_points_vs_oracle():
    """ctu_clean_state, ctu_fill_boundary (outflow, periodic, symmetry), ctu_estdt, pack/unpack"""
    torch = _torch()
    import oracle_lib as O
    from castro_b200 import lib, abi
    from cases import sedov_params
    par = sedov_params(2)
    O.params_finalize(par)
    n_cell = (12, 10, 8)
    hi = tuple(n - 1 for n in n_cell)
    glo = (-NG, -NG, -NG)
    rng = np.random.default_rng(7)
    for bc in [dict(), dict(periodic=(1, 1, 1), lo_bc=(0, 0, 0), hi_bc=(0, 0, 0)),
               dict(periodic=(1, 0, 0), lo_bc=(0, 3, 2), hi_bc=(0, 4, 3))]:
        geom = abi.make_geom(n_cell, **bc)
        U = O.sedov_init(par, geom, n_cell, r_init=0.2, nsub=2)
        v = U[:, NG:-NG, NG:-NG, NG:-NG]
        v[1:4] = rng.standard_normal(v[1:4].shape) * 0.3          # random momenta
        v[0] = 0.5 + rng.random(v[0].shape)
        v[7] = v[0] * (1.0 + 1e-3 * rng.standard_normal(v[0].shape))   # species slightly off rho
        v[4] = v[5] + 0.5 * (v[1] ** 2 + v[2] ** 2 + v[3] ** 2) / v[0]
        v[0, 1, 2, 3] = 1e-120                                    # triggers enforce_min_density
        U[:, :NG] = 777.0                                         # garbage in ghosts
        d = torch.from_numpy(U).cuda()
        lib.clean_state(par, (0, 0, 0), hi, d, glo, npass=1)
        O.clean_state(par, U, glo, (0, 0, 0), hi)
        lib.fill_boundary(par, geom, (0, 0, 0), hi, d)
        O.fill_boundary(par, geom, U, (0, 0, 0), hi)
        ghi = tuple(h + NG for h in hi)
        lib.clean_state(par, glo, ghi, d, glo, npass=2)
        O.clean_state(par, U, glo, glo, ghi, npass=2)
        assert_bitwise(d.cpu().numpy(), U, "clean+fill bc=%s" % (bc,))
        dt_g, loc_g = lib.estdt(par, geom, (0, 0, 0), hi, d, glo)
        dt_o, loc_o = O.estdt(par, geom, U, glo, (0, 0, 0), hi)
        assert dt_g == dt_o and loc_g == loc_o
    # pack / unpack round trip of a sub-box
    lo, hi2 = (-2, 1, 0), (5, 6, 3)
    n = par.num_state * np.prod([hi2[k] - lo[k] + 1 for k in range(3)])
    buf = torch.empty(int(n), dtype=torch.float64, device="cuda")
    lib.halo_pack(d, glo, lo, hi2, buf)
    ref = U[:, lo[2] + NG:hi2[2] + NG + 1, lo[1] + NG:hi2[1] + NG + 1, lo[0] + NG:hi2[0] + NG + 1]
    np.testing.assert_array_equal(buf.cpu().numpy().reshape(ref.shape), ref)
    d2 = torch.zeros_like(d)
    lib.halo_unpack(d2, glo, lo, hi2, buf)
    sub = d2.cpu().numpy()[:, lo[2] + NG:hi2[2] + NG + 1, lo[1] + NG:hi2[1] + NG + 1, lo[0] + NG:hi2[0] + NG + 1]
    np.testing.assert_array_equal(sub, ref)
    assert float(d2.sum()) == float(ref.sum())


def test_slab_count_does_not_change_results():
    """size-independent property: the z-slab decomposition of the workspace is invisible in the results"""
    _torch()
    from castro_b200 import lib
    par, geom, prob, g, snaps = load_case("sedov32_hllc")
    lib.params_finalize(par)
    outs = []
    for wsb in (0, 24 << 20):
        sim = lib.Sim(par, geom, (0, 0, 0), (31, 31, 31), workspace_bytes=wsb)
        sim.init_problem("Sedov", prob)
        rc, t, dt, hist = sim.advance(10)
        assert rc == 0
        outs.append((sim.download(), sim.last_slabs))
        sim.close()
    assert outs[0][1] == 1 and outs[1][1] > 1
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    assert_bitwise(outs[1][0], g["U10"], "sedov32 multi-slab")


def test_register_staged_kernels_match_golden():
    """CTU_NO_TMA=1 selects the register-staged kernels (the path other passive counts / EOSs take); it must agree bit for
    bit with the golden states as well.  Run in a subprocess because the switch is read once per process."""
    import os
    import subprocess
    import sys
    _torch()
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from cases import load_case\n"
        "from castro_b200 import lib\n"
        "for name in ('sedov16_hllc', 'sedov_24x16x8_cgf', 'sedov16_hllc_eos', 'sedov16_hllc_hybrid', 'sedov_24x16x8_cg', 'sedov16_hllc_limit', 'sedov16_plm_hllc', 'sedov16_plm_cgf_lim1', 'sedov16_cgf_rhoe', 'sedov16_cgf_tempfix'):\n"
        "    par, geom, prob, g, snaps = load_case(name)\n"
        "    lib.params_finalize(par)\n"
        "    n = tuple(int(x) for x in g['n_cell'])\n"
        "    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))\n"
        "    sim.init_problem('Sedov', prob)\n"
        "    rc, t, dt, hist = sim.advance(snaps[-1])\n"
        "    assert rc == 0\n"
        "    assert np.array_equal(sim.download(), g['U%%d' %% snaps[-1]]), name\n"
        "print('ok')\n") % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, CTU_NO_TMA="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_kernel_family_selection():
    """Which option sets run the tile-staged (TMA) kernels: the defaults, Colella & Glaz, piecewise-linear reconstruction, the
    hydrostatic-pressure-aware reconstructions (their own instantiations of the trace kernel), the hybrid Riemann solver and the flux
    limiter ("extended" instantiations); ppm_temp_fix, transverse_reset_rhoe and combinations of several of these stay on the
    register-staged family.  Each case also lands bit for bit on its golden."""
    _torch()
    from castro_b200 import lib
    from cases import load_rt_case
    expect = {"sedov16_hllc": True, "sedov16_cg": True, "sedov16_plm_hllc": True, "sedov16_plm_cgf_lim1": True, "sedov16_hllc_eos": False,
              "sedov16_hllc_hybrid": True, "sedov16_cgf_hybrid": True, "sedov16_hllc_limit": True, "sedov16_cgf_tempfix": False,
              "sedov16_combo": False}
    for name, tile in expect.items():
        par, geom, prob, g, snaps = load_case(name)
        lib.params_finalize(par)
        n = tuple(int(x) for x in g["n_cell"])
        sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))
        sim.init_problem("Sedov", prob)
        rc, t, dt, hist = sim.advance(snaps[-1])
        assert rc == 0
        assert sim.tile_staged == tile, name
        assert_bitwise(sim.download(), g["U%d" % snaps[-1]], name)
        sim.close()
    for name in ("rt16_pslope", "rt16_wellbal", "rt16_plm_pslope"):      # with traced / half-step gravity sources
        par, geom, g, snaps = load_rt_case(name)
        lib.params_finalize(par)
        n = tuple(int(x) for x in g["n_cell"])
        sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))
        sim.init_problem("RT", dict(frac=0.5, rho_1=1.0, rho_2=2.0, p0_base=5.0))
        rc, t, dt, hist = sim.advance(snaps[-1])
        assert rc == 0 and sim.tile_staged, name
        assert_bitwise(sim.download(), g["U%d" % snaps[-1]][:par.num_state], name)
        sim.close()


def test_persistent_trans2_variant_matches_golden():
    """CTU_PERSIST=1 selects the persistent / prefetching variant of the transverse-flux kernel (dynamic tile scheduler, exact replays
    through a list); not the default (it measured slower), but it must stay bit-identical.  Subprocess: the switch is read once."""
    import os
    import subprocess
    import sys
    _torch()
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from cases import load_case\n"
        "from castro_b200 import lib\n"
        "for name in ('sedov16_hllc_long', 'sedov32_hllc', 'sedov_24x16x8_cgf', 'sedov16_periodic_hllc'):\n"
        "    par, geom, prob, g, snaps = load_case(name)\n"
        "    lib.params_finalize(par)\n"
        "    n = tuple(int(x) for x in g['n_cell'])\n"
        "    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))\n"
        "    sim.init_problem('Sedov', prob)\n"
        "    rc, t, dt, hist = sim.advance(snaps[-1])\n"
        "    assert rc == 0\n"
        "    assert np.array_equal(sim.download(), g['U%%d' %% snaps[-1]]), name\n"
        "    assert name != 'sedov16_hllc_long' or lib.workspace_replays(sim.ws_handle) > 0\n"
        "print('ok')\n") % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, CTU_PERSIST="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("name,chunk", [("sedov16_hllc", 4), ("sedov16_periodic_hllc", 8), ("sedov_24x16x8_cgf", 4), ("sedov32_hllc", 32)])
def test_host_streaming_step_matches_golden(name, chunk, monkeypatch):
    """ctu_sim_step_host: state in host memory, streamed through the device in z-chunks with copy/compute overlap"""
    _torch()
    from castro_b200 import lib
    from castro_b200.dist import initial_dt, new_dt
    monkeypatch.setenv("CTU_HOST_CHUNK", str(chunk))
    par, geom, prob, g, snaps = load_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("Sedov", prob)
    h = sim.download()
    h2 = np.empty_like(h)
    t = 0.0
    dt = initial_dt(par, sim.estdt(), t, -1.0)
    nsteps = min(snaps[-1], 20)
    times = [0.0]
    for step in range(1, nsteps + 1):
        rc, est, nbad = sim.step_host(h, h2, (t + dt) - t)
        assert rc == 0
        h, h2 = h2, h
        t += dt
        times.append(t)
        dt = new_dt(par, est[1], dt, t, -1.0)
        if step in snaps:
            assert_bitwise(h, g["U%d" % step], "%s host-streamed step %d" % (name, step))
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    if chunk < n_cell[2]:
        assert sim.last_slabs > 1
    sim.close()


@pytest.mark.parametrize("name", ["kh16_cgf", "kh16_hllc", "kh16_cgf_nolim", "kh16_4spec_cgf", "kh16_4spec_hllc"])
def test_kh_two_passives_matches_reference_golden(name):
    """KH as shipped (advected scalar + species, periodic): problem hook vs the reference's initial plotfile, then 10 steps
    of the full time loop vs the reference's final state -- exercises the NP = 2 instances of the tile-staged kernels"""
    _torch()
    from castro_b200 import lib
    from cases import load_kh_case
    par, geom, prob, g, snaps = load_kh_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("KH", prob)
    assert_bitwise(sim.download(), g["U0"], name + " initial data")
    n = snaps[-1]
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0
    assert_bitwise(sim.download(), g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)
    sim.close()


@pytest.mark.parametrize("nadv,nspec,solver,opt", [(1, 1, 2, ""), (1, 1, 0, ""), (0, 2, 2, ""), (0, 3, 2, ""), (1, 3, 0, ""), (0, 4, 2, ""),
                                                   (1, 4, 0, ""), (2, 4, 2, ""), (0, 3, 2, "ext"), (1, 3, 0, "ext"), (1, 4, 2, "ext"),
                                                   (2, 4, 0, "ext"), (0, 3, 1, "cg"), (1, 4, 1, "cg"), (0, 3, 2, "plm"), (0, 4, 0, "plm"),
                                                   (1, 4, 2, "plm"), (1, 2, 0, "pslope"), (1, 4, 2, "pslope")])
def test_passives_vs_oracle(nadv, nspec, solver, opt):
    """evolved 3-D Sedov state with non-trivial advected scalars / several species: one ctu_hydro_advance vs the oracle.
    Up to five passives run the tile-staged kernels (NP = 2, 3, 4, 5 here), (2,4) = six passives the register-staged ones.
    opt = "ext": with the hybrid Riemann solver and the density-floor flux limiter (the "extended" tile instances; small_dens is raised
    into the range of the data so that the limiter acts); "cg": the Colella & Glaz solver; "plm": piecewise-linear reconstruction;
    "pslope": the hydrostatic-pressure-aware PPM (without gravity: the same numbers through the other trace instance) -- the tile
    instances of each option for 3, 4 and 5 passives."""
    torch = _torch()
    import oracle_lib as O
    from castro_b200 import lib, abi
    from cases import sedov_params
    par = sedov_params(solver)
    par.nadv, par.nspec = nadv, nspec
    for n in range(nspec):
        par.aion[n] = 1.0 + 3.0 * n
        par.zion[n] = 1.0 + n
    if opt == "ext":
        par.hybrid_riemann, par.limit_fluxes_on_small_dens, par.small_dens = 1, 1, 0.3
    elif opt == "cg":
        par.riemann_solver = 1
    elif opt == "plm":
        par.ppm_type = 0
    elif opt == "pslope":
        par.use_pslope = 1
    O.params_finalize(par)
    n_cell = (20, 12, 10)
    geom = abi.make_geom(n_cell, probhi=(1.25, 0.75, 0.625))
    ns = par.num_state
    # evolve a single-species blast with the oracle, then paint passives onto it
    p1 = sedov_params(solver)
    O.params_finalize(p1)
    U1 = O.sedov_init(p1, geom, n_cell, r_init=0.1, nsub=4)
    st, t, dt, _ = O.run(p1, geom, U1, 6)
    assert st == 0
    U = np.zeros((ns,) + U1.shape[1:])
    U[:7] = U1[:7]
    rng = np.random.default_rng(11)
    rho = U[0]
    for a in range(nadv):
        U[7 + a] = rho * (0.3 + 0.4 * rng.random(rho.shape))
    X = 0.1 + rng.random((nspec,) + rho.shape)
    X /= X.sum(axis=0)
    for n in range(nspec):
        U[7 + nadv + n] = rho * X[n]
    hi = tuple(n - 1 for n in n_cell)
    glo = (-NG, -NG, -NG)
    O.clean_state(par, U, glo, (0, 0, 0), hi)
    O.fill_boundary(par, geom, U, (0, 0, 0), hi)
    O.clean_state(par, U, glo, glo, tuple(h + NG for h in hi))
    Sn_ref = np.ascontiguousarray(U[:, NG:-NG, NG:-NG, NG:-NG])
    st_ref, dt_ref, loc_ref, nbad_ref = O.hydro_advance(par, geom, (0, 0, 0), hi, U, glo, Sn_ref, (0, 0, 0), dt, do_clean=1)
    dSb = torch.from_numpy(U).cuda()
    dSn = torch.from_numpy(np.ascontiguousarray(U[:, NG:-NG, NG:-NG, NG:-NG])).cuda()
    ws = lib.Workspace(0)
    rc, est, loc, nbad = lib.hydro_advance(ws, par, geom, (0, 0, 0), hi, dSb, glo, dSn, (0, 0, 0), dt, n_clean=1)
    torch.cuda.synchronize()
    assert rc == st_ref and nbad == nbad_ref
    assert ws.tile_staged == (nadv + nspec <= 5)
    assert_bitwise(dSn.cpu().numpy(), Sn_ref, "S_new nadv=%d nspec=%d %s" % (nadv, nspec, opt))
    assert est[1] == dt_ref


@pytest.mark.parametrize("name", ["rt16_cgf", "rt16_hllc_gst2", "rt16_pslope", "rt16_wellbal", "rt16_plm_pslope", "rt16_combo"])
def test_rt_gravity_matches_reference_golden(name):
    """Rayleigh-Taylor with constant gravity: old-time source + src_to_prim + traced sources inside
    the hydro + new-time corrector (grav_source_type 4 from the mass fluxes, or 2), 20 steps of the full time loop started from
    the reference's own initial plotfile, bit-compared with the reference's final state (periodic x/y, slip walls in z)."""
    _torch()
    from castro_b200 import lib
    from cases import load_rt_case
    par, geom, g, snaps = load_rt_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("RT", dict(frac=0.5, rho_1=1.0, rho_2=2.0, p0_base=5.0))       # the RT problem hook ...
    assert_bitwise(sim.download(), g["U0"], name + " initial data")                 # ... against the reference's initial plotfile
    n = snaps[-1]
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0 and len(hist) == n
    assert_bitwise(sim.download(), g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)
    sim.close()


@pytest.mark.parametrize("nspec,solver,predictor", [(1, 2, 0), (1, 0, 1), (3, 2, 1)])
def test_hydro_advance_with_sources_vs_oracle(nspec, solver, predictor):
    """one ctu_hydro_advance with a non-trivial old-time source in ALL seven conserved components (and a source corrector when
    castro.source_term_predictor = 1) vs the oracle: exercises every branch of src_to_prim and of the traced sources.
    nspec = 1 runs the tile-staged kernels, nspec = 3 the register-staged ones."""
    torch = _torch()
    import oracle_lib as O
    from castro_b200 import lib, abi
    from cases import sedov_params
    par = sedov_params(solver)
    par.nspec = nspec
    par.source_term_predictor = predictor
    for n in range(nspec):
        par.aion[n] = 1.0 + 3.0 * n
        par.zion[n] = 1.0 + n
    O.params_finalize(par)
    n_cell = (20, 12, 10)
    geom = abi.make_geom(n_cell, probhi=(1.25, 0.75, 0.625))
    ns = par.num_state
    p1 = sedov_params(solver)
    O.params_finalize(p1)
    U1 = O.sedov_init(p1, geom, n_cell, r_init=0.1, nsub=4)
    st, t, dt, _ = O.run(p1, geom, U1, 6)
    assert st == 0
    U = np.zeros((ns,) + U1.shape[1:])
    U[:7] = U1[:7]
    rng = np.random.default_rng(5)
    rho = U[0]
    X = 0.1 + rng.random((nspec,) + rho.shape)
    X /= X.sum(axis=0)
    for n in range(nspec):
        U[7 + n] = rho * X[n]
    hi = tuple(n - 1 for n in n_cell)
    glo = (-NG, -NG, -NG)
    O.clean_state(par, U, glo, (0, 0, 0), hi)
    O.fill_boundary(par, geom, U, (0, 0, 0), hi)
    O.clean_state(par, U, glo, glo, tuple(h + NG for h in hi))
    # sources on grow(valid, 3): a few per cent of the state per unit time, plus exact zeros in part of the box (untraced zones)
    slo = (-3, -3, -3)
    Ug3 = U[:7, 1:-1, 1:-1, 1:-1]
    osrc = np.ascontiguousarray(Ug3 * (0.05 * (rng.random(Ug3.shape) - 0.5)) / dt)
    osrc[6] = 0.0
    osrc[:, :, :, :6] = 0.0
    scorr = np.ascontiguousarray(Ug3 * (0.02 * (rng.random(Ug3.shape) - 0.5)) / dt ** 2) if predictor else None
    Sn0 = np.ascontiguousarray(U[:, NG:-NG, NG:-NG, NG:-NG])
    Sn0[:7] += dt * osrc[:, 3:-3, 3:-3, 3:-3]
    O.clean_state(par, Sn0, (0, 0, 0), (0, 0, 0), hi)
    Sn_ref = Sn0.copy()
    mlo = [(0, 0, 0)] * 3
    mf_ref = [np.zeros((1,) + tuple(n_cell[2 - a] + (1 if 2 - a == d else 0) for a in range(3))) for d in range(3)]
    st_ref, dt_ref, loc_ref, nbad_ref = O.hydro_advance(par, geom, (0, 0, 0), hi, U, glo, Sn_ref, (0, 0, 0), dt, do_clean=1,
                                                        old_src=osrc, src_corr=scorr, src_lo=slo, mass_fluxes=mf_ref, mass_los=mlo)
    dSb = torch.from_numpy(U).cuda()
    dSn = torch.from_numpy(Sn0).cuda()
    dos = torch.from_numpy(osrc).cuda()
    dsc = torch.from_numpy(scorr).cuda() if predictor else None
    dmf = [torch.zeros(m.shape, dtype=torch.float64, device="cuda") for m in mf_ref]
    ws = lib.Workspace(0)
    rc, est, loc, nbad = lib.hydro_advance(ws, par, geom, (0, 0, 0), hi, dSb, glo, dSn, (0, 0, 0), dt, n_clean=1, old_src=dos,
                                           src_corr=dsc, src_lo=slo, mass_fluxes=dmf, mass_los=mlo)
    torch.cuda.synchronize()
    assert rc == st_ref and nbad == nbad_ref
    assert_bitwise(dSn.cpu().numpy(), Sn_ref, "S_new with sources nspec=%d solver=%d" % (nspec, solver))
    for d in range(3):
        assert np.array_equal(dmf[d].cpu().numpy(), mf_ref[d]), "mass flux %d" % d
    assert est[1] == dt_ref
    # the sources must matter: without them the answer is different
    dSn2 = torch.from_numpy(Sn0).cuda()
    lib.hydro_advance(ws, par, geom, (0, 0, 0), hi, dSb, glo, dSn2, (0, 0, 0), dt, n_clean=1)
    assert not torch.equal(dSn2, dSn)


def _run_reference_plotfile(tmp_path, n_cell, solver, nsteps, extra=()):
    """runs the UNMODIFIED reference executable (oracle/_ref, test infrastructure; the OpenMP build for the larger boxes -- its results
    do not depend on the tiling) on bench_inputs/inputs.sedov3d and returns (plotfile dir, fcompare path, inputs path) or skips"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    big = n_cell[0] * n_cell[1] * n_cell[2] > 32 ** 3
    exe = os.path.join(root, "oracle", "_ref", "Castro3d_sedov_omp.ex" if big else "Castro3d_sedov.ex")
    if not os.path.exists(exe):
        exe = os.path.join(root, "oracle", "_ref", "Castro3d_sedov.ex")
    fcmp = os.path.join(root, "oracle", "_ref", "fcompare.ex")
    inputs = os.path.join(root, "bench_inputs", "inputs.sedov3d")
    if not (os.path.exists(exe) and os.path.exists(fcmp)):
        pytest.skip("oracle/_ref not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference is mounted)")
    ref_args = [exe, inputs, "amr.n_cell=%d %d %d" % tuple(n_cell), "castro.riemann_solver=%d" % solver,
                "amr.max_grid_size=%d" % (32 if big else max(n_cell)), "amr.blocking_factor=4",
                "max_step=%d" % nsteps, "amr.plot_files_output=1", "amr.plot_int=%d" % nsteps, "amr.plot_file=%s" % (tmp_path / "ref_plt"),
                "amr.derive_plot_vars=NONE", "castro.v=0", "amr.v=0"] + list(extra)
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    out = subprocess.run(ref_args, cwd=tmp_path, capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    ref_pf = str(tmp_path / ("ref_plt%05d" % nsteps))
    assert os.path.isdir(ref_pf)
    return ref_pf, fcmp, inputs


def _fcompare_exact(fcmp, ref_pf, my_pf, ncomp=8):
    import re
    import subprocess
    res = subprocess.run([fcmp, "--rel_tol", "1e-12", "-a", ref_pf, my_pf], capture_output=True, text=True)
    assert res.returncode == 0 and "PLOTFILE AGREE" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    rows = re.findall(r"^\s*(\w+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s*$", res.stdout, flags=re.M)
    assert len(rows) == ncomp, res.stdout
    for name, abs_err, rel_err in rows:
        assert float(abs_err) == 0.0 and float(rel_err) == 0.0, "fcompare %s: abs %s rel %s" % (name, abs_err, rel_err)


# (solver, n_cell, steps, workspace bytes): the 64^3 and 72x40x32 boxes span several 32-wide TMA tiles in x, tiles that straddle the
# box edge in every direction and odd tile-origin parities; the last two force the multi-slab path (workspace bound) at that size
@pytest.mark.parametrize("solver,n_cell,nsteps,wsb", [
    (2, (24, 24, 24), 10, 0), (0, (24, 24, 24), 25, 0),
    (2, (64, 64, 64), 10, 0), (0, (64, 64, 64), 10, 0),
    (2, (72, 40, 32), 10, 0), (0, (72, 40, 32), 10, 0),
    (2, (64, 64, 64), 10, 96 << 20), (0, (72, 40, 32), 10, 40 << 20)])
def test_plotfile_vs_reference_with_fcompare(solver, n_cell, nsteps, wsb, tmp_path):
    """the north star's own acceptance procedure: run the UNMODIFIED reference executable (oracle/_ref, test infrastructure) and
    this library on the same inputs file, write a plotfile from each, and let the reference's comparator
    (external/amrex/Tools/Plotfile/fcompare) judge: every conserved field must agree to rel-Linf <= 1e-12 (it is 0)."""
    _torch()
    from castro_b200 import lib
    # cubic zones (the reference refuses dx != dy != dz, Castro.cpp:1030-1044): the domain is n_cell / min(n_cell) units long
    hi = "geometry.prob_hi=%r %r %r" % tuple(float(n) / min(n_cell) for n in n_cell)
    ref_pf, fcmp, inputs = _run_reference_plotfile(tmp_path, n_cell, solver, nsteps, extra=[hi])
    # the same inputs file through the library's own ParmParse reader
    over = "amr.n_cell=%d %d %d castro.riemann_solver=%d " % (n_cell + (solver,)) + hi
    par, geom, nc, max_step, stop_time, prob = lib.parse_inputs(inputs, over)
    assert tuple(nc) == tuple(n_cell)
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in nc), workspace_bytes=wsb)
    sim.init_problem("Sedov", {k: (int(v) if k == "nsub" else float(v)) for k, v in (kv.split("=") for kv in prob.split())})
    rc, t, dt, hist = sim.advance(nsteps)
    assert rc == 0 and len(hist) == nsteps
    if wsb:
        assert sim.last_slabs > 1, "meant to exercise the multi-slab path"
    my_pf = str(tmp_path / "b200_plt")
    sim.write_plotfile(my_pf, t, nsteps)
    sim.close()
    _fcompare_exact(fcmp, ref_pf, my_pf)


@pytest.mark.parametrize("solver,n_cell,nsteps,opts", [
    (2, (64, 64, 64), 30, "castro.hybrid_riemann=1"),
    (0, (72, 40, 32), 30, "castro.hybrid_riemann=1 castro.limit_fluxes_on_small_dens=1 castro.small_dens=0.3"),
    (2, (72, 40, 32), 12, "castro.ppm_type=0"),
    (1, (64, 64, 64), 12, "")])
def test_option_instances_of_the_tile_kernels_with_fcompare(solver, n_cell, nsteps, opts, tmp_path):
    """the acceptance procedure above for the non-default instantiations of the tile-staged kernels at a size where their tiles have
    neighbours, seams and odd origins: hybrid Riemann solver, hybrid + flux limiter, piecewise-linear reconstruction, Colella & Glaz"""
    _torch()
    from castro_b200 import lib
    hi = "geometry.prob_hi=%r %r %r" % tuple(float(n) / min(n_cell) for n in n_cell)
    extra = [hi] + opts.split()
    ref_pf, fcmp, inputs = _run_reference_plotfile(tmp_path, n_cell, solver, nsteps, extra=extra)
    over = "amr.n_cell=%d %d %d castro.riemann_solver=%d " % (n_cell + (solver,)) + hi + " " + opts
    par, geom, nc, max_step, stop_time, prob = lib.parse_inputs(inputs, over)
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in nc))
    sim.init_problem("Sedov", {k: (int(v) if k == "nsub" else float(v)) for k, v in (kv.split("=") for kv in prob.split())})
    rc, t, dt, hist = sim.advance(nsteps)
    assert rc == 0 and len(hist) == nsteps and sim.tile_staged
    my_pf = str(tmp_path / "b200_plt")
    sim.write_plotfile(my_pf, t, nsteps)
    sim.close()
    _fcompare_exact(fcmp, ref_pf, my_pf)


def test_reference_executable_with_stub_binding(tmp_path):
    """The drop-in boundary compiled and run for real: oracle/_ref/Castro3d_sedov_stub.ex is the reference's own GPU build
    (USE_CUDA=TRUE) with ONE source file swapped -- integration/Castro_ctu_hydro.cpp, which binds Castro::construct_ctu_hydro_source to
    ctu_hydro_advance of libcastro_ctu.so (oracle/build_ref.sh stub).  Driver, FillPatch, clean_state, estdt, dt control and plotfile
    I/O are the reference's.  Its plotfile must agree with the stock CPU executable's to the north star's 1e-12 (not bitwise: the
    reference's own device code around the call is compiled by nvcc with FMA contraction)."""
    import os
    import re
    import subprocess
    _torch()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stub = os.path.join(root, "oracle", "_ref", "Castro3d_sedov_stub.ex")
    if not os.path.exists(stub):
        pytest.skip("oracle/_ref/Castro3d_sedov_stub.ex not built (oracle/build_ref.sh gpu && oracle/build_ref.sh stub)")
    n_cell, nsteps = (64, 64, 64), 10
    for solver in (2, 0):
        d = tmp_path / ("s%d" % solver)
        d.mkdir()
        ref_pf, fcmp, inputs = _run_reference_plotfile(d, n_cell, solver, nsteps)
        args = [stub, inputs, "amr.n_cell=%d %d %d" % n_cell, "castro.riemann_solver=%d" % solver, "amr.max_grid_size=32",
                "amr.blocking_factor=4", "max_step=%d" % nsteps, "amr.plot_files_output=1", "amr.plot_int=%d" % nsteps,
                "amr.plot_file=%s" % (d / "stub_plt"), "amr.derive_plot_vars=NONE", "castro.v=1", "amr.v=0"]
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "castro_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        out = subprocess.run(args, cwd=d, capture_output=True, text=True, env=env)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
        assert "[libcastro_ctu]" in out.stdout, "the run did not go through the binding"
        res = subprocess.run([fcmp, "--rel_tol", "1e-12", str(ref_pf), str(d / ("stub_plt%05d" % nsteps))], capture_output=True, text=True)
        assert res.returncode == 0 and "PLOTFILE AGREE" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
        rows = re.findall(r"^\s*(\w+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s*$", res.stdout, flags=re.M)
        assert len(rows) == 8 and max(float(r[2]) for r in rows) <= REL_TOL, res.stdout


def test_plotfile_reader_and_grid_diag_vs_reference(tmp_path):
    """the reference executable writes a plotfile per step and its grid_diag.out (castro.sum_interval = 1);
    ctu_sim_read_plotfile must load each plotfile bit for bit (several grids per level), and ctu_sim_diag of that state must
    reproduce the reference's integrated diagnostics (volume-weighted sums: 1e-12 relative to the column's scale, the summation
    order differs; maxima exactly)."""
    import os
    import subprocess
    _torch()
    from castro_b200 import lib
    from plotfile import read_plotfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "Castro3d_sedov.ex")
    inputs = os.path.join(root, "bench_inputs", "inputs.sedov3d")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    n, nsteps = 24, 6
    args = [exe, inputs, "amr.n_cell=%d %d %d" % (n, n, n), "amr.max_grid_size=16", "amr.blocking_factor=8", "max_step=%d" % nsteps,
            "amr.plot_files_output=1", "amr.plot_int=1", "amr.plot_file=%s" % (tmp_path / "plt"), "amr.derive_plot_vars=NONE",
            "castro.sum_interval=1", "castro.v=1", "amr.v=0", "amr.data_log=grid_diag.out"]
    out = subprocess.run(args, cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = [ln.split() for ln in open(tmp_path / "grid_diag.out") if not ln.startswith("#")]
    assert len(rows) == nsteps + 1 and len(rows[0]) == 20
    over = "amr.n_cell=%d %d %d" % (n, n, n)
    par, geom, nc, _, _, prob = lib.parse_inputs(inputs, over)
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), (n - 1,) * 3)
    mine = str(tmp_path / "grid_diag_b200.out")
    import ctypes as C
    for s in range(nsteps + 1):
        pf = str(tmp_path / ("plt%05d" % s))
        h, U = read_plotfile(pf)
        t, st = sim.read_plotfile(pf)
        assert st == s and t == h["time"]
        assert np.array_equal(sim.download(), U[:par.num_state]), "plotfile reader, step %d" % s
        d = sim.diag((0.5, 0.5, 0.5))
        ref = [float(x) for x in rows[s][2:]]
        assert int(rows[s][0]) == s
        scale = [ref[0], ref[9], ref[9], ref[9], ref[9], ref[9], ref[9], ref[9], ref[9], ref[9], 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, ref[16], ref[17]]
        for c in range(16):
            assert abs(d[c] - ref[c]) <= 1e-12 * abs(scale[c]) + 1e-300, "grid_diag column %d at step %d: %r vs %r" % (c + 3, s, d[c], ref[c])
        assert d[16] == ref[16] and d[17] == ref[17]
        assert lib.lib().ctu_write_grid_diag(mine.encode(), s, t, (C.c_double * 18)(*d)) == 0
    got = [ln for ln in open(mine)]
    want = [ln for ln in open(tmp_path / "grid_diag.out")]
    assert got[0].split() == want[0].split() and got[1].split() == want[1].split(), "header of grid_diag.out"
    assert len(got) == len(want) and len(got[2]) == len(want[2])
    sim.close()


@pytest.mark.parametrize("name", ["sedov16_hllc_retry", "sedov16_cgf_retry"])
def test_retry_subcycling_matches_reference(name):
    """castro.use_retry (Castro::retry_advance_ctu / subcycle_advance_ctu): the first advance fails the post-advance dt validity check
    and is redone as two subcycles of dt / 2 from the untouched old state; states and times must match the reference's, and the
    same run without use_retry must stop with CTU_STATUS_DT_POST like the reference aborts"""
    _torch()
    from castro_b200 import lib
    par, geom, prob, g, snaps = load_case(name)
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), (15, 15, 15))
    sim.init_problem("Sedov", prob)
    rc, t, dt, hist = sim.advance(1)
    assert rc == 0 and sim.retries >= 1 and sim.last_subcycles == 2
    assert_bitwise(sim.download(), g["U1"], name + " step 1 (two subcycles)")
    rc, t, dt, hist2 = sim.advance(7, time=t, dt=dt)
    assert rc == 0
    assert_bitwise(sim.download(), g["U8"], name + " step 8")
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist, hist2])), g["times"][:9], rtol=1e-14, atol=0)
    sim.close()
    par.use_retry = 0
    sim = lib.Sim(par, geom, (0, 0, 0), (15, 15, 15))
    sim.init_problem("Sedov", prob)
    rc, t, dt, hist = sim.advance(3)
    assert rc == 16 and len(hist) == 1 and sim.retries == 0
    sim.close()


@pytest.mark.parametrize("name", ["sodstellar_x_cg", "sodstellar_x_cgf", "sodstellar_y_hllc", "sodstellar_z_cgf",
                                  "sodstellar_z_grav_cgf", "sodstellar_x_grav_hllc_gst2"])
def test_helmholtz_eos_sod_stellar_matches_reference(name):
    """the Helmholtz (stellar) EOS behind eos_* (ctu_helm.cuh) on the reference's Sod_stellar problem (rho 1e7 | 1e6
    g/cc, T 1e8 | 1e6 K, C12 / O16 / Mg24 network, non-constant Gamma_1 through every stage).  The device evaluates log10 / pow / cbrt
    with the CUDA math library, so this is a tolerance test: the north star's fcompare criterion, max|a - b| / max|a| <= 1e-12 per
    conserved field after 20 steps (Temp is a Newton iterate with tolerance 1e-8: 1e-9)."""
    import os
    _torch()
    from castro_b200 import lib
    from cases import load_sodstellar_case, HELM_TABLE
    if not os.path.exists(HELM_TABLE):
        pytest.skip("oracle/_ref/helm_table.dat missing (copied from Microphysics by build() where /root/reference is mounted)")
    lib.helm_load(HELM_TABLE)
    par, geom, prob, g, snaps = load_sodstellar_case(name)
    lib.params_finalize(par)
    assert par.small_temp == 1000.0 and par.small_dens == 1.e-12
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("Sod", prob)

    def check(U, ref, what):
        rl = rel_linf(U, ref, vector_momentum="grav" in name)
        for m, r in enumerate(rl):
            tol = 1e-9 if m == 6 else REL_TOL
            assert r <= tol, "%s: component %d rel-Linf %g > %g (all: %s)" % (what, m, r, tol, rl)
        return rl
    check(sim.download(), g["U0"], name + " initial data")
    n = snaps[-1]
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0 and len(hist) == n
    rl = check(sim.download(), g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-12, atol=0)
    print(name, "rel-Linf per component after %d steps:" % n, ["%.1e" % r for r in rl])
    sim.close()


def test_helmholtz_tile_staged_kernels_match_register_staged_bitwise(tmp_path):
    """The Helmholtz-EOS instances of the tile-staged kernels (Gamma_1 staged beside the sound speed, reconstructed and traced; RS = 4:
    CGF / HLLC, RS = 5: Colella & Glaz; EK = 1 update kernel) against the register-staged family (CTU_NO_TMA=1, a subprocess because the
    switch is read once per process), which test_helmholtz_eos_sod_stellar_matches_reference pins to the reference: the same device
    arithmetic in the same order, so the two families agree bit for bit -- with and without gravity, tube along x, y and z."""
    import os
    import subprocess
    import sys
    _torch()
    from castro_b200 import lib
    from cases import load_sodstellar_case, HELM_TABLE
    if not os.path.exists(HELM_TABLE):
        pytest.skip("oracle/_ref/helm_table.dat missing")
    names = ["sodstellar_x_cg", "sodstellar_x_cgf", "sodstellar_y_hllc", "sodstellar_z_cgf", "sodstellar_z_grav_cgf", "sodstellar_x_grav_hllc_gst2"]
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from cases import load_sodstellar_case, HELM_TABLE\n"
        "from castro_b200 import lib\n"
        "lib.helm_load(HELM_TABLE)\n"
        "out = {}\n"
        "for name in %r:\n"
        "    par, geom, prob, g, snaps = load_sodstellar_case(name)\n"
        "    lib.params_finalize(par)\n"
        "    n = tuple(int(x) for x in g['n_cell'])\n"
        "    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))\n"
        "    sim.init_problem('Sod', prob)\n"
        "    rc, t, dt, hist = sim.advance(snaps[-1])\n"
        "    assert rc == 0 and not sim.tile_staged\n"
        "    out[name] = sim.download(); out[name + '_dt'] = np.array(hist)\n"
        "np.savez(%r, **out)\n"
        "print('ok')\n") % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))), names,
                            str(tmp_path / "reg.npz"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CTU_NO_TMA="1"), capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
    reg = np.load(str(tmp_path / "reg.npz"))
    lib.helm_load(HELM_TABLE)
    for name in names:
        par, geom, prob, g, snaps = load_sodstellar_case(name)
        lib.params_finalize(par)
        n = tuple(int(x) for x in g["n_cell"])
        sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n))
        sim.init_problem("Sod", prob)
        rc, t, dt, hist = sim.advance(snaps[-1])
        assert rc == 0 and sim.tile_staged, name
        assert np.array_equal(np.array(hist), reg[name + "_dt"]), name
        assert_bitwise(sim.download(), reg[name], name + " tile-staged vs register-staged")
        sim.close()


def test_helmholtz_eos_device_matches_host_and_reference_values():
    """the device copy of the Helmholtz EOS against its host copy on a sweep of (rho, T) over the table (1e-12 relative) and round
    trips through the Newton modes (rho, e) and (rho, p)"""
    import os
    _torch()
    from castro_b200 import lib
    from cases import HELM_TABLE
    if not os.path.exists(HELM_TABLE):
        pytest.skip("oracle/_ref/helm_table.dat missing")
    lib.helm_load(HELM_TABLE)
    par = lib.default_params()
    par.eos_kind, par.nspec = 1, 3
    for m, (a, z) in enumerate(((12.0, 6.0), (16.0, 8.0), (24.0, 12.0))):
        par.aion[m], par.zion[m] = a, z
    lib.params_finalize(par)
    rng = np.random.default_rng(7)
    n = 4000
    rho = 10.0 ** rng.uniform(-3.0, 9.5, n)
    T = 10.0 ** rng.uniform(4.0, 9.8, n)
    st = np.stack([rho, T, np.zeros(n), np.zeros(n)], axis=1)
    for abar, ye in ((12.0, 0.5), (13.7, 0.5), (4.0, 0.5), (1.3, 0.8)):
        h = lib.helm_eval(par, 0, abar, ye, st, where=0)
        d = lib.helm_eval(par, 0, abar, ye, st, where=1)
        for col in (2, 3, 4, 5, 6, 7):
            err = np.max(np.abs(d[:, col] - h[:, col]) / np.abs(h[:, col]))
            assert err < 1e-11, "device vs host, column %d: %g" % (col, err)
        # (rho, e) -> T and (rho, p) -> T: the temperature the Newton iteration lands on reproduces the wanted energy / pressure
        # (in degenerate matter e and p barely depend on T, so T itself is not recoverable there -- the wanted quantity is)
        guess = np.stack([rho, 1.3 * T, h[:, 2], h[:, 3]], axis=1)
        for mode, col in ((2, 2), (1, 3)):
            r = lib.helm_eval(par, mode, abar, ye, guess, where=1)
            back = lib.helm_eval(par, 0, abar, ye, np.stack([rho, r[:, 1], np.zeros(n), np.zeros(n)], axis=1), where=0)
            err = np.abs(back[:, col] - h[:, col]) / np.abs(h[:, col])
            assert np.max(err) < 1e-7, "Newton mode %d: the converged T does not reproduce the input: %g" % (mode, np.max(err))


@pytest.mark.parametrize("name", ["sod1d_cgf", "sod1d_cg", "sod1d_test2_cgf", "sod1d_test3_plm"])
def test_sod_one_dimensional_matches_reference(name):
    """BASELINE config[0] literally: Sod in ONE dimension (geom.dim = 1; the reference built with DIM=1).  Problem hook vs the
    reference's initial plotfile, then the full time loop (sod1d_cgf: to stop_time = 0.2, 81 steps) vs the reference's states and
    times, bit for bit; HLLC is refused in 1-D like the reference refuses it (Castro.cpp:410-413)."""
    _torch()
    from castro_b200 import lib
    from cases import load_sod1d_case
    par, geom, prob, g, snaps = load_sod1d_case(name)
    lib.params_finalize(par)
    nx = int(g["n_cell"][0])
    sim = lib.Sim(par, geom, (0, 0, 0), (nx - 1, 0, 0))
    sim.init_problem("Sod", prob)
    assert_bitwise(sim.download(), g["U0"], name + " initial data")
    n = snaps[-1]
    stop = 0.2 if name == "sod1d_cgf" else -1.0
    rc, t, dt, hist = sim.advance(n, stop_time=stop)
    assert rc == 0 and len(hist) == n
    assert_bitwise(sim.download(), g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)
    sim.close()
    par.riemann_solver = 2
    with pytest.raises(lib.CtuError):
        lib.Sim(par, geom, (0, 0, 0), (nx - 1, 0, 0))


def test_host_streaming_mixed_boundaries():
    """ctu_sim_step_host promises the bits of upload + pre + fill + hydro + download.  Periodic x/y with a physical boundary in z
    (reflecting wall below, outflow above) is the case where a z-ghost zone is filled from an x/y-ghost zone of the boundary plane:
    an off-centre blast puts non-trivial data into those edge and corner zones."""
    _torch()
    from castro_b200 import lib, abi
    from castro_b200.dist import initial_dt, new_dt
    from cases import sedov_params
    import os
    n_cell = (24, 16, 20)
    for solver in (2, 0):
        par = sedov_params(solver)
        par.init_shrink = 0.3
        lib.params_finalize(par)
        geom = abi.make_geom(n_cell, probhi=(1.5, 1.0, 1.25), periodic=(1, 1, 0), lo_bc=(0, 0, 3), hi_bc=(0, 0, 2))
        prob = dict(r_init=0.12, p_ambient=1.e-2, exp_energy=1.0, dens_ambient=1.0, nsub=4)
        hi = tuple(n - 1 for n in n_cell)
        ref = lib.Sim(par, geom, (0, 0, 0), hi)
        ref.init_problem("Sedov", prob)
        U = ref.download()
        # move the blast into a corner of the periodic directions, next to the lower z boundary
        U = np.roll(U, (-7, -6, -11), axis=(1, 2, 3)).copy()
        ref.upload(U)
        os.environ["CTU_HOST_CHUNK"] = "4"
        try:
            sim = lib.Sim(par, geom, (0, 0, 0), hi)
            sim.upload(U)
            h, h2 = U.copy(), np.empty_like(U)
            t = 0.0
            dt = initial_dt(par, sim.estdt(), t, -1.0)
            rc, t_ref, dt_ref, hist = ref.advance(12)
            assert rc == 0
            for _ in range(12):
                rc, est, nbad = sim.step_host(h, h2, (t + dt) - t)
                assert rc == 0
                h, h2 = h2, h
                t += dt
                dt = new_dt(par, est[1], dt, t, -1.0)
            assert sim.last_slabs > 1
        finally:
            del os.environ["CTU_HOST_CHUNK"]
        got = ref.download()
        assert np.abs(got[1:4, :, :, :]).max() > 0 and np.abs(got[3, :2]).max() > 0, "the blast must have reached the z boundary"
        assert_bitwise(h, got, "host-streamed step with periodic x/y + physical z (solver %d)" % solver)
        assert t == t_ref
        sim.close()
        ref.close()


def test_mathfast_sequences_match_ieee_operators():
    """the branch-free division / reciprocal / sqrt sequences of the flux-stage kernels (MathFast, ctu_physics.cuh) restate the fast
    path of nvcc's IEEE operators; every result they do not flag for the exact replay must equal the operator's bit for bit --
    random operands over the whole exponent range plus zeros, subnormals, infinities and NaNs"""
    _torch()
    import ctypes as C
    from castro_b200 import lib
    L = lib.lib()
    for lo, hi in ((-20, 20), (-200, 200), (-500, 500), (-1022, 1023)):
        c = (C.c_ulonglong * 5)()
        rc = L.ctu_selftest_mathfast(lo, hi, 128, c)
        assert rc == 0
        assert c[4] == 1024 * 128 * 128
        assert c[0] == 0 and c[1] == 0 and c[2] == 0, "exp [%d,%d]: mismatches div %d rcp %d sqrt %d" % (lo, hi, c[0], c[1], c[2])
        assert c[3] < 3 * c[4], "every operand flagged?"
        if hi <= 200:
            assert c[3] < 0.3 * c[4], "ordinary operands must stay on the fast path (flagged %d of %d)" % (c[3], 3 * c[4])


def test_full_size_properties_sedov256():
    """BASELINE config[1] at its full size (3-D Sedov 256^3, HLLC), through size-independent properties: (a) the z-slab
    decomposition of the workspace does not change a single bit; (b) the host-streamed step equals the device-resident step bit for
    bit; (c) mass and total energy are conserved to rounding while the blast is far from the outflow boundaries; (d) the solution
    keeps the x<->y<->z permutation symmetry of the initial data to 1e-12 (the scheme's operation order is not symmetric, so this
    is a tolerance test); (e) no exceptional zone: status 0 and finite everywhere."""
    _torch()
    from castro_b200 import lib, abi
    from castro_b200.dist import initial_dt, new_dt
    from cases import sedov_params
    n = 256
    par = sedov_params(2)
    lib.params_finalize(par)
    geom = abi.make_geom((n, n, n))
    prob = dict(r_init=0.01, p_ambient=1.e-5, exp_energy=1.0, dens_ambient=1.0, nsub=10)
    nsteps = 4
    # (a) one slab vs several
    sim = lib.Sim(par, geom, (0, 0, 0), (n - 1,) * 3)
    sim.init_problem("Sedov", prob)
    U0 = sim.download()
    rc, t, dt, hist = sim.advance(nsteps)
    assert rc == 0
    U1 = sim.download()
    slabs1 = sim.last_slabs
    sim.close()
    sim = lib.Sim(par, geom, (0, 0, 0), (n - 1,) * 3, workspace_bytes=6 << 30)
    sim.init_problem("Sedov", prob)
    rc, t2, dt2, hist2 = sim.advance(nsteps)
    assert rc == 0 and sim.last_slabs > slabs1
    U2 = sim.download()
    assert np.array_equal(U1, U2) and np.array_equal(hist, hist2)
    # (b) host-streamed step from the same initial data
    sim.upload(U0)
    h, h2 = U0.copy(), np.empty_like(U0)
    tt = 0.0
    dth = initial_dt(par, sim.estdt(), tt, -1.0)
    for _ in range(nsteps):
        rc, est, nbad = sim.step_host(h, h2, (tt + dth) - tt)
        assert rc == 0
        h, h2 = h2, h
        tt += dth
        dth = new_dt(par, est[1], dth, tt, -1.0)
    sim.close()
    assert np.array_equal(h, U1) and tt == t
    # (c) conservation (sums in extended precision)
    for comp, name in ((0, "mass"), (4, "total energy")):
        s0, s1 = np.sum(U0[comp], dtype=np.longdouble), np.sum(U1[comp], dtype=np.longdouble)
        assert abs(float((s1 - s0) / s0)) < 1e-13, "%s not conserved: %r -> %r" % (name, s0, s1)
    # (d) permutation symmetry x<->y and y<->z (array axes are [z, y, x])
    for comp in (0, 4, 5):
        a = U1[comp]
        sc = np.abs(a).max()
        assert np.abs(a - a.transpose(0, 2, 1)).max() <= 1e-12 * sc
        assert np.abs(a - a.transpose(1, 0, 2)).max() <= 1e-12 * sc
    sc = np.abs(U1[1]).max()
    assert np.abs(U1[1] - U1[2].transpose(0, 2, 1)).max() <= 1e-12 * sc      # xmom(x,y,z) = ymom(y,x,z)
    assert np.abs(U1[2] - U1[3].transpose(1, 0, 2)).max() <= 1e-12 * sc      # ymom(x,y,z) = zmom(x,z,y)
    # (e)
    assert np.isfinite(U1).all()


@pytest.mark.parametrize("name", ["sod_x_cgf", "sod_y_cgf", "sod_z_cgf", "sod_z_hllc", "sod_x_symm_ppm", "sod_y_symm_plm"])
def test_sod_shock_tube_matches_reference_golden(name):
    """BASELINE config[0] (Sod shock tube, PPM + CGF; one HLLC variant) in 3-D with the tube along x, y and z: the "Sod" problem hook
    vs the reference's initial plotfile, then 30 steps of the full time loop vs the reference's final state, bit for bit; and the
    tube must stay one-dimensional (no variation across it, no transverse momentum)."""
    _torch()
    from castro_b200 import lib
    from cases import load_sod_case
    par, geom, prob, g, snaps = load_sod_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("Sod", prob)
    if "U0" in g.files:
        assert_bitwise(sim.download(), g["U0"], name + " initial data")
    n = snaps[-1]
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0
    U = sim.download()
    assert_bitwise(U, g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)
    d = prob["idir"] - 1                       # tube direction; array axes are [comp, z, y, x]
    ax = 3 - d
    prof = np.moveaxis(U, ax, 1).reshape(U.shape[0], U.shape[ax], -1)
    assert np.all(prof == prof[:, :, :1]), "the tube is not one-dimensional any more"
    for m in (1, 2, 3):
        if m != 1 + d:
            assert np.all(U[m] == 0.0)
    sim.close()


def _sod_exact(x, t, gamma=1.4, rl=1.0, pl=1.0, rr=0.125, pr=0.1, x0=0.5):
    """exact solution of the Sod Riemann problem (left rarefaction, contact, right shock), density / velocity / pressure at (x, t)"""
    g = gamma
    cl, cr = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p):   # u*_L(p) - u*_R(p)
        ul = 2.0 * cl / (g - 1.0) * (1.0 - (p / pl) ** ((g - 1.0) / (2.0 * g)))
        ur = (p - pr) * np.sqrt(2.0 / ((g + 1.0) * rr) / (p + (g - 1.0) / (g + 1.0) * pr))
        return ul - ur
    lo, hi = pr, pl
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0.0:
            lo = mid
        else:
            hi = mid
    ps = 0.5 * (lo + hi)
    us = 2.0 * cl / (g - 1.0) * (1.0 - (ps / pl) ** ((g - 1.0) / (2.0 * g)))
    rsl = rl * (ps / pl) ** (1.0 / g)
    rsr = rr * ((ps / pr + (g - 1.0) / (g + 1.0)) / ((g - 1.0) / (g + 1.0) * ps / pr + 1.0))
    shock = us * rsr / (rsr - rr) if rsr != rr else 0.0
    csl = cl * (ps / pl) ** ((g - 1.0) / (2.0 * g))
    xi = (x - x0) / t
    rho = np.where(xi < -cl, rl, np.where(xi < us - csl, rl * (2.0 / (g + 1.0) - (g - 1.0) / ((g + 1.0) * cl) * xi) ** (2.0 / (g - 1.0)),
                   np.where(xi < us, rsl, np.where(xi < shock, rsr, rr))))
    return rho


def test_sod_converges_to_exact_solution():
    """physical sanity beyond bitwise parity (the reference ships Verification/sod-exact.out for the same purpose): the 128-zone
    Sod tube at t = 0.2 against the exact Riemann solution, L1 density error < 1 %, for CGF and HLLC"""
    _torch()
    from castro_b200 import lib, abi
    for solver in (0, 2):
        par = abi.default_params()
        par.riemann_solver = solver
        par.cfl, par.init_shrink, par.change_max = 0.9, 0.1, 1.05
        par.eos_gamma, par.eos_assume_neutral = 1.4, 1
        lib.params_finalize(par)
        n_cell = (128, 4, 4)
        geom = abi.make_geom(n_cell, probhi=(1.0, 0.03125, 0.03125), lo_bc=(2, 4, 4), hi_bc=(2, 4, 4))
        sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
        sim.init_problem("Sod", dict(p_l=1.0, u_l=0.0, rho_l=1.0, p_r=0.1, u_r=0.0, rho_r=0.125, idir=1, frac=0.5))
        rc, t, dt, hist = sim.advance(10000, stop_time=0.2)
        assert rc == 0 and abs(t - 0.2) < 1e-14 and 50 < len(hist) < 400
        rho = sim.download()[0, 0, 0, :]
        x = (np.arange(128) + 0.5) / 128.0
        err = np.abs(rho - _sod_exact(x, t)).mean()
        assert err < 0.01, "solver %d: L1 density error %g" % (solver, err)
        sim.close()


def test_exact_replay_path_is_exercised():
    """the flux-stage kernels compute with the branch-free division / sqrt sequences and replay a tile with the IEEE operators when
    an operand leaves their range (MathFast / MathExact).  The Sedov goldens contain such operands (round-off dust in the subnormal
    range around the blast): this run must have replayed tiles AND still be bit-identical to the reference; a quiescent uniform
    state must not replay anything (zero numerators are handled in the fast path)."""
    _torch()
    from castro_b200 import lib
    par, geom, prob, g, snaps = load_case("sedov16_hllc_long")
    lib.params_finalize(par)
    sim = lib.Sim(par, geom, (0, 0, 0), (15, 15, 15))
    sim.init_problem("Sedov", prob)
    rc, t, dt, hist = sim.advance(snaps[-1])
    assert rc == 0
    assert_bitwise(sim.download(), g["U%d" % snaps[-1]], "sedov16_hllc_long")
    assert lib.workspace_replays(sim.ws_handle) > 0, "expected some tiles to take the exact-replay path"
    sim.close()
    # uniform medium at rest: every quotient with a zero numerator stays on the fast path, and the state does not move
    sim = lib.Sim(par, geom, (0, 0, 0), (15, 15, 15))
    U0 = np.zeros((par.num_state, 16, 16, 16))
    U0[0] = 1.0
    U0[4] = U0[5] = 1.0 / (par.eos_gamma - 1.0)
    U0[6] = 1.0
    U0[7] = 1.0
    sim.upload(U0)
    rc, t, dt, hist = sim.advance(3)
    assert rc == 0
    assert lib.workspace_replays(sim.ws_handle) == 0
    U1 = sim.download()
    for m in (0, 1, 2, 3, 4, 5, 7):
        assert np.array_equal(U0[m], U1[m]), "a uniform state at rest must stay exactly as it is (comp %d)" % m
    sim.close()


from cases import CASES_2D


@pytest.mark.parametrize("name", sorted(CASES_2D))
def test_two_dimensional_matches_reference(name):
    """The 2-D Cartesian path (geom.dim = 2; the reference built with DIM=2): the reference's unmodified problem headers compiled for
    AMREX_SPACEDIM = 2 give the initial plotfile, then the full time loop against the reference's states and times, bit for bit
    (Sedov cylindrical blast in Cartesian coordinates: CGF / HLLC / Colella & Glaz, PLM, EOS edge states, symmetry + wall + outflow
    boundaries, a non-square box; Kelvin-Helmholtz: periodic, two passives)."""
    _torch()
    import os
    from castro_b200 import lib
    from cases import load_2d_case
    par, geom, prob, g, snaps = load_2d_case(name)
    lib.params_finalize(par)
    nx, ny = int(g["n_cell"][0]), int(g["n_cell"][1])
    sim = lib.Sim(par, geom, (0, 0, 0), (nx - 1, ny - 1, 0))
    hook = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "castro_b200", "hooks",
                        "libctu_hook_%s.so" % ("sedov2d" if name.startswith("sedov") else "rt2d" if name.startswith("rt") else "kh2d"))
    if os.path.exists(hook):
        sim.init_problem("hook:" + hook, prob)
        assert_bitwise(sim.download(), g["U0"], name + " initial data")
    else:
        sim.upload(g["U0"])
    t, dt, step, times = 0.0, -1.0, 0, [0.0]
    for s in snaps[1:]:
        rc, t, dt, hist = sim.advance(s - step, time=t, dt=dt)
        assert rc == 0 and len(hist) == s - step
        times += list(times[-1] + np.cumsum(hist))
        step = s
        assert_bitwise(sim.download(), g["U%d" % s], "%s step %d" % (name, s))
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    assert not sim.tile_staged
    sim.close()


def test_two_dimensional_unsupported_options_are_refused():
    _torch()
    from castro_b200 import lib
    from cases import load_2d_case
    par, geom, prob, g, snaps = load_2d_case("sedov2d_cgf")
    for name, k, v in (("sedov2d_cgf", "ppm_temp_fix", 2), ("sedov2d_rz_cgf", "hybrid_riemann", 1), ("sedov2d_rz_cgf", "limit_fluxes_on_small_dens", 1),
                       ("sedov2d_rz_cgf", "do_grav", 1)):
        par, geom, prob, g, snaps = load_2d_case(name)
        setattr(par, k, v)
        lib.params_finalize(par)
        with pytest.raises(lib.CtuError):
            lib.Sim(par, geom, (0, 0, 0), (31, 31, 0))


from cases import SEDOV1D_CASES


@pytest.mark.parametrize("name", sorted(SEDOV1D_CASES))
def test_one_dimensional_curvilinear_matches_reference(name):
    """1-D spherical / cylindrical radial coordinate (geom.dim = 1, geom.coord = 2 / 1; the reference's DIM=1 Sedov executable on
    inputs.1d.sph / inputs.1d.cyl): the reference's unmodified Sedov headers compiled for AMREX_SPACEDIM = 1 give the initial plotfile,
    then 80-120 steps of the full time loop against the reference's states and times, bit for bit (geometric sources traced by PPM
    / added by PLM, r-dependent areas and volumes, pressure-free radial momentum flux + grad p, curvilinear div u)."""
    _torch()
    import os
    from castro_b200 import lib
    from cases import load_sedov1d_case
    par, geom, prob, g, snaps = load_sedov1d_case(name)
    lib.params_finalize(par)
    nx = int(g["n_cell"][0])
    sim = lib.Sim(par, geom, (0, 0, 0), (nx - 1, 0, 0))
    hook = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "castro_b200", "hooks", "libctu_hook_sedov1d.so")
    if os.path.exists(hook):
        sim.init_problem("hook:" + hook, prob)
        assert_bitwise(sim.download(), g["U0"], name + " initial data")
    else:
        sim.upload(g["U0"])
    t, dt, step, times = 0.0, -1.0, 0, [0.0]
    for s in snaps[1:]:
        rc, t, dt, hist = sim.advance(s - step, time=t, dt=dt)
        assert rc == 0 and len(hist) == s - step
        times += list(times[-1] + np.cumsum(hist))
        step = s
        assert_bitwise(sim.download(), g["U%d" % s], "%s step %d" % (name, s))
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    sim.close()
    # curvilinear coordinates exist in one dimension only
    from castro_b200 import abi
    g3 = abi.make_geom((16, 16, 16), coord=1)
    with pytest.raises(lib.CtuError):
        lib.Sim(par, g3, (0, 0, 0), (15, 15, 15))


@pytest.mark.parametrize("name", ["sedov2d_cg_24x16", "sedov1d_sph_cgf"])
def test_plotfile_io_in_one_and_two_dimensions(name, tmp_path):
    """ctu_sim_write_plotfile / ctu_sim_read_plotfile follow the run's dimensionality (boxes, prob_lo / hi and dx with AMREX_SPACEDIM
    entries, coord_sys in the Header): the file parses with the reader that parsed the reference's own 1-D / 2-D plotfiles when the
    goldens were made, and reads back bit for bit"""
    _torch()
    from castro_b200 import lib
    from plotfile import read_plotfile
    from cases import load_2d_case, load_sedov1d_case
    par, geom, prob, g, snaps = load_2d_case(name) if "2d" in name else load_sedov1d_case(name)
    lib.params_finalize(par)
    hi = (int(g["n_cell"][0]) - 1, int(g["n_cell"][1]) - 1, 0)
    sim = lib.Sim(par, geom, (0, 0, 0), hi)
    sim.upload(g["U0"])
    rc, t, dt, hist = sim.advance(5)
    assert rc == 0
    pf = tmp_path / "plt00005"
    sim.write_plotfile(pf, t, 5)
    h, U = read_plotfile(str(pf))
    assert h["dim"] == geom.dim and h["step"] == 5 and h["time"] == pytest.approx(t, rel=1e-14)
    assert np.array_equal(U, sim.download())
    other = lib.Sim(par, geom, (0, 0, 0), hi)
    t2, st2 = other.read_plotfile(pf)
    assert st2 == 5 and np.array_equal(other.download(), sim.download())
    sim.close()
    other.close()


@pytest.mark.parametrize("name", ["sodstellar2d_cg", "sodstellar2d_hllc", "sodstellar2d_grav_cgf", "sodstellar1d_cg", "sodstellar1d_cgf_plm"])
def test_helmholtz_eos_in_one_and_two_dimensions(name):
    """the Helmholtz EOS on Sod_stellar in the dimensionalities the problem ships with (2-D: its GNUmakefile; 1-D: its README), against
    the reference's DIM=2 / DIM=1 executables: fcompare's criterion, rel-Linf <= 1e-12 per conserved field (Temp 1e-9) after 20 steps (the round-off of the device math library grows to 1.6e-12 by step 30 of the 1-D tube)"""
    import os
    _torch()
    from castro_b200 import lib
    from cases import load_sodstellar_lowd_case, HELM_TABLE
    if not os.path.exists(HELM_TABLE):
        pytest.skip("oracle/_ref/helm_table.dat missing (copied from Microphysics by build() where /root/reference is mounted)")
    lib.helm_load(HELM_TABLE)
    par, geom, prob, g, snaps = load_sodstellar_lowd_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.upload(g["U0"])

    def check(U, ref, what):
        rl = rel_linf(U, ref, vector_momentum="grav" in name)
        for m, r in enumerate(rl):
            tol = 1e-9 if m == 6 else REL_TOL
            assert r <= tol, "%s: component %d rel-Linf %g > %g (all: %s)" % (what, m, r, tol, rl)
        return rl
    n = snaps[-1]
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0 and len(hist) == n
    rl = check(sim.download(), g["U%d" % n], "%s step %d" % (name, n))
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-12, atol=0)
    print(name, "rel-Linf per component after %d steps:" % n, ["%.1e" % r for r in rl])
    sim.close()
