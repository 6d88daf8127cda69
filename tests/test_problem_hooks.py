"""Source-compatible problem hooks (north star: "keeps Castro's problem-setup hooks"): the reference's UNMODIFIED
Exec/hydro_tests/<problem>/problem_initialize.H + problem_initialize_state_data.H are compiled against castro_b200/hooks/shim by
tools/build_problem_hook.sh (done by build() where /root/reference is mounted; the .so files travel to the GPU box).
CPU tests: hook output + the tail of Castro::initData (oracle) == the reference's initial plotfile, bit for bit.
GPU test: ctu_problem_init("hook:<lib>") == the reference's initial plotfile and == the library's built-in problem kernels."""
import ctypes as C
import os

import numpy as np
import pytest

from castro_b200 import abi
from cases import GOLDEN, NOH_CASES, load_case, load_kh_case, load_noh_case, load_sod_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOOKS = os.path.join(ROOT, "castro_b200", "hooks")


def hook_path(name):
    p = os.path.join(HOOKS, "libctu_hook_%s.so" % name)
    if not os.path.exists(p):
        pytest.skip("%s not built (tools/build_problem_hook.sh; needs /root/reference)" % p)
    return p


def prob_string(d):
    # arrays (problem.center = 0.5 0.5 0.0) are written v1,v2,v3
    return " ".join("%s=%s" % (k, ",".join(repr(x) for x in v) if isinstance(v, (tuple, list)) else repr(v)) for k, v in d.items())


# (hook library, problem name of the built-in kernels, golden with the reference's initial plotfile, loader)
HOOK_CASES = [("sedov", "Sedov", "sedov16_hllc", "sedov"), ("sedov", "Sedov", "sedov16_periodic_hllc", "sedov"),
              ("sod", "Sod", "sod_x_cgf", "sod"), ("sod", "Sod", "sod_z_hllc", "sod"), ("sod", "Sod", "sod_y_symm_plm", "sod"),
              ("kh", "KH", "kh16_cgf", "kh"), ("kh", "KH", "kh16_hllc", "kh"), ("kh4", "KH", "kh16_4spec_cgf", "kh"),
              ("kh4", "KH", "kh16_4spec_hllc", "kh"), ("rt", "RT", "rt16_cgf", "rt")]


def load_any(golden, kind):
    if kind == "sedov":
        par, geom, prob, g, snaps = load_case(golden)
    elif kind == "sod":
        par, geom, prob, g, snaps = load_sod_case(golden)
    elif kind == "rt":      # (built from build_ref.sh's scratch copy: the shipped RT header needs `problem::` in front of one L_x to compile in 3-D)
        from cases import load_rt_case
        par, geom, g, snaps = load_rt_case(golden)
        prob = dict(frac=0.5, rho_1=1.0, rho_2=2.0, p0_base=5.0)
    else:
        par, geom, prob, g, snaps = load_kh_case(golden)
    return par, geom, prob, g


@pytest.mark.parametrize("hook,builtin,golden,kind", HOOK_CASES)
def test_reference_hook_reproduces_reference_initial_plotfile(hook, builtin, golden, kind):
    import oracle_lib as O
    par, geom, prob, g = load_any(golden, kind)
    O.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    L = C.CDLL(hook_path(hook))
    assert L.ctu_hook_num_state() == par.num_state
    U = np.full((par.num_state, n_cell[2], n_cell[1], n_cell[0]), np.nan)
    lo, hi = abi.Int3(0, 0, 0), abi.Int3(*[n - 1 for n in n_cell])
    L.ctu_hook_init.argtypes = [C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.c_void_p, abi.Int3, abi.Int3]
    rc = L.ctu_hook_init(prob_string(prob).encode(), C.byref(par), C.byref(geom), lo, hi, U.ctypes.data, lo, hi)
    assert rc == 0
    f = O.fab(U, (0, 0, 0))
    O.lib().orc_init_tail(C.byref(par), C.byref(geom), lo, hi, C.byref(f))
    assert np.array_equal(U, g["U0"]), "hook %s vs the reference's plt00000 of %s" % (hook, golden)


def test_hook_rejects_wrong_network_and_unknown_parameter():
    par, geom, prob, g = load_any("sedov16_hllc", "sedov")
    L = C.CDLL(hook_path("kh4"))
    L.ctu_hook_init.argtypes = [C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.c_void_p, abi.Int3, abi.Int3]
    U = np.zeros((12, 4, 4, 4))
    lo, hi = abi.Int3(0, 0, 0), abi.Int3(3, 3, 3)
    assert L.ctu_hook_init(b"", C.byref(par), C.byref(geom), lo, hi, U.ctypes.data, lo, hi) == -1      # NumSpec mismatch
    L = C.CDLL(hook_path("sedov"))
    L.ctu_hook_init.argtypes = [C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.c_void_p, abi.Int3, abi.Int3]
    U = np.zeros((8, 4, 4, 4))
    assert L.ctu_hook_init(b"no_such_parameter=1", C.byref(par), C.byref(geom), lo, hi, U.ctypes.data, lo, hi) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("hook,builtin,golden,kind", HOOK_CASES)
def test_problem_init_through_hook_on_gpu(hook, builtin, golden, kind):
    """ctu_problem_init("hook:<lib>") (hook on the host, upload, initData tail on the device) against the reference's initial
    plotfile and against the built-in problem kernels"""
    import torch
    assert torch.cuda.is_available()
    from castro_b200 import lib
    par, geom, prob, g = load_any(golden, kind)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    hi = tuple(n - 1 for n in n_cell)
    a = lib.Sim(par, geom, (0, 0, 0), hi)
    a.init_problem("hook:" + hook_path(hook), prob)
    b = lib.Sim(par, geom, (0, 0, 0), hi)
    b.init_problem(builtin, prob)
    Ua, Ub = a.download(), b.download()
    assert np.array_equal(Ua, g["U0"])
    assert np.array_equal(Ua, Ub)
    a.close()
    b.close()


# ---- the boundary hook: Exec/hydro_tests/Noh (initial data AND outer ghost zones come from the reference's unmodified headers) --------
HOOK_ARGS = [C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.c_void_p, abi.Int3, abi.Int3]
BC_ARGS = [C.c_double, C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.c_void_p, abi.Int3, abi.Int3]


@pytest.mark.parametrize("name", sorted(NOH_CASES))
def test_oracle_with_reference_bc_hook_matches_noh_golden(name):
    """CPU: the oracle's time loop with problem_bc_fill (the reference's header behind ctu_hook_bc_fill) after every ghost fill lands
    bit for bit on the reference executable's Noh run, dt history included"""
    import oracle_lib as O
    par, geom, prob, g, snaps = load_noh_case(name)
    O.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    L = C.CDLL(hook_path("noh"))
    assert L.ctu_hook_has_bc_fill() == 1
    assert C.CDLL(hook_path("sedov")).ctu_hook_has_bc_fill() == 0
    L.ctu_hook_init.argtypes = HOOK_ARGS
    U0 = np.full((par.num_state, n_cell[2], n_cell[1], n_cell[0]), np.nan)
    lo, hi = abi.Int3(0, 0, 0), abi.Int3(*[n - 1 for n in n_cell])
    assert L.ctu_hook_init(b"", C.byref(par), C.byref(geom), lo, hi, U0.ctypes.data, lo, hi) == 0
    f = O.fab(U0, (0, 0, 0))
    O.lib().orc_init_tail(C.byref(par), C.byref(geom), lo, hi, C.byref(f))
    assert np.array_equal(U0, g["U0"])
    ng = 4
    U = np.zeros((U0.shape[0],) + tuple(n + 2 * ng for n in U0.shape[1:]))
    U[:, ng:-ng, ng:-ng, ng:-ng] = U0
    O.lib().orc_set_bc_hook(C.cast(L.ctu_hook_bc_fill, C.c_void_p))
    try:
        t, dt, step, times = 0.0, -1.0, 0, [0.0]
        for s in snaps[1:]:
            st, t, dt, hist = O.run(par, geom, U, s - step, time=t, dt=dt)
            assert st == 0
            times += list(times[-1] + np.cumsum(hist))
            step = s
            assert np.array_equal(U[:, ng:-ng, ng:-ng, ng:-ng], g["U%d" % s]), "%s step %d" % (name, s)
    finally:
        O.lib().orc_set_bc_hook(None)
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    # without the hook the run is a different one (the test above is not vacuous)
    V = np.zeros_like(U)
    V[:, ng:-ng, ng:-ng, ng:-ng] = U0
    O.run(par, geom, V, snaps[-1])
    assert not np.array_equal(V[:, ng:-ng, ng:-ng, ng:-ng], g["U%d" % snaps[-1]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(NOH_CASES))
def test_noh_with_bc_hook_on_gpu(name):
    """ctu_sim_init_problem("hook:<noh>") keeps the library's problem_bc_fill; ctu_sim_advance must then reproduce the reference's
    Noh run bit for bit (inflow face = outflow fill + hook; hook-written ghost zones cleaned once)"""
    import torch
    assert torch.cuda.is_available()
    from castro_b200 import lib
    par, geom, prob, g, snaps = load_noh_case(name)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(n - 1 for n in n_cell))
    sim.init_problem("hook:" + hook_path("noh"), prob)
    assert np.array_equal(sim.download(), g["U0"])
    t, dt, step, times = 0.0, -1.0, 0, [0.0]
    for s in snaps[1:]:
        st, t, dt, hist = sim.advance(s - step, time=t, dt=dt)
        assert st == 0
        times += list(times[-1] + np.cumsum(hist))
        step = s
        assert np.array_equal(sim.download(), g["U%d" % s]), "%s step %d" % (name, s)
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
    sim.close()


# ---- the same problem headers compiled for AMREX_SPACEDIM = 2 (libctu_hook_sedov2d / kh2d) ---------------------------------------------
HOOK_CASES_2D = [("sedov2d", "sedov2d_cgf"), ("sedov2d", "sedov2d_cg_24x16"), ("sedov2d", "sedov2d_symm_cgf"), ("sedov2d", "sedov2d_rz_cgf"), ("kh2d", "kh2d_cgf"),
                 ("kh2d", "kh2d_hllc"), ("rt2d", "rt2d_cgf")]


@pytest.mark.parametrize("hook,golden", HOOK_CASES_2D)
def test_reference_hook_reproduces_reference_initial_plotfile_2d(hook, golden):
    import oracle_lib as O
    from cases import load_2d_case
    par, geom, prob, g, snaps = load_2d_case(golden)
    O.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    L = C.CDLL(hook_path(hook))
    U = np.full((par.num_state, 1, n_cell[1], n_cell[0]), np.nan)
    lo, hi = abi.Int3(0, 0, 0), abi.Int3(n_cell[0] - 1, n_cell[1] - 1, 0)
    L.ctu_hook_init.argtypes = HOOK_ARGS
    assert L.ctu_hook_init(prob_string(prob).encode(), C.byref(par), C.byref(geom), lo, hi, U.ctypes.data, lo, hi) == 0
    f = O.fab(U, (0, 0, 0))
    O.lib().orc_init_tail(C.byref(par), C.byref(geom), lo, hi, C.byref(f))
    assert np.array_equal(U, g["U0"]), "hook %s vs the reference's plt00000 of %s" % (hook, golden)


@pytest.mark.parametrize("golden", ["sedov1d_sph_cgf", "sedov1d_cyl_cgf"])
def test_reference_hook_reproduces_reference_initial_plotfile_1d_curvilinear(golden):
    """the Sedov headers compiled for AMREX_SPACEDIM = 1 in spherical / cylindrical coordinates (volume-weighted sub-zone sampling)"""
    import oracle_lib as O
    from cases import load_sedov1d_case
    par, geom, prob, g, snaps = load_sedov1d_case(golden)
    O.params_finalize(par)
    nx = int(g["n_cell"][0])
    L = C.CDLL(hook_path("sedov1d"))
    U = np.full((par.num_state, 1, 1, nx), np.nan)
    lo, hi = abi.Int3(0, 0, 0), abi.Int3(nx - 1, 0, 0)
    L.ctu_hook_init.argtypes = HOOK_ARGS
    assert L.ctu_hook_init(prob_string(prob).encode(), C.byref(par), C.byref(geom), lo, hi, U.ctypes.data, lo, hi) == 0
    f = O.fab(U, (0, 0, 0))
    O.lib().orc_init_tail(C.byref(par), C.byref(geom), lo, hi, C.byref(f))
    assert np.array_equal(U, g["U0"]), "hook vs the reference's plt00000 of %s" % golden
