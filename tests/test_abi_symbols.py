"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every function that
include/castro_ctu.h declares, the ctypes mirror matches the C struct sizes, and compute entry points fail loudly
(CTU_ERR_NO_DEVICE) instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "castro_ctu.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from castro_b200 import lib
    L = lib.lib()
    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/castro_ctu.h but not exported: %s" % missing


def test_struct_layout_matches_header(tmp_path):
    from castro_b200 import abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "castro_ctu.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(ctu_fab_t), '
                   'sizeof(ctu_geom_t), sizeof(ctu_params_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    assert [int(x) for x in out] == [C.sizeof(abi.Fab), C.sizeof(abi.Geom), C.sizeof(abi.Params)]


def test_params_defaults_and_finalize_without_gpu():
    from castro_b200 import lib, abi
    p = lib.default_params()
    q = abi.default_params()
    for name, _ in abi.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    lib.params_finalize(p)
    assert p.small_dens == 1.e-100 and p.small_pres > 0 and p.small_ener > 0


def test_no_cpu_fallback():
    """without a CUDA device the compute entry points must refuse (no silent CPU path)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from castro_b200 import lib, abi
    p = lib.params_finalize(lib.default_params())
    g = abi.make_geom((8, 8, 8))
    with pytest.raises(lib.CtuError, match="no CUDA device"):
        lib.Sim(p, g, (0, 0, 0), (7, 7, 7))
    with pytest.raises(lib.CtuError, match="no CUDA device"):
        lib.Workspace()


def test_parse_inputs_file():
    from castro_b200 import lib
    p, g, nc, ms, stop, prob = lib.parse_inputs(os.path.join(ROOT, "bench_inputs", "inputs.sedov3d"),
                                                "amr.n_cell=48 32 16 castro.cfl=0.4")
    assert nc == (48, 32, 16) and p.cfl == 0.4 and p.riemann_solver == 2 and p.eos_gamma == 1.4
    assert list(g.lo_bc) == [2, 2, 2] and g.dx[0] == 1.0 / 48 and ms == 10
    assert "r_init=0.01" in prob and "nsub=10" in prob


def test_parse_inputs_all_hydro_options():
    """every castro.* / gravity.* key of the hot path that the library honours is picked up from an inputs file (the RT file carries
    the gravity keys) and from command-line style overrides, with the reference's defaults for the rest"""
    from castro_b200 import lib
    p, g, nc, ms, stop, prob = lib.parse_inputs(os.path.join(ROOT, "bench_inputs", "inputs.rt3d"), "")
    assert p.do_grav == 1 and p.const_grav == -1.0 and p.grav_source_type == 4
    assert list(g.periodic) == [1, 1, 0] and list(g.lo_bc) == [0, 0, 4] and nc == (16, 16, 32)
    assert (p.ppm_type, p.riemann_solver, p.hybrid_riemann, p.use_pslope, p.ppm_well_balanced, p.ppm_temp_fix) == (1, 0, 0, 0, 0, 0)
    assert (p.plm_iorder, p.plm_limiter, p.riemann_shock_maxiter, p.riemann_cg_blend) == (2, 2, 12, 2)
    assert p.riemann_pstar_tol == 1.0e-5 and p.shock_detection_threshold == 2.0 / 3.0 and p.shock_detection_include_sources == 1
    assert p.pslope_cutoff_density == -1.e20 and p.limit_fluxes_on_small_dens == 0
    over = ("castro.ppm_type=0 castro.plm_limiter=1 castro.riemann_solver=1 castro.riemann_cg_blend=1 castro.riemann_shock_maxiter=20 "
            "castro.hybrid_riemann=1 castro.shock_detection_threshold=0.5 castro.use_pslope=1 castro.ppm_well_balanced=1 "
            "castro.pslope_cutoff_density=1.e-3 castro.ppm_temp_fix=2 castro.limit_fluxes_on_small_dens=1 castro.transverse_use_eos=1 "
            "castro.transverse_reset_rhoe=1 castro.source_term_predictor=1 castro.grav_source_type=2 gravity.const_grav=-9.8 "
            "castro.first_order_hydro=1 castro.use_flattening=0 castro.ppm_do_limiting=0 castro.difmag=0.2")
    p, g, nc, ms, stop, prob = lib.parse_inputs(os.path.join(ROOT, "bench_inputs", "inputs.rt3d"), over)
    assert (p.ppm_type, p.plm_limiter, p.riemann_solver, p.riemann_cg_blend, p.riemann_shock_maxiter) == (0, 1, 1, 1, 20)
    assert (p.hybrid_riemann, p.use_pslope, p.ppm_well_balanced, p.ppm_temp_fix, p.limit_fluxes_on_small_dens) == (1, 1, 1, 2, 1)
    assert (p.transverse_use_eos, p.transverse_reset_rhoe, p.source_term_predictor, p.grav_source_type) == (1, 1, 1, 2)
    assert (p.first_order_hydro, p.use_flattening, p.ppm_do_limiting) == (1, 0, 0)
    assert p.shock_detection_threshold == 0.5 and p.pslope_cutoff_density == 1.e-3 and p.const_grav == -9.8 and p.difmag == 0.2


def test_parse_inputs_refuses_keys_that_leave_the_supported_path():
    """keys that switch the reference onto another path (AMR levels, SDC integrators, sponge, reactions, ...) must not be
    ignored silently"""
    from castro_b200 import lib
    inputs = os.path.join(ROOT, "bench_inputs", "inputs.sedov3d")
    lib.parse_inputs(inputs, "castro.cfl=0.4")
    for bad in ("amr.max_level=1", "castro.time_integration_method=3", "castro.do_sponge=1", "castro.hybrid_hydro=1", "castro.do_react=1",
                "castro.use_reconstructed_gamma1=1", "castro.speed_limit=1.e9"):
        with pytest.raises(lib.CtuError):
            lib.parse_inputs(inputs, bad)


def test_parse_inputs_dimensionality_and_coordinates(tmp_path):
    """the number of entries of amr.n_cell gives the dimensionality of the run (the reference fixes AMREX_SPACEDIM at build time);
    geometry.coord_sys and array-valued problem parameters come through (files laid out like Exec/hydro_tests/Sedov/inputs.2d.sph_in_cylcoords,
    inputs.1d.sph and KH/inputs.2d)"""
    from castro_b200 import lib
    f = tmp_path / "inputs.2d"
    f.write_text("max_step = 7\ngeometry.is_periodic = 0 0\ngeometry.coord_sys = 1\ngeometry.prob_lo = 0 0\ngeometry.prob_hi = 0.5 0.25\n"
                 "amr.n_cell = 16 8\ncastro.lo_bc = 3 3\ncastro.hi_bc = 2 2\ncastro.cfl = 0.5\namr.max_level = 0\n"
                 "problem.center = 0.5 0.5 0.0\nproblem.r_init = 0.01\n")
    p, g, nc, ms, stop, prob = lib.parse_inputs(str(f))
    assert g.dim == 2 and g.coord == 1 and nc == (16, 8, 1) and ms == 7
    assert list(g.lo_bc)[:2] == [3, 3] and g.dx[0] == 0.5 / 16 and g.dx[1] == 0.25 / 8 and g.domhi[2] == 0
    assert "center=0.5,0.5,0.0" in prob and "r_init=0.01" in prob
    f1 = tmp_path / "inputs.1d"
    f1.write_text("geometry.coord_sys = 2\ngeometry.prob_lo = 0\ngeometry.prob_hi = 1\namr.n_cell = 32\ncastro.lo_bc = 3\ncastro.hi_bc = 2\n")
    p, g, nc, ms, stop, prob = lib.parse_inputs(str(f1))
    assert g.dim == 1 and g.coord == 2 and nc == (32, 1, 1) and g.dx[0] == 1.0 / 32
    p, g, nc, ms, stop, prob = lib.parse_inputs(os.path.join(ROOT, "bench_inputs", "inputs.sedov3d"))
    assert g.dim == 3 and g.coord == 0


def test_every_translation_unit_is_in_the_makefile():
    """the library is built from one translation unit per option family and passive count (ctu_np*, ctu_rec*, ctu_ext*, ctu_helm*):
    every .cu file of castro_b200/csrc must be listed in the Makefile's SRCS (the launchers declare their instantiations `extern`, so a
    forgotten unit would only show up as an unresolved symbol when the library is loaded)."""
    import glob
    import re
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "castro_b200", "csrc")
    mk = open(os.path.join(csrc, "Makefile")).read()
    srcs = set(re.search(r"^SRCS := (.*)$", mk, re.M).group(1).split())
    on_disk = {os.path.basename(f) for f in glob.glob(os.path.join(csrc, "*.cu"))}
    assert on_disk == srcs, (sorted(on_disk - srcs), sorted(srcs - on_disk))
