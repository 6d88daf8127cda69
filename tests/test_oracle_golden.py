"""Pins the CPU oracle (oracle/ctu_oracle.cpp) against golden states produced by the UNMODIFIED reference
executable (oracle/make_golden.py).  Bit-exact: the oracle restates the reference's arithmetic operation
for operation, and the reference binary contains no FMA instructions."""
import numpy as np
import pytest

import oracle_lib as O
from cases import CASES, load_case

NG = 4


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_bitwise(name):
    par, geom, prob, g, snaps = load_case(name)
    O.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    U = O.sedov_init(par, geom, n_cell, **prob)
    t, dt, step = 0.0, -1.0, 0
    times = [0.0]
    for s in snaps:
        if s > step:
            st, t, dt, hist = O.run(par, geom, U, s - step, time=t, dt=dt)
            assert st == 0
            tt = times[-1]
            for h in hist:
                tt += h
                times.append(tt)
            step = s
        v = U[:, NG:-NG, NG:-NG, NG:-NG]
        ref = g["U%d" % s]
        for n in range(ref.shape[0]):
            nbad = int((v[n] != ref[n]).sum())
            assert nbad == 0, "%s step %d comp %d: %d zones differ, max abs %g" % (
                name, s, n, nbad, np.abs(v[n] - ref[n]).max())
    # plotfile headers print the time with 15 significant digits
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", ["kh16_cgf", "kh16_hllc", "kh16_cgf_nolim", "kh16_4spec_cgf", "kh16_4spec_hllc"])
def test_oracle_two_passives_matches_kh_golden(name):
    """NUM_STATE = 9 (advected scalar + species): the oracle, started from the reference's own initial plotfile, must land
    bit for bit on the reference's state after 10 steps (pins the passive indexing and the periodic time loop)."""
    import oracle_lib as O
    from cases import load_kh_case
    par, geom, prob, g, snaps = load_kh_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    U = np.zeros((U0.shape[0],) + tuple(n + 2 * ng for n in U0.shape[1:]))
    U[:, ng:-ng, ng:-ng, ng:-ng] = U0
    n = snaps[-1]
    st, t, dt, hist = O.run(par, geom, U, n)
    assert st == 0 and len(hist) == n
    assert np.array_equal(U[:, ng:-ng, ng:-ng, ng:-ng], g["U%d" % n])
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", ["rt16_cgf", "rt16_hllc_gst2", "rt16_pslope", "rt16_wellbal", "rt16_plm_pslope", "rt16_combo"])
def test_oracle_gravity_sources_match_rt_golden(name):
    """Constant gravity (Rayleigh-Taylor): old-time source predictor, src_to_prim + traced sources inside the hydro, and the
    new-time corrector (grav_source_type 4 from the mass fluxes, or 2) -- 20 steps from the reference's own initial plotfile."""
    from cases import load_rt_case
    par, geom, g, snaps = load_rt_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    U = np.zeros((U0.shape[0],) + tuple(n + 2 * ng for n in U0.shape[1:]))
    U[:, ng:-ng, ng:-ng, ng:-ng] = U0
    n = snaps[-1]
    st, t, dt, hist = O.run(par, geom, U, n)
    assert st == 0 and len(hist) == n
    got, ref = U[:, ng:-ng, ng:-ng, ng:-ng], g["U%d" % n]
    for m in range(ref.shape[0]):
        nbad = int((got[m] != ref[m]).sum())
        assert nbad == 0, "%s comp %d: %d zones differ, max abs %g" % (name, m, nbad, np.abs(got[m] - ref[m]).max())
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", ["sod_x_cgf", "sod_y_cgf", "sod_z_cgf", "sod_z_hllc", "sod_x_symm_ppm", "sod_y_symm_plm"])
def test_oracle_sod_shock_tube_matches_golden(name):
    """BASELINE config[0]: the Sod shock tube (PPM + CGF; one HLLC variant), run in 3-D with the tube along x, y and z -- the oracle,
    started from the reference's own initial plotfile, must land bit for bit on the reference's state after 30 steps (pins the three
    directional code paths, outflow along the tube and slip walls across it)."""
    from cases import load_sod_case
    par, geom, prob, g, snaps = load_sod_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    U = np.zeros((U0.shape[0],) + tuple(n + 2 * ng for n in U0.shape[1:]))
    U[:, ng:-ng, ng:-ng, ng:-ng] = U0
    n = snaps[-1]
    st, t, dt, hist = O.run(par, geom, U, n)
    assert st == 0 and len(hist) == n
    got, ref = U[:, ng:-ng, ng:-ng, ng:-ng], g["U%d" % n]
    for m in range(ref.shape[0]):
        nbad = int((got[m] != ref[m]).sum())
        assert nbad == 0, "%s comp %d: %d zones differ, max abs %g" % (name, m, nbad, np.abs(got[m] - ref[m]).max())
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", ["sod1d_cgf", "sod1d_cg", "sod1d_test2_cgf", "sod1d_test3_plm"])
def test_oracle_sod_1d_matches_golden(name):
    """BASELINE config[0] literally -- Sod in ONE dimension (the reference built with DIM=1; Castro_ctu_hydro.cpp:464-481: trace in x,
    one Riemann solve per face, no transverse terms): the oracle's 1-D path, started from the reference's initial plotfile, against
    the reference's states (sod1d_cgf runs to stop_time = 0.2 in 81 steps) and times, bit for bit"""
    import oracle_lib as O
    from cases import load_sod1d_case
    par, geom, prob, g, snaps = load_sod1d_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    nx = U0.shape[3]
    U = np.zeros((U0.shape[0], 1 + 2 * ng, 1 + 2 * ng, nx + 2 * ng))
    U[:, ng, ng, ng:-ng] = U0[:, 0, 0, :]
    n = snaps[-1]
    stop = 0.2 if name == "sod1d_cgf" else -1.0
    st, t, dt, hist = O.run(par, geom, U, n, stop_time=stop)
    assert st == 0 and len(hist) == n
    assert np.array_equal(U[:, ng, ng, ng:-ng], g["U%d" % n][:, 0, 0, :])
    np.testing.assert_allclose(np.cumsum(np.concatenate([[0.0], hist])), g["times"][:n + 1], rtol=1e-14, atol=0)
    if name == "sod1d_cgf":
        assert t == 0.2


from cases import CASES_2D


@pytest.mark.parametrize("name", sorted(CASES_2D))
def test_oracle_two_dimensional_matches_golden(name):
    """The 2-D Cartesian path (the reference built with DIM=2; Castro_ctu_hydro.cpp:530-650: one transverse correction per direction in
    the area / volume form of trans.cpp, 2-D div u, two-node artificial viscosity average, dx dy volumes): the oracle, started from the
    reference's initial plotfile, against the reference's states and times, bit for bit"""
    import oracle_lib as O
    from cases import load_2d_case
    par, geom, prob, g, snaps = load_2d_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    ny, nx = U0.shape[2], U0.shape[3]
    U = np.zeros((U0.shape[0], 1 + 2 * ng, ny + 2 * ng, nx + 2 * ng))
    U[:, ng, ng:-ng, ng:-ng] = U0[:, 0]
    t, dt, step, times = 0.0, -1.0, 0, [0.0]
    for s in snaps[1:]:
        st, t, dt, hist = O.run(par, geom, U, s - step, time=t, dt=dt)
        assert st == 0 and len(hist) == s - step
        times += list(times[-1] + np.cumsum(hist))
        step = s
        assert np.array_equal(U[:, ng, ng:-ng, ng:-ng], g["U%d" % s][:, 0]), "%s step %d" % (name, s)
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)


from cases import SEDOV1D_CASES


@pytest.mark.parametrize("name", sorted(SEDOV1D_CASES))
def test_oracle_one_dimensional_curvilinear_matches_golden(name):
    """1-D spherical / cylindrical coordinates (geometry.coord_sys 2 / 1): geometric source terms in the PPM / PLM tracing, r-dependent
    areas and volumes in consup_hydro, the radial momentum flux without pressure + the grad p term, the curvilinear div u -- the
    oracle, started from the reference's initial plotfile, against the reference's states and times, bit for bit"""
    import oracle_lib as O
    from cases import load_sedov1d_case
    par, geom, prob, g, snaps = load_sedov1d_case(name)
    O.params_finalize(par)
    U0 = g["U0"]
    ng = 4
    nx = U0.shape[3]
    U = np.zeros((U0.shape[0], 1 + 2 * ng, 1 + 2 * ng, nx + 2 * ng))
    U[:, ng, ng, ng:-ng] = U0[:, 0, 0, :]
    t, dt, step, times = 0.0, -1.0, 0, [0.0]
    for s in snaps[1:]:
        st, t, dt, hist = O.run(par, geom, U, s - step, time=t, dt=dt)
        assert st == 0 and len(hist) == s - step
        times += list(times[-1] + np.cumsum(hist))
        step = s
        assert np.array_equal(U[:, ng, ng, ng:-ng], g["U%d" % s][:, 0, 0, :]), "%s step %d" % (name, s)
    np.testing.assert_allclose(np.array(times), g["times"][:len(times)], rtol=1e-14, atol=0)
