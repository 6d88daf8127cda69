"""GPU tests of the multi-box drivers: the C++ one over NCCL (ctu_dist_*, castro_b200.lib.Dist -- what bench.py runs at N > 1) and
the portable Python one (castro_b200/dist.py with the CUDA box engine).  Both must reproduce the reference executable's golden states
bit for bit; with >= 2 visible GPUs the NCCL halo exchange, the all-reduce and the host-streamed multi-box step are run too."""
import os
import socket
import sys

import numpy as np
import pytest

from cases import load_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_distsim_single_box_matches_golden():
    import torch
    assert torch.cuda.is_available()
    from castro_b200 import lib
    from castro_b200.dist import DistSim, CudaBoxEngine
    for name in ("sedov16_hllc_long", "sedov16_periodic_hllc"):
        par, geom, prob, g, snaps = load_case(name)
        lib.params_finalize(par)
        n_cell = tuple(int(x) for x in g["n_cell"])
        sim = DistSim(par, geom, n_cell, CudaBoxEngine)
        sim.init_problem("Sedov", prob)
        n = snaps[-1]
        st, hist = sim.advance(n)
        assert st == 0 and len(hist) == n
        assert np.array_equal(sim.engine.download(), g["U%d" % n])
        np.testing.assert_allclose(np.cumsum([0.0] + hist), g["times"][:n + 1], rtol=1e-14, atol=0)
        sim.close()


def test_cpp_dist_single_box_matches_golden():
    """ctu_dist_* with one box: the same time loop as ctu_sim_advance, no collectives"""
    import torch
    assert torch.cuda.is_available()
    from castro_b200 import lib
    for name in ("sedov16_hllc_long", "sedov16_periodic_hllc"):
        par, geom, prob, g, snaps = load_case(name)
        lib.params_finalize(par)
        d = lib.Dist(par, geom, (1, 1, 1))
        d.init_problem("Sedov", prob)
        n = snaps[-1]
        st, hist = d.advance(n)
        assert st == 0 and len(hist) == n
        assert np.array_equal(d.sim.download(), g["U%d" % n])
        np.testing.assert_allclose(np.cumsum([0.0] + hist), g["times"][:n + 1], rtol=1e-14, atol=0)
        # host-streamed step through the multi-box entry point (one box: no exchange)
        d2 = lib.Dist(par, geom, (1, 1, 1))
        d2.init_problem("Sedov", prob)
        from castro_b200.dist import initial_dt, new_dt
        h = d2.sim.download()
        h2 = np.empty_like(h)
        t, dt = 0.0, initial_dt(par, d2.estdt_global(), 0.0, -1.0)
        for _ in range(n):
            st, est = d2.step_host(h, h2, (t + dt) - t)
            assert st == 0
            h, h2 = h2, h
            t += dt
            dt = new_dt(par, est[1], dt, t, -1.0)
        assert np.array_equal(h, g["U%d" % n])
        d.close()
        d2.close()


def _worker_cpp(rank, world, port, kind, name, nsteps, dims, out, streamed):
    """C++ multi-box driver on `world` GPUs; kind = sedov | rt"""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CTU_HOST_CHUNK"] = "4"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from castro_b200 import lib
        from castro_b200.dist import rank_coords, box_of, initial_dt, new_dt
        if kind == "rt":
            from cases import load_rt_case
            par, geom, g, snaps = load_rt_case(name)
            prob = None
        elif kind == "2d":
            from cases import load_2d_case
            par, geom, prob, g, snaps = load_2d_case(name)
            prob = None
        elif kind == "helm":
            from cases import load_sodstellar_case, HELM_TABLE
            lib.helm_load(HELM_TABLE)
            par, geom, prob, g, snaps = load_sodstellar_case(name)
            prob = None
        else:
            par, geom, prob, g, snaps = load_case(name)
        lib.params_finalize(par)
        n_cell = tuple(int(x) for x in g["n_cell"])
        d = lib.Dist(par, geom, dims, rank=rank, world=world, comm=dist)
        lo, hi = d.lo, d.hi
        assert (lo, hi) == box_of(rank_coords(rank, dims), dims, n_cell)
        if prob is None:
            d.sim.upload(np.ascontiguousarray(g["U0"][:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]))
        else:
            d.init_problem("Sedov", prob)
        if streamed:
            h = d.sim.download()
            h2 = np.empty_like(h)
            t, dt = 0.0, initial_dt(par, d.estdt_global(), 0.0, -1.0)
            hist = []
            for _ in range(nsteps):
                st, est = d.step_host(h, h2, (t + dt) - t)
                assert st == 0
                h, h2 = h2, h
                hist.append(dt)
                t += dt
                dt = new_dt(par, est[1], dt, t, -1.0)
            mine = torch.from_numpy(np.ascontiguousarray(h)).cuda()
        else:
            st, hist = d.advance(nsteps)
            assert st == 0 and len(hist) == nsteps
            mine = torch.from_numpy(d.sim.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            full = np.empty((par.num_state,) + tuple(n_cell[a] for a in (2, 1, 0)))
            for r in range(world):
                blo, bhi = box_of(rank_coords(r, dims), dims, n_cell)
                full[:, blo[2]:bhi[2] + 1, blo[1]:bhi[1] + 1, blo[0]:bhi[0] + 1] = parts[r].cpu().numpy()
            np.savez(out, U=full, times=np.cumsum([0.0] + list(hist)))
        d.close()
    finally:
        dist.destroy_process_group()


def _spawn2(tmp_path, *args):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "o.npz")
    mp.spawn(_worker_cpp, args=(2, port) + args[:4] + (out,) + args[4:], nprocs=2, join=True)
    return np.load(out)


@pytest.mark.parametrize("name,nsteps,dims,streamed", [
    ("sedov16_hllc", 5, (2, 1, 1), False), ("sedov16_periodic_hllc", 10, (1, 1, 2), False), ("sedov16_periodic_hllc", 10, (2, 1, 1), False),
    ("sedov16_hllc", 5, (2, 1, 1), True), ("sedov16_hllc", 5, (1, 1, 2), True), ("sedov16_periodic_hllc", 10, (1, 1, 2), True),
    ("sedov16_periodic_hllc", 10, (1, 2, 1), True), ("sedov32_hllc", 10, (1, 1, 2), True)])
def test_two_gpus_cpp_dist(tmp_path, name, nsteps, dims, streamed):
    """C++ multi-box driver on two GPUs (resident time loop and host-streamed step) against the single-run goldens, bit for bit"""
    r = _spawn2(tmp_path, "sedov", name, nsteps, dims, streamed)
    par, geom, prob, g, snaps = load_case(name)
    assert np.array_equal(r["U"], g["U%d" % nsteps])
    np.testing.assert_allclose(r["times"], g["times"][:nsteps + 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name,dims", [("rt16_cgf", (1, 1, 2)), ("rt16_hllc_gst2", (2, 1, 1))])
def test_two_gpus_cpp_dist_gravity(tmp_path, name, dims):
    """Rayleigh-Taylor with constant gravity on two boxes through the C++ driver"""
    from cases import load_rt_case
    par, geom, g, snaps = load_rt_case(name)
    n = snaps[-1]
    r = _spawn2(tmp_path, "rt", name, n, dims, False)
    assert np.array_equal(r["U"], g["U%d" % n])
    np.testing.assert_allclose(r["times"], g["times"][:n + 1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name,dims", [("sodstellar_z_grav_cgf", (1, 1, 2)), ("sodstellar_x_grav_hllc_gst2", (2, 1, 1)), ("sodstellar_y_hllc", (1, 2, 1))])
def test_two_gpus_cpp_dist_helmholtz(tmp_path, name, dims):
    """The Helmholtz EOS (with and without constant gravity: the combination BASELINE config[4] names) on two boxes through the C++
    driver: bit for bit the single-box run on one GPU from the same initial state, and the reference's golden to the 1e-12 of the
    single-box test."""
    import torch
    from castro_b200 import lib
    from cases import load_sodstellar_case, HELM_TABLE
    if not os.path.exists(HELM_TABLE):
        pytest.skip("oracle/_ref/helm_table.dat missing")
    par, geom, prob, g, snaps = load_sodstellar_case(name)
    n = snaps[-1]
    r = _spawn2(tmp_path, "helm", name, n, dims, False)
    lib.helm_load(HELM_TABLE)
    lib.params_finalize(par)
    n_cell = tuple(int(x) for x in g["n_cell"])
    sim = lib.Sim(par, geom, (0, 0, 0), tuple(x - 1 for x in n_cell))
    sim.upload(np.ascontiguousarray(g["U0"]))
    rc, t, dt, hist = sim.advance(n)
    assert rc == 0
    whole = sim.download()
    sim.close()
    assert np.array_equal(r["U"], whole), "decomposed run differs from the single-box run"
    np.testing.assert_array_equal(r["times"], np.cumsum([0.0] + list(hist)))
    ref = g["U%d" % n]
    scale = [max(np.abs(ref[m]).max(), 1e-300) for m in range(ref.shape[0])]
    scale[1] = scale[2] = scale[3] = max(scale[1:4])
    for m in range(ref.shape[0]):
        assert np.abs(r["U"][m] - ref[m]).max() / scale[m] <= (1e-9 if m == 6 else 1e-12), (name, m)


@pytest.mark.parametrize("name,dims", [("sedov2d_hllc", (2, 1, 1)), ("kh2d_cgf", (1, 2, 1)), ("rt2d_cgf", (1, 2, 1)), ("sedov2d_rz_cgf", (2, 1, 1))])
def test_two_gpus_cpp_dist_two_dimensional(tmp_path, name, dims):
    """2-D runs (Cartesian, periodic, with gravity, R-Z) on two boxes through the C++ driver against the single-run goldens, bit for bit"""
    from cases import load_2d_case
    par, geom, prob, g, snaps = load_2d_case(name)
    n = snaps[-1]
    r = _spawn2(tmp_path, "2d", name, n, dims, False)
    assert np.array_equal(r["U"], g["U%d" % n])
    np.testing.assert_allclose(r["times"], g["times"][:n + 1], rtol=1e-14, atol=0)


def _worker(rank, world, port, name, nsteps, dims, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from castro_b200 import lib
        from castro_b200.dist import DistSim, CudaBoxEngine, rank_coords, box_of
        par, geom, prob, g, snaps = load_case(name)
        lib.params_finalize(par)
        n_cell = tuple(int(x) for x in g["n_cell"])
        sim = DistSim(par, geom, n_cell, CudaBoxEngine, comm=dist, rank=rank, world=world, dims=dims)
        sim.init_problem("Sedov", prob)
        st, hist = sim.advance(nsteps)
        assert st == 0
        mine = torch.from_numpy(sim.engine.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            full = np.empty((par.num_state,) + tuple(n_cell[d] for d in (2, 1, 0)))
            for r in range(world):
                lo, hi = box_of(rank_coords(r, sim.dims), sim.dims, n_cell)
                full[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = parts[r].cpu().numpy()
            np.savez(out, U=full)
        sim.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,nsteps,dims", [("sedov16_hllc", 5, (2, 1, 1)), ("sedov16_periodic_hllc", 10, (1, 1, 2))])
def test_two_gpus_nccl_halo_exchange(tmp_path, name, nsteps, dims):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "o.npz")
    mp.spawn(_worker, args=(2, port, name, nsteps, dims, out), nprocs=2, join=True)
    par, geom, prob, g, snaps = load_case(name)
    assert np.array_equal(np.load(out)["U"], g["U%d" % nsteps])


def _worker_rt(rank, world, port, name, dims, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cases import load_rt_case
        from castro_b200 import lib
        from castro_b200.dist import DistSim, CudaBoxEngine, rank_coords, box_of
        par, geom, g, snaps = load_rt_case(name)
        lib.params_finalize(par)
        n_cell = tuple(int(x) for x in g["n_cell"])
        sim = DistSim(par, geom, n_cell, CudaBoxEngine, comm=dist, rank=rank, world=world, dims=dims)
        lo, hi = sim.lo, sim.hi
        sim.engine.sim.upload(np.ascontiguousarray(g["U0"][:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]))
        n = snaps[-1]
        st, hist = sim.advance(n)
        assert st == 0 and len(hist) == n
        mine = torch.from_numpy(sim.engine.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            full = np.empty((par.num_state,) + tuple(n_cell[d] for d in (2, 1, 0)))
            for r in range(world):
                blo, bhi = box_of(rank_coords(r, sim.dims), sim.dims, n_cell)
                full[:, blo[2]:bhi[2] + 1, blo[1]:bhi[1] + 1, blo[0]:bhi[0] + 1] = parts[r].cpu().numpy()
            np.savez(out, U=full, times=np.cumsum([0.0] + hist))
        sim.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,dims", [("rt16_cgf", (1, 1, 2)), ("rt16_hllc_gst2", (2, 1, 1))])
def test_two_gpus_gravity_sources(tmp_path, name, dims):
    """Rayleigh-Taylor with constant gravity on two boxes (cut across the gravity direction / across a periodic direction): the
    old-time source on the 3 ghost layers is computed from the exchanged ghost zones of Sborder, so the two-box run must land on the
    reference's single-run golden state bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from cases import load_rt_case
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "o.npz")
    mp.spawn(_worker_rt, args=(2, port, name, dims, out), nprocs=2, join=True)
    par, geom, g, snaps = load_rt_case(name)
    n = snaps[-1]
    r = np.load(out)
    assert np.array_equal(r["U"], g["U%d" % n])
    np.testing.assert_allclose(r["times"], g["times"][:n + 1], rtol=1e-14, atol=0)
