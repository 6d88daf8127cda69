"""GPU tests of the multi-box driver (castro_b200/dist.py) with the CUDA box engine: the Python dt control + Sim phases must
reproduce the reference executable's golden states bit for bit; with >= 2 visible GPUs the NCCL halo exchange is run too."""
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
