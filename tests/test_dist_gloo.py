"""Host-side multi-box logic on CPU: world_size 2 (and 4) over gloo, with the CPU checker as the box engine.
The assembled N-box state must be BITWISE identical to the reference executable's single-run golden state."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import load_case

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, nsteps, dims, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_lib as O
        from oracle_engine import OracleBoxEngine
        from castro_b200.dist import DistSim
        if "2d" in name:          # 2-D runs start from the reference's own initial plotfile
            from cases import load_2d_case
            par, geom, prob, g, snaps = load_2d_case(name)
        else:
            par, geom, prob, g, snaps = load_case(name)
        O.params_finalize(par)
        n_cell = tuple(int(x) for x in g["n_cell"])
        sim = DistSim(par, geom, n_cell, OracleBoxEngine, comm=dist, rank=rank, world=world, dims=dims)
        if "2d" in name:
            sim.init_problem("upload", dict(U0=g["U0"]))
        else:
            sim.init_problem("Sedov", prob)
        status, hist = sim.advance(nsteps)
        assert status == 0
        mine = torch.from_numpy(sim.engine.download())
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            from castro_b200.dist import rank_coords, box_of
            full = np.empty((par.num_state,) + tuple(n_cell[d] for d in (2, 1, 0)))
            for r in range(world):
                lo, hi = box_of(rank_coords(r, sim.dims), sim.dims, n_cell)
                full[:, lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = parts[r].numpy()
            np.savez(out, U=full, times=np.cumsum([0.0] + hist))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,nsteps,world,dims", [
    ("sedov16_hllc", 5, 2, (2, 1, 1)),
    ("sedov16_periodic_hllc", 10, 2, (1, 2, 1)),     # both y-neighbours are the same peer: message-order check
    ("sedov16_cgf", 5, 4, (2, 1, 2)),
    # 2-D runs decomposed in x, in y and in both (kh2d: periodic, every box is its own neighbour's neighbour twice)
    ("sedov2d_hllc", 40, 2, (2, 1, 1)), ("kh2d_cgf", 20, 2, (1, 2, 1)), ("kh2d_hllc", 20, 4, (2, 2, 1)), ("sedov2d_rz_cgf", 80, 2, (1, 2, 1)),
])
def test_multibox_matches_reference_golden(tmp_path, name, nsteps, world, dims):
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), name, nsteps, dims, out), nprocs=world, join=True)
    r = np.load(out)
    if "2d" in name:
        from cases import load_2d_case
        par, geom, prob, g, snaps = load_2d_case(name)
    else:
        par, geom, prob, g, snaps = load_case(name)
    ref = g["U%d" % nsteps]
    assert np.array_equal(r["U"], ref), "multi-box run differs from the reference golden state"
    np.testing.assert_allclose(r["times"], g["times"][:nsteps + 1], rtol=1e-14, atol=0)


def test_halo_plan_symmetry():
    """every send has exactly one matching receive on the peer, with equal size"""
    from castro_b200.dist import halo_plan, DIRECTIONS
    for dims, per in [((2, 2, 2), (1, 1, 1)), ((2, 2, 1), (0, 0, 0)), ((2, 1, 1), (1, 0, 1)), ((4, 2, 1), (1, 1, 0))]:
        world = dims[0] * dims[1] * dims[2]
        n_cell = tuple(16 * d for d in dims)
        plans = [halo_plan(r, dims, n_cell, per) for r in range(world)]
        for r, plan in enumerate(plans):
            for (d, peer, slo, shi, rlo, rhi) in plan:
                md = tuple(-x for x in d)
                match = [m for m in plans[peer] if m[0] == md and m[1] == r]
                assert len(match) == 1, (dims, per, r, d)
                m = match[0]
                sz = lambda lo, hi: tuple(hi[a] - lo[a] + 1 for a in range(3))
                assert sz(slo, shi) == sz(m[4], m[5])


@pytest.mark.parametrize("dims,n_cell,periodic", [((2, 1, 1), (16, 16, 16), (0, 0, 0)), ((2, 1, 1), (16, 16, 16), (1, 1, 1)),
                                                   ((2, 2, 1), (32, 16, 8), (1, 0, 1)), ((2, 2, 2), (16, 16, 16), (1, 1, 1)),
                                                   ((2, 2, 2), (16, 32, 16), (0, 0, 0)), ((1, 1, 2), (8, 8, 16), (1, 1, 0)),
                                                   ((3, 1, 2), (24, 8, 16), (1, 1, 1))])
def test_cpp_halo_plan_matches_python_plan(dims, n_cell, periodic):
    """the neighbour table of the C++ multi-box layer (ctu_dist_halo_plan, the one the GPUs run) against castro_b200.dist.halo_plan
    (the one the gloo tests above exercise end to end): same messages in the same send order, same receive order, and every
    message has its mirror image on the peer (sizes agree, pairs of ranks match their messages in posting order)"""
    from castro_b200 import lib
    from castro_b200.dist import halo_plan, DIRECTIONS
    world = dims[0] * dims[1] * dims[2]
    plans = {}
    for rank in range(world):
        msgs, ro = lib.dist_halo_plan(dims, rank, n_cell, periodic)
        ref = halo_plan(rank, dims, n_cell, periodic)
        assert msgs == [(tuple(m[0]), m[1], tuple(m[2]), tuple(m[3]), tuple(m[4]), tuple(m[5])) for m in ref]
        want = sorted(range(len(ref)), key=lambda i: DIRECTIONS.index(tuple(-x for x in ref[i][0])))
        assert ro == want
        plans[rank] = (msgs, ro)
    # NCCL matches the k-th send from a to b with the k-th receive b posts from a
    for a in range(world):
        for b in range(world):
            sends = [m for m in plans[a][0] if m[1] == b]
            recvs = [plans[b][0][i] for i in plans[b][1] if plans[b][0][i][1] == a]
            assert len(sends) == len(recvs)
            for s_, r_ in zip(sends, recvs):
                assert tuple(-x for x in s_[0]) == r_[0]
                assert [s_[3][d] - s_[2][d] for d in range(3)] == [r_[5][d] - r_[4][d] for d in range(3)]
