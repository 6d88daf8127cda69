"""Box decomposition of a uniform single-level domain across the GPUs of one node.

One process per GPU, one box per process.  Replaces, for amr.max_level = 0 and equal boxes,
  - AmrLevel::FillPatch -> FabArray::FillBoundary   (Castro_advance.cpp:169; AMReX_FabArrayCommI.H:11-320):
        ghost exchange of NUM_STATE x 4 layers on 6 faces + 12 edges + 8 corners (the CTU transverse terms read
        corners), packed per peer by the CUDA pack kernels and moved as ONE NCCL send/recv group;
  - ParallelAllReduce::Min(dt) and ReduceIntMax(status)  (Castro.cpp:1640, 3387): one small all-reduce.
The per-zone arithmetic never depends on the decomposition, so an N-box run is bitwise identical to the 1-box run.

The time loop mirrors ctu_sim_advance (castro_b200/csrc/ctu_sim.cu), i.e. Amr::coarseTimeStep -> Castro::advance
with the dt control of Castro::computeInitialDt / computeNewDt (Castro.cpp:1559-1573, 1751-1992).

The box engine is pluggable so that the host logic (neighbour table, regions, message order, dt control) can be
exercised on CPU with world_size 2 over gloo (tests/test_dist_gloo.py drives it with the CPU checker as engine);
the product engine is CudaBoxEngine and nothing in this module imports the oracle.
"""
import itertools
import sys

import numpy as np

NG = 4

# decompositions of SURVEY §8(d).3: (1,1,1),(2,1,1),(2,2,1),(2,2,2)
_DIMS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def proc_grid(n):
    if n in _DIMS:
        return _DIMS[n]
    # generic fallback: peel factors of 2, then put the rest on x
    dims = [1, 1, 1]
    d = 0
    m = n
    while m % 2 == 0:
        dims[d % 3] *= 2
        m //= 2
        d += 1
    dims[0] *= m
    return tuple(dims)


def rank_coords(rank, dims):
    px, py, pz = dims
    return (rank % px, (rank // px) % py, rank // (px * py))


def coords_rank(c, dims):
    return c[0] + dims[0] * (c[1] + dims[1] * c[2])


def box_of(coords, dims, n_cell):
    """equal split of the domain [0, n_cell) in each dimension"""
    lo, hi = [], []
    for d in range(3):
        assert n_cell[d] % dims[d] == 0, "n_cell must divide evenly over the process grid"
        n = n_cell[d] // dims[d]
        lo.append(coords[d] * n)
        hi.append(coords[d] * n + n - 1)
    return tuple(lo), tuple(hi)


DIRECTIONS = [d for d in itertools.product((-1, 0, 1), repeat=3) if d != (0, 0, 0)]


def halo_plan(rank, dims, n_cell, periodic, ng=NG):
    """-> list of (direction, peer, send_lo, send_hi, recv_lo, recv_hi), in the index space of this rank's box.

    For direction d this rank RECEIVES its ghost region on side d from the box at coords+d (that box sends the valid
    slab on its side -d) and SENDS its own valid slab on side d to that same peer.  Directions leaving a non-periodic
    domain have no peer (physical BC fill); directions whose peer is this rank itself (periodic, one box across) are
    left to the local periodic wrap of ctu_fill_boundary."""
    c = rank_coords(rank, dims)
    lo, hi = box_of(c, dims, n_cell)
    plan = []
    for d in DIRECTIONS:
        pc = []
        ok = True
        for a in range(3):
            x = c[a] + d[a]
            if x < 0 or x >= dims[a]:
                if not periodic[a]:
                    ok = False
                    break
                x %= dims[a]
            pc.append(x)
        if not ok:
            continue
        peer = coords_rank(pc, dims)
        if peer == rank:
            continue
        slo, shi, rlo, rhi = [], [], [], []
        for a in range(3):
            if d[a] == 0:
                slo.append(lo[a]); shi.append(hi[a]); rlo.append(lo[a]); rhi.append(hi[a])
            elif d[a] > 0:
                slo.append(hi[a] - ng + 1); shi.append(hi[a]); rlo.append(hi[a] + 1); rhi.append(hi[a] + ng)
            else:
                slo.append(lo[a]); shi.append(lo[a] + ng - 1); rlo.append(lo[a] - ng); rhi.append(lo[a] - 1)
        plan.append((d, peer, tuple(slo), tuple(shi), tuple(rlo), tuple(rhi)))
    return plan


# ---- dt control (host scalars; must reproduce the reference bit for bit) ---------------------------------------
def est_time_step(par, raw):
    """Castro::estTimeStep given the raw CFL minimum (Castro.cpp:1576-1668)"""
    if par.fixed_dt > 0.0:
        return par.fixed_dt
    estdt_hydro = par.max_dt / par.cfl
    estdt_hydro = min(estdt_hydro, raw) * par.cfl
    r = par.max_dt
    if estdt_hydro < r:
        r = estdt_hydro
    return r


def initial_dt(par, raw, time, stop_time):
    """computeInitialDt + initialTimeStep (Castro.cpp:1559-1573, 1947-1992)"""
    init_dt = par.initial_dt if par.initial_dt > 0.0 else par.init_shrink * est_time_step(par, raw)
    dt_0 = min(1.0e+100, init_dt)
    eps = 0.001 * dt_0
    if stop_time >= 0.0 and (time + dt_0) > (stop_time - eps):
        dt_0 = stop_time - time
    return dt_0


def new_dt(par, raw, dt_old, time, stop_time):
    """computeNewDt (Castro.cpp:1751-1945)"""
    dt_min = est_time_step(par, raw)
    if par.fixed_dt <= 0.0:
        dt_min = min(dt_min, par.change_max * dt_old)
    dt_0 = min(1.0e+100, dt_min)
    eps = sys.float_info.epsilon
    if stop_time >= 0.0 and (time + dt_0) >= (stop_time - eps):
        dt_0 = stop_time - time
    return dt_0


# ---- engines -------------------------------------------------------------------------------------------------------
class CudaBoxEngine:
    """One box on the current CUDA device: every operation is a C-ABI call into libcastro_ctu.so."""

    def __init__(self, par, geom, lo, hi, workspace_bytes=0):
        import torch
        from . import lib
        self.torch, self.lib = torch, lib
        self.par, self.geom, self.lo, self.hi = par, geom, tuple(lo), tuple(hi)
        self.sim = lib.Sim(par, geom, lo, hi, workspace_bytes=workspace_bytes)
        self.ncomp = par.num_state
        self.device = torch.device("cuda", torch.cuda.current_device())

    def init_problem(self, problem, prob):
        self.sim.init_problem(problem, prob)

    def new_buffer(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.device)

    def pack(self, lo, hi, buf):
        self.sim.pack(lo, hi, buf)

    def unpack(self, lo, hi, buf):
        self.sim.unpack(lo, hi, buf)

    def step_pre(self):
        self.sim.step_pre()

    def step_fill(self):
        self.sim.step_fill()

    def step_hydro(self, dt):
        rc, est, nbad = self.sim.step_hydro(dt)
        return rc, est

    def estdt(self):
        return self.sim.estdt()

    def reduce_buffer(self, vals):
        return self.torch.tensor(vals, dtype=self.torch.float64, device=self.device)

    def download(self):
        return self.sim.download()

    @property
    def launches(self):
        return self.sim.launches

    def close(self):
        self.sim.close()


class DistSim:
    """One box per rank + halo exchange + dt/status all-reduce.  `comm` is torch.distributed (already initialised)
    or None for a single box."""

    def __init__(self, par, geom, n_cell, engine_factory, comm=None, rank=0, world=1, dims=None):
        self.par, self.geom = par, geom
        self.n_cell = tuple(n_cell)
        self.comm, self.rank, self.world = comm, rank, world
        self.dims = dims or proc_grid(world)
        assert self.dims[0] * self.dims[1] * self.dims[2] == world
        self.coords = rank_coords(rank, self.dims)
        self.lo, self.hi = box_of(self.coords, self.dims, n_cell)
        self.engine = engine_factory(par, geom, self.lo, self.hi)
        periodic = tuple(int(geom.periodic[d]) for d in range(3))
        self.plan = halo_plan(rank, self.dims, n_cell, periodic)
        nc = par.num_state
        self.msgs = []
        for (d, peer, slo, shi, rlo, rhi) in self.plan:
            n = nc * int(np.prod([shi[a] - slo[a] + 1 for a in range(3)]))
            self.msgs.append(dict(d=d, peer=peer, slo=slo, shi=shi, rlo=rlo, rhi=rhi, n=n,
                                  sbuf=self.engine.new_buffer(n), rbuf=self.engine.new_buffer(n)))
        self.halo_bytes_per_step = 8 * sum(m["n"] for m in self.msgs)
        self.time, self.dt, self.step = 0.0, -1.0, 0
        self.exchange_launches = 0

    # -- halo exchange: pack -> one grouped send/recv -> unpack ---------------------------------------------------
    def exchange(self):
        if not self.msgs:
            return
        dist = self.comm
        for m in self.msgs:
            self.engine.pack(m["slo"], m["shi"], m["sbuf"])
        ops = []
        # sends in DIRECTIONS order; the matching receive on the peer is its direction -d, so receives are posted in
        # the order of -d: messages between one pair of ranks then match up in order (NCCL has no tags)
        for m in self.msgs:
            ops.append(dist.P2POp(dist.isend, m["sbuf"], m["peer"]))
        for r in sorted(self.msgs, key=lambda m: DIRECTIONS.index(tuple(-x for x in m["d"]))):
            ops.append(dist.P2POp(dist.irecv, r["rbuf"], r["peer"]))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for m in self.msgs:
            self.engine.unpack(m["rlo"], m["rhi"], m["rbuf"])
        self.exchange_launches += 2 * len(self.msgs)

    def _allreduce_min(self, vals):
        if self.comm is None or self.world == 1:
            return list(vals)
        t = self.engine.reduce_buffer(list(vals))
        self.comm.all_reduce(t, op=self.comm.ReduceOp.MIN)
        return [float(x) for x in t.tolist()]

    def init_problem(self, problem, prob):
        self.engine.init_problem(problem, prob)

    def advance(self, nsteps, stop_time=-1.0):
        """-> (status, dt history)."""
        par = self.par
        hist = []
        status = 0
        if self.dt <= 0.0:
            raw = self._allreduce_min([self.engine.estdt()])[0]
            self.dt = initial_dt(par, raw, self.time, stop_time)
        for _ in range(nsteps):
            if stop_time >= 0.0 and self.time >= stop_time:
                break
            self.engine.step_pre()
            self.exchange()
            self.engine.step_fill()
            dt_sub = (self.time + self.dt) - self.time
            rc, est = self.engine.step_hydro(dt_sub)
            # min over ranks of both dt estimates; the status travels as -status so that one MIN all-reduce serves
            red = self._allreduce_min([est[0], est[1], -float(rc)])
            rc = int(-red[2])
            if rc > 0:
                status |= rc
                break
            hist.append(self.dt)
            self.time += self.dt
            self.step += 1
            if par.check_dt_after_advance:
                if par.change_max * est_time_step(par, red[0]) < dt_sub:
                    status |= 16
                    break
            self.dt = new_dt(par, red[1], self.dt, self.time, stop_time)
        return status, hist

    def close(self):
        self.engine.close()
