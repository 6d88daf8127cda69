"""TEST INFRASTRUCTURE: a box engine for castro_b200.dist.DistSim backed by the CPU checker (oracle/liboracle.so), so the
host-side multi-box logic (neighbour table, message order, dt control) can be tested on CPU with world_size > 1 over gloo.
Never imported by the product package."""
import numpy as np
import torch

import oracle_lib as O

NG = 4


class OracleBoxEngine:
    def __init__(self, par, geom, lo, hi):
        self.par, self.geom, self.lo, self.hi = par, geom, tuple(lo), tuple(hi)
        ngd = (NG, NG, 0 if geom.dim == 2 else NG)          # a 2-D run has no ghost zones in z
        self.glo = tuple(l - g for l, g in zip(lo, ngd))
        self.ghi = tuple(h + g for h, g in zip(hi, ngd))
        shp = tuple(hi[d] - lo[d] + 1 + 2 * ngd[d] for d in (2, 1, 0))
        self.U = np.zeros((par.num_state,) + shp)
        self.launches = 0

    def _sl(self, lo, hi):
        return tuple(slice(lo[d] - self.glo[d], hi[d] - self.glo[d] + 1) for d in (2, 1, 0))

    def init_problem(self, problem, prob):
        if problem == "upload":          # prob["U0"]: the whole domain's initial state [ncomp, nz, ny, nx] (a reference plotfile)
            kz, jy, ix = self._sl(self.lo, self.hi)
            U0 = prob["U0"]
            self.U[:, kz, jy, ix] = U0[:, self.lo[2]:self.hi[2] + 1, self.lo[1]:self.hi[1] + 1, self.lo[0]:self.hi[0] + 1]
            return
        assert problem == "Sedov"
        n_cell = tuple(self.geom.domhi[d] - self.geom.domlo[d] + 1 for d in range(3))
        self.U = O.sedov_init(self.par, self.geom, n_cell, lo=self.lo, hi=self.hi, **prob)

    def new_buffer(self, n):
        return torch.empty(n, dtype=torch.float64)

    def pack(self, lo, hi, buf):
        kz, jy, ix = self._sl(lo, hi)
        buf.copy_(torch.from_numpy(np.ascontiguousarray(self.U[:, kz, jy, ix]).reshape(-1)))

    def unpack(self, lo, hi, buf):
        kz, jy, ix = self._sl(lo, hi)
        v = self.U[:, kz, jy, ix]
        v[...] = buf.numpy().reshape(v.shape)

    def step_pre(self):
        O.clean_state(self.par, self.U, self.glo, self.lo, self.hi)

    def step_fill(self):
        O.fill_boundary(self.par, self.geom, self.U, self.lo, self.hi, lo_arr=self.glo)
        O.clean_state(self.par, self.U, self.glo, self.glo, self.ghi)

    def step_hydro(self, dt):
        kz, jy, ix = self._sl(self.lo, self.hi)
        snew = np.ascontiguousarray(self.U[:, kz, jy, ix])
        st, est0, _, nbad = O.hydro_advance(self.par, self.geom, self.lo, self.hi, self.U, self.glo, snew, self.lo, dt, do_clean=1)
        if st:
            return st, (est0, est0)
        O.clean_state(self.par, snew, self.lo, self.lo, self.hi)
        est1, _ = O.estdt(self.par, self.geom, snew, self.lo, self.lo, self.hi)
        self.U[:, kz, jy, ix] = snew
        return 0, (est0, est1)

    def estdt(self):
        return O.estdt(self.par, self.geom, self.U, self.glo, self.lo, self.hi)[0]

    def reduce_buffer(self, vals):
        return torch.tensor(vals, dtype=torch.float64)

    def download(self):
        kz, jy, ix = self._sl(self.lo, self.hi)
        return np.ascontiguousarray(self.U[:, kz, jy, ix])

    def close(self):
        pass
