"""Loader + thin Python mirror of the C ABI (include/castro_ctu.h) of libcastro_ctu.so.

PyTorch is used only as plumbing (device memory, streams, torch.distributed); every number is computed by the
hand-written sm_100a kernels behind the C ABI.  There is no CPU or PyTorch fallback: if the shared library is
missing or no CUDA device is usable, the calls raise."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# CTU_LIB: developer A/B switch (tools/build_variant.sh); the product library is libcastro_ctu.so
SO_PATH = os.environ.get("CTU_LIB") or os.path.join(_HERE, "libcastro_ctu.so")
_LIB = None


class CtuError(RuntimeError):
    pass


_ERR = {-1: "bad argument", -2: "no CUDA device (there is no CPU fallback)", -3: "CUDA error", -4: "out of memory",
        -5: "unsupported parameter combination", -6: "I/O error"}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise CtuError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                           "`make -C castro_b200/csrc` (no fallback path exists)" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.ctu_version.restype = C.c_char_p
        L.ctu_abi_sizeof_params.restype = C.c_size_t
        L.ctu_abi_sizeof_geom.restype = C.c_size_t
        if L.ctu_abi_sizeof_params() != C.sizeof(abi.Params) or L.ctu_abi_sizeof_geom() != C.sizeof(abi.Geom):
            raise CtuError("castro_b200/abi.py and %s disagree on the layout of ctu_params_t / ctu_geom_t: rebuild the library" % SO_PATH)
        L.ctu_sim_last_subcycles.argtypes = [C.c_void_p]
        L.ctu_sim_retries.argtypes = [C.c_void_p]
        L.ctu_workspace_bytes.restype = C.c_size_t
        L.ctu_sim_workspace.restype = C.c_void_p
        L.ctu_workspace_create.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.ctu_workspace_destroy.argtypes = [C.c_void_p]
        L.ctu_workspace_bytes.argtypes = [C.c_void_p]
        L.ctu_workspace_last_slabs.argtypes = [C.c_void_p]
        L.ctu_workspace_last_tile_staged.argtypes = [C.c_void_p]
        L.ctu_workspace_launches.argtypes = [C.c_void_p]
        L.ctu_workspace_replays.argtypes = [C.c_void_p]
        L.ctu_workspace_replays.restype = C.c_long
        L.ctu_workspace_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.ctu_stage_name.restype = C.c_char_p
        L.ctu_stage_name.argtypes = [C.c_int]
        L.ctu_workspace_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
        L.ctu_sim_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3,
                                     C.c_size_t, C.c_void_p]
        for fn in ("ctu_sim_destroy", "ctu_sim_step_pre", "ctu_sim_step_fill"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ctu_sim_workspace.argtypes = [C.c_void_p]
        L.ctu_sim_launches.argtypes = [C.c_void_p]
        L.ctu_sim_state.argtypes = [C.c_void_p, C.POINTER(abi.Fab)]
        L.ctu_sim_download.argtypes = [C.c_void_p, C.c_void_p]
        L.ctu_sim_upload.argtypes = [C.c_void_p, C.c_void_p]
        L.ctu_sim_init_problem.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.ctu_sim_set_time.argtypes = [C.c_void_p, C.c_double]
        L.ctu_sim_write_plotfile.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int]
        L.ctu_sim_step_hydro.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_sim_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_sim_advance.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_hydro_advance.argtypes = [C.c_void_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3,
                                        C.POINTER(abi.Fab), C.POINTER(abi.Fab), C.POINTER(abi.Fab), C.POINTER(abi.Fab),
                                        C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.c_void_p]
        L.ctu_estdt.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.POINTER(abi.Fab),
                                C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p]
        L.ctu_clean_state.argtypes = [C.POINTER(abi.Params), abi.Int3, abi.Int3, C.POINTER(abi.Fab), C.c_int, C.c_void_p]
        L.ctu_fill_boundary.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3, C.POINTER(abi.Fab),
                                        C.c_int, C.c_void_p]
        L.ctu_halo_pack.argtypes = [C.POINTER(abi.Fab), abi.Int3, abi.Int3, C.c_void_p, C.c_void_p]
        L.ctu_halo_unpack.argtypes = [C.POINTER(abi.Fab), abi.Int3, abi.Int3, C.c_void_p, C.c_void_p]
        L.ctu_dist_unique_id.argtypes = [C.c_char_p]
        L.ctu_dist_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, C.c_int, C.c_int,
                                      C.c_char_p, C.c_size_t, C.c_void_p]
        L.ctu_dist_destroy.argtypes = [C.c_void_p]
        L.ctu_dist_sim.argtypes = [C.c_void_p]
        L.ctu_dist_sim.restype = C.c_void_p
        L.ctu_dist_box.argtypes = [C.c_void_p, abi.Int3, abi.Int3]
        L.ctu_dist_halo_bytes.argtypes = [C.c_void_p]
        L.ctu_dist_halo_bytes.restype = C.c_long
        L.ctu_dist_exchange.argtypes = [C.c_void_p]
        L.ctu_dist_allreduce_min.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.ctu_dist_advance.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_dist_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_dist_halo_plan.argtypes = [abi.Int3, C.c_int, abi.Int3, abi.Int3, abi.Int3, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ctu_sim_read_plotfile.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ctu_sim_diag.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.ctu_write_grid_diag.argtypes = [C.c_char_p, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.ctu_eos_helm_load.argtypes = [C.c_char_p]
        L.ctu_eos_helm_eval.argtypes = [C.POINTER(abi.Params), C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ctu_selftest_mathfast.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
        L.ctu_problem_init.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), abi.Int3, abi.Int3,
                                       C.POINTER(abi.Fab), C.c_void_p]
        L.ctu_parse_inputs.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(abi.Params), C.POINTER(abi.Geom), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc < 0:
        raise CtuError("%s failed: %s (%d)" % (what or "ctu call", _ERR.get(rc, "?"), rc))
    return rc


NSTAGE = 8


def stage_names():
    return [lib().ctu_stage_name(i).decode() for i in range(NSTAGE)]


def workspace_set_profiling(ws_handle, on):
    check(lib().ctu_workspace_set_profiling(ws_handle, 1 if on else 0), "ctu_workspace_set_profiling")


def workspace_stage_times(ws_handle, reset=False):
    """-> {kernel family: (total ms, launches)} accumulated since the last reset"""
    ms = (C.c_double * NSTAGE)()
    nl = (C.c_int * NSTAGE)()
    check(lib().ctu_workspace_stage_times(ws_handle, ms, nl, 1 if reset else 0), "ctu_workspace_stage_times")
    return {n: (ms[i], nl[i]) for i, n in enumerate(stage_names())}


def workspace_replays(ws_handle):
    """tiles recomputed with the IEEE division / sqrt operators since the workspace was created (diagnostic)"""
    return int(lib().ctu_workspace_replays(ws_handle))


def default_params():
    p = abi.Params()
    lib().ctu_params_defaults(C.byref(p))
    return p


def params_finalize(p):
    check(lib().ctu_params_finalize(C.byref(p)), "ctu_params_finalize")
    return p


def parse_inputs(path, overrides=""):
    p = default_params()
    g = abi.Geom()
    nc = (C.c_int * 3)()
    ms = C.c_int(-1)
    stop = C.c_double(-1.0)
    buf = C.create_string_buffer(4096)
    check(lib().ctu_parse_inputs(path.encode() if path else None, overrides.encode(), C.byref(p), C.byref(g), nc, C.byref(ms),
                                 C.byref(stop), buf, 4096), "ctu_parse_inputs")
    return p, g, tuple(nc), ms.value, stop.value, buf.value.decode()


def tensor_fab(t, lo):
    """torch CUDA float64 tensor [ncomp, nz, ny, nx] -> Fab view (Array4 layout)."""
    assert t.is_cuda and t.is_contiguous() and str(t.dtype) == "torch.float64"
    nc, nz, ny, nx = t.shape
    return abi.Fab.make(t.data_ptr(), lo, (lo[0] + nx - 1, lo[1] + ny - 1, lo[2] + nz - 1), nc)


def prob_string(d):
    # arrays (problem.center = 0.5 0.5 0.0) are written v1,v2,v3
    return " ".join("%s=%s" % (k, ",".join(repr(x) for x in v) if isinstance(v, (tuple, list)) else repr(v)) for k, v in d.items())


class Workspace:
    def __init__(self, max_bytes=0):
        self.h = C.c_void_p()
        check(lib().ctu_workspace_create(C.byref(self.h), max_bytes), "ctu_workspace_create")

    def close(self):
        if self.h:
            lib().ctu_workspace_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def bytes(self):
        return lib().ctu_workspace_bytes(self.h)

    @property
    def last_slabs(self):
        return lib().ctu_workspace_last_slabs(self.h)

    @property
    def tile_staged(self):
        """True when the last ctu_hydro_advance on this workspace ran the tile-staged (TMA) kernels"""
        return bool(lib().ctu_workspace_last_tile_staged(self.h))

    @property
    def launches(self):
        return lib().ctu_workspace_launches(self.h)


def hydro_advance(ws, par, geom, vlo, vhi, Sborder, sb_lo, S_new, sn_lo, dt, n_clean=1, from_sborder=False, fluxes=None,
                  flux_los=None, mass_fluxes=None, mass_los=None, want_loc=False, stream=0, old_src=None, src_corr=None, src_lo=None):
    """Sborder, S_new, fluxes: torch CUDA tensors [ncomp, nz, ny, nx].  old_src / src_corr: optional NSRC = 7 component
    tensors on grow(valid, 3) with low corner src_lo.  Returns (status, dt_est[2], loc, n_bad)."""
    fos = tensor_fab(old_src, src_lo) if old_src is not None else None
    fsc = tensor_fab(src_corr, src_lo) if src_corr is not None else None
    fb = tensor_fab(Sborder, sb_lo)
    fn = tensor_fab(S_new, sn_lo)
    est = (C.c_double * 2)()
    loc = (C.c_int * 3)()
    nbad = (C.c_int * 2)()
    FabP3 = C.POINTER(abi.Fab) * 3
    keep = []

    def mk(ts, los):
        if ts is None:
            return None
        arr = FabP3()
        for d in range(3):
            f = tensor_fab(ts[d], los[d])
            keep.append(f)
            arr[d] = C.pointer(f)
        return C.cast(arr, C.c_void_p)

    flags = (n_clean & 3) | (4 if from_sborder else 0)
    rc = lib().ctu_hydro_advance(ws.h, C.byref(par), C.byref(geom), abi.Int3(*vlo), abi.Int3(*vhi), C.byref(fb),
                                 C.byref(fos) if fos is not None else None, C.byref(fsc) if fsc is not None else None,
                                 C.byref(fn), mk(fluxes, flux_los), mk(mass_fluxes, mass_los), dt, flags, est,
                                 loc if want_loc else None, nbad, stream)
    check(rc, "ctu_hydro_advance")
    return rc, (est[0], est[1]), tuple(loc), tuple(nbad)


def estdt(par, geom, lo, hi, S, s_lo, stream=0):
    f = tensor_fab(S, s_lo)
    dt = C.c_double(0)
    loc = (C.c_int * 3)()
    check(lib().ctu_estdt(C.byref(par), C.byref(geom), abi.Int3(*lo), abi.Int3(*hi), C.byref(f), C.byref(dt), loc, stream), "ctu_estdt")
    return dt.value, tuple(loc)


def clean_state(par, lo, hi, S, s_lo, npass=1, stream=0):
    f = tensor_fab(S, s_lo)
    check(lib().ctu_clean_state(C.byref(par), abi.Int3(*lo), abi.Int3(*hi), C.byref(f), npass, stream), "ctu_clean_state")


def fill_boundary(par, geom, vlo, vhi, S, ng=abi.CTU_NGROW, stream=0):
    f = tensor_fab(S, tuple(l - ng for l in vlo))
    check(lib().ctu_fill_boundary(C.byref(par), C.byref(geom), abi.Int3(*vlo), abi.Int3(*vhi), C.byref(f), ng, stream),
          "ctu_fill_boundary")


def halo_pack(S, s_lo, lo, hi, buf, stream=0):
    f = tensor_fab(S, s_lo)
    check(lib().ctu_halo_pack(C.byref(f), abi.Int3(*lo), abi.Int3(*hi), buf.data_ptr(), stream), "ctu_halo_pack")


def halo_unpack(S, s_lo, lo, hi, buf, stream=0):
    f = tensor_fab(S, s_lo)
    check(lib().ctu_halo_unpack(C.byref(f), abi.Int3(*lo), abi.Int3(*hi), buf.data_ptr(), stream), "ctu_halo_unpack")


def problem_init(problem, prob, par, geom, lo, hi, S, s_lo, stream=0):
    f = tensor_fab(S, s_lo)
    check(lib().ctu_problem_init(problem.encode(), prob_string(prob).encode(), C.byref(par), C.byref(geom), abi.Int3(*lo),
                                 abi.Int3(*hi), C.byref(f), stream), "ctu_problem_init")


def dist_halo_plan(dims, rank, n_cell, periodic):
    """the C++ neighbour table of one rank (ctu_dist_halo_plan; host logic only) -> (messages, receive order);
    message = (direction, peer, send_lo, send_hi, recv_lo, recv_hi) like castro_b200.dist.halo_plan"""
    out = (C.c_int * (16 * 26))()
    ro = (C.c_int * 26)()
    n = check(lib().ctu_dist_halo_plan(abi.Int3(*dims), rank, abi.Int3(0, 0, 0), abi.Int3(*[x - 1 for x in n_cell]), abi.Int3(*periodic),
                                      26, out, ro), "ctu_dist_halo_plan")
    msgs = []
    for m in range(n):
        o = out[16 * m:16 * m + 16]
        msgs.append((tuple(o[0:3]), o[3], tuple(o[4:7]), tuple(o[7:10]), tuple(o[10:13]), tuple(o[13:16])))
    return msgs, list(ro[:n])


class Sim:
    """Single-box driver (one per rank): owns the ghosted ping-pong state buffers on the device."""

    def __init__(self, par, geom, lo, hi, workspace_bytes=0, stream=0, _borrowed=None):
        self.par, self.geom = par, geom
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:
            self.h = C.c_void_p(_borrowed)          # the box of a Dist (owned by it)
        else:
            check(lib().ctu_sim_create(C.byref(self.h), C.byref(par), C.byref(geom), abi.Int3(*lo), abi.Int3(*hi), workspace_bytes,
                                       stream), "ctu_sim_create")
        self.shape = (par.num_state,) + tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0))

    def close(self):
        if self.h and self._owned:
            lib().ctu_sim_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_problem(self, problem, prob):
        check(lib().ctu_sim_init_problem(self.h, problem.encode(), prob_string(prob).encode()), "ctu_sim_init_problem")

    def state_fab(self):
        f = abi.Fab()
        check(lib().ctu_sim_state(self.h, C.byref(f)), "ctu_sim_state")
        return f

    def download(self, out=None):
        """valid zones -> host ndarray [ncomp, nz, ny, nx]"""
        if out is None:
            out = np.empty(self.shape)
        check(lib().ctu_sim_download(self.h, out.ctypes.data), "ctu_sim_download")
        return out

    def write_plotfile(self, path, time, step):
        """single-level AMReX plotfile of the current state (state variables only), readable by the reference's fcompare"""
        check(lib().ctu_sim_write_plotfile(self.h, str(path).encode(), float(time), int(step)), "ctu_sim_write_plotfile")

    def read_plotfile(self, path):
        """single-level plotfile (the reference's or this library's) -> the sim's box; returns (time, step)"""
        t, st = C.c_double(0), C.c_int(0)
        check(lib().ctu_sim_read_plotfile(self.h, str(path).encode(), C.byref(t), C.byref(st)), "ctu_sim_read_plotfile")
        return t.value, st.value

    def diag(self, center, local_only=False):
        """columns 3-20 of the reference's grid_diag.out for the current state (Castro::sum_integrated_quantities)"""
        out = (C.c_double * 18)()
        check(lib().ctu_sim_diag(self.h, (C.c_double * 3)(*center), 1 if local_only else 0, out), "ctu_sim_diag")
        return list(out)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self.shape
        check(lib().ctu_sim_upload(self.h, arr.ctypes.data), "ctu_sim_upload")

    def pack(self, lo, hi, buf, stream=0):
        """region [lo,hi] x all comps of the current state -> contiguous torch CUDA buffer"""
        f = self.state_fab()
        check(lib().ctu_halo_pack(C.byref(f), abi.Int3(*lo), abi.Int3(*hi), buf.data_ptr(), stream), "ctu_halo_pack")

    def unpack(self, lo, hi, buf, stream=0):
        f = self.state_fab()
        check(lib().ctu_halo_unpack(C.byref(f), abi.Int3(*lo), abi.Int3(*hi), buf.data_ptr(), stream), "ctu_halo_unpack")

    def estdt(self, stream=0):
        f = self.state_fab()
        dt = C.c_double(0)
        check(lib().ctu_estdt(C.byref(self.par), C.byref(self.geom), abi.Int3(*self.lo), abi.Int3(*self.hi), C.byref(f),
                              C.byref(dt), None, stream), "ctu_estdt")
        return dt.value

    def step_pre(self):
        check(lib().ctu_sim_step_pre(self.h), "ctu_sim_step_pre")

    def step_fill(self):
        check(lib().ctu_sim_step_fill(self.h), "ctu_sim_step_fill")

    def step_hydro(self, dt):
        est = (C.c_double * 2)()
        nbad = (C.c_int * 2)()
        rc = check(lib().ctu_sim_step_hydro(self.h, dt, est, nbad), "ctu_sim_step_hydro")
        return rc, (est[0], est[1]), tuple(nbad)

    def step_host(self, host_in, host_out, dt):
        """one step with the state in host memory (numpy [ncomp, nz, ny, nx], ideally pinned), streamed in z-chunks"""
        assert host_in.shape == self.shape and host_out.shape == self.shape
        assert host_in.flags["C_CONTIGUOUS"] and host_out.flags["C_CONTIGUOUS"]
        est = (C.c_double * 2)()
        nbad = (C.c_int * 2)()
        rc = check(lib().ctu_sim_step_host(self.h, host_in.ctypes.data, host_out.ctypes.data, dt, est, nbad), "ctu_sim_step_host")
        return rc, (est[0], est[1]), tuple(nbad)

    def advance(self, nsteps, stop_time=-1.0, time=0.0, dt=-1.0):
        t = C.c_double(time)
        d = C.c_double(dt)
        hist = (C.c_double * max(nsteps, 1))()
        done = C.c_int(0)
        rc = check(lib().ctu_sim_advance(self.h, nsteps, stop_time, C.byref(t), C.byref(d), hist, C.byref(done)), "ctu_sim_advance")
        return rc, t.value, d.value, np.array(hist[:done.value])

    @property
    def ws_handle(self):
        return C.c_void_p(lib().ctu_sim_workspace(self.h))

    @property
    def launches(self):
        return lib().ctu_sim_launches(self.h)

    @property
    def last_slabs(self):
        return lib().ctu_workspace_last_slabs(lib().ctu_sim_workspace(self.h))

    @property
    def tile_staged(self):
        """True if the last advance ran the tile-staged (TMA) kernel family"""
        return bool(lib().ctu_workspace_last_tile_staged(lib().ctu_sim_workspace(self.h)))

    @property
    def last_subcycles(self):
        return lib().ctu_sim_last_subcycles(self.h)

    @property
    def retries(self):
        return lib().ctu_sim_retries(self.h)


class Dist:
    """Multi-box driver in C++ over NCCL (ctu_dist_*): one process per GPU, one equal box per process.  `comm` is an initialised
    torch.distributed (used once, as plumbing, to carry the ncclUniqueId from rank 0 to the other ranks) or None for one box."""

    def __init__(self, par, geom, dims, rank=0, world=1, comm=None, workspace_bytes=0, stream=0):
        self.par, self.geom, self.dims, self.rank, self.world = par, geom, tuple(dims), rank, world
        idbuf = C.create_string_buffer(128)
        if world > 1:
            import torch
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                check(lib().ctu_dist_unique_id(idbuf), "ctu_dist_unique_id")
                t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
            comm.broadcast(t, src=0)
            idbuf = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        self.h = C.c_void_p()
        check(lib().ctu_dist_create(C.byref(self.h), C.byref(par), C.byref(geom), abi.Int3(*dims), rank, world, idbuf, workspace_bytes,
                                    stream), "ctu_dist_create")
        lo, hi = abi.Int3(), abi.Int3()
        check(lib().ctu_dist_box(self.h, lo, hi), "ctu_dist_box")
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.sim = Sim(par, geom, self.lo, self.hi, _borrowed=lib().ctu_dist_sim(self.h))
        self.time, self.dt = 0.0, -1.0

    def close(self):
        if self.h:
            self.sim.close()
            lib().ctu_dist_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_problem(self, problem, prob):
        self.sim.init_problem(problem, prob)

    @property
    def halo_bytes_per_step(self):
        return int(lib().ctu_dist_halo_bytes(self.h))

    def advance(self, nsteps, stop_time=-1.0):
        """-> (status, dt history); time and dt are kept on the object"""
        t, d = C.c_double(self.time), C.c_double(self.dt)
        hist = (C.c_double * max(nsteps, 1))()
        done = C.c_int(0)
        rc = check(lib().ctu_dist_advance(self.h, nsteps, stop_time, C.byref(t), C.byref(d), hist, C.byref(done)), "ctu_dist_advance")
        self.time, self.dt = t.value, d.value
        return rc, list(hist[:done.value])

    def estdt_global(self):
        v = (C.c_double * 1)(self.sim.estdt())
        check(lib().ctu_dist_allreduce_min(self.h, v, 1), "ctu_dist_allreduce_min")
        return v[0]

    def step_host(self, host_in, host_out, dt):
        """one step of this rank's box with the state in host memory (streamed z-chunks, halo exchange inside the pipeline)"""
        assert host_in.shape == self.sim.shape and host_out.shape == self.sim.shape
        est = (C.c_double * 2)()
        st = C.c_int(0)
        check(lib().ctu_dist_step_host(self.h, host_in.ctypes.data, host_out.ctypes.data, dt, est, C.byref(st)), "ctu_dist_step_host")
        return st.value, (est[0], est[1])


def helm_load(path):
    """Helmholtz EOS table (helm_table.dat of Microphysics); must precede params_finalize with eos_kind = 1"""
    check(lib().ctu_eos_helm_load(str(path).encode()), "ctu_eos_helm_load")


def helm_eval(par, input, abar, ye, states, where=0):
    """states: ndarray [n, 4] = rho, T, e, p -> ndarray [n, 8] = rho, T, e, p, cs, gam1, dpde, dpdr_e (host or device copy of the EOS)"""
    a = np.ascontiguousarray(states, dtype=np.float64)
    out = np.empty((a.shape[0], 8))
    check(lib().ctu_eos_helm_eval(C.byref(par), input, abar, ye, a.shape[0], a.ctypes.data, out.ctypes.data, where), "ctu_eos_helm_eval")
    return out
