"""Minimal reader for single-level AMReX plotfiles (format restated from the reference's output, see
SURVEY.md §8c; writer is external/amrex/Src/Base/AMReX_PlotFileUtil.cpp + AMReX_VisMF.cpp).

Test infrastructure: used to turn the reference executable's plotfiles into golden fixtures and to
compare states.  Returns (header dict, ndarray[ncomp, nz, ny, nx]) assembled over all boxes."""
import os
import re
import numpy as np


def read_header(pf):
    with open(os.path.join(pf, "Header")) as f:
        lines = [l.rstrip("\n") for l in f]
    it = iter(lines)
    version = next(it)
    ncomp = int(next(it))
    names = [next(it) for _ in range(ncomp)]
    dim = int(next(it))
    time = float(next(it))
    finest = int(next(it))
    problo = [float(x) for x in next(it).split()]
    probhi = [float(x) for x in next(it).split()]
    next(it)  # ref ratios
    dline = next(it)
    if dim == 3:
        dom = re.findall(r"\((-?\d+),(-?\d+),(-?\d+)\)", dline)
        domlo = [int(x) for x in dom[0]]
        domhi = [int(x) for x in dom[1]]
    elif dim == 2:      # ((lox,loy) (hix,hiy) (0,0))
        dom = re.findall(r"\((-?\d+),(-?\d+)\)", dline)
        domlo = [int(dom[0][0]), int(dom[0][1]), 0]
        domhi = [int(dom[1][0]), int(dom[1][1]), 0]
        problo, probhi = problo + [0.0], probhi + [0.0]
    else:       # 1-D: ((lo) (hi) (0))
        dom = re.findall(r"\((-?\d+)\)", dline)
        domlo = [int(dom[0]), 0, 0]
        domhi = [int(dom[1]), 0, 0]
        problo, probhi = problo + [0.0, 0.0], probhi + [0.0, 0.0]
    steps = int(next(it).split()[0])
    dx = [float(x) for x in next(it).split()]
    return dict(version=version, ncomp=ncomp, names=names, dim=dim, time=time, finest=finest, problo=problo,
                probhi=probhi, domlo=domlo, domhi=domhi, step=steps, dx=dx)


def read_plotfile(pf):
    h = read_header(pf)
    assert h["finest"] == 0 and h["dim"] in (1, 2, 3)
    lev = os.path.join(pf, "Level_0")
    with open(os.path.join(lev, "Cell_H")) as f:
        txt = f.read()
    if h["dim"] == 3:
        boxes = re.findall(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \(0,0,0\)\)", txt)
    elif h["dim"] == 2:
        boxes = [(a, b, "0", c, d, "0") for a, b, c, d in re.findall(r"\(\((-?\d+),(-?\d+)\) \((-?\d+),(-?\d+)\) \(0,0\)\)", txt)]
    else:
        boxes = [(a, "0", "0", b, "0", "0") for a, b in re.findall(r"\(\((-?\d+)\) \((-?\d+)\) \(0\)\)", txt)]
    fods = re.findall(r"FabOnDisk: (\S+) (\d+)", txt)
    assert len(boxes) == len(fods)
    nx = [h["domhi"][d] - h["domlo"][d] + 1 for d in range(3)]
    out = np.full((h["ncomp"], nx[2], nx[1], nx[0]), np.nan)
    for b, (fn, off) in zip(boxes, fods):
        lo = [int(x) for x in b[:3]]
        hi = [int(x) for x in b[3:]]
        n = [hi[d] - lo[d] + 1 for d in range(3)]
        with open(os.path.join(lev, fn), "rb") as f:
            f.seek(int(off))
            hdr = f.readline()
            assert hdr.startswith(b"FAB")
            nc = int(hdr.strip().split()[-1])
            data = np.frombuffer(f.read(8 * nc * n[0] * n[1] * n[2]), dtype="<f8")
        data = data.reshape(nc, n[2], n[1], n[0])
        o = [lo[d] - h["domlo"][d] for d in range(3)]
        out[:, o[2]:o[2] + n[2], o[1]:o[1] + n[1], o[0]:o[0] + n[0]] = data
    assert not np.isnan(out).any()
    return h, out
