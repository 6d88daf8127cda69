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
