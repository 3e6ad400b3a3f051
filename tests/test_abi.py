"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/waiwera_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from waiwera_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "waiwera_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(L):
    names = declared_symbols()
    assert len(names) >= 40
    for nm in names:
        assert hasattr(L, nm), "include/waiwera_b200.h declares %s but the library does not export it" % nm


def test_binding_covers_header():
    from waiwera_b200 import _lib
    assert set(declared_symbols()) <= set(_lib.SIGNATURES), set(declared_symbols()) - set(_lib.SIGNATURES)


def test_struct_layouts_match_header(L):
    """ctypes mirrors of the POD structs have the sizes the C compiler gives the header's structs"""
    import subprocess
    import tempfile
    from waiwera_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "waiwera_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(wb_relperm), sizeof(wb_cappress), sizeof(wb_params),
  sizeof(wb_ksp_opts), sizeof(wb_newton_opts), sizeof(wb_newton_result)); return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(_lib.Relperm), C.sizeof(_lib.Cappress), C.sizeof(_lib.Params), C.sizeof(_lib.KspOpts),
                     C.sizeof(_lib.NewtonOpts), C.sizeof(_lib.NewtonResult)]


def test_no_cpu_fallback(L):
    """without a GPU the product fails loudly instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from waiwera_b200 import flow
    h = C.c_void_p()
    rc = L.wb_create(C.byref(flow.make_params()), 0, C.byref(h))
    assert rc < 0
    assert b"no CPU fallback" in L.wb_last_error()


def test_product_does_not_use_oracle():
    """nothing under waiwera_b200/ or include/ imports, links or calls the test oracle"""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle\.h|libwaiwera_oracle|wo_[a-z_]+\s*\()")
    for base in ("waiwera_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, fn)).read()
                    assert not pat.search(txt), os.path.join(dp, fn)


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "waiwera_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(wb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_fortran_interface_matches_header():
    """fortran/waiwera_b200.F90 (not compilable here: no Fortran compiler) binds every exported function with
    the header's argument count, and its constants equal the header's #defines"""
    f90 = open(os.path.join(ROOT, "fortran", "waiwera_b200.F90")).read()
    f90 = re.sub(r"&\s*\n\s*", " ", f90)
    bound = {}
    for m in re.finditer(r"function\s+(wb_[a-z0-9_]+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(wb_[a-z0-9_]+)\"\)", f90):
        assert m.group(1) == m.group(3)
        args = m.group(2).strip()
        bound[m.group(1)] = 0 if not args else len(args.split(","))
    protos = _header_prototypes()
    assert set(protos) == set(bound), set(protos) ^ set(bound)
    for nm, n in protos.items():
        assert bound[nm] == n, (nm, bound[nm], n)
    hdr = open(os.path.join(ROOT, "include", "waiwera_b200.h")).read()
    defines = dict(re.findall(r"#define\s+(WB_[A-Z0-9_]+)\s+(-?\d+)\b", hdr))
    consts = dict(re.findall(r"\b(WB_[A-Z0-9_]+)\s*=\s*(-?\d+)", f90))
    assert len(consts) >= 20
    for k, v in consts.items():
        assert defines.get(k) == v, (k, v, defines.get(k))
