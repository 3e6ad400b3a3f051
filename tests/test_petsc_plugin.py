"""Seam B1: petsc/wb_petsc_plugin.c is a real source file.  This image has no PETSc and no
MPI, so it is compile-checked against petsc/stub/ (declarations of exactly the PETSc / MPI names it uses, nothing
implemented), and the object's undefined symbols are checked: everything it needs is either a PETSc / MPI name declared
in the stub or an entry point the built libwaiwera_b200.so exports."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_petsc_plugin_compiles_and_binds_the_c_abi(tmp_path):
    src = os.path.join(ROOT, "petsc", "wb_petsc_plugin.c")
    obj = str(tmp_path / "wb_petsc_plugin.o")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fPIC", "-c", src, "-o", obj,
           "-I", os.path.join(ROOT, "petsc", "stub"), "-I", os.path.join(ROOT, "include")]
    subprocess.run(cmd, check=True)
    nm = subprocess.run(["nm", "-u", obj], check=True, capture_output=True, text=True).stdout
    undefined = set(re.findall(r"\bU\s+(\w+)", nm))
    wb = {s for s in undefined if s.startswith("wb_")}
    # the engine entry points the plug-in binds
    assert {"wb_create", "wb_comm_unique_id", "wb_comm_init", "wb_set_halo", "wb_set_global_offset", "wb_mat_create",
            "wb_mat_set_values", "wb_mat_destroy", "wb_pc_setup", "wb_pc_refactor", "wb_pc_apply", "wb_pc_destroy",
            "wb_ksp_solve", "wb_last_error"} <= wb
    from waiwera_b200 import build as wbuild
    lib = wbuild.build()
    exported = subprocess.run(["nm", "-D", "--defined-only", lib], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT\s+(\w+)", exported))
    assert wb <= exported, wb - exported
    stub = open(os.path.join(ROOT, "petsc", "stub", "petscksp.h")).read()
    for s in undefined - wb:
        assert s in ("memcpy", "memset", "_GLOBAL_OFFSET_TABLE_", "__stack_chk_fail") or re.search(r"\b%s\b" % s, stub), "undeclared external: " + s
    defined = subprocess.run(["nm", "--defined-only", obj], check=True, capture_output=True, text=True).stdout
    for s in ("PCCreate_WB", "KSPCreate_WB", "PetscDLLibraryRegister_wb_petsc_plugin"):
        assert re.search(r"\bT\s+%s\b" % s, defined), s
