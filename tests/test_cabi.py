"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/corleone_b200.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import corleone_b200 as cb
import problems as P
from corleone_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "corleone_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 12
    L = ctypes.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/corleone_b200.h but not exported"
    assert set(names) == set(native.EXPORTS)
    assert native.lib().cb200_version() == 201


def test_struct_layouts_match_header():
    assert ctypes.sizeof(native.Desc) == 14 * 4 + 17 * 8 + 2 * 4 + 2 * 8 + 2 * 8 + 8 + 2 * 4   # ... + saveat, n_saveat, reserved
    assert ctypes.sizeof(native.Out) == 12 * 8 + 8 + 8 + 8 + 8 + 8 + 8 + 3 * 8 + 2 * 4
    assert ctypes.sizeof(native.Sizes) == 9 * 8
    # the library reports sizeof and the offsets of the trailing fields of every struct a binding restates by hand
    # (julia/CorleoneB200.jl asserts the same numbers at load time)
    for which, T, fields in ((0, native.Desc, ("u0", "model_data_len", "observation_grid", "reltol", "saveat", "n_saveat")),
                             (1, native.Out, ("objective_row", "jac_vals", "status", "defects_gathered", "comm_b0", "resident")),
                             (2, native.Sizes, ("n_traj_sens",)), (3, native.CommInfo, ("epoch", "error"))):
        assert native.abi_layout(which) == [ctypes.sizeof(T)] + [getattr(T, f).offset for f in fields], which


def test_bad_descriptor_is_rejected_before_any_device_work():
    L = native.lib()
    d = native.Desc()
    h = ctypes.c_void_p()
    d.struct_size = 4
    assert L.cb200_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"struct_size" in L.cb200_last_error()
    d.struct_size = ctypes.sizeof(native.Desc)
    d.model = 99
    assert L.cb200_create(ctypes.byref(d), ctypes.byref(h)) == -3
    # consistent sizes but no tables: refused with a message, not a segfault
    d.model, d.nx, d.np_all, d.n_tunable_p, d.n_controls, d.n_intervals, d.steps_per_piece = 4, 1, 1, 1, 0, 1, 1
    assert L.cb200_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"NULL" in L.cb200_last_error()
    # adaptive mode needs positive tolerances; unknown method ids are refused
    d.method, d.abstol, d.reltol = 2, 0.0, 1e-3
    assert L.cb200_create(ctypes.byref(d), ctypes.byref(h)) == -1 and b"abstol" in L.cb200_last_error()
    d.method = 7
    assert L.cb200_create(ctypes.byref(d), ctypes.byref(h)) == -3


def test_no_cpu_fallback():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(1))
    ps, st = cb.setup(None, lay)
    with pytest.raises(native.NativeError, match="no CPU fallback"):
        lay(None, ps, st)
    assert native.probe_fp64_tflops(0, 10) < 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "corleone.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "configs.py", f"{f} mentions the oracle"
    # measurement helpers are not tests either: nothing under tools/ may import the checker
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", f)).read()
            assert "pyoracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f"tools/{f} uses the oracle"


def _build_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "cabi_lin1d")
    libdir = os.path.join(ROOT, "corleone.jl_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "cabi_lin1d.c"), "-o", exe, "-L", libdir,
                           "-lcorleone_b200", "-lm", f"-Wl,-rpath,{libdir}"])
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_plain_c_client_compiles_and_fails_loudly_without_gpu(tmp_path):
    """the header is plain C99; tests/c/cabi_lin1d.c links against the library alone.  Without a device the
    program must stop at cb200_create with the no-fallback message (exit code 2), never compute on the CPU."""
    res = _build_c_client(tmp_path)
    assert res.returncode in (0, 2), res.stdout + res.stderr
    if res.returncode == 2:
        assert "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_plain_c_client(tmp_path):
    """tests/c/cabi_lin1d.c: a C program with nothing but the header and the shared library (what a ccall sees)"""
    res = _build_c_client(tmp_path)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "cabi_lin1d ok" in res.stdout


def test_julia_binding_restates_the_header_structs_field_for_field():
    """julia/CorleoneB200.jl cannot run here (no Julia in the image): at least its hand-written structs must list the
    header's fields in the header's order with matching widths (it also asserts cb200_abi_layout at run time)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "corleone_b200.h")).read()
    jl = open(os.path.join(ROOT, "julia", "CorleoneB200.jl")).read()

    def c_fields(name):
        end = re.search(r"\}\s*" + name + r"\s*;", hdr).start()
        body = hdr[hdr.rfind("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ptr = "*" in decl
            typ = decl.replace("const", "").replace("*", " ").split()[0]
            for nm in decl.replace("*", " ").split(typ, 1)[1].split(","):
                out.append((nm.strip(), "ptr" if ptr else typ))
        return out

    def jl_fields(name):
        body = re.search(r"^struct " + name + r"\b.*?\n(.*?)^end", jl, re.S | re.M).group(1)
        out = []
        for line in body.splitlines():
            m = re.match(r"\s*(\w+)::([\w{}]+)", line)
            if m:
                out.append((m.group(1), m.group(2)))
        return out

    width = {"int32_t": "Int32", "uint32_t": "UInt32", "int64_t": "Int64", "double": "Float64"}
    for cname, jname in (("cb200_desc", "Desc"), ("cb200_out", "Out"), ("cb200_sizes", "Sizes"), ("cb200_comm_info", "CommInfo")):
        c, j = c_fields(cname), jl_fields(jname)
        assert [n for n, _ in c] == [n for n, _ in j], (cname, [n for n, _ in c], [n for n, _ in j])
        for (n, ct), (_, jt) in zip(c, j):
            assert jt.startswith("Ptr{") if ct == "ptr" else jt == width[ct], (cname, n, ct, jt)


def test_new_entry_points_reject_bad_arguments_without_a_device():
    """version 200 / 201 entry points (data plane, resident results, sample times): NULL handles and buffers are refused with
    CB200_ERR_ARG and a message -- no device work, no crash"""
    L = native.lib()
    L.cb200_comm_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.cb200_comm_unique_id.argtypes = [ctypes.c_void_p]
    L.cb200_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
    L.cb200_comm_gathered.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.cb200_resident.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
    L.cb200_sample_times.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    info = native.CommInfo()
    assert L.cb200_comm_query(None, ctypes.byref(info)) == -1 and b"null" in L.cb200_last_error()
    assert L.cb200_comm_unique_id(None) == -1
    buf = ctypes.create_string_buffer(native.COMM_ID_BYTES)
    assert L.cb200_comm_init(None, 2, 0, buf, 0) == -1
    p = ctypes.c_void_p()
    assert L.cb200_comm_gathered(None, ctypes.byref(p), None) == -1
    assert L.cb200_resident(None, native.WANT_SENS, ctypes.byref(p), None) == -1
    assert L.cb200_sample_times(None, None) == -1
    assert L.cb200_comm_destroy(None) == 0                     # destroying nothing is fine
    assert native.abi_layout(7) == []                          # unknown struct id
    assert native.lib().cb200_version() == 201
