"""CPU tests of the C-ABI library: it loads, exports every symbol
include/gorender_b200.h declares, and refuses to run without a CUDA device
(no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gorender_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gorender_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gorender_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert lib.grb_abi_version() == _cabi.GRB_ABI_VERSION == 3


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.grb_object) == 4 + 64 + 64
    assert C.sizeof(_cabi.grb_triangle_rec) == 64
    assert C.sizeof(_cabi.grb_frame_stats) == 24
    assert C.sizeof(_cabi.grb_draw_params) == 64 + 12 + 4 + 8 + 12 + 8 + 4
    assert _cabi.grb_mesh_desc.bbox.offset == 72 and C.sizeof(_cabi.grb_mesh_desc) == 72 + 128


def test_no_cpu_fallback():
    """Without a device the library fails loudly instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _cabi.load()
    h = C.c_void_p()
    rc = lib.grb_context_create(0, C.byref(h))
    assert rc == 2 and not h.value
    assert b"no CUDA device" in lib.grb_last_error(None)
    import gorender_b200 as g

    with pytest.raises(_cabi.GorenderError):
        g.Device(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "gorender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_binding" not in text and "liboracle" not in text and "gorender_oracle" not in text, f


def test_graft_entry_build():
    """The driver's "does it build" check: compiles (or finds up to date) every native piece and loads the library."""
    import __graft_entry__ as ge

    ge.build()


def test_option_bits_match_header():
    """The GRB_OPT_* / GRB_TEX_* / GRB_BOX_* values of the ctypes binding and of the oracle's header are the header's."""
    src = open(os.path.join(ROOT, "include", "gorender_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    bits = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"\b(GRB_OPT_[A-Z_]+)\s*=\s*1u\s*<<\s*(\d+)", src)}
    assert len(bits) == 11
    for name, value in bits.items():
        assert getattr(_cabi, name) == value, name
    orc = open(os.path.join(ROOT, "oracle", "gorender_oracle.h")).read()
    orc = re.sub(r"/\*.*?\*/", "", orc, flags=re.S)
    obits = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"\bORC_OPT_([A-Z_]+)\s*=\s*1u\s*<<\s*(\d+)", orc)}
    assert {("GRB_OPT_" + k): v for k, v in obits.items()} == bits     # tests feed Renderer.options() to both sides
    assert int(re.search(r"#define GRB_ABI_VERSION (\d+)", src).group(1)) == _cabi.GRB_ABI_VERSION
    assert int(re.search(r"#define GRB_TILE (\d+)", src).group(1)) == _cabi.GRB_TILE


def test_go_shim_uses_only_what_the_header_declares():
    """The cgo shim cannot be compiled here (no Go toolchain), so at least every C identifier it touches must exist in the
    header it is written against, with the number of arguments the header declares, and its braces must balance."""
    hdr = open(os.path.join(ROOT, "include", "gorender_b200.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(grb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr_nc, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    consts = set(re.findall(r"\b(GRB_[A-Z0-9_]+)\b", hdr_nc))
    types = set(re.findall(r"\btypedef struct (grb_[a-z_]+)", hdr_nc)) | set(re.findall(r"}\s*(grb_[a-z_]+)\s*;", hdr_nc))
    go_dir = os.path.join(ROOT, "go")
    seen_calls = 0
    for fn in sorted(os.listdir(go_dir)):
        if not fn.endswith(".go"):
            continue
        src = open(os.path.join(go_dir, fn)).read()
        code = re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))
        assert code.count("{") == code.count("}"), fn
        assert code.count("(") == code.count(")"), fn
        for name in set(re.findall(r"\bC\.(GRB_[A-Z0-9_]+)\b", code)):
            assert name in consts, f"{fn}: C.{name} is not in the header"
        for name in set(re.findall(r"\bC\.(grb_[a-z0-9_]+)\b", code)):
            assert name in decls or name in types, f"{fn}: C.{name} is not in the header"
        # calls: count top-level commas between the parentheses
        for m in re.finditer(r"\bC\.(grb_[a-z0-9_]+)\(", code):
            name = m.group(1)
            if name not in decls:
                continue
            depth, i, commas, empty = 1, m.end(), 0, True
            while depth:
                ch = code[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    commas += 1
                if depth and not ch.isspace():
                    empty = False
                i += 1
            nargs = 0 if empty else commas + 1
            assert nargs == decls[name], f"{fn}: C.{name} called with {nargs} arguments, the header declares {decls[name]}"
            seen_calls += 1
    assert seen_calls >= 12
