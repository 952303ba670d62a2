"""The C-ABI library builds for sm_100a, loads without a GPU and exports every function include/dstd_b200.h declares
(no compute calls here), and the ctypes structs mirror the header's structs field by field."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from dstd_gcn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dstd_b200.h")


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dstd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in _lib.SYMBOLS"
    assert sorted(_lib.SYMBOLS) == names


def test_host_only_entry_points(lib):
    assert b"sm_100a" in lib.dstd_version()
    assert lib.dstd_kernel_launch_count() >= 0
    assert lib.dstd_gc_fwd_workspace_bytes(4, 64, 64, 35, 22, 2) > 0
    assert lib.dstd_gc_bwd_workspace_bytes(4, 64, 64, 35, 22, 2) > lib.dstd_gc_fwd_workspace_bytes(4, 64, 64, 35, 22, 2)
    assert lib.dstd_bn_act_workspace_bytes(4, 64, 35, 22) > 0
    # argument validation happens before any CUDA call, so it is testable without a device
    a = _lib.GcFwdArgs()
    assert lib.dstd_gc_forward(ctypes.byref(a), None) == -1
    assert b"gc_forward" in lib.dstd_last_error()


def _header_struct_fields(name):
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", src).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"\[[^\]]*\]", "", decl)
        parts = decl.replace("*", " ").split(",")
        first = parts[0].split()
        fields.append(first[-1])
        fields += [p.strip() for p in parts[1:]]
    return fields


@pytest.mark.parametrize("cname,pycls", [("dstd_view", _lib.View), ("dstd_branch", _lib.Branch),
                                         ("dstd_branch_grad", _lib.BranchGrad), ("dstd_gc_fwd_args", _lib.GcFwdArgs),
                                         ("dstd_gc_bwd_args", _lib.GcBwdArgs), ("dstd_bn_act_fwd_args", _lib.BnFwdArgs),
                                         ("dstd_bn_act_bwd_args", _lib.BnBwdArgs),
                                         ("dstd_chmix_fwd_args", _lib.ChmixFwdArgs),
                                         ("dstd_chmix_bwd_args", _lib.ChmixBwdArgs)])
def test_ctypes_structs_mirror_header(cname, pycls):
    assert [f[0] for f in pycls._fields_] == _header_struct_fields(cname)


def test_missing_library_is_a_loud_error(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load_library(str(tmp_path / "nope.so"))
