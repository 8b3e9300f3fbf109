"""The C-ABI library loads on a CPU-only box and exports exactly the symbols include/genhancer_b200.h declares
(no compute calls here -- those are the `-m gpu` tests)."""
import ctypes
import os
import re

from conftest import ROOT

from genhancer_b200 import _lib

HEADER = os.path.join(ROOT, "include", "genhancer_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gh_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 25
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in the header but not exported by the library"
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)


def test_loader_declares_signatures_and_version():
    h = _lib.lib()
    assert h.gh_version() == 100
    assert h.gh_last_error() is not None


def test_struct_layouts_match_header_field_order():
    src = open(HEADER).read()
    for cname, cls in (("gh_gemm_args", _lib.GemmArgs), ("gh_rows_view", _lib.RowsView),
                       ("gh_attn_tensor", _lib.AttnTensor), ("gh_attn_out", _lib.AttnOut),
                       ("gh_conv_args", _lib.ConvArgs), ("gh_copy_desc", _lib.CopyDesc)):
        body = re.search(r"typedef struct \{([^{}]*)\} " + cname + ";", src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            decl = re.sub(r"^(const\s+)?[a-z0-9_]+\s*\*?\s*", "", stmt, count=1)
            fields += [f.strip().lstrip("*").strip() for f in decl.split(",")]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_product_path_has_no_cpu_fallback():
    import pytest
    import torch
    from genhancer_b200 import kernels as K
    with pytest.raises(_lib.GhError, match="no CPU fallback"):
        K.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
