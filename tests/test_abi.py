"""CPU: the C-ABI library loads and exports exactly what include/dunk_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dunk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dunk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "dunk_knn_match_hamming" in syms and "dunk_ctx_create" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(dunk):
    lib = ctypes.CDLL(dunk._lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"libdunk_b200.so lacks: {missing}"


def test_binding_table_covers_header(dunk):
    assert sorted(dunk._lib.SIGNATURES) == header_symbols()


def test_struct_layouts(dunk):
    assert dunk.KEYPOINT_DTYPE.itemsize == 28   # cv::KeyPoint
    assert dunk.DMATCH_DTYPE.itemsize == 16     # cv::DMatch
    assert dunk.TOP2_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_gpu(dunk):
    """Without a CUDA device the product path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dunk.DunkError) as e:
        dunk.Context(0)
    assert e.value.code == dunk._lib.ERR_CUDA and "no CPU fallback" in e.value.message


def test_product_never_imports_oracle():
    """the product path may mention the oracle in comments, never import / include / link it"""
    pkg = os.path.join(ROOT, "cubesat-apds_b200")
    pat_py = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle|import\s+cv2|from\s+cv2)", re.M)
    pat_c = re.compile(r"#\s*include\s*[<\"][^>\"]*oracle", re.M)
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            path = os.path.join(dp, fn)
            if fn.endswith(".py"):
                assert not pat_py.search(open(path).read()), f"{fn} imports the oracle / cv2"
            elif fn.endswith((".cu", ".h", ".cuh")) or fn == "Makefile":
                txt = open(path).read()
                assert not pat_c.search(txt), f"{fn} includes oracle sources"
                if fn == "Makefile":
                    assert "oracle" not in txt
