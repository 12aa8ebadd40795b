"""CPU tests (-m "not gpu"): the C-ABI library loads without a GPU, exports every symbol
include/cm_b200.h declares, refuses to compute without CUDA (no CPU fallback), and the product
never touches oracle/."""
import ctypes
import os
import re

import pytest

import cm_b200 as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(cm.LIB_PATH)
    names = declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/cm_b200.h but not exported: {missing}"
    # and the Python binding covers the whole header
    assert sorted(cm.SYMBOLS) == names


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cm.CmError) as e:
        cm.init(0, 1, 0)
    assert "no CPU fallback" in str(e.value)
    assert cm.rank_me() == -1
    h = ctypes.c_void_p()
    rc = cm.lib().cm_create(ctypes.byref(h), cm.F64, 100, 100, 1, 1, 64, 64, 128)
    assert rc == 4 and not h.value  # CM_ERR_STATE: runtime not initialised


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.join(ROOT, "convergent-matrix_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle[/_]|cm_oracle|libcm_ref|cmref_|cmo_", text) and "nothing here imports" not in text:
                    offenders.append(os.path.join(base, f))
    for f in ("include/cm_b200.h", "include/convergent_matrix.hpp", "include/local_matrix.hpp"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p) and re.search(r"cm_oracle|libcm_ref|cmref_|cmo_", open(p).read()):
            offenders.append(p)
    assert not offenders, offenders


def test_library_has_sm100a_code_only():
    """the kernels are built for sm_100a and nothing else (no multi-arch dispatch)"""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", cm.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
