"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/fe_abi.h declares, its PODs have the reference's wire sizes, and it fails loudly (no CPU
fallback) when no CUDA device exists.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fe_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from front_end_b200 import lib as L
    lib = C.CDLL(L.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libfe_b200.so does not export %s" % n
    assert set(names) == set(L.EXPORTS), "ctypes table and header disagree"
    assert L.load().fe_abi_version() == 6


def test_python_constants_match_the_header_enums():
    """front_end_b200/lib.py restates the header's enum values by hand: they must not drift apart."""
    from front_end_b200 import lib as L
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fe_abi.h")).read(), flags=re.S)
    enums = dict((k, int(v)) for k, v in re.findall(r"\b(FE_[A-Z0-9_]+)\s*=\s*(-?\d+)", src))
    for name, value in enums.items():
        py = name[3:]                                   # FE_DESC_FREAK -> DESC_FREAK; error codes keep their FE_ prefix
        if hasattr(L, name):
            assert getattr(L, name) == value, name
        elif hasattr(L, py):
            assert getattr(L, py) == value, name
    for must in ("DESC_FREAK", "DESC_BRIEF64", "NORM_HAMMING2", "MASK_WINDOW", "MATCH_CROSSCHECK", "FE_ERR_UNSUPPORTED"):
        assert hasattr(L, must)
    assert enums["FE_DESC_FREAK"] == 6 and enums["FE_NORM_L2"] == 4


def test_wire_layouts_match_reference_messages():
    from front_end_b200 import lib as L
    # msg/kPoint.msg: 5 x float32 + 2 x int32 ; msg/cvMatch.msg: 3 x uint32 + float32
    assert L.KPOINT.itemsize == 28 and L.KPOINT.names == ("x", "y", "size", "angle", "response", "octave", "class_id")
    assert L.MATCH.itemsize == 16 and L.MATCH.names == ("queryIdx", "trainIdx", "imgIdx", "distance")
    assert C.sizeof(L.MatchCfg) == 48 and C.sizeof(L.Config) == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import front_end_b200 as fe
    with pytest.raises(fe.FeError) as e:
        fe.FrontEnd()
    assert e.value.code == fe.lib.FE_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "front_end_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "import cv2" not in txt, f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/fe_abi.h is a C header (C99, -pedantic clean) and a plain C program links against libfe_b200.so: the
    boundary carries no C++ or torch types.  Without a GPU fe_create must answer FE_ERR_NO_DEVICE (no CPU fallback)."""
    import subprocess
    import torch
    src = tmp_path / "c_abi.c"
    src.write_text('#include <stdio.h>\n#include "fe_abi.h"\n'
                   "int main(void) {\n"
                   "    fe_config cfg; fe_ctx *ctx = NULL; int st;\n"
                   "    fe_default_config(&cfg);\n"
                   "    if (sizeof(fe_kpoint) != 28 || sizeof(fe_match) != 16 || fe_abi_version() != FE_ABI_VERSION) return 10;\n"
                   "    if (cfg.nonmax != 1 || cfg.n_features != 5000 || cfg.edge_threshold != 31 || cfg.fast_type != FE_FAST_9_16) return 11;\n"
                   "    cfg.max_width = 64; cfg.max_height = 64;\n"
                   "    st = fe_create(&cfg, &ctx);\n"
                   '    printf("%d %s\\n", st, st == FE_OK ? "ok" : fe_last_error(NULL));\n'
                   "    if (ctx) fe_destroy(ctx);\n"
                   "    return st == FE_OK ? 0 : (st == FE_ERR_NO_DEVICE ? 3 : 4);\n}\n")
    exe = tmp_path / "c_abi"
    libdir = os.path.join(ROOT, "front_end_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lfe_b200", "-Wl,-rpath," + libdir])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == (0 if torch.cuda.is_available() else 3), (p.returncode, p.stdout, p.stderr)
