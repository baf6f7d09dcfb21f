"""The drop-in boundary: libsblas_spmv.so loads, exports every symbol the headers in
include/ declare (C-ABI names, the reference's C names and its C++-mangled names), and
its compute entry points fail loudly -- never fall back -- when no GPU is present."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import sblas_b200 as sb
from conftest import ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sblas_|spMV_|get_)\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = sb.lib()
    names = _declared("sblas_spmv.h") + _declared("sblas_device.h") + _declared("sblas_synth.h") + \
        _declared("sblas_ingest.h") + _declared("spmv_kernel.h")
    assert len(names) > 40
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_reference_mangled_names_exported():
    """the unmodified harness links against C++-mangled names."""
    out = subprocess.run(["nm", "-D", "--defined-only", sb.LIB_PATH], capture_output=True, text=True).stdout
    for sym in ("_Z18spMV_mgpu_baselineiixPdS_PxPiS_S_S_i", "_Z12spMV_mgpu_v1iixPdS_PxPiS_S_S_ii",
                "_Z12spMV_mgpu_v2iixPdS_PxPiS_S_S_iixi", "_Z18get_row_from_indexiPxx", "_Z8get_timev",
                "_Z20get_gpu_availble_memi"):
        assert re.search(r"\bT %s\b" % re.escape(sym), out), sym


def test_no_oracle_or_cpu_fallback_in_product():
    """Nothing under s-blas_b200/ may reference the oracle."""
    for base, _, files in os.walk(os.path.join(ROOT, "s-blas_b200")):
        for f in files:
            if f.endswith((".c", ".cu", ".cpp", ".h", ".py")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in txt and "import oracle" not in txt and "spmv_oracle" not in txt, f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rp = np.array([0, 1, 2], np.int64)
    y = np.zeros(2)
    rc = sb.spMV_mgpu_v1(2, 2, 2, 1.0, np.ones(2), rp, np.array([0, 1], np.int32), np.ones(2), 0.0, y, 1, 1)
    assert rc != 0 and "no CPU fallback" in sb.last_error()
    assert (y == 0).all()
    assert sb.spMV_mgpu_v2(2, 2, 2, 1.0, np.ones(2), rp, np.array([0, 1], np.int32), np.ones(2), 0.0, y, 1, 1, 0, 1) == -1


def test_time_helper():
    t0 = sb.get_time()
    assert sb.get_time() >= t0 > 1.0e9


def test_every_public_header_stands_alone_in_c_and_cxx(tmp_path):
    """A maintainer includes one header at a time from C (gcc) or C++ (g++, as the reference's .cu files do): each
    header of include/ must compile on its own in both languages, with warnings as errors, and twice in a row
    (include guards).  sblas_device.h and spmm/spmv_kernel.h need the CUDA runtime types, so the toolkit's include
    directory is on the path like in the reference's Makefiles."""
    import glob
    import subprocess
    inc = os.path.join(ROOT, "include")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    for h in sorted(glob.glob(os.path.join(inc, "*.h"))):
        name = os.path.basename(h)
        for compiler, ext, std in (("gcc", "c", "-std=gnu11"), ("g++", "cpp", "-std=c++14")):
            src = str(tmp_path / ("inc_%s.%s" % (name.replace(".", "_"), ext)))
            open(src, "w").write('#include "%s"\n#include "%s"\nint main(void) { return 0; }\n' % (name, name))
            p = subprocess.run([compiler, std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, "-I", cuda_inc, src],
                               capture_output=True, text=True)
            assert p.returncode == 0, (name, compiler, p.stderr[:2000])
