"""CPU: the C-ABI library loads without a GPU and exports every function include/fastllama_b200.h declares; without a
CUDA device the entry points fail loudly (FL_ERR_CUDA), they never fall back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fastllama_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(fl):
    lib = fl.lib()
    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(fl.EXPORTED_SYMBOLS) == names, set(fl.EXPORTED_SYMBOLS) ^ set(names)


def test_library_links_no_oracle_and_no_blas(fl):
    import subprocess
    out = subprocess.run(["ldd", fl.lib_path()], capture_output=True, text=True).stdout
    for banned in ("ref_port", "libref", "cublas", "openblas", "mkl"):
        assert banned not in out, out


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_fails_loudly_without_a_gpu(fl):
    with pytest.raises(fl.FlError) as ei:
        fl.Engine(512, 704, 2, 4, 4, 1000)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)
    x = np.ones(64, np.float32)
    with pytest.raises(fl.FlError):
        fl.ops.rmsnorm(x, x)


def test_invalid_arguments_are_rejected_before_touching_the_device(fl):
    lib = fl.lib()
    out = C.c_void_p()
    assert lib.fl_create(None, 0, C.byref(out)) == -1                      # FL_ERR_INVALID
    cfg = fl.FlConfig(512, 704, 2, 3, 3, 128, 1000, 1024, fl.Q_INT8, 64, 1, 0)
    assert lib.fl_create(C.byref(cfg), 0, C.byref(out)) == -1              # head_size * n_heads != dim
    cfg = fl.FlConfig(512, 704, 2, 4, 4, 128, 1000, 1024, fl.Q_INT16, 32, 1, 0)
    assert lib.fl_create(C.byref(cfg), 0, C.byref(out)) == -4              # FL_ERR_UNSUPPORTED: int16 has no group 32
    assert lib.fl_last_error(None)
    assert lib.fl_step_bytes(None, 0) == 0 and lib.fl_launch_count(None) == 0
