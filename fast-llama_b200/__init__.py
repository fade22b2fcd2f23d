"""fast-llama_b200 — B200-native decode engine behind the C-ABI in include/fastllama_b200.h.

This Python package is only the ctypes view of libfastllama_b200.so used by tests/, bench.py and
__graft_entry__.py.  The product is the shared library (CUDA, sm_100a); `shard` holds the host logic of the multi-GPU request sharding.
There is no CPU fallback: importing works anywhere, but every call needs the built library and a CUDA device.

The directory name has a hyphen (the repo's required layout); import it with

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("fast_llama_b200", ".../fast-llama_b200/__init__.py")
"""
from .binding import (Engine, Sampler, FlConfig, FlError, lib, lib_path, Q_INT8, Q_INT16,
                      T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS,
                      FLAG_NO_GRAPH, FLAG_NO_PDL, FLAG_NO_MEGAKERNEL, FLAG_PROFILE, FLAG_NO_TC, FLAG_RELAXED, ops, EXPORTED_SYMBOLS)

from . import shard, loaders, flm, gguf_file, tokenizer, convert
from .generate import generate_text

__all__ = ["shard", "loaders", "flm", "gguf_file", "tokenizer", "convert", "generate_text", "Engine", "Sampler", "FlConfig", "FlError", "lib", "lib_path", "ops", "EXPORTED_SYMBOLS"]
