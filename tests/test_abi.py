"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and the ctypes mirrors have the C struct sizes.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

import sad_monte_carlo_b200 as pkg
from sad_monte_carlo_b200 import _abi, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "sadmc_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sadmc_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_all_exported_and_bound(gpu_lib):
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(gpu_lib, s), "libsadmc_gpu.so does not export %s" % s
        assert s in _capi.PROTOTYPES, "no ctypes prototype for %s" % s
    assert sorted(_capi.PROTOTYPES) == syms


def test_struct_sizes_match_the_compiled_library(gpu_lib):
    assert gpu_lib.sadmc_sizeof_config() == C.sizeof(_abi.Config)
    assert gpu_lib.sadmc_sizeof_walker_state() == C.sizeof(_abi.WalkerState)
    assert gpu_lib.sadmc_sizeof_binning_state() == C.sizeof(_abi.BinningState)
    assert gpu_lib.sadmc_sizeof_replica_state() == C.sizeof(_abi.ReplicaState)
    assert gpu_lib.sadmc_sizeof_zeno_replica_state() == C.sizeof(_abi.ZenoReplicaState)
    assert gpu_lib.sadmc_abi_version() == _abi.ABI_VERSION


def test_create_without_a_gpu_fails_loudly_not_silently(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = _abi.make_config("ising", N=32, n_walkers=4)
    h = C.c_void_p()
    rc = gpu_lib.sadmc_create(C.byref(cfg), C.byref(h))
    assert rc == _abi.ERR_CUDA
    assert b"no CPU fallback" in gpu_lib.sadmc_last_error()
    with pytest.raises(RuntimeError):
        pkg.WalkerEngine(cfg)


def test_product_never_references_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "sad_monte_carlo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                if "oracle/" in open(os.path.join(d, f), errors="ignore").read().replace("the oracle", ""):
                    bad.append(f)
    assert not bad, bad
