"""sad_monte_carlo_b200 -- B200-native walker engine for flat-histogram Monte Carlo.

Only the hot path of droundy/sad-monte-carlo lives here: the propose / dE /
accept loop plus SAD / WL / 1/t-WL / SAMC bookkeeping, for thousands of
independent walkers, as hand-written sm_100a CUDA behind the C ABI declared in
include/sadmc_gpu.h.  There is no CPU fallback: loading fails loudly when
libsadmc_gpu.so has not been built.
"""
import ctypes as _C
import os as _os

from . import _abi
from ._abi import make_config, Config, WalkerState  # noqa: F401

_HERE = _os.path.dirname(_os.path.abspath(__file__))
# SADMC_GPU_LIB: another build of the same library (kernel experiments); never a different implementation
LIB_PATH = _os.environ.get("SADMC_GPU_LIB") or _os.path.join(_HERE, "libsadmc_gpu.so")
_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not _os.path.exists(LIB_PATH):
            raise RuntimeError(
                "sad_monte_carlo_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIB_PATH)
        from ._capi import bind
        _lib = bind(_C.CDLL(LIB_PATH))
    return _lib


def __getattr__(name):
    if name in ("WalkerEngine", "EnergyMC"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
