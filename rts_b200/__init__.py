"""rts_b200 — B200-native (sm_100a) ray-tracing radar path of ymartin101/RTS behind a C-ABI.

The compute path is librts_b200.so (hand-written CUDA, rts_b200/csrc); this package is the thin
Python binding used by the tests and bench.py.  There is no CPU fallback.
"""
from .abi import PulseSpec, Target, RAY_RECORD, BIN_DTYPE  # noqa: F401
from . import lib  # noqa: F401
from .lib import Engine, RtsError  # noqa: F401
