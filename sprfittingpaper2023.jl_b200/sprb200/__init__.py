"""sprb200 -- Python host side of libsprb200.so (B200-native SSA surrogate engine).

Mirrors the reference's interface for the hot path (BioPhysParams, SimParams, outputters,
terminators, run_spr_sim, build_surrogate_*) on top of the C ABI in include/sprb200.h."""
from . import _lib  # noqa: F401
from ._lib import Engine, SimSpec, SprB200Error, make_grid, make_run  # noqa: F401
