"""arrowspace-rs_b200: B200-native (sm_100a) lambda-tau build + lambda-aware search.

The directory name carries a hyphen (it mirrors the reference's repository name), so it is
imported through the root-level shim ``arrowspace_b200`` (``import arrowspace_b200 as asb``).
"""
from . import _build  # noqa: F401
from .host import (  # noqa: F401
    ABI_SYMBOLS, ArrowItem, ArrowSpace, ArrowSpaceBuilder, ArrowSpaceError, ClusteredOutput, Context, GraphLaplacian,
    GraphParams, TauMode, TAUDEFAULT, TAU_FLOOR, default_context, load_library,
)
from . import host, heuristics, parallel, synth  # noqa: F401
