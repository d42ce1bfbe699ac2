"""natrium_b200 -- B200-native (sm_100a) implementation of NATriuM's per-timestep hot path.

Semi-Lagrangian streaming (block-sparse SpMV with the pre-assembled interpolation matrices)
fused with collision, device-resident DistributionFunctions, NCCL ghost exchange; reached
through the C ABI in include/natrium_b200.h.  There is no CPU fallback: the CUDA library must
be built (``__graft_entry__.build()``) and a GPU must be present to create a context.
"""
from . import _capi, harness, host, mrt, stencils  # noqa: F401
from ._capi import CollisionException, Context, NatriumB200Error  # noqa: F401
from .host import (CFDSolver, CompressibleCFDSolver, DistributionFunctions, ExponentialFilter, PseudoEntropicStabilizer,  # noqa: F401
                   SemiLagrangian, SolverConfiguration, selectCollision)
from .stencils import Stencil  # noqa: F401
