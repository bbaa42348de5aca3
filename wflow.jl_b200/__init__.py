"""wflow.jl_b200 -- B200 (sm_100a) implementation of Wflow.jl's `sbm` per-timestep hot path
(SBM vertical land update + kinematic-wave routing) behind a C ABI (include/wflow_b200.h).

The directory name is not a valid dotted module name; load it by path:
`__graft_entry__.load_pkg()` registers it as module `wflow_jl_b200` (tests/conftest.py and
bench.py use that helper).
"""
from . import _lib, partition, synthetic  # noqa: F401
from ._lib import build, header_symbols  # noqa: F401
from .model import SbmModel, ShardGroup, WflowB200Error  # noqa: F401
from .network import build_network_artifacts  # noqa: F401
