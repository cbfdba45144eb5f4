"""cs_pathplan_b200 -- B200-native batched minimum-snap trajectory solver.

Drop-in for the reference's ``TrajectoryGeneratorTool`` (math_util/minimum_snap.{hpp,cpp}) behind a C ABI
(``include/msnap.h``, implemented in ``cs_pathplan_b200/csrc`` as hand-written sm_100a CUDA), plus this thin
Python host layer used by the tests and the benchmark.
"""
from .api import (  # noqa: F401
    AltitudeParams,
    shipped_altitude_params,
    BatchResult,
    Bezier,
    BezierConfig,
    MinimumSnapConfig,
    TrajectoryGeneratorTool,
    load_minimum_snap_config,
    shipped_config,
)
from .sharding import shard_bounds, shard_batch, shard_rows  # noqa: F401

__all__ = [
    "AltitudeParams",
    "shipped_altitude_params",
    "BatchResult",
    "Bezier",
    "BezierConfig",
    "MinimumSnapConfig",
    "TrajectoryGeneratorTool",
    "load_minimum_snap_config",
    "shipped_config",
    "shard_bounds",
    "shard_batch",
    "shard_rows",
]
