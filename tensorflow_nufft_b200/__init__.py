"""B200-native NUFFT engine behind the `tfft.nufft` operator surface of mrphys/tensorflow-nufft.

Public names mirror `tensorflow_nufft/__init__.py` of the reference: `nufft`, `nudft`, `interp`,
`spread`, `Options`, `DebuggingOptions`, `FftwOptions`, `FftwPlanningRigor`, `PointsRange`.
"""
from tensorflow_nufft_b200.python.ops.nufft_ops import (  # noqa: F401
    nufft, nudft, interp, spread, clear_plan_cache, set_engine_defaults, set_points_reuse)
from tensorflow_nufft_b200.python.ops.nufft_options import (  # noqa: F401
    Options, DebuggingOptions, FftwOptions, FftwPlanningRigor, PointsRange)

__version__ = "0.1.0"
