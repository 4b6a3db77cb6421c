"""Options for `nufft`, mirroring `tfft.Options` of the reference
(tensorflow_nufft/python/ops/nufft_options.py:24-273): same class and attribute names, same
defaults (points_range EXTENDED, max_batch_size None, debugging.check_points_range False,
fftw.planning_rigor AUTO). The proto round trip is replaced by `to_engine_kwargs()`, which
yields the already-decoded scalars the C ABI takes (include/b200nufft.h: b200nufft_opts).
"""
import dataclasses
import enum
import typing


class FftwPlanningRigor(enum.IntEnum):
  """FFTW planning rigor. Accepted for API compatibility; the GPU path has no FFTW."""
  AUTO = 0
  ESTIMATE = 1
  MEASURE = 2
  PATIENT = 3
  EXHAUSTIVE = 4


class PointsRange(enum.IntEnum):
  """Supported range of the nonuniform points: [-pi, pi], [-3pi, 3pi] or unbounded."""
  STRICT = 0
  EXTENDED = 1
  INFINITE = 2


@dataclasses.dataclass
class DebuggingOptions:
  check_points_range: bool = False


@dataclasses.dataclass
class FftwOptions:
  planning_rigor: FftwPlanningRigor = FftwPlanningRigor.AUTO

  def __post_init__(self):
    self.planning_rigor = FftwPlanningRigor(self.planning_rigor)


@dataclasses.dataclass
class Options:
  """Advanced options for `nufft` (same fields as `tfft.Options`)."""
  debugging: DebuggingOptions = dataclasses.field(default_factory=DebuggingOptions)
  fftw: FftwOptions = dataclasses.field(default_factory=FftwOptions)
  max_batch_size: typing.Optional[int] = None
  points_range: PointsRange = PointsRange.EXTENDED

  def __post_init__(self):
    self._validate()

  def _validate(self):
    if isinstance(self.points_range, str):
      self.points_range = PointsRange[self.points_range.upper()]
    self.points_range = PointsRange(self.points_range)
    if self.max_batch_size is not None:
      if not isinstance(self.max_batch_size, int) or isinstance(self.max_batch_size, bool):
        raise ValueError("max_batch_size must be an integer or None")
      if self.max_batch_size < 0:
        raise ValueError("max_batch_size must be non-negative")
    if not isinstance(self.debugging, DebuggingOptions):
      raise ValueError("debugging must be a DebuggingOptions")
    if not isinstance(self.fftw, FftwOptions):
      raise ValueError("fftw must be a FftwOptions")

  def to_engine_kwargs(self):
    self._validate()
    return {
        "points_range": int(self.points_range),
        "check_points_range": int(bool(self.debugging.check_points_range)),
        "max_batch_size": int(self.max_batch_size or 0),
    }
