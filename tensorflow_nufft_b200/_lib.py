"""ctypes binding of libb200nufft.so (C ABI: include/b200nufft.h).

There is NO fallback: if the shared library is missing, or no CUDA device is present, every
entry point raises. Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C tensorflow_nufft_b200/csrc`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200nufft.so")

OK, INVALID_ARGUMENT, UNIMPLEMENTED, RESOURCE_EXHAUSTED, INTERNAL = range(5)
COMPLEX64, COMPLEX128 = 0, 1


class Opts(ctypes.Structure):
  _fields_ = [
      ("points_range", ctypes.c_int),
      ("check_points_range", ctypes.c_int),
      ("max_batch_size", ctypes.c_int),
      ("spread_only", ctypes.c_int),
      ("fseries_mode", ctypes.c_int),
      ("num_threads_compat", ctypes.c_int),
      ("bin_dims", ctypes.c_int * 3),
      ("max_subproblem_size", ctypes.c_int),
      ("spread_method", ctypes.c_int),
      ("interp_method", ctypes.c_int),
      ("profile", ctypes.c_int),
      ("upsampling", ctypes.c_int),
      ("reuse_points", ctypes.c_int),
      ("external_workspace", ctypes.c_int),
      ("reserved", ctypes.c_int * 8),
  ]


ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int)
FREE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int)


class Allocator(ctypes.Structure):
  """b200nufft_allocator: device-memory callbacks (alloc, free, user)."""
  _fields_ = [("alloc", ALLOC_FN), ("free", FREE_FN), ("user", ctypes.c_void_p)]


class Info(ctypes.Structure):
  _fields_ = [
      ("kernel_width", ctypes.c_int),
      ("kernel_beta", ctypes.c_double),
      ("kernel_c", ctypes.c_double),
      ("upsampling_factor", ctypes.c_double),
      ("kernel_scale", ctypes.c_double),
      ("fine_dims", ctypes.c_int * 3),
      ("bin_dims", ctypes.c_int * 3),
      ("num_bins", ctypes.c_int * 3),
      ("batch_size", ctypes.c_int),
      ("num_threads_compat", ctypes.c_int),
      ("num_points", ctypes.c_int64),
      ("subproblem_bound", ctypes.c_int64),
      ("spread_method", ctypes.c_int),
      ("interp_method", ctypes.c_int),
      ("fft_method", ctypes.c_int),
  ]


class NufftError(RuntimeError):
  """Engine error; `.code` is the C ABI return code."""

  def __init__(self, code, message):
    super().__init__(message)
    self.code = code


class InvalidArgumentError(NufftError, ValueError):
  pass


_lib = None

# (name, restype, argtypes) for every symbol include/b200nufft.h declares.
_P = ctypes.c_void_p
SIGNATURES = [
    ("b200nufft_default_opts", None, [ctypes.POINTER(Opts)]),
    ("b200nufft_plan_create", ctypes.c_int,
     [ctypes.POINTER(_P), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
      ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Opts), ctypes.c_int]),
    ("b200nufft_plan_create_ex", ctypes.c_int,
     [ctypes.POINTER(_P), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
      ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Opts), ctypes.c_int, ctypes.POINTER(Allocator)]),
    ("b200nufft_plan_destroy", None, [_P]),
    ("b200nufft_plan_acquire", ctypes.c_int,
     [ctypes.POINTER(_P), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
      ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(Opts), ctypes.c_int, ctypes.POINTER(Allocator)]),
    ("b200nufft_plan_release", None, [_P]),
    ("b200nufft_plan_cache_clear", None, []),
    ("b200nufft_plan_cache_stats", None, [ctypes.POINTER(ctypes.c_int64)]),
    ("b200nufft_workspace_bytes", ctypes.c_size_t, [_P, ctypes.c_int64]),
    ("b200nufft_bind_workspace", ctypes.c_int, [_P, _P, ctypes.c_size_t, ctypes.c_int64]),
    ("b200nufft_unbind_workspace", ctypes.c_int, [_P]),
    ("b200nufft_reserve", ctypes.c_int, [_P, ctypes.c_int64]),
    ("b200nufft_debug_alloc_counts", None, [ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    ("b200nufft_get_reuse_stats", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64)]),
    ("b200nufft_set_points", ctypes.c_int, [_P, ctypes.c_int64, _P, _P, _P, _P]),
    ("b200nufft_set_points_interleaved", ctypes.c_int, [_P, ctypes.c_int64, _P, _P]),
    ("b200nufft_execute", ctypes.c_int, [_P, _P, _P, _P]),
    ("b200nufft_interp", ctypes.c_int, [_P, _P, _P, _P]),
    ("b200nufft_spread", ctypes.c_int, [_P, _P, _P, _P]),
    ("b200nufft_get_sort", ctypes.c_int,
     [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32)]),
    ("b200nufft_binsort", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _P, _P, _P, ctypes.POINTER(ctypes.c_int),
      ctypes.POINTER(ctypes.c_int), ctypes.c_int, _P, _P, _P, _P]),
    ("b200nufft_fold_rescale", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _P, _P, ctypes.c_int, _P]),
    ("b200nufft_copy_to_host", ctypes.c_int, [_P, _P, ctypes.c_size_t]),
    ("b200nufft_get_info", ctypes.c_int, [_P, ctypes.POINTER(Info)]),
    ("b200nufft_get_fseries", ctypes.c_int, [_P, ctypes.c_int, _P]),
    ("b200nufft_get_timings", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_float)]),
    ("b200nufft_launch_count", ctypes.c_int64, [_P]),
    ("b200nufft_last_error", ctypes.c_char_p, [_P]),
    ("b200nufft_last_create_error", ctypes.c_char_p, []),
    ("b200nufft_host_kernel_width", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double]),
    ("b200nufft_host_next_smooth_int", ctypes.c_int, [ctypes.c_int]),
    ("b200nufft_host_fseries", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P]),
    ("b200nufft_host_scale_factor", ctypes.c_double, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    ("b200nufft_host_gauss_legendre", ctypes.c_int,
     [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    ("b200nufft_version", ctypes.c_char_p, []),
]


def lib():
  """Loads the shared library (once). Raises if it has not been built: no fallback."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          f"{LIB_PATH} is missing: the CUDA engine has not been built and there is no CPU "
          "fallback. Run `make -C tensorflow_nufft_b200/csrc`.")
    L = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SIGNATURES:
      fn = getattr(L, name)
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = L
  return _lib


def raise_for(code, message):
  if code == OK:
    return
  if code == INVALID_ARGUMENT:
    raise InvalidArgumentError(code, message)
  if code == UNIMPLEMENTED:
    raise NufftError(code, "Unimplemented: " + message)
  if code == RESOURCE_EXHAUSTED:
    raise NufftError(code, "ResourceExhausted: " + message)
  raise NufftError(code, "Internal: " + message)


def make_opts(**opt_kwargs):
  """b200nufft_opts from keyword arguments (field names of the C struct, plus names for the
  reserved[] A/B switches)."""
  opts = Opts()
  lib().b200nufft_default_opts(ctypes.byref(opts))
  reserved = {"no_tma": 0, "coils_per_cta": 1, "no_preclear": 2, "no_pack": 3, "full_fft": 4,
              "fft_mode": 4,
              "no_tma_flush": 5, "no_zrange": 6, "no_point_major": 7, "otf_weights": 7}
  for k, v in opt_kwargs.items():
    if k == "bin_dims":
      for i, b in enumerate(v):
        opts.bin_dims[i] = int(b)
    elif k in reserved:
      opts.reserved[reserved[k]] = int(v)
    else:
      if not hasattr(opts, k):
        raise TypeError(f"unknown option {k}")
      setattr(opts, k, int(v))
  return opts


def alloc_counts():
  """(allocations, frees) issued by the library so far (cudaMalloc or allocator callbacks)."""
  a, f = ctypes.c_int64(), ctypes.c_int64()
  lib().b200nufft_debug_alloc_counts(ctypes.byref(a), ctypes.byref(f))
  return a.value, f.value


def plan_cache_stats():
  out = (ctypes.c_int64 * 3)()
  lib().b200nufft_plan_cache_stats(out)
  return {"hits": out[0], "misses": out[1], "idle": out[2]}


def plan_cache_clear():
  lib().b200nufft_plan_cache_clear()


class Plan:
  """Owns one b200nufft_plan handle. Pointers are raw device addresses (ints).

  cached=True takes the handle from the library's process-level plan cache
  (b200nufft_plan_acquire) and close() gives it back (b200nufft_plan_release) instead of
  destroying it. allocator: an `Allocator` (kept alive by this object)."""

  def __init__(self, transform_type, grid_dims, fft_sign, num_transforms, tol, dtype_code,
               device=0, cached=False, allocator=None, **opt_kwargs):
    L = lib()
    opts = make_opts(**opt_kwargs)
    self.rank = len(grid_dims)
    gd = (ctypes.c_int64 * 3)(*([int(g) for g in grid_dims] + [1] * (3 - self.rank)))
    h = _P()
    self._allocator = allocator
    aptr = ctypes.byref(allocator) if allocator is not None else None
    args = (ctypes.byref(h), int(transform_type), self.rank, gd, int(fft_sign), int(num_transforms),
            float(tol), int(dtype_code), ctypes.byref(opts), int(device))
    if cached:
      rc = L.b200nufft_plan_acquire(*args, aptr)
    elif allocator is not None:
      rc = L.b200nufft_plan_create_ex(*args, aptr)
    else:
      rc = L.b200nufft_plan_create(*args)
    if rc != OK:
      raise_for(rc, L.b200nufft_last_create_error().decode())
    self._h = h
    self.cached = bool(cached)
    self.type = int(transform_type)
    self.num_transforms = int(num_transforms)
    self.grid_dims = [int(g) for g in grid_dims]
    self.dtype_code = int(dtype_code)
    self.device = int(device)

  def _check(self, rc):
    if rc != OK:
      raise_for(rc, lib().b200nufft_last_error(self._h).decode())

  def set_points(self, num_points, x, y, z, stream):
    self._check(lib().b200nufft_set_points(self._h, int(num_points), x, y, z, stream))

  def set_points_interleaved(self, num_points, pts, stream):
    self._check(lib().b200nufft_set_points_interleaved(self._h, int(num_points), pts, stream))

  def execute(self, c, f, stream):
    self._check(lib().b200nufft_execute(self._h, c, f, stream))

  def interp(self, c, f, stream):
    self._check(lib().b200nufft_interp(self._h, c, f, stream))

  def spread(self, c, f, stream):
    self._check(lib().b200nufft_spread(self._h, c, f, stream))

  def info(self):
    inf = Info()
    self._check(lib().b200nufft_get_info(self._h, ctypes.byref(inf)))
    return inf

  def sort_pointers(self):
    idx, bs, bz = _P(), _P(), _P()
    n = ctypes.c_int32()
    self._check(lib().b200nufft_get_sort(self._h, ctypes.byref(idx), ctypes.byref(bs),
                                         ctypes.byref(bz), ctypes.byref(n)))
    return idx.value, bs.value, bz.value, n.value

  def sort_arrays(self):
    """Host copies (numpy int32) of idx[M], bin_start[bins], bin_sizes[bins]."""
    import numpy as np
    idx, bs, bz, n = self.sort_pointers()
    m = int(self.info().num_points)
    out = []
    for ptr, cnt in ((idx, m), (bs, n), (bz, n)):
      a = np.empty(cnt, np.int32)
      if cnt:
        self._check(lib().b200nufft_copy_to_host(a.ctypes.data, ptr, 4 * cnt))
      out.append(a)
    return out

  def fseries(self, dim, real_np_dtype):
    import numpy as np
    nf = self.info().fine_dims[dim]
    out = np.empty(nf // 2 + 1, real_np_dtype)
    self._check(lib().b200nufft_get_fseries(self._h, dim, out.ctypes.data))
    return out

  def timings(self):
    out = (ctypes.c_float * 4)()
    self._check(lib().b200nufft_get_timings(self._h, out))
    return {"spread_interp_ms": out[0], "fft_ms": out[1], "deconv_ms": out[2], "set_points_ms": out[3]}

  def launch_count(self):
    return int(lib().b200nufft_launch_count(self._h))

  def workspace_bytes(self, num_points):
    return int(lib().b200nufft_workspace_bytes(self._h, int(num_points)))

  def bind_workspace(self, ptr, nbytes, num_points):
    self._check(lib().b200nufft_bind_workspace(self._h, ptr, int(nbytes), int(num_points)))

  def unbind_workspace(self):
    self._check(lib().b200nufft_unbind_workspace(self._h))

  def reserve(self, num_points):
    self._check(lib().b200nufft_reserve(self._h, int(num_points)))

  def reuse_stats(self):
    out = (ctypes.c_int64 * 2)()
    self._check(lib().b200nufft_get_reuse_stats(self._h, out))
    return {"skipped": out[0], "full": out[1]}

  def close(self):
    if getattr(self, "_h", None):
      if self.cached:
        lib().b200nufft_plan_release(self._h)
      else:
        lib().b200nufft_plan_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass
