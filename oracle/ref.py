"""ctypes front end to oracle/_ref/libref.so -- the reference's OWN CPU plan
(`tensorflow::nufft::Plan<CPUDevice, F>`, /root/reference/tensorflow_nufft/cc/kernels/nufft_plan.cc)
compiled unmodified by oracle/ref_build/Makefile.

TEST INFRASTRUCTURE ONLY. Importers allowed: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs. The product path (tensorflow_nufft_b200) never imports this.

Modes:
  mode="auto"       what `tfft.nufft` does on /cpu:0 (Horner evaluation, automatic sigma)
  mode="gpuparams"  the same CPU code driven with Plan<GPUDevice>'s choices
                    (sigma = 2.0, direct exp(sqrt) evaluation; nufft_plan.cu.cc:1849-1857):
                    the parity target for the CUDA engine.
  mode="lowups_direct"  sigma = 1.25 (the CPU plan's choice for large grids, nufft_plan.h:745-752)
                    with direct evaluation: the parity target of the engine's opts.upsampling = 1.
`tol` is cast through float32 exactly like the op attr (nufft_ops.cc:214, nufft_kernels.cc:361)
unless tol_is_exact=True.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
_lib = None

POINTS_RANGE = {"strict": 0, "extended": 1, "infinite": 2}


def available():
  return os.path.exists(_LIB_PATH)


def lib():
  global _lib
  if _lib is None:
    if not available():
      raise RuntimeError(
          f"{_LIB_PATH} not built; run `make -C oracle/ref_build` where /root/reference exists")
    L = ctypes.CDLL(_LIB_PATH)
    L.ref_plan_create.restype = ctypes.c_void_p
    L.ref_plan_create.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
        ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    L.ref_set_points.restype = ctypes.c_int
    L.ref_set_points.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    L.ref_run.restype = ctypes.c_int
    L.ref_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_char_p, ctypes.c_int]
    L.ref_get_params.restype = None
    L.ref_get_params.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                 ctypes.POINTER(ctypes.c_double)]
    L.ref_get_fseries.restype = ctypes.c_int
    L.ref_get_fseries.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ref_get_sort.restype = ctypes.c_int
    L.ref_get_sort.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.ref_plan_destroy.restype = None
    L.ref_plan_destroy.argtypes = [ctypes.c_void_p]
    L.ref_kernel_fseries.restype = None
    L.ref_kernel_fseries.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    L.ref_next_smooth_int.restype = ctypes.c_int
    L.ref_next_smooth_int.argtypes = [ctypes.c_int]
    L.ref_scale_factor.restype = ctypes.c_double
    L.ref_scale_factor.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_double]
    _lib = L
  return _lib


def op_tol(tol, real_dtype, tol_is_exact=False):
  """The tolerance the plan sees: the op attr is a 32-bit float (nufft_ops.cc:214)."""
  if tol_is_exact:
    return float(tol)
  return float(np.float32(tol))


class RefPlan:
  """One reference CPU plan. Layout contract = PlanBase's (nufft_plan.h:223-256): grid_dims are
  x-fastest, points are `rank` separate contiguous arrays, c is [T][M], f is [T][N] x-fastest."""

  def __init__(self, transform_type, grid_dims, fft_sign, num_transforms, tol, dtype,
               mode="gpuparams", points_range="extended", check_points_range=False,
               max_batch_size=0, num_threads=0, spread_only=False, tol_is_exact=False):
    L = lib()
    self.cdtype = np.dtype(dtype)
    assert self.cdtype in (np.complex64, np.complex128)
    self.is_double = int(self.cdtype == np.complex128)
    self.rdtype = np.dtype(np.float64 if self.is_double else np.float32)
    self.rank = len(grid_dims)
    self.grid_dims = [int(g) for g in grid_dims]
    self.type = int(transform_type)
    self.T = int(num_transforms)
    if mode == "auto":
      upsampfac, kerevalmeth = 0.0, 0
    elif mode == "gpuparams":
      upsampfac, kerevalmeth = 2.0, 1
    elif mode == "horner2":   # sigma=2 but Horner: isolates the evaluator difference
      upsampfac, kerevalmeth = 2.0, 2
    elif mode == "lowups_direct":   # sigma=1.25 with direct evaluation: parity target of opts.upsampling=1
      upsampfac, kerevalmeth = 1.25, 1
    else:
      raise ValueError(mode)
    err = ctypes.create_string_buffer(512)
    gd = (ctypes.c_int * 3)(*(self.grid_dims + [1] * (3 - self.rank)))
    self._h = L.ref_plan_create(
        self.is_double, self.type, self.rank, gd, int(fft_sign), self.T,
        op_tol(tol, self.rdtype, tol_is_exact), POINTS_RANGE[points_range],
        int(check_points_range), int(max_batch_size), upsampfac, kerevalmeth, int(num_threads),
        int(spread_only), err, 512)
    if not self._h:
      raise ValueError(err.value.decode())
    self._pts = None
    self.M = 0
    iout = (ctypes.c_int * 8)()
    dout = (ctypes.c_double * 8)()
    L.ref_get_params(self._h, iout, dout)
    self.kernel_width = iout[0]
    self.fine_dims = [iout[1 + d] for d in range(self.rank)]
    self.batch_size = iout[4]
    self.num_threads = iout[5]
    self.kerevalmeth = iout[6]
    self.beta, self.c, self.sigma, self.kernel_scale, self.half_width = (dout[i] for i in range(5))
    self.spread_only = bool(spread_only)

  def set_points(self, points):
    """points: [rank][M] array (coordinate 0 = fastest grid axis). Copied; the copy is mutated
    in place by the reference (folded + rescaled) and is readable afterwards as self.folded."""
    L = lib()
    pts = np.ascontiguousarray(np.array(points, dtype=self.rdtype, copy=True))
    assert pts.shape[0] == self.rank
    self.M = pts.shape[1]
    self._pts = pts
    err = ctypes.create_string_buffer(512)
    ptrs = [pts[d].ctypes.data for d in range(self.rank)] + [None] * (3 - self.rank)
    rc = L.ref_set_points(self._h, self.M, ptrs[0], ptrs[1], ptrs[2], err, 512)
    if rc:
      raise ValueError(err.value.decode())
    return self

  @property
  def folded(self):
    return self._pts

  def sort_indices(self):
    out = np.empty(self.M, np.int32)
    did = lib().ref_get_sort(self._h, self.M, out.ctypes.data)
    return out, bool(did)

  def fseries(self, dim):
    out = np.empty(self.fine_dims[dim] // 2 + 1, self.rdtype)
    rc = lib().ref_get_fseries(self._h, dim, out.ctypes.data)
    if rc:
      raise RuntimeError("no fseries (spread-only plan?)")
    return out

  def _run(self, op, c, f):
    err = ctypes.create_string_buffer(512)
    rc = lib().ref_run(self._h, op, c.ctypes.data, f.ctypes.data, err, 512)
    if rc:
      raise RuntimeError(err.value.decode())

  def execute(self, src):
    """type 1: src = c[T][M] -> f[T][N]; type 2: src = f[T][N] -> c[T][M]. N is x-fastest."""
    N = int(np.prod(self.grid_dims))
    src = np.ascontiguousarray(src, dtype=self.cdtype)
    if self.type == 1:
      c = src.reshape(self.T, self.M)
      f = np.zeros((self.T, N), self.cdtype)
      self._run(0, c, f)
      return f
    f = src.reshape(self.T, N)
    c = np.zeros((self.T, self.M), self.cdtype)
    self._run(0, c, f)
    return c

  def interp(self, f):
    N = int(np.prod(self.grid_dims))
    f = np.ascontiguousarray(f, dtype=self.cdtype).reshape(self.T, N)
    c = np.zeros((self.T, self.M), self.cdtype)
    self._run(1, c, f)
    return c

  def spread(self, c):
    N = int(np.prod(self.grid_dims))
    c = np.ascontiguousarray(c, dtype=self.cdtype).reshape(self.T, self.M)
    f = np.zeros((self.T, N), self.cdtype)
    self._run(2, c, f)
    return f

  def close(self):
    if getattr(self, "_h", None):
      lib().ref_plan_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def kernel_fseries(nf, ns, beta, c, num_threads, dtype):
  """Reference kernel_fseries_1d (nufft_util.cc:71-117)."""
  dt = np.dtype(dtype)
  out = np.empty(nf // 2 + 1, dt)
  lib().ref_kernel_fseries(int(dt == np.float64), nf, ns, beta, c, num_threads, out.ctypes.data)
  return out


def next_smooth_int(n):
  return lib().ref_next_smooth_int(int(n))


def scale_factor(rank, ns, beta, c, dtype):
  return lib().ref_scale_factor(int(np.dtype(dtype) == np.float64), rank, ns, beta, c)
