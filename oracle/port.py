"""ctypes front end to oracle/liboracle.so, the plain-C restatement of the reference algorithm
(oracle/nufft_oracle.c). TEST INFRASTRUCTURE ONLY -- see the header of nufft_oracle.c.
Same plan-level layout as oracle/ref.py: grid_dims x-fastest, points [rank][M]."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_P = ctypes.c_void_p


def available():
  return os.path.exists(_LIB_PATH)


def lib():
  global _lib
  if _lib is None:
    if not available():
      raise RuntimeError(f"{_LIB_PATH} not built; run `make -C oracle`")
    L = ctypes.CDLL(_LIB_PATH)
    L.oracle_next_smooth_int.restype = ctypes.c_int
    L.oracle_next_smooth_int.argtypes = [ctypes.c_int]
    L.oracle_gauss_legendre.restype = None
    L.oracle_gauss_legendre.argtypes = [ctypes.c_int, _P, _P]
    for sfx in ("f32", "f64"):
      f = getattr(L, f"oracle_kernel_params_{sfx}")
      f.restype = ctypes.c_int
      f.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
      f = getattr(L, f"oracle_fold_rescale_{sfx}")
      f.restype = None
      f.argtypes = [ctypes.c_long, _P, _P, ctypes.c_int, ctypes.c_int]
      f = getattr(L, f"oracle_binsort_{sfx}")
      f.restype = None
      f.argtypes = [ctypes.c_int, ctypes.c_long, _P, _P, _P, ctypes.c_int, _P, _P, _P]
      f = getattr(L, f"oracle_kernel_fseries_{sfx}")
      f.restype = None
      f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, _P]
      f = getattr(L, f"oracle_scale_factor_{sfx}")
      f.restype = ctypes.c_double
      f.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
      f = getattr(L, f"oracle_nufft_{sfx}")
      f.restype = ctypes.c_int
      f.argtypes = [ctypes.c_int, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                    ctypes.c_int, ctypes.c_long, _P, _P, _P, _P]
    _lib = L
  return _lib


def _sfx(dtype):
  return "f64" if np.dtype(dtype) in (np.float64, np.complex128) else "f32"


def kernel_params(tol, dtype, sigma=2.0):
  b, c = ctypes.c_double(), ctypes.c_double()
  ns = getattr(lib(), f"oracle_kernel_params_{_sfx(dtype)}")(float(tol), float(sigma), ctypes.byref(b), ctypes.byref(c))
  return ns, b.value, c.value


def next_smooth_int(n):
  return lib().oracle_next_smooth_int(int(n))


def gauss_legendre(n):
  x = np.empty(n)
  w = np.empty(n)
  lib().oracle_gauss_legendre(n, x.ctypes.data, w.ctypes.data)
  return x, w


def fold_rescale(x, nf, points_range="extended"):
  x = np.ascontiguousarray(x)
  out = np.empty_like(x)
  getattr(lib(), f"oracle_fold_rescale_{_sfx(x.dtype)}")(x.size, x.ctypes.data, out.ctypes.data,
                                                         {"strict": 0, "extended": 1, "infinite": 2}[points_range], int(nf))
  return out


def binsort(folded, fine_dims, bin_dims, rounding=0):
  folded = np.ascontiguousarray(folded)
  rank, M = folded.shape
  nb = 1
  for d in range(rank):
    nb *= ((fine_dims[d] + bin_dims[d] - 1) // bin_dims[d]) if rounding == 0 else (fine_dims[d] // bin_dims[d] + 1)
  idx = np.empty(max(M, 1), np.int32)
  bs = np.empty(nb, np.int32)
  bz = np.empty(nb, np.int32)
  nf = np.asarray(fine_dims, np.int32)
  bd = np.asarray(bin_dims, np.int32)
  getattr(lib(), f"oracle_binsort_{_sfx(folded.dtype)}")(rank, M, folded.ctypes.data, nf.ctypes.data, bd.ctypes.data,
                                                      rounding, idx.ctypes.data, bs.ctypes.data, bz.ctypes.data)
  return idx[:M], bs, bz


def kernel_fseries(nf, ns, beta, c, num_threads, dtype):
  out = np.empty(nf // 2 + 1, np.dtype(dtype))
  getattr(lib(), f"oracle_kernel_fseries_{_sfx(dtype)}")(nf, ns, beta, c, num_threads, out.ctypes.data)
  return out


def scale_factor(rank, beta, c, dtype):
  return getattr(lib(), f"oracle_scale_factor_{_sfx(dtype)}")(rank, beta, c)


def nufft(src, points, grid_dims, transform_type, fft_sign, tol, dtype, points_range="extended",
          num_threads=1, tol_is_exact=False):
  """src [T][M] (type 1) or [T][N] (type 2); points [rank][M] radians; grid_dims x-fastest."""
  cd = np.dtype(dtype)
  rd = np.float64 if cd == np.complex128 else np.float32
  pts = np.ascontiguousarray(points, dtype=rd)
  rank, M = pts.shape
  N = int(np.prod(grid_dims))
  src = np.ascontiguousarray(src, dtype=cd)
  T = src.shape[0]
  if transform_type == 1:
    c = src.reshape(T, M).copy()
    f = np.zeros((T, N), cd)
  else:
    f = src.reshape(T, N).copy()
    c = np.zeros((T, M), cd)
  gd = np.asarray(grid_dims, np.int32)
  info = np.zeros(8, np.int32)
  t = float(tol) if tol_is_exact else float(np.float32(tol))
  rc = getattr(lib(), f"oracle_nufft_{_sfx(cd)}")(int(transform_type), rank, gd.ctypes.data, int(fft_sign), T, t,
                                                {"strict": 0, "extended": 1, "infinite": 2}[points_range],
                                                int(num_threads), M, pts.ctypes.data, c.ctypes.data, f.ctypes.data,
                                                info.ctypes.data)
  if rc:
    raise ValueError("oracle_nufft: bad arguments")
  return f if transform_type == 1 else c
