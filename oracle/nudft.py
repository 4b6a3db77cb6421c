"""Direct non-uniform DFT in float64 (numpy). TEST INFRASTRUCTURE ONLY.

Matches the reference's own test oracle `_nudft_matrix` (tensorflow_nufft/python/ops/
nufft_ops.py:293-321): A[j, k] = exp(-+ i sum_d k_d x_{j,d}), k_d = -N_d/2 ... N_d/2 - 1.
Plan-level layout (PlanBase, nufft_plan.h:223-256): points [rank][M] with coordinate 0 the
fastest-varying grid axis; grids flattened x-fastest.
"""
import numpy as np


def _phase(points, grid_dims):
  rank = len(grid_dims)
  ks = [np.arange(n, dtype=np.float64) - (n // 2) for n in grid_dims]
  # x-fastest flattening: index = k0 + n0*(k1 + n1*k2)
  mesh = np.meshgrid(*ks[::-1], indexing="ij")  # slowest first
  ph = np.zeros((points.shape[1], mesh[0].size))
  for d in range(rank):
    kd = mesh[rank - 1 - d].reshape(-1)
    ph += np.outer(points[d].astype(np.float64), kd)
  return ph


def nudft_plan_layout(src, points, grid_dims, transform_type, fft_sign):
  """src: [T][M] (type 1) or [T][N] (type 2); returns [T][N] or [T][M] complex128."""
  points = np.asarray(points, dtype=np.float64)
  A = np.exp(1j * float(np.sign(fft_sign)) * _phase(points, grid_dims))  # [M, N]
  src = np.asarray(src, dtype=np.complex128)
  if transform_type == 1:
    return src @ A
  return src @ A.T
