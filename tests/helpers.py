"""Shared helpers for the parity tests: synthetic point sets, numpy
restatements of the integer/bit-exact stages, error metrics."""
import numpy as np

PI32 = np.float32(3.14159265358979329)
INV2PI32 = np.float32(0.159154943091895336)


def rel_l2(a, b):
  a = np.asarray(a).astype(np.complex128).ravel()
  b = np.asarray(b).astype(np.complex128).ravel()
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def radial_points(spokes=200, samples=500, dtype=np.float32):
  """cfg 1/5: 2D radial, theta_s = pi s / spokes, r_i = -pi + 2 pi i / samples. [M, 2] TF layout."""
  th = np.pi * np.arange(spokes) / spokes
  r = -np.pi + 2 * np.pi * np.arange(samples) / samples
  kx = np.outer(np.cos(th), r).ravel()
  ky = np.outer(np.sin(th), r).ravel()
  return np.stack([kx, ky], -1).astype(dtype)


def spiral_points(interleaves=32, samples=62500, turns=48, dtype=np.float32):
  """cfg 2: Archimedean spiral, k = pi t exp(i (2 pi turns t + 2 pi l / interleaves))."""
  t = np.arange(samples) / samples
  out = []
  for l in range(interleaves):
    ph = 2 * np.pi * turns * t + 2 * np.pi * l / interleaves
    out.append(np.stack([np.pi * t * np.cos(ph), np.pi * t * np.sin(ph)], -1))
  return np.concatenate(out, 0).astype(dtype)


def uniform_points(m, rank, seed, dtype=np.float32):
  rng = np.random.default_rng(seed)
  return rng.uniform(-np.pi, np.pi, (m, rank)).astype(dtype)


def stack_of_stars_points(nz=125, spokes=125, samples=256, dtype=np.float32):
  """cfg 4: kz planes uniform in [-pi, pi), golden-angle spokes in-plane. [M, 3] (z, y, x) order
  is irrelevant for the synthetic set; returns columns (kz, ky, kx)."""
  kz = -np.pi + 2 * np.pi * np.arange(nz) / nz
  ang = np.deg2rad(111.246) * np.arange(spokes)
  r = -np.pi + 2 * np.pi * np.arange(samples) / samples
  kx = np.outer(np.cos(ang), r).ravel()
  ky = np.outer(np.sin(ang), r).ravel()
  pts = np.empty((nz, spokes * samples, 3), dtype)
  pts[:, :, 0] = kz[:, None]
  pts[:, :, 1] = ky[None, :]
  pts[:, :, 2] = kx[None, :]
  return pts.reshape(-1, 3)


def random_complex(shape, seed, dtype=np.complex64):
  rng = np.random.default_rng(seed)
  return (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(dtype)


def fold_rescale_np(x, nf, points_range="extended"):
  """FoldAndRescale functors (reference nufft_plan.h:676-734) in the array's own precision."""
  x = np.asarray(x)
  dt = x.dtype.type
  pi = dt(3.14159265358979329)
  twopi = dt(6.283185307179586)
  inv2pi = dt(0.159154943091895336)
  if points_range == "strict":
    s = x + pi
  elif points_range == "extended":
    s = np.where(x > pi, x - pi, np.where(x < -pi, x + dt(3.0) * pi, x + pi))
  else:
    s = np.fmod(x + pi, twopi)
    s = np.where(s < 0, s + twopi, s)
  return (s * inv2pi * dt(nf)).astype(x.dtype)


def binsort_np(folded, fine_dims, bin_dims, rounding=0):
  """Stable restatement of the reference bin-sort. rounding 0 = GPU rule (CalcBinSizeNoGhost*,
  nufft_plan.cu.cc:160-231): floor(x / bin) clamped into [0, ceil(nf/bin)); rounding 1 = CPU rule
  (binsort_singlethread, nufft_plan.cc:475-531): int(x / bin), nf/bin + 1 boxes. folded: [rank][M].
  Returns (idx, bin_start, bin_sizes)."""
  rank = len(fine_dims)
  key = np.zeros(folded.shape[1], np.int64)
  mul = 1
  nbtot = 1
  for d in range(rank):
    x = folded[d]
    q = x / x.dtype.type(bin_dims[d])
    if rounding == 0:
      nb = (fine_dims[d] + bin_dims[d] - 1) // bin_dims[d]
      b = np.floor(q).astype(np.int64)
      b = np.where(b >= nb, b - 1, b)
      b = np.where(b < 0, 0, b)
    else:
      nb = fine_dims[d] // bin_dims[d] + 1
      b = q.astype(np.int64)
    key += mul * b
    mul *= nb
    nbtot *= nb
  idx = np.argsort(key, kind="stable").astype(np.int32)
  sizes = np.bincount(key, minlength=nbtot).astype(np.int32)
  start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
  return idx, start, sizes


def nudft_samples(src, pts_tf, grid, ttype, sign, n_samples=256, seed=0):
  """Float64 direct NUDFT of a random SAMPLE of the outputs (cheap at any size): returns
  (flat output indices, truth[T][n_samples]). src: [T, M] (type 1) or [T, *grid] (type 2);
  pts_tf: [M, rank] in the op's layout; grid in TF order; output flattened in TF (row-major) order."""
  rng = np.random.default_rng(seed)
  pts = np.asarray(pts_tf, np.float64)
  M, rank = pts.shape
  T = src.shape[0]
  ks = [np.arange(n, dtype=np.float64) - (n // 2) for n in grid]
  src = np.asarray(src, np.complex128)
  if ttype == 2:
    sel = rng.choice(M, size=min(n_samples, M), replace=False)
    f = src.reshape((T,) + tuple(grid))
    out = np.empty((T, sel.size), np.complex128)
    for i, j in enumerate(sel):
      ph = [np.exp(1j * sign * pts[j, d] * ks[d]) for d in range(rank)]
      v = f
      for d in range(rank - 1, -1, -1):   # contract the last axis each time
        v = v @ ph[d]
      out[:, i] = v
    return sel, out
  N = int(np.prod(grid))
  sel = rng.choice(N, size=min(n_samples, N), replace=False)
  kidx = np.unravel_index(sel, grid)
  out = np.empty((T, sel.size), np.complex128)
  for i in range(sel.size):
    ph = np.zeros(M)
    for d in range(rank):
      ph += pts[:, d] * ks[d][kidx[d][i]]
    out[:, i] = src.reshape(T, M) @ np.exp(1j * sign * ph)
  return sel, out
