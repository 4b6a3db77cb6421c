"""GPU parity tests: the CUDA engine (through the C ABI) against the reference's own CPU plan
(oracle/_ref/libref.so, mode "gpuparams"), numpy restatements of the bit-exact stages, and a
float64 NUDFT. Tolerances are BASELINE.json's gates: complex64 rel-L2 <= max(2 tol, 1e-6),
complex128 <= 2 tol; bin-sort and fold are bit-exact."""
import ctypes
import os

import numpy as np
import pytest

from tests import helpers as H

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NTHREADS = os.cpu_count() or 1


def _engine():
  import tensorflow_nufft_b200 as tfft
  tfft.set_engine_defaults(num_threads_compat=NTHREADS)
  return tfft


def _ref():
  from oracle import ref
  if not ref.available():
    pytest.skip("oracle/_ref/libref.so not built")
  return ref


@pytest.mark.parametrize("points_range", ["strict", "extended", "infinite"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fold_rescale_bit_exact(points_range, dtype):
  from tensorflow_nufft_b200 import _lib
  L = _lib.lib()
  rng = np.random.default_rng(7)
  span = {"strict": np.pi, "extended": 3 * np.pi, "infinite": 40.0}[points_range]
  x = rng.uniform(-span, span, 200000).astype(dtype)
  edge = np.array([-np.pi, np.pi, 0.0, np.nextafter(dtype(np.pi), dtype(0)), -np.nextafter(dtype(np.pi), dtype(0))], dtype)
  x[:edge.size] = edge
  for nf in (512, 540, 1024, 30):
    xin = torch.from_numpy(x).cuda()
    out = torch.empty_like(xin)
    rc = L.b200nufft_fold_rescale(int(dtype == np.float64), {"strict": 0, "extended": 1, "infinite": 2}[points_range],
                                  x.size, xin.data_ptr(), out.data_ptr(), nf, None)
    assert rc == 0
    torch.cuda.synchronize()
    want = H.fold_rescale_np(x, nf, points_range)
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32 if dtype == np.float32 else np.uint64),
                          want.view(np.uint32 if dtype == np.float32 else np.uint64))


def _gpu_binsort(folded, fine_dims, bin_dims, rounding):
  from tensorflow_nufft_b200 import _lib
  L = _lib.lib()
  rank = len(fine_dims)
  M = folded.shape[1]
  is_double = int(folded.dtype == np.float64)
  d = [torch.from_numpy(np.ascontiguousarray(folded[i])).cuda() for i in range(rank)]
  nb = 1
  for i in range(rank):
    nb *= ((fine_dims[i] + bin_dims[i] - 1) // bin_dims[i]) if rounding == 0 else (fine_dims[i] // bin_dims[i] + 1)
  idx = torch.empty(max(M, 1), dtype=torch.int32, device="cuda")
  bs = torch.empty(nb, dtype=torch.int32, device="cuda")
  bz = torch.empty(nb, dtype=torch.int32, device="cuda")
  fd = (ctypes.c_int * 3)(*(list(fine_dims) + [1] * (3 - rank)))
  bd = (ctypes.c_int * 3)(*(list(bin_dims) + [1] * (3 - rank)))
  ptr = [t.data_ptr() for t in d] + [None] * (3 - rank)
  rc = L.b200nufft_binsort(is_double, rank, M, ptr[0], ptr[1], ptr[2], fd, bd, rounding,
                           idx.data_ptr(), bs.data_ptr(), bz.data_ptr(), None)
  assert rc == 0
  torch.cuda.synchronize()
  return idx.cpu().numpy()[:M], bs.cpu().numpy(), bz.cpu().numpy()


def _pointsets_2d():
  return {
      "radial": H.radial_points(200, 500),
      "spiral": H.spiral_points(8, 20000),
      "uniform": H.uniform_points(150001, 2, 3),
      "tiny": H.uniform_points(5, 2, 4),
      "edges": np.array([[-np.pi, np.pi], [np.pi, -np.pi], [0, 0], [np.pi, np.pi]], np.float32),
  }


@pytest.mark.parametrize("name", ["radial", "spiral", "uniform", "tiny", "edges"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_binsort_2d_bit_exact_gpu_rule(name, dtype):
  pts = _pointsets_2d()[name].astype(dtype)
  fine = [512, 540]
  folded = np.stack([H.fold_rescale_np(pts[:, 1 - d], fine[d]) for d in range(2)])
  got = _gpu_binsort(folded, fine, [32, 32], 0)
  want = H.binsort_np(folded, fine, [32, 32], 0)
  for g, w, what in zip(got, want, ("idx_nupts", "bin_start_pts", "bin_sizes")):
    assert np.array_equal(g, w), what


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_binsort_3d_bit_exact_gpu_rule(dtype):
  pts = H.uniform_points(300000, 3, 5, dtype)
  pts[:1000] = H.stack_of_stars_points(10, 10, 10, dtype)
  fine = [256, 270, 64]
  folded = np.stack([H.fold_rescale_np(pts[:, 2 - d], fine[d]) for d in range(3)])
  got = _gpu_binsort(folded, fine, [16, 16, 2], 0)
  want = H.binsort_np(folded, fine, [16, 16, 2], 0)
  for g, w, what in zip(got, want, ("idx_nupts", "bin_start_pts", "bin_sizes")):
    assert np.array_equal(g, w), what


def test_binsort_1d_and_empty():
  x = H.uniform_points(70000, 1, 6)
  folded = np.stack([H.fold_rescale_np(x[:, 0], 4096)])
  got = _gpu_binsort(folded, [4096], [1024], 0)
  want = H.binsort_np(folded, [4096], [1024], 0)
  for g, w in zip(got, want):
    assert np.array_equal(g, w)
  got = _gpu_binsort(np.zeros((2, 0), np.float32), [64, 64], [32, 32], 0)
  assert got[0].size == 0 and np.all(got[2] == 0) and np.all(got[1] == 0)


@pytest.mark.parametrize("rank", [2, 3])
def test_binsort_matches_compiled_reference_cpu_sort(rank):
  """Engine kernel with the CPU geometry (16,4,4; nf/bin+1 boxes; truncation) against the
  compiled binsort_singlethread/_multithread (deterministic real reference code)."""
  ref = _ref()
  n = 64 if rank == 2 else 24
  pts = H.uniform_points(200000, rank, 11)
  plan_pts = np.ascontiguousarray(pts[:, ::-1].T)
  for nthr in (1, 4):
    rp = ref.RefPlan(2, [n] * rank, -1, 1, 1e-6, np.complex64, mode="gpuparams", num_threads=nthr)
    rp.set_points(plan_pts)
    want, did = rp.sort_indices()
    assert did
    got = _gpu_binsort(rp.folded, rp.fine_dims, [16, 4, 4][:rank], 1)
    assert np.array_equal(got[0], want)


CASES = [
    # (rank, grid, M, T, type, dtype, tol, points kind)
    (2, (64, 48), 5000, 1, 2, np.complex64, 1e-6, "uniform"),
    (2, (64, 48), 5000, 3, 1, np.complex64, 1e-6, "uniform"),
    (2, (256, 256), 100000, 1, 2, np.complex64, 1e-6, "radial"),
    (2, (256, 256), 100000, 1, 1, np.complex64, 1e-6, "radial"),
    (2, (128, 128), 40000, 2, 1, np.complex64, 1e-4, "spiral"),
    (2, (128, 128), 40000, 2, 2, np.complex64, 1e-3, "spiral"),
    (3, (32, 24, 40), 30000, 2, 1, np.complex64, 1e-6, "uniform"),
    (3, (32, 24, 40), 30000, 2, 2, np.complex64, 1e-6, "uniform"),
    (3, (64, 64, 64), 200000, 1, 1, np.complex64, 1e-6, "sos"),
    (3, (64, 64, 64), 200000, 1, 2, np.complex64, 1e-6, "sos"),
    (1, (256,), 3000, 2, 1, np.complex64, 1e-6, "uniform"),
    (1, (256,), 3000, 2, 2, np.complex64, 1e-6, "uniform"),
    (2, (64, 64), 20000, 1, 2, np.complex128, 1e-12, "uniform"),
    (2, (64, 64), 20000, 2, 1, np.complex128, 1e-12, "uniform"),
    (3, (16, 20, 24), 5000, 1, 2, np.complex128, 1e-9, "uniform"),
    (3, (16, 20, 24), 5000, 1, 1, np.complex128, 1e-9, "uniform"),
    (2, (64, 64), 20000, 1, 1, np.complex64, 1e-7, "uniform"),   # ns = 8: float row-lane kernels
    (2, (64, 64), 20000, 3, 2, np.complex64, 1e-7, "uniform"),
    (2, (90, 50), 20000, 2, 1, np.complex64, 5e-8, "edges"),     # ns = 9 (tol clipped at 6e-8)
    (2, (90, 50), 20000, 2, 2, np.complex64, 5e-8, "edges"),
    (3, (16, 20, 12), 5000, 1, 2, np.complex64, 1e-7, "uniform"),  # 3D ns = 8: generic kernels
    # fine grids that are not multiples of the bin size (partial last bins, wrap inside a tile)
    (2, (135, 77), 30000, 3, 1, np.complex64, 1e-6, "uniform"),
    (2, (135, 77), 30000, 3, 2, np.complex64, 1e-6, "uniform"),
    (3, (25, 27, 15), 20000, 2, 1, np.complex64, 1e-6, "uniform"),
    (3, (25, 27, 15), 20000, 2, 2, np.complex64, 1e-6, "uniform"),
    (2, (9, 7), 500, 5, 1, np.complex64, 1e-5, "uniform"),       # fine grid smaller than one tile
    (2, (9, 7), 500, 5, 2, np.complex64, 1e-5, "uniform"),
    (3, (5, 6, 7), 400, 1, 1, np.complex64, 1e-4, "uniform"),
    (3, (5, 6, 7), 400, 1, 2, np.complex64, 1e-4, "uniform"),
    # coil counts that select the 8 / 4 / 2 / 1 coils-per-CTA variants of the 2D spreader and
    # interpolator (8 | T, 4 | T, 2 | T, odd)
    (2, (96, 72), 30000, 4, 1, np.complex64, 1e-6, "spiral"),
    (2, (96, 72), 30000, 4, 2, np.complex64, 1e-6, "spiral"),
    (2, (96, 72), 30000, 6, 1, np.complex64, 1e-6, "uniform"),
    (2, (96, 72), 30000, 6, 2, np.complex64, 1e-6, "uniform"),
    (2, (96, 72), 30000, 8, 1, np.complex64, 1e-6, "radial"),
    (2, (96, 72), 30000, 8, 2, np.complex64, 1e-6, "radial"),
    (2, (80, 112), 20000, 16, 1, np.complex64, 1e-5, "uniform"),
    (2, (80, 112), 20000, 16, 2, np.complex64, 1e-5, "uniform"),
    (2, (80, 112), 20000, 7, 1, np.complex64, 1e-2, "uniform"),   # ns = 4 through the window kernels
    (2, (80, 112), 20000, 7, 2, np.complex64, 1e-2, "uniform"),
    (3, (24, 40, 32), 30000, 4, 2, np.complex64, 1e-5, "sos"),
    (3, (24, 40, 32), 30000, 4, 1, np.complex64, 1e-5, "sos"),
    # more transforms than one batch (32): full batches + a remainder batch with an odd coil count
    (2, (40, 48), 6000, 37, 1, np.complex64, 1e-6, "uniform"),
    (2, (40, 48), 6000, 37, 2, np.complex64, 1e-6, "uniform"),
    (3, (16, 12, 20), 5000, 34, 1, np.complex64, 1e-5, "uniform"),
    (3, (16, 12, 20), 5000, 34, 2, np.complex64, 1e-5, "uniform"),
    # points exactly on the fold boundaries (+-pi, 0) mixed into a uniform set, every tile kernel
    (2, (64, 48), 4000, 2, 1, np.complex64, 1e-6, "edges"),
    (2, (64, 48), 4000, 2, 2, np.complex64, 1e-6, "edges"),
    (3, (20, 24, 16), 4000, 1, 1, np.complex64, 1e-6, "edges"),
    (3, (20, 24, 16), 4000, 1, 2, np.complex64, 1e-6, "edges"),
    (2, (64, 48), 4000, 2, 1, np.complex128, 1e-12, "edges"),
    (2, (64, 48), 4000, 2, 2, np.complex128, 1e-12, "edges"),
    # complex128 row-lane kernels: every (record width, lanes per point) instantiation
    (2, (48, 40), 10000, 3, 1, np.complex128, 1e-5, "uniform"),    # ns = 6:  PX 8,  8 lanes
    (2, (48, 40), 10000, 3, 2, np.complex128, 1e-5, "uniform"),
    (2, (48, 40), 10000, 1, 1, np.complex128, 1e-7, "spiral"),     # ns = 8:  PX 12, 8 lanes
    (2, (48, 40), 10000, 1, 2, np.complex128, 1e-7, "spiral"),
    (2, (48, 40), 10000, 2, 1, np.complex128, 1e-9, "radial"),     # ns = 10/11: PX 12, 16 lanes
    (2, (48, 40), 10000, 2, 2, np.complex128, 1e-9, "radial"),
    (2, (135, 77), 20000, 2, 1, np.complex128, 1e-13, "uniform"),  # ns = 15: PX 16, 16 lanes, odd grid
    (2, (135, 77), 20000, 2, 2, np.complex128, 1e-13, "uniform"),
    (2, (9, 7), 300, 2, 1, np.complex128, 1e-12, "uniform"),       # fine grid smaller than one tile
    (2, (9, 7), 300, 2, 2, np.complex128, 1e-12, "uniform"),
]


def _points(kind, M, rank, rdtype, seed):
  if kind == "uniform":
    return H.uniform_points(M, rank, seed, rdtype)
  if kind == "radial":
    return H.radial_points(max(1, M // 500), min(M, 500), rdtype) if M < 100000 else H.radial_points(200, M // 200, rdtype)
  if kind == "spiral":
    return H.spiral_points(8, M // 8, 24, rdtype)
  if kind == "edges":
    pts = H.uniform_points(M, rank, seed, rdtype)
    pi = rdtype(np.pi)
    special = [pi, -pi, rdtype(0), np.nextafter(pi, rdtype(0)), -np.nextafter(pi, rdtype(0))]
    k = 0
    for a in special:
      for b in special:
        pts[k, :] = a
        pts[k, -1] = b
        k += 1
    return pts
  if kind == "sos":
    return H.stack_of_stars_points(20, 100, max(1, M // 2000), rdtype)
  raise ValueError(kind)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-t{c[4]}-{np.dtype(c[5]).name}-{c[6]:g}-{c[7]}-M{c[2]}-T{c[3]}")
@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_nufft_matches_reference_cpu_plan(case, direction):
  rank, grid, M, T, ttype, cdtype, tol, kind = case
  tfft = _engine()
  ref = _ref()
  rdtype = np.float32 if cdtype == np.complex64 else np.float64
  pts = _points(kind, M, rank, rdtype, 100 + rank)
  M = pts.shape[0]
  N = int(np.prod(grid))
  src = H.random_complex((T, M) if ttype == 1 else (T,) + tuple(grid), 200 + rank, cdtype)
  out = tfft.nufft(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid_shape=grid,
                   transform_type=f"type_{ttype}", fft_direction=direction, tol=tol)
  out = out.cpu().numpy()
  assert out.shape == ((T,) + tuple(grid) if ttype == 1 else (T, M))
  sign = -1 if direction == "forward" else 1
  rp = ref.RefPlan(ttype, list(grid[::-1]), sign, T, tol, cdtype, mode="gpuparams", num_threads=NTHREADS)
  rp.set_points(np.ascontiguousarray(pts[:, ::-1].T))
  want = rp.execute(src.reshape(T, -1))
  err = H.rel_l2(out.reshape(T, -1), want)
  gate = max(2 * tol, 1e-6) if cdtype == np.complex64 else 2 * float(np.float32(tol))
  assert err <= gate, f"rel L2 {err:.3e} > gate {gate:.1e}"


@pytest.mark.parametrize("rank,grid,M", [(1, (16,), 40), (2, (6, 8), 48), (3, (4, 8, 6), 192)])
@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("cdtype,tol,gate", [(np.complex64, 1e-6, 2e-5), (np.complex128, 1e-12, 2e-11)])
def test_small_cases_match_direct_nudft(rank, grid, M, ttype, cdtype, tol, gate):
  from oracle import nudft
  tfft = _engine()
  rdtype = np.float32 if cdtype == np.complex64 else np.float64
  pts = H.uniform_points(M, rank, 31, rdtype)
  src = H.random_complex((M,) if ttype == 1 else tuple(grid), 32, cdtype)
  out = tfft.nufft(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid_shape=grid,
                   transform_type=f"type_{ttype}", fft_direction="forward", tol=tol).cpu().numpy()
  truth = nudft.nudft_plan_layout(src.reshape(1, -1), np.ascontiguousarray(pts[:, ::-1].T), list(grid[::-1]), ttype, -1)
  assert H.rel_l2(out.reshape(-1), truth[0]) <= gate


@pytest.mark.parametrize("rank", [2, 3])
def test_tile_and_global_kernels_agree(rank):
  """The shared-memory tile kernels and the point-driven global kernels are two implementations
  of the same sums; they must agree to float rounding."""
  from tensorflow_nufft_b200.python.ops import nufft_ops
  grid = (96, 80) if rank == 2 else (32, 40, 24)
  M = 60000
  pts = H.uniform_points(M, rank, 41)
  res = {}
  for ttype in (1, 2):
    src = H.random_complex((2, M) if ttype == 1 else (2,) + grid, 42)
    for meth in (1, 2, 3, 4):
      out = nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid, f"type_{ttype}",
                              "forward", 1e-6, None, "nufft",
                              engine_kwargs={"spread_method": meth, "interp_method": meth})
      res[(ttype, meth)] = out.cpu().numpy()
    assert H.rel_l2(res[(ttype, 2)], res[(ttype, 1)]) < 5e-7
    assert H.rel_l2(res[(ttype, 3)], res[(ttype, 1)]) < 5e-7
    assert H.rel_l2(res[(ttype, 4)], res[(ttype, 1)]) < 5e-7


@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("grid", [(24, 20, 32), (9, 12, 7)])
def test_pruned_3d_fft_matches_single_plan(ttype, grid):
  """3D plans run the fine-grid FFT as 2D cuFFT transforms on the populated z-planes plus a strided
  1D cuFFT along z; the result must equal the single 3D cuFFT plan to rounding."""
  from tensorflow_nufft_b200.python.ops import nufft_ops
  M = 20000
  pts = H.uniform_points(M, 3, 51)
  src = H.random_complex((2, M) if ttype == 1 else (2,) + grid, 52)
  res = []
  for full in (0, 1):
    out = nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid, f"type_{ttype}",
                            "backward", 1e-6, None, "nufft", engine_kwargs={"full_fft": full})
    res.append(out.cpu().numpy())
  assert H.rel_l2(res[0], res[1]) < 5e-7


@pytest.mark.parametrize("rank", [2, 3])
def test_window_sorted_plan_keeps_reference_bins(rank):
  """Type-1 plans refine the sort key to (bin, stencil window). The bin offsets and sizes must
  still be the reference's bit-exactly, and every bin must hold exactly the reference's point set
  (the order inside a bin is by window, then by point index)."""
  from tensorflow_nufft_b200 import _lib
  grid = (128, 96) if rank == 2 else (32, 48, 40)
  pts = H.uniform_points(120000, rank, 77)
  pts[:2000] = H.radial_points(20, 100)[:, :rank] if rank == 2 else pts[:2000]
  plan = _lib.Plan(1, tuple(reversed(grid)), -1, 1, 1e-6, _lib.COMPLEX64)
  dp = torch.from_numpy(pts).cuda()
  plan.set_points_interleaved(pts.shape[0], dp.data_ptr(), None)
  torch.cuda.synchronize()
  idx, start, sizes = plan.sort_arrays()
  info = plan.info()
  fine = [info.fine_dims[d] for d in range(rank)]
  bins = [info.bin_dims[d] for d in range(rank)]
  folded = np.stack([H.fold_rescale_np(pts[:, rank - 1 - d], fine[d]) for d in range(rank)])
  widx, wstart, wsizes = H.binsort_np(folded, fine, bins, 0)
  assert np.array_equal(start, wstart) and np.array_equal(sizes, wsizes)
  assert sorted(idx.tolist()) == list(range(pts.shape[0]))
  ends = wstart + wsizes
  for b in np.nonzero(wsizes)[0][:2000]:
    assert np.array_equal(np.sort(idx[wstart[b]:ends[b]]), widx[wstart[b]:ends[b]])
  plan.close()
