"""GPU tests at the operator level, re-stating the reference's own test cases
(tensorflow_nufft/python/ops/nufft_ops_test.py) on the torch-hosted mirror: values and gradients
against the dense NUDFT for every batch-broadcast pattern, points_range handling, range checking,
interp/spread invariants, and size-independent properties at BASELINE sizes."""
import itertools
import os

import numpy as np
import pytest

from tests import helpers as H

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _tfft():
  import tensorflow_nufft_b200 as tfft
  tfft.set_engine_defaults(num_threads_compat=os.cpu_count() or 1)
  return tfft


GRIDS = {1: [8], 2: [6, 8], 3: [4, 8, 6]}
BATCHES = [([], []), ([2, 4], []), ([4], [4]), ([2, 4], [2, 1]), ([2, 4], [1, 4]), ([4], [1, 4]), ([], [4])]


@pytest.mark.parametrize("rank", [1, 2, 3])
@pytest.mark.parametrize("source_batch,points_batch", BATCHES)
@pytest.mark.parametrize("transform_type", ["type_1", "type_2"])
@pytest.mark.parametrize("fft_direction", ["forward", "backward"])
@pytest.mark.parametrize("cdtype", [torch.complex64, torch.complex128])
def test_nufft_values_and_gradients(rank, source_batch, points_batch, transform_type, fft_direction, cdtype):
  """test_nufft (nufft_ops_test.py:87-221): forward result and gradients w.r.t. source and points
  against the NUDFT's autodiff, all batch patterns. The reference's GPU tolerance is 1e-1 (CPU 1e-3);
  we hold 2e-4 relative for complex64 and 1e-9 for complex128."""
  tfft = _tfft()
  grid = GRIDS[rank]
  M = int(np.prod(grid))
  rdtype = torch.float32 if cdtype == torch.complex64 else torch.float64
  g = torch.Generator().manual_seed(rank * 7 + len(source_batch) * 3 + len(points_batch))
  src_shape = source_batch + ([M] if transform_type == "type_1" else grid)
  src = torch.complex(torch.rand(src_shape, generator=g, dtype=rdtype) - 0.5,
                      torch.rand(src_shape, generator=g, dtype=rdtype) - 0.5).cuda().requires_grad_(True)
  pts = ((torch.rand(points_batch + [M, rank], generator=g, dtype=rdtype) - 0.5) * 2 * np.pi).cuda().requires_grad_(True)
  tol = 1e-6 if cdtype == torch.complex64 else 1e-12
  out = tfft.nufft(src, pts, grid_shape=grid, transform_type=transform_type, fft_direction=fft_direction, tol=tol)
  ref = tfft.nudft(src, pts, grid_shape=grid, transform_type=transform_type, fft_direction=fft_direction)
  assert out.shape == ref.shape
  scale = ref.abs().max().item()
  atol = (2e-4 if cdtype == torch.complex64 else 1e-9) * scale
  assert (out - ref).abs().max().item() <= atol
  # gradients with a complex upstream multiplier (nufft_ops_test.py:160-196)
  up = torch.complex(torch.rand(out.shape, generator=g, dtype=rdtype) - 0.5,
                     torch.rand(out.shape, generator=g, dtype=rdtype) - 0.5).cuda()
  loss = torch.real((out * up).sum())
  gs, gp = torch.autograd.grad(loss, [src, pts])
  loss_r = torch.real((ref * up).sum())
  gs_r, gp_r = torch.autograd.grad(loss_r, [src, pts])
  assert gs.shape == src.shape and gp.shape == pts.shape
  gtol = 5e-4 if cdtype == torch.complex64 else 1e-8
  assert (gs - gs_r).abs().max().item() <= gtol * max(gs_r.abs().max().item(), 1.0)
  assert (gp - gp_r).abs().max().item() <= gtol * max(gp_r.abs().max().item(), 1.0)


def test_nufft_with_options_same_result():
  """test_nufft_with_options (nufft_ops_test.py:65-84): max_batch_size=2 gives the same result."""
  tfft = _tfft()
  src = torch.from_numpy(H.random_complex((5, 2000), 1)).cuda()
  pts = torch.from_numpy(H.uniform_points(2000, 2, 2)).cuda()
  a = tfft.nufft(src, pts, grid_shape=(32, 24), transform_type="type_1")
  o = tfft.Options(max_batch_size=2)
  o.fftw.planning_rigor = tfft.FftwPlanningRigor.PATIENT
  b = tfft.nufft(src, pts, grid_shape=(32, 24), transform_type="type_1", options=o)
  assert H.rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6


def test_different_batch_ranks():
  """test_nufft_different_batch_ranks (nufft_ops_test.py:351-417)."""
  tfft = _tfft()
  grid = [6, 8]
  M = 48
  for sb, pb in (([2, 4], [1]), ([4], [2, 1])):
    for tt in ("type_1", "type_2"):
      src = torch.from_numpy(H.random_complex(tuple(sb) + ((M,) if tt == "type_1" else tuple(grid)), 3)).cuda()
      pts = torch.from_numpy(H.uniform_points(int(np.prod(pb)) * M, 2, 4).reshape(tuple(pb) + (M, 2))).cuda()
      out = tfft.nufft(src, pts, grid_shape=grid, transform_type=tt)
      ref = tfft.nudft(src, pts, grid_shape=grid, transform_type=tt)
      assert out.shape == ref.shape
      assert (out - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


def test_points_range_modes_agree():
  """test_nufft_points_range (nufft_ops_test.py:506-566): shifted points give the unshifted result."""
  tfft = _tfft()
  rng = np.random.default_rng(5)
  grid = (24, 20)
  M = 3000
  pts = H.uniform_points(M, 2, 6)
  src = torch.from_numpy(H.random_complex(grid, 7)).cuda()
  base = tfft.nufft(src, torch.from_numpy(pts).cuda(), options=tfft.Options(points_range="strict")).cpu().numpy()
  shift = (rng.integers(-1, 2, (M, 2)) * 2 * np.pi).astype(np.float32)
  ext = tfft.nufft(src, torch.from_numpy(pts + shift).cuda(), options=tfft.Options(points_range="extended")).cpu().numpy()
  assert H.rel_l2(ext, base) < 1e-4
  shift = (rng.integers(-5, 6, (M, 2)) * 2 * np.pi).astype(np.float32)
  inf = tfft.nufft(src, torch.from_numpy(pts + shift).cuda(), options=tfft.Options(points_range="infinite")).cpu().numpy()
  assert H.rel_l2(inf, base) < 1e-4


def test_check_points_range_raises():
  """test_nufft_check_points_range (nufft_ops_test.py:569-620)."""
  tfft = _tfft()
  src = torch.from_numpy(H.random_complex((16, 16), 1)).cuda()
  pts = H.uniform_points(500, 2, 2)
  o = tfft.Options(points_range="strict")
  o.debugging.check_points_range = True
  tfft.nufft(src, torch.from_numpy(pts).cuda(), options=o)          # in range: fine
  bad = pts.copy()
  bad[17, 1] = 4.0
  with pytest.raises(ValueError, match="outside expected range"):
    tfft.nufft(src, torch.from_numpy(bad).cuda(), options=o)
  o2 = tfft.Options(points_range="extended")
  o2.debugging.check_points_range = True
  tfft.nufft(src, torch.from_numpy(bad).cuda(), options=o2)         # inside [-3pi, 3pi]
  bad[3, 0] = -10.0
  with pytest.raises(ValueError, match="outside expected range"):
    tfft.nufft(src, torch.from_numpy(bad).cuda(), options=o2)


@pytest.mark.parametrize("grid", [(128, 128), (128, 128, 128)])
@pytest.mark.parametrize("cdtype", [np.complex64, np.complex128])
def test_interp_of_constant_is_constant(grid, cdtype):
  """test_interp (nufft_ops_test.py:224-252)."""
  tfft = _tfft()
  rd = np.float32 if cdtype == np.complex64 else np.float64
  pts = H.uniform_points(20000, len(grid), 8, rd)
  src = torch.full(grid, 1.0 + 0.5j, dtype=torch.complex64 if cdtype == np.complex64 else torch.complex128).cuda()
  out = tfft.interp(src, torch.from_numpy(pts).cuda(), tol=1e-4).cpu().numpy()
  assert np.allclose(out, 1.0 + 0.5j, rtol=1e-4, atol=1e-4)


def test_interp_batch_and_spread_batch():
  """test_interp_batch / test_spread_batch (nufft_ops_test.py:287-348)."""
  tfft = _tfft()
  grid = (64, 96)
  pts = torch.from_numpy(H.uniform_points(4 * 10000, 2, 9).reshape(4, 10000, 2)).cuda()
  src = torch.ones((4,) + grid, dtype=torch.complex64).cuda() * torch.arange(1, 5).reshape(4, 1, 1).cuda()
  out = tfft.interp(src, pts, tol=1e-4).cpu().numpy()
  for b in range(4):
    assert np.allclose(out[b], b + 1, rtol=1e-4, atol=1e-4)
  ones = torch.ones((4, 10000), dtype=torch.complex64).cuda()
  sp = tfft.spread(ones, pts, grid, tol=1e-4).cpu().numpy()
  assert sp.shape == (4,) + grid
  dens = 10000 / np.prod(grid)
  assert abs(sp.real.mean() / dens - 1.0) < 1e-3


@pytest.mark.parametrize("grid", [(64, 64), (64, 64, 64)])
def test_spread_of_ones(grid):
  """test_spread (nufft_ops_test.py:255-284): spreading ones on a regular point lattice gives a
  flat grid of mean 1."""
  tfft = _tfft()
  axes = [np.linspace(-np.pi, np.pi, n, endpoint=False) for n in grid]
  pts = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, len(grid)).astype(np.float32)
  src = torch.ones(pts.shape[0], dtype=torch.complex64).cuda()
  out = tfft.spread(src, torch.from_numpy(pts).cuda(), grid).cpu().numpy()
  assert out.real.min() > 0.0 and out.real.max() < 3.0
  assert abs(out.real.mean() - 1.0) < 1e-4


def test_interp_3d_many_points_is_deterministic():
  """test_interp_3d_many_points (nufft_ops_test.py:420-435): 3M points on 128^3, repeated; the
  reference notes non-deterministic behaviour, this engine must be bit-reproducible."""
  tfft = _tfft()
  grid = (128, 128, 128)
  pts = torch.from_numpy(H.uniform_points(3000000, 3, 10)).cuda()
  src = torch.ones(grid, dtype=torch.complex64).cuda()
  first = None
  for _ in range(3):
    out = tfft.interp(src, pts)
    assert torch.allclose(out, torch.ones_like(out), rtol=1e-3, atol=1e-3)
    if first is None:
      first = out.clone()
    assert torch.equal(out, first)


def test_point_set_reuse_is_content_based_and_needs_no_host_sync():
  """SURVEY 8f-4, in the C library: with opts.reuse_points the raw coordinates are fingerprinted on
  the device; an unchanged set (even at another address) skips the sort, ANY change of content
  -- including writes behind torch's version counter -- redoes it."""
  from tensorflow_nufft_b200 import _lib
  src = torch.from_numpy(H.random_complex((4, 32, 40), 21)).cuda()
  pts = torch.from_numpy(H.uniform_points(5000, 2, 22)).cuda()
  out = torch.empty((4, 5000), dtype=torch.complex64, device="cuda")
  st = torch.cuda.current_stream().cuda_stream
  plan = _lib.Plan(2, (40, 32), -1, 4, float(np.float32(1e-6)), _lib.COMPLEX64, device=0, reuse_points=1)

  def run(p):
    plan.set_points_interleaved(p.shape[0], p.data_ptr(), st)
    plan.execute(out.data_ptr(), src.data_ptr(), st)
    return out.clone()

  a = run(pts)
  b = run(pts)
  assert plan.reuse_stats() == {"skipped": 1, "full": 1}
  c = run(pts.clone())                      # other address, same content -> skipped
  assert plan.reuse_stats() == {"skipped": 2, "full": 1}
  assert torch.equal(a, b) and torch.equal(a, c)
  alias = torch.as_strided(pts, (1,), (1,), 7)   # write through a view created behind the API
  alias.fill_(0.123)
  d = run(pts)
  assert plan.reuse_stats() == {"skipped": 2, "full": 2}
  ref_plan = _lib.Plan(2, (40, 32), -1, 4, float(np.float32(1e-6)), _lib.COMPLEX64, device=0)
  ref_plan.set_points_interleaved(5000, pts.data_ptr(), st)
  want = torch.empty_like(out)
  ref_plan.execute(want.data_ptr(), src.data_ptr(), st)
  assert torch.equal(d, want) and not torch.equal(d, a)
  e = run(pts[:4000])                       # different M (a prefix of the same memory): never skipped
  assert plan.reuse_stats() == {"skipped": 2, "full": 3}
  e2 = run(pts[:4000])
  assert plan.reuse_stats() == {"skipped": 3, "full": 3} and torch.equal(e[:, :4000], e2[:, :4000])
  plan.close()
  ref_plan.close()


def test_point_set_reuse_through_the_operator_and_plan_cache():
  tfft = _tfft()
  from tensorflow_nufft_b200 import _lib
  tfft.clear_plan_cache()
  src = torch.from_numpy(H.random_complex((4, 32, 40), 21)).cuda()
  pts = torch.from_numpy(H.uniform_points(5000, 2, 22)).cuda()
  base = tfft.nufft(src, pts)
  h0 = _lib.plan_cache_stats()
  tfft.set_points_reuse(True)
  try:
    a = tfft.nufft(src, pts)
    b = tfft.nufft(src, pts)
    pts.mul_(0.5)
    c = tfft.nufft(src, pts)
  finally:
    tfft.set_points_reuse(False)
  want = tfft.nufft(src, pts)
  h1 = _lib.plan_cache_stats()
  assert torch.equal(a, base) and torch.equal(a, b) and torch.equal(c, want) and not torch.equal(a, c)
  # 4 calls after the first: one new plan (the reuse_points variant), three served from the cache
  assert h1["misses"] - h0["misses"] == 1 and h1["hits"] - h0["hits"] == 3


def test_conjugate_views_are_resolved_before_the_engine_sees_them():
  """torch.conj(x) is a lazy view with the parent's data_ptr."""
  tfft = _tfft()
  src = torch.from_numpy(H.random_complex((2, 24, 20), 3)).cuda()
  pts = torch.from_numpy(H.uniform_points(700, 2, 4)).cuda()
  lazy = torch.conj(src)
  assert lazy.is_conj()
  got = tfft.nufft(lazy, pts)
  want = tfft.nufft(lazy.clone().resolve_conj(), pts)
  assert torch.equal(got, want)
  assert not torch.equal(got, tfft.nufft(src, pts))
  c = torch.from_numpy(H.random_complex((2, 700), 5)).cuda()
  got1 = tfft.nufft(c.conj(), pts, grid_shape=(24, 20), transform_type="type_1")
  want1 = tfft.nufft(torch.conj(c).resolve_conj().clone(), pts, grid_shape=(24, 20), transform_type="type_1")
  # type 1 sums through global reductions: same values, bits depend on arrival order
  assert H.rel_l2(got1.cpu().numpy(), want1.cpu().numpy()) < 1e-6
  assert H.rel_l2(got1.cpu().numpy(), tfft.nufft(c, pts, grid_shape=(24, 20), transform_type="type_1").cpu().numpy()) > 0.1
  assert torch.equal(tfft.interp(torch.conj(src), pts), tfft.interp(torch.conj(src).resolve_conj(), pts))
  # host tensors take the same path
  assert torch.equal(tfft.nufft(torch.conj(src.cpu()), pts.cpu()).cuda(), want)


@pytest.mark.parametrize("stream_min", [0, 1 << 40])
def test_host_tensors_round_trip(stream_min, monkeypatch):
  tfft = _tfft()
  from tensorflow_nufft_b200.python.ops import nufft_ops
  monkeypatch.setattr(nufft_ops, "_HOST_STREAM_MIN_BYTES", stream_min)   # streamed and plain host paths
  src = torch.from_numpy(H.random_complex((3, 24, 20), 11))
  pts = torch.from_numpy(H.uniform_points(777, 2, 12))
  out_h = tfft.nufft(src, pts)
  out_d = tfft.nufft(src.cuda(), pts.cuda())
  assert not out_h.is_cuda and out_d.is_cuda
  assert torch.equal(out_h, out_d.cpu())


@pytest.mark.parametrize("ttype", ["type_1", "type_2"])
def test_host_pipelined_chunks_match_device_path(ttype, monkeypatch):
  """Host-resident batches are streamed (H2D / transform / D2H overlapped); the result must be
  bit-identical to the all-on-device call."""
  tfft = _tfft()
  from tensorflow_nufft_b200.python.ops import nufft_ops
  monkeypatch.setattr(nufft_ops, "_HOST_STREAM_MIN_BYTES", 0)
  grid = (40, 36)
  M = 5000
  T = 20   # two full chunks of 8 and a remainder of 4
  pts = torch.from_numpy(H.uniform_points(M, 2, 13))
  src = torch.from_numpy(H.random_complex((T, M) if ttype == "type_1" else (T,) + grid, 14))
  out_h = tfft.nufft(src.pin_memory(), pts, grid_shape=grid, transform_type=ttype)
  out_d = tfft.nufft(src.cuda(), pts.cuda(), grid_shape=grid, transform_type=ttype)
  assert out_h.shape == out_d.shape and not out_h.is_cuda
  if ttype == "type_2":
    assert torch.equal(out_h, out_d.cpu())
  else:  # type-1 grid sums go through REDG: order-dependent rounding between runs
    assert H.rel_l2(out_h.numpy(), out_d.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("ttype", ["type_1", "type_2"])
def test_host_pipelined_single_transform_chunks(ttype, monkeypatch):
  """Large transforms are streamed one per chunk (cfg4: a 256^3 grid per coil); forced here on a
  small 3D case by shrinking the chunk byte target."""
  tfft = _tfft()
  from tensorflow_nufft_b200.python.ops import nufft_ops
  monkeypatch.setattr(nufft_ops, "_HOST_CHUNK_BYTES", 1024)
  monkeypatch.setattr(nufft_ops, "_HOST_STREAM_MIN_BYTES", 0)
  grid = (12, 16, 10)
  M = 3000
  T = 3
  pts = torch.from_numpy(H.uniform_points(M, 3, 15))
  src = torch.from_numpy(H.random_complex((T, M) if ttype == "type_1" else (T,) + grid, 16))
  assert nufft_ops._host_chunk(src, T, M, grid) == 1
  out_h = tfft.nufft(src.pin_memory(), pts, grid_shape=grid, transform_type=ttype)
  out_d = tfft.nufft(src.cuda(), pts.cuda(), grid_shape=grid, transform_type=ttype)
  assert out_h.shape == out_d.shape and not out_h.is_cuda
  assert H.rel_l2(out_h.numpy(), out_d.cpu().numpy()) < 1e-6


def test_empty_inputs():
  tfft = _tfft()
  out = tfft.nufft(torch.zeros((0,), dtype=torch.complex64).cuda(), torch.zeros((0, 2), dtype=torch.float32).cuda(),
                   grid_shape=(8, 8), transform_type="type_1")
  assert out.shape == (8, 8) and out.abs().max().item() == 0.0
  out = tfft.nufft(torch.zeros((8, 8), dtype=torch.complex64).cuda(), torch.zeros((0, 2), dtype=torch.float32).cuda())
  assert out.shape == (0,)
  out = tfft.nufft(torch.zeros((0, 8, 8), dtype=torch.complex64).cuda(), torch.zeros((5, 2), dtype=torch.float32).cuda())
  assert out.shape == (0, 5)


# ---- size-independent properties at BASELINE sizes (the oracle cannot run these in seconds) ----

def _inner(a, b):
  return torch.sum(a.to(torch.complex128) * torch.conj(b.to(torch.complex128)))


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_full_size_adjointness_and_linearity(name):
  """<A x, y> = <x, A^H y> with A = type-2 forward, A^H = type-1 backward, and linearity, at the
  BASELINE grid/point counts (cfg2 with 2 coils). Rounding-level agreement."""
  tfft = _tfft()
  if name == "cfg2":
    grid, pts = (512, 512), H.spiral_points(32, 62500)
  else:
    grid, pts = (128, 128, 128), H.uniform_points(8000000, 3, 3)
  M = pts.shape[0]
  dp = torch.from_numpy(pts).cuda()
  x = torch.from_numpy(H.random_complex((2,) + grid, 21)).cuda()
  y = torch.from_numpy(H.random_complex((2, M), 22)).cuda()
  Ax = tfft.nufft(x, dp, transform_type="type_2", fft_direction="forward")
  Ahy = tfft.nufft(y, dp, grid_shape=grid, transform_type="type_1", fft_direction="backward")
  lhs, rhs = _inner(Ax, y), _inner(x, Ahy)
  assert abs(lhs - rhs) / abs(lhs) < 2e-5
  y2 = torch.from_numpy(H.random_complex((2, M), 23)).cuda()
  lin = tfft.nufft(2.0 * y + 3.0 * y2, dp, grid_shape=grid, transform_type="type_1", fft_direction="backward")
  Ahy2 = tfft.nufft(y2, dp, grid_shape=grid, transform_type="type_1", fft_direction="backward")
  resid = torch.linalg.norm((lin - (2.0 * Ahy + 3.0 * Ahy2)).reshape(-1)) / torch.linalg.norm(lin.reshape(-1))
  assert float(resid) < 5e-6


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4_half"])
def test_full_size_parity_against_reference_cpu_plan(name):
  """BASELINE configs at their full point counts and grids (fewer coils), engine vs the compiled
  reference CPU plan driven with the GPU plan's parameters: rel-L2 <= max(2 tol, 1e-6) = 2e-6."""
  tfft = _tfft()
  from oracle import port, ref
  if not ref.available() and not port.available():
    pytest.skip("no oracle library built")
  nthr = os.cpu_count() or 1
  if name == "cfg2":
    grid, pts, T, tt, direction = (512, 512), H.spiral_points(32, 62500), 2, 1, "backward"
  elif name == "cfg3":
    grid, pts, T, tt, direction = (128, 128, 128), H.uniform_points(8000000, 3, 3), 1, 1, "forward"
  else:  # cfg4 at half the grid size per dim (the 512^3 fine grid is too slow for the CPU oracle's FFT)
    grid, pts, T, tt, direction = (128, 128, 128), H.stack_of_stars_points(125, 125, 256), 2, 2, "forward"
  M = pts.shape[0]
  src = H.random_complex((T, M) if tt == 1 else (T,) + grid, 51)
  out = tfft.nufft(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid_shape=grid,
                   transform_type=f"type_{tt}", fft_direction=direction, tol=1e-6).cpu().numpy()
  plan_pts = np.ascontiguousarray(pts[:, ::-1].T)
  sign = -1 if direction == "forward" else 1
  if ref.available():
    rp = ref.RefPlan(tt, list(grid[::-1]), sign, T, 1e-6, np.complex64, mode="gpuparams", num_threads=nthr)
    rp.set_points(plan_pts)
    want = rp.execute(src.reshape(T, -1))
  else:  # the pinned plain-C restatement
    want = port.nufft(src.reshape(T, -1), plan_pts, list(grid[::-1]), tt, sign, 1e-6, np.complex64, num_threads=nthr)
  err = H.rel_l2(out.reshape(T, -1), want)
  assert err <= 2e-6, f"{name}: rel L2 {err:.3e}"


def test_full_size_cfg1_against_reference_and_sort_properties():
  """BASELINE config 1 (256^2, 100k radial points) against the compiled reference; plus bin-sort
  invariants read back through the parity hook: permutation, sortedness by bin, offsets = scan."""
  tfft = _tfft()
  from oracle import ref
  from tensorflow_nufft_b200 import _lib
  pts = H.radial_points(200, 500)
  src = H.random_complex((256, 256), 31)
  out = tfft.nufft(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), tol=1e-6).cpu().numpy()
  if ref.available():
    rp = ref.RefPlan(2, [256, 256], -1, 1, 1e-6, np.complex64, mode="gpuparams", num_threads=os.cpu_count() or 1)
    rp.set_points(np.ascontiguousarray(pts[:, ::-1].T))
    want = rp.execute(src.reshape(1, -1))[0]
    assert H.rel_l2(out, want) <= 2e-6
  plan = _lib.Plan(2, (256, 256), -1, 1, 1e-6, _lib.COMPLEX64)
  dp = torch.from_numpy(pts).cuda()
  plan.set_points_interleaved(pts.shape[0], dp.data_ptr(), None)
  torch.cuda.synchronize()
  M = pts.shape[0]
  got_idx, got_start, got_sizes = plan.sort_arrays()
  nb = got_sizes.size
  info = plan.info()
  folded = np.stack([H.fold_rescale_np(pts[:, 1 - d], info.fine_dims[d]) for d in range(2)])
  want_idx, want_start, want_sizes = H.binsort_np(folded, [info.fine_dims[0], info.fine_dims[1]],
                                                  [info.bin_dims[0], info.bin_dims[1]], 0)
  assert sorted(got_idx.tolist()) == list(range(M))                       # a permutation
  assert int(got_sizes.sum()) == M
  assert np.array_equal(np.cumsum(got_sizes)[:-1], got_start[1:]) and got_start[0] == 0   # offsets = exscan(sizes)
  assert np.array_equal(got_idx, want_idx)                                # stable, bit-exact
  assert np.array_equal(got_start, want_start) and np.array_equal(got_sizes, want_sizes)
  plan.close()
