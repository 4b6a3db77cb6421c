"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol the
header declares, the host parameter maths matches the reference's golden tables, the operator
mirror reproduces the reference's validation errors, Options behave like tfft.Options."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def L():
  from tensorflow_nufft_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip("libb200nufft.so not built (run __graft_entry__.build())")
  return _lib.lib()


def test_cabi_exports_every_declared_symbol(L):
  from tensorflow_nufft_b200 import _lib
  hdr = open(os.path.join(ROOT, "include", "b200nufft.h")).read()
  declared = sorted(set(re.findall(r"\b(b200nufft_[a-z_0-9]+)\s*\(", hdr)))
  assert len(declared) >= 20
  raw = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared:
    assert hasattr(raw, name), f"{name} declared in include/b200nufft.h but not exported"
  bound = {s[0] for s in _lib.SIGNATURES}
  assert set(declared) == bound, set(declared) ^ bound
  assert b"sm_100a" in L.b200nufft_version()


def test_no_cuda_device_fails_loudly(L):
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from tensorflow_nufft_b200 import _lib
  with pytest.raises(_lib.NufftError, match="no CUDA device|CUDA"):
    _lib.Plan(2, (16, 16), -1, 1, 1e-6, _lib.COMPLEX64)
  import tensorflow_nufft_b200 as tfft
  src = torch.zeros((16, 16), dtype=torch.complex64)
  pts = torch.zeros((10, 2), dtype=torch.float32)
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    tfft.nufft(src, pts)


def test_host_kernel_width_and_smooth_int_match_reference(L):
  p = np.load(os.path.join(GOLD, "params.npz"))
  for is_double, tol, ns, beta, c, nf0, nf1 in p["ptab"]:
    assert L.b200nufft_host_kernel_width(int(is_double), float(np.float32(tol)), 2.0) == int(ns)
  for n, want in p["smooth"]:
    assert L.b200nufft_host_next_smooth_int(int(n)) == int(want)


def test_host_fseries_float_bit_exact_vs_reference(L):
  p = np.load(os.path.join(GOLD, "params.npz"))
  n = 0
  for key in p.files:
    if not key.startswith("fser_"):
      continue
    _, dname, nf, ns, nt = key.split("_")
    dt = np.dtype(dname).type
    out = np.empty(int(nf) // 2 + 1, dt)
    assert L.b200nufft_host_fseries(int(dt == np.float64), int(nf), int(ns), 0, int(nt), out.ctypes.data) == 0
    want = p[key]
    if dt == np.float32:
      assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), key
    else:
      used = slice(0, int(nf) // 4 + 1)
      assert np.max(np.abs(out[used] - want[used]) / np.abs(want[used])) < 2e-12, key
    # accurate (double) mode agrees with the double reference on the used modes
    acc = np.empty_like(out)
    assert L.b200nufft_host_fseries(int(dt == np.float64), int(nf), int(ns), 1, 1, acc.ctypes.data) == 0
    if dt == np.float64:
      used = slice(0, int(nf) // 4 + 1)
      assert np.max(np.abs(acc[used] - want[used]) / np.abs(want[used])) < 2e-12, key
    n += 1
  assert n > 20


def test_host_scale_factor_matches_reference(L):
  p = np.load(os.path.join(GOLD, "params.npz"))
  for is_double, rank, ns, want in p["scale"]:
    got = L.b200nufft_host_scale_factor(int(is_double), int(rank), int(ns))
    assert got == want or abs(got - want) / abs(want) < 1e-15


def test_host_gauss_legendre(L):
  for n in (8, 46):
    x = np.empty(n)
    w = np.empty(n)
    assert L.b200nufft_host_gauss_legendre(n, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    xr, wr = np.polynomial.legendre.leggauss(n)
    assert np.max(np.abs(x - xr)) < 5e-15 and np.max(np.abs(w - wr) / wr) < 5e-13


# ---- operator mirror: validation happens before any device work (nufft_kernels.cc:58-130) ----

def _t(shape, dtype):
  import torch
  return torch.zeros(shape, dtype=dtype)


def test_type_1_no_grid_shape_raises():
  import torch
  import tensorflow_nufft_b200 as tfft
  with pytest.raises(ValueError, match="grid_shape must be provided for type-1 transforms"):
    tfft.nufft(_t((4, 48), torch.complex64), _t((48, 2), torch.float32), transform_type="type_1")


def test_type_1_invalid_grid_shape_raises():
  import torch
  import tensorflow_nufft_b200 as tfft
  with pytest.raises(ValueError, match="grid_shape must have length 2"):
    tfft.nufft(_t((48,), torch.complex64), _t((48, 2), torch.float32), grid_shape=(6, 8, 2), transform_type="type_1")


def test_type_1_incompatible_source_points_raises():
  import torch
  import tensorflow_nufft_b200 as tfft
  with pytest.raises(ValueError, match="must have equal samples dimensions"):
    tfft.nufft(_t((40,), torch.complex64), _t((48, 2), torch.float32), grid_shape=(6, 8), transform_type="type_1")


def test_other_validation_messages():
  import torch
  import tensorflow_nufft_b200 as tfft
  with pytest.raises(ValueError, match="Dimension must be 1, 2 or 3"):
    tfft.nufft(_t((4, 4, 4, 4), torch.complex64), _t((10, 4), torch.float32))
  with pytest.raises(ValueError, match="must have type"):
    tfft.nufft(_t((8, 8), torch.complex64), _t((10, 2), torch.float64))
  with pytest.raises(ValueError, match="rank of at least 2"):
    tfft.nufft(_t((8, 8), torch.complex64), _t((10,), torch.float32))
  with pytest.raises(ValueError, match="Incompatible shapes"):
    tfft.nufft(_t((3, 8, 8), torch.complex64), _t((2, 10, 2), torch.float32))
  with pytest.raises(ValueError, match="transform_type"):
    tfft.nufft(_t((8, 8), torch.complex64), _t((10, 2), torch.float32), transform_type="type_3")


def test_options_mirror_reference_defaults_and_validation():
  import tensorflow_nufft_b200 as tfft
  o = tfft.Options()
  assert o.points_range == tfft.PointsRange.EXTENDED
  assert o.max_batch_size is None
  assert o.debugging.check_points_range is False
  assert o.fftw.planning_rigor == tfft.FftwPlanningRigor.AUTO
  assert o.to_engine_kwargs() == {"points_range": 1, "check_points_range": 0, "max_batch_size": 0}
  o.max_batch_size = 2
  o.debugging.check_points_range = True
  o.points_range = tfft.PointsRange.INFINITE
  assert o.to_engine_kwargs() == {"points_range": 2, "check_points_range": 1, "max_batch_size": 2}
  with pytest.raises(ValueError):
    tfft.Options(max_batch_size=-1).to_engine_kwargs()
  with pytest.raises(ValueError):
    tfft.Options(points_range=7)
  assert tfft.Options(points_range="strict").points_range == tfft.PointsRange.STRICT


def test_nudft_reference_implementation_matches_float64_sum():
  """`nudft` (reference nufft_ops.py:235-321) restated in torch, against the numpy oracle."""
  import torch
  import tensorflow_nufft_b200 as tfft
  from oracle import nudft as onudft
  pts = H.uniform_points(30, 2, 1, np.float64)
  src = H.random_complex((6, 8), 2, np.complex128)
  got = tfft.nudft(torch.from_numpy(src), torch.from_numpy(pts), transform_type="type_2").numpy()
  want = onudft.nudft_plan_layout(src.reshape(1, -1), np.ascontiguousarray(pts[:, ::-1].T), [8, 6], 2, -1)[0]
  assert H.rel_l2(got, want) < 1e-13
  c = H.random_complex((30,), 3, np.complex128)
  got = tfft.nudft(torch.from_numpy(c), torch.from_numpy(pts), grid_shape=(6, 8), transform_type="type_1",
                   fft_direction="backward").numpy()
  want = onudft.nudft_plan_layout(c.reshape(1, -1), np.ascontiguousarray(pts[:, ::-1].T), [8, 6], 1, +1)[0]
  assert H.rel_l2(got.reshape(-1), want) < 1e-13


def test_shard_bounds_cover_and_balance():
  from tensorflow_nufft_b200 import sharding
  for T in (1, 2, 7, 16, 32, 33):
    for world in (1, 2, 3, 4, 8):
      spans = [sharding.shard_bounds(T, world, r) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == T
      for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0]
      sizes = [e - b for b, e in spans]
      assert max(sizes) - min(sizes) <= 1


def test_host_stream_chunking_policy():
  """Host-resident batches are streamed in chunks of 8 transforms when a transform is small
  (cfg2: 16 MB of strengths) and of one transform when it is already >= the 128 MB chunk target
  (cfg4: a 256^3 grid)."""
  import torch
  from tensorflow_nufft_b200.python.ops import nufft_ops
  small = torch.zeros((2, 4), dtype=torch.complex64)
  assert nufft_ops._host_chunk(small, 32, 2_000_000, (512, 512)) == 8          # cfg2
  assert nufft_ops._host_chunk(small, 16, 4_000_000, (256, 256, 256)) == 1     # cfg4
  assert nufft_ops._host_chunk(small, 4, 8_000_000, (128, 128, 128)) == 2      # 64 MB per transform
  big = torch.zeros((2, 4), dtype=torch.complex128)
  assert nufft_ops._host_chunk(big, 32, 2_000_000, (512, 512)) == 4            # 32 MB per transform
  assert nufft_ops._host_chunk(small, 3, 10, (8, 8)) == 8


def test_point_set_reuse_is_opt_in_and_lives_in_the_c_library(L):
  """The unchanged-points shortcut is a b200nufft_opts field (device-side fingerprint), off by
  default; the Python mirror only forwards the switch."""
  from tensorflow_nufft_b200 import _lib
  from tensorflow_nufft_b200.python.ops import nufft_ops
  assert _lib.make_opts().reuse_points == 0
  assert _lib.make_opts(reuse_points=1).reuse_points == 1
  assert os.environ.get("B200NUFFT_REUSE_POINTS", "0") != "0" or not nufft_ops._REUSE_POINTS
  # the plan cache is the library's, not Python's
  stats = _lib.plan_cache_stats()
  assert set(stats) == {"hits", "misses", "idle"}
  _lib.plan_cache_clear()
  assert _lib.plan_cache_stats()["idle"] == 0
  a, f = _lib.alloc_counts()
  assert a >= f >= 0


def test_tf_glue_type_checks_against_reference_headers():
  """The OpKernel-side glue is parsed and type-checked (g++ -fsyntax-only) against the reference's
  nufft_plan.h and stand-in TF headers; skipped where the reference tree is absent (GPU box)."""
  if not os.path.isdir("/root/reference/tensorflow_nufft"):
    pytest.skip("reference tree not present")
  import __graft_entry__ as g
  g.check_tf_glue()
