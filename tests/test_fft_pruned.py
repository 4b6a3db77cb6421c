"""The engine's own pruned FFT passes (csrc/fft_pruned.cuh).

CPU: tests/fft_pruned_host.cc runs the kernels' own load / butterfly / store functions thread by
thread (the header is plain C++ outside the __global__ wrappers) and is compared with numpy.fft on
zero-padded / cropped grids, factors included. GPU: the plan with the own passes against the same
plan on cuFFT + amplify / deconvolve kernels.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
  out = str(tmp_path_factory.mktemp("fft") / "fft_pruned_host.so")
  subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include",
                  "-o", out, os.path.join(ROOT, "tests", "fft_pruned_host.cc")], check=True)
  return ctypes.CDLL(out)


def _host_case(L, ttype, n, N, sign, ntr=1, seed=0):
  """n, N: fine sizes / modes, x first. Returns the relative L2 error against numpy (float64)."""
  rank = len(n)
  rng = np.random.default_rng(seed)
  n_ = (ctypes.c_int * 3)(*n, *([1] * (3 - rank)))
  N_ = (ctypes.c_int * 3)(*N, *([1] * (3 - rank)))
  fac = [rng.uniform(0.5, 2.0, nd // 2 + 1).astype(np.float32) for nd in n] + [np.ones(1, np.float32)] * (3 - rank)
  fshape, wshape = (ntr,) + tuple(N[::-1]), (ntr,) + tuple(n[::-1])
  A = np.ones(tuple(N[::-1]), np.float64)
  for d in range(rank):
    k = np.abs(np.arange(N[d]) - N[d] // 2)
    sh = [1] * rank
    sh[rank - 1 - d] = N[d]
    A = A * fac[d][k].astype(np.float64).reshape(sh)

  def widx(d):
    k = np.arange(N[d]) - N[d] // 2
    return np.where(k >= 0, k, n[d] + k)

  ix = (slice(None),) + np.ix_(*[widx(d) for d in range(rank - 1, -1, -1)])
  axes = tuple(range(1, rank + 1))
  fftn = (lambda a: np.fft.fftn(a, axes=axes)) if sign < 0 else (lambda a: np.fft.ifftn(a, axes=axes) * np.prod(n))
  ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
  if ttype == 2:
    f = (rng.standard_normal(fshape) + 1j * rng.standard_normal(fshape)).astype(np.complex64)
    fw = np.full(wshape, np.nan + 1j * np.nan, np.complex64)   # the passes must not read what they did not write
    assert L.fft_pruned_host(2, rank, n_, N_, sign, ntr, ptr(fw), ptr(f), *[ptr(x) for x in fac]) == 0
    big = np.zeros(wshape, np.complex128)
    big[ix] = f / A
    return H.rel_l2(fw, fftn(big))
  fw = (rng.standard_normal(wshape) + 1j * rng.standard_normal(wshape)).astype(np.complex64)
  want = fftn(fw.astype(np.complex128))[ix] / A
  f = np.full(fshape, np.nan + 1j * np.nan, np.complex64)
  assert L.fft_pruned_host(1, rank, n_, N_, sign, ntr, ptr(fw), ptr(f), *[ptr(x) for x in fac]) == 0
  return H.rel_l2(f, want)


HOST_CASES = [
    (2, (64, 64), (32, 32), -1, 1), (1, (64, 64), (32, 32), 1, 2), (2, (128, 64), (64, 30), 1, 1), (1, (256, 128), (96, 51), -1, 1),
    (2, (64, 64, 64), (32, 32, 32), -1, 2), (1, (64, 64, 64), (32, 32, 32), 1, 1), (2, (128, 64, 256), (64, 31, 100), 1, 1),
    (1, (64, 128, 64), (32, 64, 17), -1, 1), (2, (512, 64), (256, 32), -1, 1), (1, (1024, 64), (512, 20), 1, 1),
    (2, (64, 1024), (32, 500), -1, 1), (1, (64, 512), (64, 256), 1, 1), (2, (256, 256), (128, 128), 1, 1),
]


@pytest.mark.parametrize("ttype,n,N,sign,ntr", HOST_CASES)
def test_pruned_passes_match_numpy_on_cpu(host_lib, ttype, n, N, sign, ntr):
  """Every pass length 64 .. 1024 on both kinds of axis (strided bundles / rows), both transform types
  and signs, odd mode counts on the slow axes, batches: float-rounding agreement with numpy."""
  assert _host_case(host_lib, ttype, n, N, sign, ntr) < 4e-7


def test_pruned_passes_reject_ineligible_sizes(host_lib):
  n_ = (ctypes.c_int * 3)(96, 64, 1)
  N_ = (ctypes.c_int * 3)(48, 32, 1)
  z = np.zeros(8, np.float32)
  p = z.ctypes.data_as(ctypes.c_void_p)
  assert host_lib.fft_pruned_host(2, 2, n_, N_, -1, 1, p, p, p, p, p) == 1   # 96 is not a power of two
  n_ = (ctypes.c_int * 3)(128, 64, 1)
  N_ = (ctypes.c_int * 3)(48, 32, 1)
  assert host_lib.fft_pruned_host(2, 2, n_, N_, -1, 1, p, p, p, p, p) == 1   # x modes not a multiple of 32


GPU_CASES = [
    # (grid in TF order, num points, transforms)
    ((32, 32), 20000, 1), ((32, 64), 20000, 5), ((256, 32), 30000, 2), ((512, 512), 100000, 3), ((128, 256), 50000, 40),
    ((32, 32, 32), 30000, 3), ((64, 32, 128), 30000, 2), ((128, 128, 128), 200000, 1), ((32, 512, 32), 20000, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("sign", ["forward", "backward"])
@pytest.mark.parametrize("grid,M,T", GPU_CASES)
def test_own_fft_matches_cufft_path(ttype, sign, grid, M, T):
  import torch
  from tensorflow_nufft_b200 import _lib
  from tensorflow_nufft_b200.python.ops import nufft_ops
  rank = len(grid)
  pts = H.uniform_points(M, rank, 77)
  src = H.random_complex((T, M) if ttype == 1 else (T,) + grid, 78)
  res = {}
  for mode in (0, 1, 2):
    out = nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid, f"type_{ttype}", sign, 1e-6,
                            None, "nufft", engine_kwargs={"fft_mode": mode})
    res[mode] = out.cpu().numpy()
  assert np.isfinite(res[0]).all()
  assert H.rel_l2(res[0], res[1]) < 6e-7
  assert H.rel_l2(res[2], res[1]) < 6e-7
  plan = _lib.Plan(ttype, grid[::-1], -1, T, 1e-6, 0, device=0)
  assert plan.info().fft_method == 3      # the own passes are what the default plan runs on these sizes
  plan.close()


@pytest.mark.gpu
def test_own_fft_not_selected_when_ineligible():
  from tensorflow_nufft_b200 import _lib
  for grid, dtype in [((48, 48), 0), ((30, 64), 0), ((64, 64), 1), ((1024, 1024), 0)]:
    plan = _lib.Plan(2, grid[::-1], -1, 1, 1e-6 if dtype == 0 else 1e-9, dtype, device=0)
    assert plan.info().fft_method in (1, 2), (grid, dtype)   # not a power of two / modes not 32k / complex128 / nf = 2048
    plan.close()
