"""GPU tests of the boundary's resource behaviour: no allocation on
the hot path, caller-owned workspace, allocator callbacks, the library's plan cache, device and
stream hygiene."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = float(np.float32(1e-6))


def _lib():
  from tensorflow_nufft_b200 import _lib as L
  return L


def _run(plan, pts, src, out, ttype):
  st = torch.cuda.current_stream().cuda_stream
  plan.set_points_interleaved(pts.shape[0], pts.data_ptr(), st)
  if ttype == 1:
    plan.execute(src.data_ptr(), out.data_ptr(), st)
  else:
    plan.execute(out.data_ptr(), src.data_ptr(), st)
  torch.cuda.synchronize()
  return out.clone()


def test_set_points_and_execute_do_not_allocate_after_reserve():
  L = _lib()
  grid = (48, 40)
  plan = L.Plan(1, grid[::-1], -1, 2, TOL, L.COMPLEX64, device=0)
  plan.reserve(60000)
  out = torch.empty((2, 48 * 40), dtype=torch.complex64, device="cuda")
  a0 = L.alloc_counts()
  for m in (1000, 20000, 60000, 300):      # growing, then shrinking point sets
    pts = torch.from_numpy(H.uniform_points(m, 2, m)).cuda()
    src = torch.from_numpy(H.random_complex((2, m), m + 1)).cuda()
    _run(plan, pts, src, out, 1)
  assert L.alloc_counts() == a0, "set_points / execute allocated or freed device memory"
  # without reserve the buffers grow geometrically: a much larger set allocates, then stays put
  pts = torch.from_numpy(H.uniform_points(200000, 2, 9)).cuda()
  src = torch.from_numpy(H.random_complex((2, 200000), 10)).cuda()
  _run(plan, pts, src, out, 1)
  a1 = L.alloc_counts()
  assert a1[0] > a0[0]
  _run(plan, pts, src, out, 1)
  assert L.alloc_counts() == a1
  plan.close()


@pytest.mark.parametrize("pow2", [False, True])
@pytest.mark.parametrize("ttype,rank", [(1, 2), (2, 2), (2, 3), (1, 3)])
def test_caller_owned_workspace_matches_plan_owned_buffers(ttype, rank, pow2):
  L = _lib()
  if pow2:   # power-of-two fine grids: the engine's own FFT passes (their tables are plan-owned, the grid is not)
    grid = (32, 64) if rank == 2 else (32, 32, 32)
  else:
    grid = (24, 20) if rank == 2 else (12, 16, 10)
  m, T = 7000, 3
  N = int(np.prod(grid))
  pts = torch.from_numpy(H.uniform_points(m, rank, 31)).cuda()
  src = torch.from_numpy(H.random_complex((T, m) if ttype == 1 else (T, N), 32)).cuda()
  out = torch.empty((T, N) if ttype == 1 else (T, m), dtype=torch.complex64, device="cuda")
  ref = L.Plan(ttype, grid[::-1], -1, T, TOL, L.COMPLEX64, device=0)
  want = _run(ref, pts, src, out, ttype)
  ref.close()

  a0 = L.alloc_counts()
  plan = L.Plan(ttype, grid[::-1], -1, T, TOL, L.COMPLEX64, device=0, external_workspace=1)
  a_created = L.alloc_counts()
  nbytes = plan.workspace_bytes(m)
  assert nbytes > 0 and plan.workspace_bytes(2 * m) > nbytes
  with pytest.raises(L.NufftError, match="bind_workspace"):
    plan.set_points_interleaved(m, pts.data_ptr(), torch.cuda.current_stream().cuda_stream)
  ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  with pytest.raises(L.NufftError, match="too small"):
    plan.bind_workspace(ws.data_ptr(), nbytes - 1024, m)
  plan.bind_workspace(ws.data_ptr(), nbytes, m)
  got = _run(plan, pts, src, out, ttype)
  assert L.alloc_counts() == a_created, "a workspace-bound plan must not allocate"
  if ttype == 2:
    assert torch.equal(got, want)
  else:   # type 1 sums through global reductions: same values, bits depend on arrival order
    assert H.rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 1e-6
  with pytest.raises(L.NufftError, match="workspace was bound for"):
    big = torch.from_numpy(H.uniform_points(2 * m, rank, 33)).cuda()
    plan.set_points_interleaved(2 * m, big.data_ptr(), torch.cuda.current_stream().cuda_stream)
  plan.unbind_workspace()
  with pytest.raises(L.NufftError):
    plan.execute(out.data_ptr(), src.data_ptr(), torch.cuda.current_stream().cuda_stream)
  plan.close()
  assert a_created[0] - a0[0] <= (14 if pow2 else 8)   # only the small fixed buffers (factors, FFT tables, scan scratch)


def test_allocator_callbacks_own_every_device_buffer():
  """b200nufft_plan_create_ex: all device memory through the caller's allocator (here torch's
  caching allocator; in TensorFlow the BFC allocator's AllocateRaw / DeallocateRaw)."""
  L = _lib()
  live = {}

  def alloc(user, nbytes, device):
    ptr = torch.cuda.caching_allocator_alloc(max(int(nbytes), 1), device)
    live[ptr] = nbytes
    return ptr

  def free(user, ptr, device):
    del live[ptr]
    torch.cuda.caching_allocator_delete(ptr)

  al = L.Allocator(L.ALLOC_FN(alloc), L.FREE_FN(free), None)
  for grid in [(32, 24), (32, 64)]:   # cuFFT plan / the engine's own FFT passes (twiddle + factor tables)
    m = 9000
    pts = torch.from_numpy(H.uniform_points(m, 2, 41)).cuda()
    src = torch.from_numpy(H.random_complex((2,) + grid, 42)).cuda()
    out = torch.empty((2, m), dtype=torch.complex64, device="cuda")
    ref = L.Plan(2, grid[::-1], -1, 2, TOL, L.COMPLEX64, device=0)
    want = _run(ref, pts, src, out, 2)
    ref.close()
    plan = L.Plan(2, grid[::-1], -1, 2, TOL, L.COMPLEX64, device=0, allocator=al)
    got = _run(plan, pts, src, out, 2)
    assert torch.equal(got, want)
    assert len(live) >= 10 and sum(live.values()) > m * 64
    plan.close()
    assert not live, "plan_destroy must return every buffer to the allocator"


def test_plan_cache_hands_out_idle_plans_only_and_evicts_lru():
  L = _lib()
  L.plan_cache_clear()
  s0 = L.plan_cache_stats()
  a = L.Plan(2, (16, 16), -1, 1, TOL, L.COMPLEX64, device=0, cached=True)
  b = L.Plan(2, (16, 16), -1, 1, TOL, L.COMPLEX64, device=0, cached=True)   # a is busy -> new plan
  assert a._h.value != b._h.value
  ha = a._h.value
  a.close()
  c = L.Plan(2, (16, 16), -1, 1, TOL, L.COMPLEX64, device=0, cached=True)   # gets a's handle back
  assert c._h.value == ha
  d = L.Plan(2, (16, 16), -1, 1, TOL, L.COMPLEX64, device=0, cached=True, points_range=0)  # other opts
  assert d._h.value not in (ha, b._h.value)
  s1 = L.plan_cache_stats()
  assert s1["hits"] - s0["hits"] == 1 and s1["misses"] - s0["misses"] == 3
  for p in (b, c, d):
    p.close()
  assert L.plan_cache_stats()["idle"] == 3
  keep = [L.Plan(2, (16, 16 + 2 * i), -1, 1, TOL, L.COMPLEX64, device=0, cached=True) for i in range(12)]
  for p in keep:
    p.close()
  assert L.plan_cache_stats()["idle"] == 8       # default capacity; the oldest were destroyed
  L.plan_cache_clear()
  assert L.plan_cache_stats()["idle"] == 0


def test_one_handle_used_from_two_streams_is_ordered():
  """set_points on one stream, execute on another, no explicit synchronisation."""
  L = _lib()
  grid, m, T = (256, 256), 400000, 4
  pts = torch.from_numpy(H.uniform_points(m, 2, 51)).cuda()
  src = torch.from_numpy(H.random_complex((T, m), 52)).cuda()
  out = torch.empty((T, 256 * 256), dtype=torch.complex64, device="cuda")
  plan = L.Plan(1, grid[::-1], -1, T, TOL, L.COMPLEX64, device=0)
  want = _run(plan, pts, src, out, 1)
  s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
  torch.cuda.synchronize()
  for _ in range(5):
    out.zero_()
    torch.cuda.synchronize()
    plan.set_points_interleaved(m, pts.data_ptr(), s1.cuda_stream)
    plan.execute(src.data_ptr(), out.data_ptr(), s2.cuda_stream)
    s2.synchronize()
    assert H.rel_l2(out.cpu().numpy(), want.cpu().numpy()) < 1e-6
  plan.close()


def test_entry_points_restore_the_callers_device():
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  L = _lib()
  torch.cuda.set_device(0)
  plan = L.Plan(2, (16, 16), -1, 1, TOL, L.COMPLEX64, device=1)
  assert torch.cuda.current_device() == 0      # c10 asks the runtime (cudaGetDevice)
  pts = torch.from_numpy(H.uniform_points(100, 2, 1)).to("cuda:1")
  with torch.cuda.device(1):
    st = torch.cuda.current_stream().cuda_stream
  plan.set_points_interleaved(100, pts.data_ptr(), st)
  assert torch.cuda.current_device() == 0
  plan.close()
  assert torch.cuda.current_device() == 0


def test_user_batch_size_is_clamped_to_the_grid_limit():
  L = _lib()
  plan = L.Plan(2, (16, 16), -1, 70000, TOL, L.COMPLEX64, device=0, max_batch_size=70000)
  assert plan.info().batch_size == 65535
  plan.close()


@pytest.mark.parametrize("grid", [(64, 48), (256, 256)])   # cuFFT plan / the engine's own FFT passes (64 KB bundles)
@pytest.mark.parametrize("ttype", [1, 2])
def test_set_points_and_execute_capture_into_a_cuda_graph(ttype, grid):
  """No allocation, no host synchronisation, no cross-stream events on the hot path: the pair is
  capturable and the replayed graph gives the same result on new input data."""
  L = _lib()
  m, T = 20000, 2
  N = grid[0] * grid[1]
  pts = torch.from_numpy(H.uniform_points(m, 2, 61)).cuda()
  src = torch.from_numpy(H.random_complex((T, m) if ttype == 1 else (T, N), 62)).cuda()
  out = torch.zeros((T, N) if ttype == 1 else (T, m), dtype=torch.complex64, device="cuda")
  plan = L.Plan(ttype, grid[::-1], -1, T, TOL, L.COMPLEX64, device=0)
  plan.reserve(m)
  side = torch.cuda.Stream()
  side.wait_stream(torch.cuda.current_stream())

  def step(stream):
    plan.set_points_interleaved(m, pts.data_ptr(), stream)
    if ttype == 1:
      plan.execute(src.data_ptr(), out.data_ptr(), stream)
    else:
      plan.execute(out.data_ptr(), src.data_ptr(), stream)

  with torch.cuda.stream(side):
    for _ in range(2):
      step(side.cuda_stream)
  side.synchronize()
  want = out.clone()
  g = torch.cuda.CUDAGraph()
  a0 = L.alloc_counts()
  with torch.cuda.graph(g, stream=side):
    step(side.cuda_stream)
  assert L.alloc_counts() == a0
  # new data in the same buffers: replay must transform IT
  pts.copy_(torch.from_numpy(H.uniform_points(m, 2, 63)).cuda())
  src.mul_(0.5)
  torch.cuda.synchronize()
  g.replay()
  torch.cuda.synchronize()
  got = out.clone()
  step(torch.cuda.current_stream().cuda_stream)   # eager, same inputs
  torch.cuda.synchronize()
  assert H.rel_l2(got.cpu().numpy(), out.cpu().numpy()) < 1e-6
  assert H.rel_l2(got.cpu().numpy(), want.cpu().numpy()) > 0.1
  plan.close()
