"""world_size-2 gloo test of the multi-GPU path's host logic (bench.py N>1 / SURVEY.md 8e): coils
are sharded contiguously with no data-path collective, each rank transforms its slab (here with
the CPU oracle standing in for the device engine), the optional final gather reassembles the
batch, and the job time is the max over ranks."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, T, ret):
  sys.path.insert(0, ROOT)
  import torch
  import torch.distributed as dist
  from oracle import port as oport
  from tensorflow_nufft_b200 import sharding
  from tests import helpers as H
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    grid = (12, 10)
    M = 200
    pts = np.ascontiguousarray(H.uniform_points(M, 2, 1)[:, ::-1].T)      # every rank: full point set
    src = H.random_complex((T, M), 2)                                     # the whole batch (synthetic)
    b, e = sharding.shard_bounds(T, world, rank)
    local = oport.nufft(src[b:e], pts, grid, 1, -1, 1e-6, np.complex64) if e > b else np.zeros((0, 120), np.complex64)
    full = sharding.gather_slabs(torch.from_numpy(local), T)
    tmax = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
      want = oport.nufft(src, pts, grid, 1, -1, 1e-6, np.complex64)
      ret["err"] = float(np.abs(full.numpy() - want).max())
      ret["tmax"] = tmax
      ret["shape"] = tuple(full.shape)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [4, 5])
def test_two_rank_batch_sharding_gloo(T):
  import torch.multiprocessing as mp
  from oracle import port as oport
  if not oport.available():
    pytest.skip("oracle/liboracle.so not built")
  ctx = mp.get_context("spawn")
  mgr = ctx.Manager()
  ret = mgr.dict()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, T, ret)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(timeout=180)
  for p in procs:
    assert p.exitcode == 0
  assert ret["shape"] == (T, 120)
  assert ret["err"] == 0.0          # same code, same inputs per coil: bit-identical slabs
  assert ret["tmax"] == 11.0
