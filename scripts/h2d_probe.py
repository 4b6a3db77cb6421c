"""Concurrent pinned host <-> device copy bandwidth with N ranks on one box (torchrun, one rank per
GPU): does the host side cap the end-to-end arm of bench.py when every rank streams its coils in?
Prints one JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both at once.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  nbytes = 512 << 20
  h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
  h_in.fill_(1)
  d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  d_b = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
  s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, reps=8):
    for _ in range(2):
      fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      fn()
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / reps

  def h2d():
    d_a.copy_(h_in, non_blocking=True)

  def d2h():
    h_out.copy_(d_b, non_blocking=True)

  def both():
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
      d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
      h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)

  res = {}
  for name, fn, mult in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    ms = timed(fn)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
      allt = [torch.zeros_like(t) for _ in range(world)]
      dist.all_gather(allt, t)
      per = [float(x[0]) for x in allt]
    else:
      per = [ms]
    res[name] = {"per_rank_gbs": [round(mult * nbytes / (m * 1e-3) / 1e9, 1) for m in per],
                 "aggregate_gbs": round(world * mult * nbytes / (max(per) * 1e-3) / 1e9, 1)}
  if rank == 0:
    print(json.dumps({"probe": "pinned_copy_bandwidth", "ranks": world, "bytes_per_copy": nbytes,
                      "host_cores": os.cpu_count(), **res}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
