"""Batch sharding of one `nufft` call across the GPUs of a box.

Transforms that share a point set (coils) are independent, so they are split contiguously across
ranks with NO collective on the data path: every rank gets the whole point set, builds its own
plan + bin-sort, and produces the contiguous `[T/G, ...]` slab of the result. The only optional
communication is a final gather of the slabs (NCCL over NVLink, or gloo in CPU tests).
"""
import torch


def shard_bounds(num_transforms, world_size, rank):
  """Contiguous, balanced [begin, end) of the transforms owned by `rank`."""
  if world_size < 1 or not 0 <= rank < world_size:
    raise ValueError("invalid rank/world_size")
  base, rem = divmod(int(num_transforms), int(world_size))
  begin = rank * base + min(rank, rem)
  end = begin + base + (1 if rank < rem else 0)
  return begin, end


def shard_sizes(num_transforms, world_size):
  return [shard_bounds(num_transforms, world_size, r)[1] - shard_bounds(num_transforms, world_size, r)[0]
          for r in range(world_size)]


def gather_slabs(local, num_transforms, group=None):
  """All-gathers the per-rank output slabs `[T_r, ...]` into `[T, ...]` (the optional final gather).
  Complex tensors travel as real views. Works with any torch.distributed backend."""
  import torch.distributed as dist
  world = dist.get_world_size(group)
  sizes = shard_sizes(num_transforms, world)
  is_complex = local.is_complex()
  loc = torch.view_as_real(local.contiguous()) if is_complex else local.contiguous()
  tail = list(loc.shape[1:])
  if len(set(sizes)) == 1:
    # equal slabs (T divisible by the world size: cfg2 32 coils, cfg4 16 coils on 1/2/4/8 GPUs):
    # one collective straight into the result, no staging copies
    out = torch.empty([num_transforms] + tail, dtype=loc.dtype, device=loc.device)
    dist.all_gather_into_tensor(out, loc, group=group)
    return torch.view_as_complex(out) if is_complex else out
  maxn = max(sizes)
  pad = torch.zeros([maxn] + tail, dtype=loc.dtype, device=loc.device)
  pad[:loc.shape[0]] = loc
  bufs = [torch.empty_like(pad) for _ in range(world)]
  dist.all_gather(bufs, pad, group=group)
  out = torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
  return torch.view_as_complex(out) if is_complex else out


def max_over_ranks(value, device=None, group=None):
  """Timing reduction used by bench.py: the step time of a job is the slowest rank's."""
  import torch.distributed as dist
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t[0])
