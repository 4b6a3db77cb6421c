"""`nufft`, `interp`, `spread`, `nudft`: the reference operator surface
(tensorflow_nufft/python/ops/nufft_ops.py:29-321) re-hosted on torch tensors, because TensorFlow
is not in this image. Same names, argument meaning, defaults and error text; the layout
bookkeeping of `NUFFTBaseOp::Compute` (tensorflow_nufft/cc/kernels/nufft_kernels.cc:54-379) is
reproduced here on the host, and all arithmetic is done by the CUDA engine through the C ABI
(include/b200nufft.h). There is no CPU fallback: tensors that live on the host are copied to the
current CUDA device and the result is copied back.
"""
import math
import os

import numpy as np
import torch

from tensorflow_nufft_b200 import _lib
from tensorflow_nufft_b200.python.ops import nufft_options

_ENGINE_DEFAULTS = {}
# Point-set reuse lives in the C library (opts.reuse_points): set_points fingerprints
# the raw coordinates ON THE DEVICE and, when they equal the set the cached plan was built from,
# every set_points kernel exits at once -- the fixed-trajectory case of iterative reconstruction
# (every CG iteration calls A and A^H with the same k-space trajectory). It is content based, so
# writes that bypass torch's version counter (DLPack aliases, raw-pointer libraries) cannot fool
# it. Opt-in: B200NUFFT_REUSE_POINTS=1 or set_points_reuse(True).
_REUSE_POINTS = os.environ.get("B200NUFFT_REUSE_POINTS", "0") != "0"
STATS = {"set_points_calls": 0}


def set_points_reuse(enabled):
  """Enables / disables the device-side unchanged-points shortcut of set_points (default: off)."""
  global _REUSE_POINTS
  _REUSE_POINTS = bool(enabled)


def set_engine_defaults(**kwargs):
  """Engine knobs applied to every new plan (e.g. num_threads_compat=8, fseries_mode=0)."""
  _ENGINE_DEFAULTS.update(kwargs)


def clear_plan_cache():
  """Destroys the idle plans of the library's process-level plan cache."""
  _lib.plan_cache_clear()


def _get_plan(key_args, opt_kwargs):
  """Takes a plan from the C library's plan cache (b200nufft_plan_acquire); the caller gives it
  back with plan.close() (b200nufft_plan_release). Keyed by every create argument."""
  ttype, grid_dims, sign, ntr, tol, dcode, device = key_args
  kw = dict(opt_kwargs)
  if _REUSE_POINTS:
    kw.setdefault("reuse_points", 1)
  return _lib.Plan(ttype, grid_dims, sign, ntr, tol, dcode, device=device, cached=True, **kw)


def _to_host(t):
  """Device -> pinned host copy (torch's caching host allocator makes the pinned buffer cheap)."""
  out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
  out.copy_(t, non_blocking=True)
  torch.cuda.current_stream().synchronize()
  return out


def _complex_dtype(real_dtype):
  return {torch.float32: torch.complex64, torch.float64: torch.complex128}[real_dtype]


def _real_dtype(complex_dtype):
  return {torch.complex64: torch.float32, torch.complex128: torch.float64}[complex_dtype]


def _op_tol(tol):
  # The op attr is a 32-bit float (nufft_ops.cc:214); the plan gets static_cast<FloatType>(tol_)
  # (nufft_kernels.cc:361).
  return float(np.float32(tol))


def _run_op(source, points, grid_shape, transform_type, fft_direction, tol, options, op_type,
            engine_kwargs=None):
  """Host mirror of NUFFTBaseOp::Compute + ::Execute (nufft_kernels.cc:54-542)."""
  if transform_type not in ("type_1", "type_2"):
    raise ValueError(
        f"transform_type attr must be 'type_1' or 'type_2', but is {transform_type}")
  if fft_direction not in ("forward", "backward"):
    raise ValueError(f"fft_direction must be 'forward' or 'backward', but is {fft_direction}")
  if not torch.is_tensor(source) or not source.is_complex():
    raise ValueError("Input `source` must have type complex64 or complex128")
  if not torch.is_tensor(points):
    raise ValueError("Input `points` must be a tensor")
  if source.dtype not in (torch.complex64, torch.complex128):
    raise ValueError(f"Input `source` must have type complex64 or complex128 but got: {source.dtype}")
  # Lazy conjugate / negative views keep the parent's data_ptr: materialise them before any raw
  # pointer reaches the engine (torch.conj(x), x.mH, ... would otherwise transform x itself).
  source = source.resolve_conj().resolve_neg()
  points = points.resolve_neg()
  rdtype = _real_dtype(source.dtype)
  if points.dtype != rdtype:
    raise ValueError(
        f"Input `points` must have type {rdtype} but got: {points.dtype}")
  if points.dim() < 2:
    raise ValueError(
        f"Input `points` must have rank of at least 2, but got shape: {list(points.shape)}")
  rank = points.shape[-1]
  num_points = points.shape[-2]
  if rank not in (1, 2, 3):
    raise ValueError(f"Dimension must be 1, 2 or 3 but is {rank}")  # nufft_ops.cc:33-41

  ttype = 1 if transform_type == "type_1" else 2
  if ttype == 1:
    if grid_shape is None:
      raise ValueError("grid_shape must be provided for type-1 transforms")
    grid_shape = [int(g) for g in grid_shape]
    if len(grid_shape) != rank:
      raise ValueError(
          f"grid_shape must have length {rank} for a {rank}D transform (as inferred from points), "
          f"but got length: {len(grid_shape)}")
    if source.dim() < 1 or source.shape[-1] != num_points:
      raise ValueError(
          "source and points must have equal samples dimensions for type-1 transforms, but got "
          f"source.shape[-1] = {source.shape[-1] if source.dim() else None} and "
          f"points.shape[-2] = {num_points}")
    source_elem_rank = 1
  else:
    if source.dim() < rank:
      raise ValueError(
          f"Input `source` must have rank of at least {rank} but received shape: {list(source.shape)}")
    grid_shape = [int(g) for g in source.shape[source.dim() - rank:]]
    source_elem_rank = rank

  source_batch = list(source.shape[:source.dim() - source_elem_rank])
  points_batch = list(points.shape[:-2])
  source_elem = list(source.shape[source.dim() - source_elem_rank:])
  nb = max(len(source_batch), len(points_batch))
  source_batch = [1] * (nb - len(source_batch)) + source_batch
  points_batch = [1] * (nb - len(points_batch)) + points_batch
  out_batch = []
  for s, p in zip(source_batch, points_batch):
    if s != p and s != 1 and p != 1:
      raise ValueError(f"Incompatible shapes: {list(source.shape)} vs. {list(points.shape)}")
    out_batch.append(max(s, p) if min(s, p) != 0 else 0)
  target_elem = grid_shape if ttype == 1 else [num_points]
  target_shape = out_batch + target_elem

  # "inner" batch dims share one point set (-> num_transforms); "outer" dims get their own
  # set_points call (nufft_kernels.cc:224-239).
  inner = [i for i in range(nb) if points_batch[i] == 1]
  outer = [i for i in range(nb) if points_batch[i] != 1]
  order = outer + inner
  num_transforms = 1
  for i in inner:
    num_transforms *= source_batch[i]
  num_calls = 1
  for i in outer:
    num_calls *= points_batch[i]

  # Device placement: the engine runs on CUDA only.
  if not torch.cuda.is_available():
    raise RuntimeError("tensorflow_nufft_b200 needs a CUDA device; there is no CPU fallback")
  host_io = not source.is_cuda
  device = source.device if source.is_cuda else torch.device("cuda", torch.cuda.current_device())
  # Host-resident inputs sharing one point set always take the streamed path: even a single chunk
  # gains the overlap of the strengths copy with the points copy + set_points.
  if (host_io and op_type == "nufft" and not outer and num_points > 0 and
      all(d != 0 for d in target_shape) and source.is_contiguous() and
      source.numel() * source.element_size() >= _HOST_STREAM_MIN_BYTES):
    return _run_host_pipelined(source, points, grid_shape, ttype, fft_direction, tol, options,
                               num_transforms, num_points, target_shape, device, engine_kwargs)
  src = source.to(device, non_blocking=True) if host_io else source
  pts = points.to(device, non_blocking=True) if not points.is_cuda else points
  if pts.device != device:
    raise ValueError("source and points must be on the same device")

  if any(d == 0 for d in target_shape):
    out = torch.zeros(target_shape, dtype=source.dtype, device=device)
    return _to_host(out) if host_io else out

  src_b = src.reshape(source_batch + source_elem)
  pts_b = pts.reshape(points_batch + [num_points, rank])
  elem_axes_s = list(range(nb, nb + len(source_elem)))
  src_p = src_b.permute(order + elem_axes_s).contiguous()
  pts_p = pts_b.permute(order + [nb, nb + 1]).contiguous().reshape(num_calls, num_points, rank)

  n_coeffs = 1
  for g in grid_shape:
    n_coeffs *= g
  src_outer_dims = [source_batch[i] for i in outer]
  pts_outer_dims = [points_batch[i] for i in outer]
  src_outer_count = 1
  for d in src_outer_dims:
    src_outer_count *= d
  src_elem_count = num_points if ttype == 1 else n_coeffs
  tgt_elem_count = n_coeffs if ttype == 1 else num_points
  src_flat = src_p.reshape(src_outer_count, num_transforms, src_elem_count)
  tgt_flat = torch.empty((num_calls, num_transforms, tgt_elem_count), dtype=source.dtype, device=device)

  options = options or nufft_options.Options()
  opt_kwargs = dict(_ENGINE_DEFAULTS)
  if op_type == "nufft":
    opt_kwargs.update(options.to_engine_kwargs())
  else:
    # Interp/Spread ops carry no options attr: default proto => STRICT (nufft_kernels.cc:455).
    opt_kwargs.update({"points_range": 0, "spread_only": 1})
  if engine_kwargs:
    opt_kwargs.update(engine_kwargs)
  dcode = _lib.COMPLEX64 if source.dtype == torch.complex64 else _lib.COMPLEX128
  sign = -1 if fft_direction == "forward" else 1
  grid_dims_xfast = tuple(reversed(grid_shape))  # nufft_kernels.cc:347-352
  dev_index = device.index if device.index is not None else torch.cuda.current_device()
  with torch.cuda.device(dev_index):
    plan = _get_plan((ttype, grid_dims_xfast, sign, num_transforms, _op_tol(tol), dcode, dev_index),
                     opt_kwargs)
    try:
      stream = torch.cuda.current_stream().cuda_stream
      # mixed-radix decode of the call index over the outer dims (nufft_kernels.cc:512-523)
      pf = [1] * len(outer)
      sf = [1] * len(outer)
      for d in range(len(outer) - 2, -1, -1):
        pf[d] = pf[d + 1] * pts_outer_dims[d + 1]
        sf[d] = sf[d + 1] * src_outer_dims[d + 1]
      for call in range(num_calls):
        plan.set_points_interleaved(num_points, pts_p[call].data_ptr(), stream)
        STATS["set_points_calls"] += 1
        rem = call
        sidx = 0
        for d in range(len(outer)):
          i_d = rem // pf[d]
          rem = rem % pf[d]
          if src_outer_dims[d] == 1:
            i_d = 0
          sidx += i_d * sf[d]
        s_ptr = src_flat[sidx].data_ptr()
        t_ptr = tgt_flat[call].data_ptr()
        if op_type == "nufft":
          if ttype == 1:
            plan.execute(s_ptr, t_ptr, stream)
          else:
            plan.execute(t_ptr, s_ptr, stream)
        elif op_type == "interp":
          plan.interp(t_ptr, s_ptr, stream)
        else:
          plan.spread(s_ptr, t_ptr, stream)
    finally:
      plan.close()   # back to the library's plan cache

  tgt_perm_shape = [out_batch[i] for i in order] + target_elem
  tgt = tgt_flat.reshape(tgt_perm_shape)
  inv = [0] * nb
  for pos, ax in enumerate(order):
    inv[ax] = pos
  tgt = tgt.permute(inv + list(range(nb, nb + len(target_elem)))).contiguous()
  tgt = tgt.reshape(target_shape)
  return _to_host(tgt) if host_io else tgt


_HOST_STREAM_MIN_BYTES = 8 << 20  # below this the plain copy-in / transform / copy-out path is used
_HOST_CHUNK = 8                   # transforms per pipelined chunk for small transforms
_HOST_CHUNK_BYTES = 128 << 20     # ... and about this many bytes per chunk for large ones


def _host_chunk(source, num_transforms, num_points, grid_shape):
  """Transforms per pipelined chunk: 8 for 2D-sized transforms (cfg2: 16 MB of strengths each),
  fewer when one transform is already tens of MB (cfg4: a 256^3 grid is 134 MB -> one per chunk,
  so that the copy of coil k+1 overlaps the transform of coil k)."""
  n_coeffs = 1
  for g in grid_shape:
    n_coeffs *= g
  per = max(num_points, n_coeffs) * source.element_size()
  return max(1, min(_HOST_CHUNK, _HOST_CHUNK_BYTES // max(per, 1)))


def _run_host_pipelined(source, points, grid_shape, ttype, fft_direction, tol, options, num_transforms,
                        num_points, target_shape, device, engine_kwargs):
  """Host-resident inputs, many transforms sharing one point set: the coils are streamed through
  the GPU in chunks (`_host_chunk`) with the H2D copy of chunk k+1, the transform of chunk k and
  the D2H copy of chunk k-1 overlapped on three CUDA streams (the PCIe copies dominate: 8 bytes
  per point-transform each way)."""
  dev_index = device.index if device.index is not None else torch.cuda.current_device()
  n_coeffs = 1
  for g in grid_shape:
    n_coeffs *= g
  src_elems = num_points if ttype == 1 else n_coeffs
  tgt_elems = n_coeffs if ttype == 1 else num_points
  T = num_transforms
  src_flat = source.reshape(T, src_elems)
  if not src_flat.is_pinned():
    src_flat = src_flat.pin_memory()
  out = torch.empty((T, tgt_elems), dtype=source.dtype, pin_memory=True)
  opt_kwargs = dict(_ENGINE_DEFAULTS)
  opt_kwargs.update((options or nufft_options.Options()).to_engine_kwargs())
  if engine_kwargs:
    opt_kwargs.update(engine_kwargs)
  dcode = _lib.COMPLEX64 if source.dtype == torch.complex64 else _lib.COMPLEX128
  sign = -1 if fft_direction == "forward" else 1
  chunk = min(T, _host_chunk(source, T, num_points, grid_shape))
  # Chunk schedule: full chunks, then the last full chunk's worth is halved so that the part of the
  # pipeline that cannot overlap (the last transform + its D2H copy) is short.
  sizes = [chunk] * (T // chunk)
  if T % chunk:
    sizes.append(T % chunk)
  if len(sizes) >= 2 and sizes[-1] == chunk and chunk % 2 == 0:
    sizes[-1:] = [chunk // 2, chunk // 2]
  with torch.cuda.device(dev_index):
    main = torch.cuda.current_stream()
    copy_in, copy_out = _side_streams(dev_index)
    d_in = [torch.empty((chunk, src_elems), dtype=source.dtype, device=device) for _ in range(2)]
    d_out = [torch.empty((chunk, tgt_elems), dtype=source.dtype, device=device) for _ in range(2)]
    # The first strengths copy may start as soon as the buffers exist: it overlaps the points copy
    # and set_points (bin-sort + stencil records) on the main stream.
    copy_in.wait_stream(main)
    plans = {}
    for n in sorted(set(sizes), reverse=True):
      plans[n] = _get_plan((ttype, tuple(reversed(grid_shape)), sign, n, _op_tol(tol), dcode, dev_index), opt_kwargs)
    pts = points.to(device, non_blocking=True).reshape(num_points, -1)
    for n, pl in plans.items():
      pl.set_points_interleaved(num_points, pts.data_ptr(), main.cuda_stream)
      STATS["set_points_calls"] += 1
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]
    out_free = [torch.cuda.Event() for _ in range(2)]
    b0 = 0
    for k, n in enumerate(sizes):
      s = k & 1
      with torch.cuda.stream(copy_in):
        if k >= 2:
          copy_in.wait_event(in_free[s])
        d_in[s][:n].copy_(src_flat[b0:b0 + n], non_blocking=True)
        in_ready[s].record(copy_in)
      main.wait_event(in_ready[s])
      if k >= 2:
        main.wait_event(out_free[s])
      pl = plans[n]
      if ttype == 1:
        pl.execute(d_in[s].data_ptr(), d_out[s].data_ptr(), main.cuda_stream)
      else:
        pl.execute(d_out[s].data_ptr(), d_in[s].data_ptr(), main.cuda_stream)
      in_free[s].record(main)
      out_ready[s].record(main)
      with torch.cuda.stream(copy_out):
        copy_out.wait_event(out_ready[s])
        out[b0:b0 + n].copy_(d_out[s][:n], non_blocking=True)
        out_free[s].record(copy_out)
      b0 += n
    main.wait_stream(copy_out)
    main.synchronize()
    for pl in plans.values():
      pl.close()   # back to the library's plan cache
  return out.reshape(target_shape)


_SIDE_STREAMS = {}


def _side_streams(dev_index):
  if dev_index not in _SIDE_STREAMS:
    _SIDE_STREAMS[dev_index] = (torch.cuda.Stream(device=dev_index), torch.cuda.Stream(device=dev_index))
  return _SIDE_STREAMS[dev_index]


def _sum_to_shape(x, shape):
  """Undo batch broadcasting (BroadcastGradientArgs + reduce_sum, nufft_ops.py:218-229)."""
  shape = list(shape)
  while x.dim() > len(shape):
    x = x.sum(0)
  for i, s in enumerate(shape):
    if s == 1 and x.shape[i] != 1:
      x = x.sum(i, keepdim=True)
  return x.reshape(shape)


class _NufftFunction(torch.autograd.Function):
  """Gradients as `_nufft_grad` registers them (nufft_ops.py:126-232)."""

  @staticmethod
  def forward(ctx, source, points, grid_shape, transform_type, fft_direction, tol, options):
    ctx.save_for_backward(source, points)
    ctx.cfg = (grid_shape, transform_type, fft_direction, tol, options)
    return _run_op(source.detach(), points.detach(), grid_shape, transform_type, fft_direction, tol,
                   options, "nufft")

  @staticmethod
  def backward(ctx, grad):
    source, points = ctx.saved_tensors
    grid_shape, transform_type, fft_direction, tol, options = ctx.cfg
    rank = points.shape[-1]
    dtype = source.dtype
    if transform_type == "type_2":
      grid_shape = list(source.shape[-rank:])
    grad_transform_type = "type_2" if transform_type == "type_1" else "type_1"
    grad_fft_direction = "forward" if fft_direction == "backward" else "backward"
    grad_source = grad_points = None
    if ctx.needs_input_grad[0]:
      grad_source = nufft(grad, points, grid_shape=grid_shape, transform_type=grad_transform_type,
                          fft_direction=grad_fft_direction, tol=tol, options=options)
      grad_source = _sum_to_shape(grad_source, source.shape)
    if ctx.needs_input_grad[1]:
      rdtype = _real_dtype(dtype)
      grid_vec = [torch.arange(n, dtype=rdtype, device=grad.device) - n / 2 for n in grid_shape]
      grid_points = torch.stack(torch.meshgrid(*grid_vec, indexing="ij"), dim=0).to(dtype)
      imag_unit = torch.tensor(-1j if fft_direction == "forward" else 1j, dtype=dtype, device=grad.device)
      gconj = torch.conj(grad)
      if transform_type == "type_2":
        gp = nufft(source.unsqueeze(-(rank + 1)) * grid_points, points.unsqueeze(-3),
                   transform_type="type_2", fft_direction=fft_direction, tol=tol, options=options)
        gp = gp * gconj.unsqueeze(-2) * imag_unit
      else:
        gp = nufft(gconj.unsqueeze(-(rank + 1)) * grid_points, points.unsqueeze(-3),
                   transform_type="type_2", fft_direction=fft_direction, tol=tol, options=options)
        gp = gp * source.unsqueeze(-2) * imag_unit
      gp = torch.real(gp).transpose(-1, -2)
      grad_points = _sum_to_shape(gp, points.shape)
    return grad_source, grad_points, None, None, None, None, None


def nufft(source,  # pylint: disable=missing-raises-doc
          points,
          grid_shape=None,
          transform_type="type_2",
          fft_direction="forward",
          tol=1e-6,
          options=None):
  """Computes the non-uniform discrete Fourier transform via NUFFT.

  Same contract as `tfft.nufft` (reference nufft_ops.py:34-123). `source`: complex64/complex128,
  `[..., M]` (type-1) or `[...] + grid_shape` (type-2). `points`: float32/float64 `[..., M, N]`,
  radians/sample; batch dims broadcast against `source`'s. `grid_shape` is required for type-1 and
  ignored for type-2. Returns `[...] + grid_shape` (type-1) or `[..., M]` (type-2).
  Differentiable w.r.t. `source` and `points`.
  """
  if grid_shape is None:
    if transform_type == "type_1":
      raise ValueError("grid_shape must be provided for type-1 transforms")
  elif torch.is_tensor(grid_shape):
    grid_shape = tuple(int(g) for g in grid_shape.tolist())
  else:
    grid_shape = tuple(int(g) for g in grid_shape)
  if transform_type == "type_2":
    grid_shape = None  # ignored by the op (nufft_ops.py:108-116)
  options = options or nufft_options.Options()
  needs_grad = torch.is_grad_enabled() and (
      (torch.is_tensor(source) and source.requires_grad) or
      (torch.is_tensor(points) and points.requires_grad))
  if needs_grad:
    return _NufftFunction.apply(source, points, grid_shape, transform_type, fft_direction, tol, options)
  return _run_op(source, points, grid_shape, transform_type, fft_direction, tol, options, "nufft")


def interp(source, points, tol=1e-6):
  """Interpolates a regular grid at arbitrary points (the `Interp` op, nufft_ops.cc:136-167):
  the interpolation step only, no FFT, no deconvolution, output scaled by kernel_scale."""
  return _run_op(source, points, None, "type_2", "forward", tol, None, "interp")


def spread(source, points, grid_shape, tol=1e-6):
  """Spreads arbitrary points onto a regular grid (the `Spread` op, nufft_ops.cc:170-201)."""
  if torch.is_tensor(grid_shape):
    grid_shape = grid_shape.tolist()
  return _run_op(source, points, tuple(int(g) for g in grid_shape), "type_1", "forward", tol, None,
                 "spread")


def nudft(source, points, grid_shape=None, transform_type="type_2", fft_direction="forward"):
  """Dense non-uniform DFT (reference nufft_ops.py:235-321); for testing. Pure torch, any device,
  differentiable. Not part of the accelerated path."""
  rank = points.shape[-1]
  if transform_type == "type_1":
    if grid_shape is None:
      raise ValueError("grid_shape must be provided for type-1 transforms")
    grid_shape = [int(g) for g in grid_shape]
    src_batch = source.shape[:-1]
  else:
    grid_shape = list(source.shape[-rank:])
    src_batch = source.shape[:-rank]
  pts_batch = points.shape[:-2]
  batch = torch.broadcast_shapes(tuple(src_batch), tuple(pts_batch))
  r_vec = [torch.arange(n, dtype=points.dtype, device=points.device) - n / 2 for n in grid_shape]
  r_grid = torch.stack(torch.meshgrid(*r_vec, indexing="ij"), dim=0).reshape(rank, -1)
  phase = torch.matmul(points, r_grid)  # [..., M, N]
  sign = 1.0 if fft_direction == "backward" else -1.0
  mat = torch.exp(torch.complex(torch.zeros_like(phase), sign * phase))
  if transform_type == "type_1":
    src = source.expand(tuple(batch) + source.shape[-1:])
    out = torch.matmul(src.unsqueeze(-2), mat.expand(tuple(batch) + mat.shape[-2:])).squeeze(-2)
    return out.reshape(tuple(batch) + tuple(grid_shape))
  src = source.reshape(tuple(src_batch) + (-1,)).expand(tuple(batch) + (math.prod(grid_shape),))
  out = torch.matmul(mat.expand(tuple(batch) + mat.shape[-2:]), src.unsqueeze(-1)).squeeze(-1)
  return out
