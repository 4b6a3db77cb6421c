#!/usr/bin/env python
"""bench.py -- NU points/sec of the tfft.nufft hot path on B200 (BASELINE.json metric).

Headline workload (N=1): BASELINE config[1] -- 2D type-1 adjoint NUFFT, 512x512 grid, 32-coil batch
sharing 2M spiral points, complex64, tol 1e-6. One "step" = one full pass of the hot path over the
batch: set_points (fold, bin-sort, stencil records) + execute (spread, FFT, deconvolve) for all
coils (the FFT is the engine's own pruned passes with the deconvolve / amplify step fused in, so
`stages_ms.deconv_ms` is ~0 and `fft_ms` carries both). `value` = coils*M / step time with inputs resident in HBM; `e2e` = the same through the public
`tfft.nufft` call with pinned HOST tensors (H2D + D2H inside the timed region).

The same JSON line also carries
  * `type2`:  BASELINE config[3] as north_star splits it -- 3D type-2, 256^3, 2 coils per GPU x 4M
              stack-of-stars points -- with its own stages, roofline and e2e;
  * `strong`: strong scaling of the two batched configs (cfg2: 32 coils, cfg4: 16 coils, split
              T/N per rank with sharding.shard_bounds, every rank bin-sorts the full point set) and
              the optional final all-gather of the result slabs timed separately (`gather_ms`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--config cfg1|cfg2|cfg3|cfg4|ref1..ref8] [--only-main]

N>1: launched by torchrun, one rank per GPU; the headline is weak scaling (every rank transforms its
own 32-coil shard of a 32*N-coil batch with the same point set, no data-path collective); time =
max over ranks. `--impl reference` times the reference's own OpenMP CPU plan (oracle/_ref/libref.so,
mode auto = what tfft.nufft does on /cpu:0) on the same workload. ref1..ref8 are the reference's own
benchmark shapes (nufft_ops_test.py:732-741).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import helpers as H  # noqa: E402  (synthetic point sets only; no oracle code)


def load_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  return 6650.0, "fallback (B200_PROFILING.md)"


def _uni(m, rank, seed):
  return lambda: H.uniform_points(m, rank, seed)


CONFIGS = {
    # ttype, fft direction, grid (TF order), coils per GPU, point sets per call (outer batch), points
    "cfg1": dict(ttype=2, direction="forward", grid=(256, 256), coils=1, points=lambda: H.radial_points(200, 500),
                 desc="2D type-2, 256x256, 100k radial points, 1 transform, complex64, tol 1e-6"),
    "cfg2": dict(ttype=1, direction="backward", grid=(512, 512), coils=32, points=lambda: H.spiral_points(32, 62500),
                 desc="2D type-1 adjoint, 512x512, 32 coils x 2M spiral points, complex64, tol 1e-6"),
    "cfg3": dict(ttype=1, direction="forward", grid=(128, 128, 128), coils=1, points=_uni(8000000, 3, 3),
                 desc="3D type-1, 128^3, 8M uniform-random points, 1 transform, complex64, tol 1e-6"),
    "cfg4": dict(ttype=2, direction="forward", grid=(256, 256, 256), coils=2,
                 points=lambda: H.stack_of_stars_points(125, 125, 256),
                 desc="3D type-2, 256^3, 2 coils/GPU x 4M stack-of-stars points, complex64, tol 1e-6"),
    # The reference's own benchmark cases (nufft_ops_test.py:732-741): 200k / 800k uniform points.
    "ref1": dict(ttype=2, direction="forward", grid=(256, 256), coils=1, points=_uni(200000, 2, 11),
                 desc="ref bench 1: 2D type-2 256^2, 200k points"),
    "ref2": dict(ttype=2, direction="forward", grid=(256, 256), coils=16, points=_uni(200000, 2, 12),
                 desc="ref bench 2: 2D type-2 256^2, batch 16 sharing 200k points"),
    "ref3": dict(ttype=2, direction="forward", grid=(256, 256), coils=1, sets=16, points=_uni(200000, 2, 13),
                 desc="ref bench 3: 2D type-2 256^2, batch 16 with OWN point sets (16 set_points per call)"),
    "ref4": dict(ttype=1, direction="forward", grid=(256, 256), coils=1, points=_uni(200000, 2, 14),
                 desc="ref bench 4: 2D type-1 256^2, 200k points"),
    "ref5": dict(ttype=1, direction="forward", grid=(256, 256), coils=16, points=_uni(200000, 2, 15),
                 desc="ref bench 5: 2D type-1 256^2, batch 16 sharing 200k points"),
    "ref6": dict(ttype=1, direction="forward", grid=(256, 256), coils=1, sets=16, points=_uni(200000, 2, 16),
                 desc="ref bench 6: 2D type-1 256^2, batch 16 with OWN point sets (16 set_points per call)"),
    "ref7": dict(ttype=2, direction="forward", grid=(128, 128, 128), coils=1, points=_uni(800000, 3, 17),
                 desc="ref bench 7: 3D type-2 128^3, 800k points"),
    "ref8": dict(ttype=1, direction="forward", grid=(128, 128, 128), coils=1, points=_uni(800000, 3, 18),
                 desc="ref bench 8: 3D type-1 128^3, 800k points"),
}
STRONG = {"cfg2": 32, "cfg4": 16}   # total coils of the strong-scaling jobs
TOL = 1e-6


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""

  FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
            "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.gpu_index = gpu_index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
           "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.25)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    sm, smmax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in self.lines:
      parts = [p.strip() for p in ln.split(",")]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1]))
        smmax.append(float(parts[2]))
      except ValueError:
        continue
      for nm, val in zip(names, parts[5:9]):
        if val.lower().startswith("active"):
          reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smmax)) if smmax else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def spread_interp_bytes(rank, M, nf_tot, coils_per_launch):
  """Algorithmic bytes of ONE spread/interp launch: coordinates once + per coil the
  strengths and the fine grid once. complex64 / float32."""
  return M * rank * 4 + coils_per_launch * (M * 8 + nf_tot * 8)


class Dist:
  """Barrier / max-over-ranks helpers that degrade to no-ops at world size 1."""

  def __init__(self, world):
    self.world = world

  def barrier(self):
    import torch
    if self.world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize()

  def max(self, *vals):
    import torch
    if self.world == 1:
      return [float(v) for v in vals]
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return [float(x) for x in t]


def kernel_name(ttype, rank):
  if ttype == 1:
    return "spread_sweep2d_f32_kernel" if rank == 2 else "spread_sweep3d_f32_kernel"
  return "interp_qw_f32_kernel"


def device_arm(name, cfg, T, dist, device, steps, warmup, seed, want_out=False):
  """C ABI, inputs resident in HBM: K timed steps (set_points + execute) between barriers, then the
  same K steps again with per-stage CUDA events read back after every step (stage averages; the
  synchronisation they need stays out of the headline timing)."""
  import torch
  from tensorflow_nufft_b200 import _lib
  pts_np = cfg["points"]()
  M, rank = pts_np.shape
  grid = cfg["grid"]
  N = int(np.prod(grid))
  ttype = cfg["ttype"]
  sets = cfg.get("sets", 1)
  sign = -1 if cfg["direction"] == "forward" else 1
  ncores = os.cpu_count() or 1
  gen = torch.Generator(device="cuda").manual_seed(seed)
  d_pts = [torch.from_numpy(pts_np if s == 0 else np.roll(pts_np, s, axis=0).copy()).cuda() for s in range(sets)]
  src_shape = (sets, T, M) if ttype == 1 else (sets, T, N)
  d_src = torch.view_as_complex(torch.rand(src_shape + (2,), generator=gen, device="cuda") - 0.5)
  d_out = torch.empty((sets, T, N) if ttype == 1 else (sets, T, M), dtype=torch.complex64, device="cuda")
  plan = _lib.Plan(ttype, tuple(reversed(grid)), sign, T, float(np.float32(TOL)), _lib.COMPLEX64, device=device,
                   profile=1, num_threads_compat=ncores)
  plan.reserve(M)
  stream = torch.cuda.current_stream().cuda_stream

  def step():
    for s in range(sets):
      plan.set_points_interleaved(M, d_pts[s].data_ptr(), stream)
      if ttype == 1:
        plan.execute(d_src[s].data_ptr(), d_out[s].data_ptr(), stream)
      else:
        plan.execute(d_out[s].data_ptr(), d_src[s].data_ptr(), stream)

  for _ in range(warmup):
    step()
  dist.barrier()
  a0 = _lib.alloc_counts()
  l0 = plan.launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  dist.barrier()
  e0.record()
  for _ in range(steps):
    step()
  e1.record()
  dist.barrier()
  ms_total = e0.elapsed_time(e1)
  launches = plan.launch_count() - l0
  allocs = _lib.alloc_counts()[0] - a0[0]
  stage = {"spread_interp_ms": 0.0, "fft_ms": 0.0, "deconv_ms": 0.0, "set_points_ms": 0.0}
  for _ in range(steps):
    step()
    tm = plan.timings()       # last set_points + execute of the step (synchronises on its events)
    for k in stage:
      stage[k] += tm[k] / steps
  info = plan.info()
  # the same step replayed from a CUDA graph (no per-launch CPU cost; matters for launch-bound sizes)
  graph_ms = None
  try:
    gplan = _lib.Plan(ttype, tuple(reversed(grid)), sign, T, float(np.float32(TOL)), _lib.COMPLEX64, device=device,
                      num_threads_compat=ncores)
    gplan.reserve(M)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())

    def gstep(st):
      for s_ in range(sets):
        gplan.set_points_interleaved(M, d_pts[s_].data_ptr(), st)
        if ttype == 1:
          gplan.execute(d_src[s_].data_ptr(), d_out[s_].data_ptr(), st)
        else:
          gplan.execute(d_out[s_].data_ptr(), d_src[s_].data_ptr(), st)

    with torch.cuda.stream(side):
      gstep(side.cuda_stream)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
      gstep(side.cuda_stream)
    for _ in range(3):
      graph.replay()
    dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
      graph.replay()
    g1.record()
    dist.barrier()
    (graph_ms,) = dist.max(g0.elapsed_time(g1) / steps)
    del graph
    gplan.close()
  except Exception as exc:  # pylint: disable=broad-except
    graph_ms = f"unavailable: {exc}"
  (ms_total,) = dist.max(ms_total)
  res = dict(M=M, rank=rank, T=T, sets=sets, grid=grid, ttype=ttype, ms_per_step=ms_total / steps, stage=stage,
             graph_ms_per_step=graph_ms,
             launches=int(launches), allocs_in_timed_region=int(allocs), info=info,
             fine=[int(x) for x in info.fine_dims[:rank]], pts_np=pts_np)
  if want_out:
    res["out"] = d_out[0]
    res["plan"] = plan
  else:
    plan.close()
  return res


def e2e_arm(cfg, T, dist, steps, warmup, seed):
  """Public API, pinned host tensors in, host tensor out: H2D + D2H inside the timed region."""
  import torch
  import tensorflow_nufft_b200 as tfft
  pts_np = cfg["points"]()
  M, rank = pts_np.shape
  grid = cfg["grid"]
  ttype = cfg["ttype"]
  sets = cfg.get("sets", 1)
  if sets > 1:
    pts_np = np.stack([np.roll(pts_np, s, axis=0) for s in range(sets)])
    src_shape = (sets, M) if ttype == 1 else (sets,) + tuple(grid)
  else:
    src_shape = (T, M) if ttype == 1 else (T,) + tuple(grid)
  h_pts = torch.from_numpy(pts_np).pin_memory()
  h_src = torch.from_numpy(H.random_complex(src_shape, seed)).pin_memory()

  def e2e_step():
    return tfft.nufft(h_src, h_pts, grid_shape=grid, transform_type=f"type_{ttype}",
                      fft_direction=cfg["direction"], tol=TOL)

  for _ in range(max(1, min(warmup, 3))):
    res = e2e_step()
  dist.barrier()
  t0 = time.perf_counter()
  for _ in range(steps):
    res = e2e_step()
  torch.cuda.synchronize()
  t_e2e = time.perf_counter() - t0
  dist.barrier()
  (ms,) = dist.max(t_e2e * 1e3)
  h2d = h_pts.numel() * h_pts.element_size() + h_src.numel() * h_src.element_size()
  d2h = res.numel() * res.element_size()
  return ms / steps, h2d, d2h


FFT_METHODS = {1: "cuFFT", 2: "cuFFT, three-plan pruned scheme", 3: "own pruned passes, amplify/deconvolve fused (in fft_ms)"}


def roofline_block(name, r, peak, peak_src):
  info = r["info"]
  nf_tot = int(info.fine_dims[0]) * int(info.fine_dims[1]) * int(info.fine_dims[2])
  per_exec = (r["T"] + info.batch_size - 1) // info.batch_size    # spread|interp launches of one execute
  n_launch = r["sets"] * per_exec
  coils_per_launch = min(r["T"], info.batch_size)
  bytes_per_launch = spread_interp_bytes(r["rank"], r["M"], nf_tot, coils_per_launch)
  launch_ms = r["stage"]["spread_interp_ms"] / per_exec             # stage time = sum over one execute's launches
  achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
  if os.path.exists(tpath):
    with open(tpath) as f:
      traffic = json.load(f).get("dram_bytes_per_launch")
  return {"bound": "hbm", "kernel": kernel_name(r["ttype"], r["rank"]),
          "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
          "traffic": traffic, "peak_source": peak_src,
          "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
          "launches_per_step": n_launch,
          "note": ("type-1 spreading is bound by on-chip read-modify-write / LSU throughput, not HBM" if r["ttype"] == 1 else
                   "type-2 gathering is bound by shared-memory load bandwidth (ns^d cells per point) and, for sparse "
                   "point sets, by L2->shared tile traffic, not HBM")}


def strong_arm(name, total_coils, dist, world, rank_id, device, steps, warmup):
  """Strong scaling: a fixed job of `total_coils` transforms split T/N per rank; every rank runs the
  full set_points (the serial fraction). The final all-gather of the result slabs over NVLink is
  timed on its own."""
  import torch
  from tensorflow_nufft_b200 import sharding
  cfg = CONFIGS[name]
  b, e = sharding.shard_bounds(total_coils, world, rank_id)
  r = device_arm(name, cfg, e - b, dist, device, steps, warmup, 3000 + rank_id, want_out=True)
  gather_ms = 0.0
  if world > 1:
    out = r["out"]
    for _ in range(2):
      full = sharding.gather_slabs(out, total_coils)
    dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
      full = sharding.gather_slabs(out, total_coils)
    g1.record()
    dist.barrier()
    (gather_ms,) = dist.max(g0.elapsed_time(g1) / steps)
    assert full.shape[0] == total_coils
    del full
  r["plan"].close()
  units = total_coils * r["M"]
  return {"workload": f"{name}: {total_coils} coils in total, {e - b} on rank 0", "coils_total": total_coils,
          "ms_per_step": r["ms_per_step"], "value": units / (r["ms_per_step"] * 1e-3),
          "gather_ms": gather_ms, "value_with_gather": units / ((r["ms_per_step"] + gather_ms) * 1e-3),
          "gather_bytes": int(total_coils * (np.prod(cfg["grid"]) if cfg["ttype"] == 1 else r["M"]) * 8),
          "stages_ms": {k: round(v, 4) for k, v in r["stage"].items()},
          "serial_fraction_set_points": r["stage"]["set_points_ms"] / max(r["ms_per_step"], 1e-9),
          "unit": "points/s", "scaling": "strong"}


def run_ours(args, rank_id, world, device):
  import torch
  import tensorflow_nufft_b200 as tfft

  torch.cuda.set_device(device)
  dist = Dist(world)
  cfg = CONFIGS[args.config]
  ncores = os.cpu_count() or 1
  tfft.set_engine_defaults(num_threads_compat=ncores)
  tfft.set_points_reuse(False)   # a "step" includes set_points
  peak, peak_src = load_peaks()
  T = cfg["coils"]

  sampler = ClockSampler(device)
  sampler.start()
  r = device_arm(args.config, cfg, T, dist, device, args.steps, args.warmup, 1000 + rank_id)
  ms_e2e, h2d, d2h = e2e_arm(cfg, T, dist, args.steps, args.warmup, 1000 + rank_id)
  clocks = sampler.stop()

  units = world * r["sets"] * T * r["M"]
  line = {
      "metric": "NU points/sec", "value": units / (r["ms_per_step"] * 1e-3), "unit": "points/s", "n_gpus": world,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": f"{args.config}: {cfg['desc']}", "coils_per_gpu": T, "points": r["M"],
                 "point_sets_per_step": r["sets"], "grid": list(r["grid"]), "fine_grid": r["fine"], "tol": TOL,
                 "kernel_width": r["info"].kernel_width, "parallelism": f"batch-shard x{world}",
                 "fft": FFT_METHODS.get(r["info"].fft_method, "?"),
                 "l2": "inputs larger than L2 (no flush needed)" if h2d > 126e6 else "working set below L2 size",
                 "step": "set_points + execute, all coils"},
      "stages_ms": {k: round(v, 4) for k, v in r["stage"].items()},
      "roofline": roofline_block(args.config, r, peak, peak_src),
      "e2e": {"value": units / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d,
              "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
      "gpu_launches": r["launches"],
      "cuda_graph_ms_per_step": r["graph_ms_per_step"],
      "device_allocations_in_timed_region": r["allocs_in_timed_region"],
      "clocks": clocks,
  }
  if args.config == "cfg2" and not args.only_main:
    # ---- type-2 block: cfg4 as north_star shards it (2 coils per GPU) ----
    c4 = CONFIGS["cfg4"]
    r4 = device_arm("cfg4", c4, c4["coils"], dist, device, args.steps, args.warmup, 2000 + rank_id)
    ms4, h2d4, d2h4 = e2e_arm(c4, c4["coils"], dist, max(2, args.steps // 2), 3, 2000 + rank_id)
    u4 = world * c4["coils"] * r4["M"]
    line["type2"] = {
        "workload": f"cfg4: {c4['desc']}", "value": u4 / (r4["ms_per_step"] * 1e-3), "unit": "points/s",
        "ms_per_step": r4["ms_per_step"], "scaling": "weak", "coils_per_gpu": c4["coils"], "fine_grid": r4["fine"],
        "kernel_width": r4["info"].kernel_width, "upsampling_factor": r4["info"].upsampling_factor,
        "fft": FFT_METHODS.get(r4["info"].fft_method, "?"),
        "stages_ms": {k: round(v, 4) for k, v in r4["stage"].items()},
        "roofline": roofline_block("cfg4", r4, peak, peak_src),
        "e2e": {"value": u4 / (ms4 * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d4,
                "d2h_bytes_per_step": d2h4, "ms_per_step": ms4},
        "gpu_launches": r4["launches"]}
    # ---- strong scaling of the batched configs (T/N coils per rank) + the final gather ----
    line["strong"] = {n: strong_arm(n, STRONG[n], dist, world, rank_id, device, args.steps, args.warmup)
                      for n in ("cfg2", "cfg4")}
  return line


def run_cpu_reference(cfg, steps, warmup, mode="auto", sample_coils=None, budget_s=150.0):
  """Times the reference CPU plan (libref.so) with all host threads: set_points + execute per step,
  with the FFT share read from the shim's timer. Coils are reduced only if the run would not fit
  `budget_s`."""
  from oracle import ref as oref
  if not oref.available():
    return None
  L = oref.lib()
  L.ref_fft_seconds.restype = ctypes.c_double
  L.ref_fft_seconds.argtypes = [ctypes.c_int]
  pts_np = cfg["points"]()
  M, rank = pts_np.shape
  grid = cfg["grid"]
  T_full = cfg["coils"]
  T = sample_coils or T_full
  ttype = cfg["ttype"]
  sign = -1 if cfg["direction"] == "forward" else 1
  N = int(np.prod(grid))
  plan_pts = np.ascontiguousarray(pts_np[:, ::-1].T)
  ncores = os.cpu_count() or 1

  def one(Tn, src):
    rp = oref.RefPlan(ttype, list(grid[::-1]), sign, Tn, TOL, np.complex64, mode=mode, num_threads=ncores)
    L.ref_fft_seconds(1)
    t1 = time.perf_counter()
    rp.set_points(plan_pts)
    t2 = time.perf_counter()
    rp.execute(src)
    t3 = time.perf_counter()
    fft = L.ref_fft_seconds(1)
    info = (rp.kernel_width, rp.sigma, list(rp.fine_dims))
    rp.close()
    return t2 - t1, t3 - t2, fft, info

  src = H.random_complex((T, M) if ttype == 1 else (T, N), 2000)
  sp, ex, fft, info = one(T, src)          # first (cold) step: also sizes the run
  if (sp + ex) * (steps + warmup) > budget_s and T > 1:
    T = max(1, int(T * budget_s / ((sp + ex) * (steps + warmup))))
    src = src[:T].copy()
  rows = []
  for it in range(warmup + steps):
    row = one(T, src)
    if it >= warmup:
      rows.append(row)
  sp = float(np.mean([r[0] for r in rows]))
  ex = float(np.mean([r[1] for r in rows]))
  fft = float(np.mean([r[2] for r in rows]))
  mean = sp + ex
  best = min(r[0] + r[1] for r in rows)
  return {"value": T * M / mean, "best": T * M / best, "unit": "points/s", "cores": ncores, "kind": "reference",
          "sample": f"{T} of {T_full} coils, all {M} points, set_points+execute, reference CPU plan mode={mode} "
                    f"(kernel width {info[0]}, sigma {info[1]}, fine grid {info[2]}); FFT = oracle/fft235.c "
                    "(SIMD Stockham, not FFTW)",
          "ms_per_step": mean * 1e3, "steps": steps, "warmup": warmup, "coils": T,
          "stages_ms": {"set_points_ms": sp * 1e3, "fft_ms": fft * 1e3, "spread_interp_deconv_ms": (ex - fft) * 1e3},
          "fft_share": fft / mean}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--only-main", action="store_true", help="skip the type2 / strong-scaling blocks")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
  cfg = CONFIGS[args.config]

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank_id = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))

  if args.impl == "reference":
    if rank_id != 0:
      return
    base = run_cpu_reference(cfg, max(1, args.steps), max(0, args.warmup), mode="auto")
    if base is None:
      print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so not built (needs /root/reference)"}))
      return
    line = {"impl": "reference", "metric": "NU points/sec", "value": base["value"], "unit": "points/s",
            "n_gpus": args.gpus, "steps": base["steps"], "warmup": base["warmup"], "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['desc']}", "coils_per_gpu": base["coils"],
                       "note": "reference OpenMP CPU plan, host cores only"},
            "stages_ms": base["stages_ms"],
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "stages_ms", "fft_share")},
            "e2e": {"value": base["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return

  import torch
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  line = run_ours(args, rank_id, world, local_rank)
  if rank_id == 0:
    if world == 1 and not args.no_cpu_baseline:
      base = run_cpu_reference(cfg, 2, 1, mode="auto", sample_coils=min(cfg["coils"], 8), budget_s=30.0)
      line["cpu_baseline"] = ({k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "stages_ms", "fft_share")}
                              if base else
                              {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "unavailable: oracle/_ref/libref.so not built"})
    print(json.dumps(line))
  if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
  main()
