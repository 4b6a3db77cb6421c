#!/usr/bin/env python
"""bench.py -- NU points/sec of the tfft.nufft hot path on B200 (BASELINE.json metric).

Default workload (N=1): BASELINE config[1] -- 2D type-1 adjoint NUFFT, 512x512 grid, 32-coil batch
sharing 2M spiral points, complex64, tol 1e-6. One "step" = one full pass of the hot path over the
batch: set_points (fold, bin-sort, stencil records) + execute (spread, cuFFT, deconvolve) for all
coils. `value` = coils*M / step time with inputs resident in HBM; `e2e` = the same through the public
`tfft.nufft` call with pinned HOST tensors (H2D + D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg1]

N>1: launched by torchrun, one rank per GPU; every rank transforms its own 32-coil shard of a
32*N-coil batch with the same point set (weak scaling, no data-path collective); time = max over ranks.
`--impl reference` times the reference's own OpenMP CPU plan (oracle/_ref/libref.so, mode auto =
what tfft.nufft does on /cpu:0) on a bounded sample of the same workload.
"""
import argparse
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


CONFIGS = {
    # name: (transform_type, fft_direction, grid (TF order), coils per GPU, points generator, cpu sample coils)
    "cfg1": dict(ttype=2, direction="forward", grid=(256, 256), coils=1, cpu_coils=1,
                 points=lambda: H.radial_points(200, 500),
                 desc="2D type-2, 256x256, 100k radial points, 1 transform, complex64, tol 1e-6"),
    "cfg2": dict(ttype=1, direction="backward", grid=(512, 512), coils=32, cpu_coils=8,
                 points=lambda: H.spiral_points(32, 62500),
                 desc="2D type-1 adjoint, 512x512, 32 coils x 2M spiral points, complex64, tol 1e-6"),
    "cfg3": dict(ttype=1, direction="forward", grid=(128, 128, 128), coils=1, cpu_coils=1,
                 points=lambda: H.uniform_points(8000000, 3, 3),
                 desc="3D type-1, 128^3, 8M uniform-random points, 1 transform, complex64, tol 1e-6"),
    "cfg4": dict(ttype=2, direction="forward", grid=(256, 256, 256), coils=2, cpu_coils=1,
                 points=lambda: H.stack_of_stars_points(125, 125, 256),
                 desc="3D type-2, 256^3, 2 coils/GPU x 4M stack-of-stars points, complex64, tol 1e-6"),
}
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


def run_ours(args, cfg, rank_id, world, device):
  import torch
  import tensorflow_nufft_b200 as tfft
  from tensorflow_nufft_b200 import _lib

  torch.cuda.set_device(device)
  pts_np = cfg["points"]()
  M, rank = pts_np.shape
  grid = cfg["grid"]
  T = cfg["coils"]
  N = int(np.prod(grid))
  ttype = cfg["ttype"]
  sign = -1 if cfg["direction"] == "forward" else 1
  src_shape = (T, M) if ttype == 1 else (T,) + tuple(grid)
  src_np = H.random_complex(src_shape, 1000 + rank_id)
  ncores = os.cpu_count() or 1
  tfft.set_engine_defaults(num_threads_compat=ncores)
  # a "step" includes set_points: the unchanged-points shortcut of the Python mirror stays off
  tfft.set_points_reuse(False)

  # ---------------- device-resident arm: C ABI, inputs already in HBM ----------------
  d_pts = torch.from_numpy(pts_np).cuda()
  d_src = torch.from_numpy(src_np).cuda()
  d_out = torch.empty((T, N) if ttype == 1 else (T, M), dtype=torch.complex64, device="cuda")
  plan = _lib.Plan(ttype, tuple(reversed(grid)), sign, T, float(np.float32(TOL)), _lib.COMPLEX64, device=device,
                   profile=1, num_threads_compat=ncores)
  stream = torch.cuda.current_stream().cuda_stream

  def step():
    plan.set_points_interleaved(M, d_pts.data_ptr(), stream)
    if ttype == 1:
      plan.execute(d_src.data_ptr(), d_out.data_ptr(), stream)
    else:
      plan.execute(d_out.data_ptr(), d_src.data_ptr(), stream)

  def barrier():
    if world > 1:
      torch.distributed.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    step()
  barrier()
  sampler = ClockSampler(device)
  sampler.start()
  l0 = plan.launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  stage = {"spread_interp_ms": 0.0, "fft_ms": 0.0, "deconv_ms": 0.0, "set_points_ms": 0.0}
  barrier()
  e0.record()
  for _ in range(args.steps):
    step()
  e1.record()
  barrier()
  ms_total = e0.elapsed_time(e1)
  launches = plan.launch_count() - l0
  # per-stage CUDA events of the last timed step (recorded on the launching stream)
  tm = plan.timings()
  for k in stage:
    stage[k] = tm[k]
  info = plan.info()

  # ---------------- end-to-end arm: public API, pinned host tensors ----------------
  h_pts = torch.from_numpy(pts_np).pin_memory()
  h_src = torch.from_numpy(src_np).pin_memory()

  def e2e_step():
    return tfft.nufft(h_src, h_pts, grid_shape=grid, transform_type=f"type_{ttype}",
                      fft_direction=cfg["direction"], tol=TOL)

  for _ in range(max(1, min(args.warmup, 3))):
    res = e2e_step()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    res = e2e_step()
  torch.cuda.synchronize()
  t_e2e = time.perf_counter() - t0
  barrier()
  clocks = sampler.stop()
  h2d = h_pts.numel() * h_pts.element_size() + h_src.numel() * h_src.element_size()
  d2h = res.numel() * res.element_size()

  # max over ranks
  times = torch.tensor([ms_total, t_e2e * 1e3], dtype=torch.float64, device="cuda")
  if world > 1:
    torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
  ms_total, ms_e2e = float(times[0]), float(times[1])

  ms_per_step = ms_total / args.steps
  total_units = world * T * M
  value = total_units / (ms_per_step * 1e-3)
  e2e_value = total_units / (ms_e2e / args.steps * 1e-3)

  nf_tot = int(info.fine_dims[0]) * int(info.fine_dims[1]) * int(info.fine_dims[2])
  n_launch = (T + info.batch_size - 1) // info.batch_size
  coils_per_launch = min(T, info.batch_size)
  bytes_per_launch = spread_interp_bytes(rank, M, nf_tot, coils_per_launch)
  launch_ms = stage["spread_interp_ms"] / n_launch
  peak, peak_src = load_peaks()
  achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, "profiles", f"traffic_{args.config}.json")
  if os.path.exists(tpath):
    with open(tpath) as f:
      traffic = json.load(f).get("dram_bytes_per_launch")

  line = {
      "metric": "NU points/sec", "value": value, "unit": "points/s", "n_gpus": world,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": f"{args.config}: {cfg['desc']}", "coils_per_gpu": T, "points": M,
                 "grid": list(grid), "fine_grid": [int(x) for x in info.fine_dims[:rank]], "tol": TOL,
                 "kernel_width": info.kernel_width, "parallelism": f"batch-shard x{world}",
                 "l2": "inputs larger than L2 (no flush needed)" if h2d > 126e6 else "working set below L2 size",
                 "step": "set_points + execute, all coils"},
      "stages_ms": {k: round(v, 4) for k, v in stage.items()},
      "roofline": {"bound": "hbm", "kernel": ("spread_ws2_f32_kernel" if rank == 2 else "spread_tile_f32_kernel") if ttype == 1 else "interp_qw_f32_kernel",
                   "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                   "traffic": traffic, "peak_source": peak_src,
                   "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": launch_ms,
                   "launches_per_step": n_launch,
                   "note": ("type-1 spreading is bound by shared-memory read-modify-write bandwidth, not HBM" if ttype == 1 else
                            "type-2 gathering is bound by shared-memory load bandwidth (ns^d cells per point) and, for sparse "
                            "point sets, by L2->shared tile traffic, not HBM")},
      "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
              "ms_per_step": ms_e2e / args.steps},
      "gpu_launches": int(launches),
      "clocks": clocks,
  }
  plan.close()
  return line


def run_cpu_reference(cfg, steps, warmup, mode="auto", sample_coils=None):
  """Times the reference CPU plan (libref.so) on a bounded sample: same points, fewer coils."""
  from oracle import ref as oref
  if not oref.available():
    return None
  pts_np = cfg["points"]()
  M, rank = pts_np.shape
  grid = cfg["grid"]
  T = sample_coils or cfg["cpu_coils"]
  ttype = cfg["ttype"]
  sign = -1 if cfg["direction"] == "forward" else 1
  N = int(np.prod(grid))
  src = H.random_complex((T, M) if ttype == 1 else (T, N), 2000)
  plan_pts = np.ascontiguousarray(pts_np[:, ::-1].T)
  ncores = os.cpu_count() or 1
  best = None
  times = []
  for it in range(warmup + steps):
    t0 = time.perf_counter()
    rp = oref.RefPlan(ttype, list(grid[::-1]), sign, T, TOL, np.complex64, mode=mode, num_threads=ncores)
    t1 = time.perf_counter()
    rp.set_points(plan_pts)
    rp.execute(src)
    t2 = time.perf_counter()
    rp.close()
    if it >= warmup:
      times.append(t2 - t1)
    del t0
  best = min(times)
  mean = float(np.mean(times))
  return {"value": T * M / mean, "best": T * M / best, "unit": "points/s", "cores": ncores, "kind": "reference",
          "sample": f"{T} of {cfg['coils']} coils, all {M} points, set_points+execute, reference CPU plan mode={mode} "
                    "(FFT = oracle/fft235.c, not FFTW)", "ms_per_step": mean * 1e3}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  args = ap.parse_args()
  cfg = CONFIGS[args.config]

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank_id = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))

  if args.impl == "reference":
    if rank_id != 0:
      return
    steps = max(1, min(args.steps, 3))
    base = run_cpu_reference(cfg, steps, min(args.warmup, 1), mode="auto")
    if base is None:
      print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so not built (needs /root/reference)"}))
      return
    line = {"impl": "reference", "metric": "NU points/sec", "value": base["value"], "unit": "points/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['desc']}", "note": "reference OpenMP CPU plan, host cores only"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
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
  line = run_ours(args, cfg, rank_id, world, local_rank)
  if rank_id == 0:
    if world == 1 and not args.no_cpu_baseline:
      base = run_cpu_reference(cfg, 2, 1, mode="auto")
      line["cpu_baseline"] = ({k: base[k] for k in ("value", "unit", "cores", "kind", "sample")} if base else
                              {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "unavailable: oracle/_ref/libref.so not built"})
    print(json.dumps(line))
  if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
  main()
