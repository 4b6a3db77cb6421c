"""GPU probe: how much do the stages of two coil groups overlap when they run on two streams?
Execute-only time of ONE plan with T coils vs TWO plans with T/2 coils on two streams (spread /
interp is bound by the SM's load-store pipes and latency, the FFT by HBM: different resources)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, ttype, grid, pts, T, reps=8):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  gen = torch.Generator(device="cuda").manual_seed(1)
  src = torch.view_as_complex(torch.rand(((T, M) if ttype == 1 else (T, N)) + (2,), generator=gen, device="cuda") - 0.5)
  out = torch.empty((T, N) if ttype == 1 else (T, M), dtype=torch.complex64, device="cuda")
  res = {}
  for groups in (1, 2, 4):
    if T % groups: continue
    Tg = T // groups
    plans = [_lib.Plan(ttype, grid[::-1], -1, Tg, float(np.float32(1e-6)), _lib.COMPLEX64, device=0) for _ in range(groups)]
    streams = [torch.cuda.Stream() for _ in range(groups)]
    main = torch.cuda.current_stream()
    for pl in plans:
      pl.set_points_interleaved(M, dp.data_ptr(), main.cuda_stream)
    torch.cuda.synchronize()

    def step():
      ev = torch.cuda.Event()
      ev.record(main)
      for g, (pl, s) in enumerate(zip(plans, streams)):
        s.wait_event(ev)
        a, b = src[g * Tg:(g + 1) * Tg], out[g * Tg:(g + 1) * Tg]
        if ttype == 1: pl.execute(a.data_ptr(), b.data_ptr(), s.cuda_stream)
        else: pl.execute(b.data_ptr(), a.data_ptr(), s.cuda_stream)
      for s in streams:
        main.wait_stream(s)

    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): step()
    e1.record()
    torch.cuda.synchronize()
    res[f"groups{groups}_ms"] = round(e0.elapsed_time(e1) / reps, 4)
    for pl in plans: pl.close()
  print(json.dumps({"case": name, "T": T, "M": M, **res}), flush=True)


if __name__ == "__main__":
  run("cfg2 type-1 512^2 32 coils", 1, (512, 512), H.spiral_points(32, 62500), 32)
  run("cfg2 mirrored type-2", 2, (512, 512), H.spiral_points(32, 62500), 32)
  run("cfg4 type-2 256^3 2 coils", 2, (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 2)
  run("cfg4 type-2 256^3 4 coils", 2, (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 4)
  run("cfg4 set as type-1, 4 coils", 1, (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 4)
