"""Scratch performance probe (GPU box): stage timings of the BASELINE configs per kernel variant."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, ttype, grid, pts, T, methods, reps=5, bin_dims=None, msub=0):
  rank = pts.shape[1]
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  c = torch.from_numpy(H.random_complex((T, M), 1)).cuda()
  f = torch.from_numpy(H.random_complex((T, N), 2)).cuda()
  for meth in methods:
    kw = dict(spread_method=min(meth, 3) if ttype == 1 else 0, interp_method=meth if ttype == 2 else 0, profile=1, max_subproblem_size=msub)
    if bin_dims: kw["bin_dims"] = bin_dims
    plan = _lib.Plan(ttype, grid[::-1], -1, T, 1e-6, 0, device=0, **kw)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      plan.execute(c.data_ptr(), f.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]:
        best = t
    inf = plan.info()
    # timings cover the LAST batch only
    nb = min(T, inf.batch_size)
    last = T - ((T - 1) // nb) * nb
    pts_s = last * M / (best["spread_interp_ms"] * 1e-3)
    print(json.dumps({"case": name, "type": ttype, "method": meth, "bins": list(inf.bin_dims), "nf": list(inf.fine_dims),
                      "M": M, "T": T, "last_batch": last, **{k: round(v, 4) for k, v in best.items()},
                      "Gpts_per_s_stage": round(pts_s / 1e9, 3)}), flush=True)
    plan.close()


if __name__ == "__main__":
  which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4"]
  if "cfg1" in which:
    run("cfg1-radial-256", 2, (256, 256), H.radial_points(200, 500), 1, (1, 2, 3))
    run("cfg1-radial-256-t1", 1, (256, 256), H.radial_points(200, 500), 1, (1, 2, 3))
  if "cfg2" in which:
    p = H.spiral_points(32, 62500)
    run("cfg2-spiral-512-T8", 1, (512, 512), p, 8, (2, 3))
    run("cfg2-spiral-512-T8-type2", 2, (512, 512), p, 8, (2, 3))
  if "cfg3" in which:
    p = H.uniform_points(8000000, 3, 3)
    run("cfg3-uniform-128", 1, (128, 128, 128), p, 1, (2,), bin_dims=(16, 16, 2))
    run("cfg3-uniform-128-ws-16x16x4", 1, (128, 128, 128), p, 1, (3,), bin_dims=(16, 16, 4))
    run("cfg3-uniform-128-ws-16x16x8", 1, (128, 128, 128), p, 1, (3,), bin_dims=(16, 16, 8))
    run("cfg3-uniform-128-ws-16x8x8", 1, (128, 128, 128), p, 1, (3,), bin_dims=(16, 8, 8))
    run("cfg3-uniform-128-ws-32x8x8", 1, (128, 128, 128), p, 1, (3,), bin_dims=(32, 8, 8))
    run("cfg3-uniform-128-type2", 2, (128, 128, 128), p, 1, (2, 3))
    run("cfg3-uniform-128-type2-bin2", 2, (128, 128, 128), p, 1, (2,), bin_dims=(16, 16, 2))
  if "cfg4" in which:
    p = H.stack_of_stars_points(125, 125, 256)
    run("cfg4-sos-256-T2", 2, (256, 256, 256), p, 2, (2, 3))
    run("cfg4-sos-256-T2-bin8", 2, (256, 256, 256), p, 2, (3,), bin_dims=(16, 16, 8))
    run("cfg4-sos-256-T2-type1", 1, (256, 256, 256), p, 2, (2,), bin_dims=(16, 16, 2))
    run("cfg4-sos-256-T2-type1-ws", 1, (256, 256, 256), p, 2, (3,))
