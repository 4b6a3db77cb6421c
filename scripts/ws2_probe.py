"""GPU probe: window-sorted spreader generations on cfg2 (and odd cases): stage time + agreement."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, pts, T, variants, reps=6, tol=1e-6):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  c = torch.from_numpy(H.random_complex((T, M), 1)).cuda()
  ref = None
  for v in variants:
    f = torch.zeros((T, N), dtype=torch.complex64, device="cuda")
    kw = dict(spread_method=v["method"], profile=1)
    if "bins" in v: kw["bin_dims"] = v["bins"]
    if "nc" in v: kw["coils_per_cta"] = v["nc"]
    if "msub" in v: kw["max_subproblem_size"] = v["msub"]
    if "var" in v: kw["kernel_variant"] = v["var"]
    if "no_tma_flush" in v: kw["no_tma_flush"] = v["no_tma_flush"]
    plan = _lib.Plan(1, grid[::-1], 1, T, tol, 0, device=0, **kw)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      plan.execute(c.data_ptr(), f.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]:
        best = t
    out = f.cpu().numpy()
    if ref is None:
      ref = out
      err = 0.0
    else:
      err = H.rel_l2(out, ref)
    inf = plan.info()
    print(json.dumps({"case": name, **v, "bins_used": list(inf.bin_dims)[:len(grid)], "T": T, "M": M,
                      **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_first": err}), flush=True)
    plan.close()


if __name__ == "__main__":
  p = H.spiral_points(32, 62500)
  V = [dict(method=3), dict(method=4, no_tma_flush=1), dict(method=4), dict(method=4, nc=4), dict(method=4, nc=4, no_tma_flush=1),
       dict(method=4, bins=(16, 16)), dict(method=4, bins=(32, 8)), dict(method=4, bins=(16, 4)), dict(method=4, bins=(32, 16))]
  if len(sys.argv) > 1: V = V[:int(sys.argv[1])]
  run("cfg2-spiral-512-T32", (512, 512), p, 32, V)
  # agreement on awkward shapes: odd grid, narrow kernels, points on the fold boundaries
  rng = np.random.default_rng(5)
  q = rng.uniform(-np.pi, np.pi, (50000, 2)).astype(np.float32)
  q[:64, 0] = np.float32(np.pi); q[64:128, 1] = -np.float32(np.pi); q[128:160] = 0
  for tol in (1e-6, 1e-4, 1e-3, 1e-2):
    run(f"odd-130x94-tol{tol}", (94, 130), q, 4, [dict(method=2), dict(method=3), dict(method=4), dict(method=4, nc=1)], reps=2, tol=tol)
  run("ext-range", (64, 64), (q * 2.9).astype(np.float32), 8, [dict(method=2), dict(method=4), dict(method=4, nc=8)], reps=2)
