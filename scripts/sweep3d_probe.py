"""GPU probe: 3D sweep spreader (method 7) against the private-tile spreader (method 2) on cfg3 and
the sparse cfg4 point set, plus agreement on awkward shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, pts, T, variants, reps=4, tol=1e-6):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  c = torch.from_numpy(H.random_complex((T, M), 1)).cuda()
  ref = None
  for v in variants:
    f = torch.zeros((T, N), dtype=torch.complex64, device="cuda")
    kw = dict(spread_method=v["method"], profile=1)
    if "bins" in v: kw["bin_dims"] = v["bins"]
    for k in ("max_subproblem_size", "no_pack", "no_tma_flush", "no_zrange", "otf_weights", "no_preclear"):
      if k in v: kw[k] = v[k]
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
    print(json.dumps({"case": name, **v, "method_used": inf.spread_method, "bins_used": list(inf.bin_dims)[:len(grid)], "T": T, "M": M,
                      **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_first": err,
                      "finite": bool(np.isfinite(out).all())}), flush=True)
    plan.close()


if __name__ == "__main__":
  mode = sys.argv[1] if len(sys.argv) > 1 else "all"
  if mode == "prof":
    run("cfg3", (128, 128, 128), H.uniform_points(8000000, 3, 3), 1, [dict(method=7)], reps=3)
    sys.exit(0)
  V = [dict(method=2), dict(method=7), dict(method=7, bins=(8, 8, 16)), dict(method=7, bins=(8, 8, 8)),
       dict(method=7, bins=(8, 8, 32)), dict(method=7, bins=(8, 16, 16)), dict(method=7, bins=(24, 8, 16))]
  run("cfg3-uniform-128^3-8M", (128, 128, 128), H.uniform_points(8000000, 3, 3), 1, V)
  run("cfg4adj-sos-256^3-4M", (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 1, V[:4] + [dict(method=7, no_zrange=1)], reps=3)
  run("ref8-uniform-128^3-800k", (128, 128, 128), H.uniform_points(800000, 3, 18), 1, V[:4], reps=3)
  rng = np.random.default_rng(5)
  q = rng.uniform(-np.pi, np.pi, (60000, 3)).astype(np.float32)
  q[:64, 0] = np.float32(np.pi); q[64:128, 1] = -np.float32(np.pi); q[128:192, 2] = np.float32(np.pi); q[192:224] = 0
  for tol in (1e-6, 1e-4, 1e-3, 1e-2):
    for T in (1, 3):
      run(f"odd-34x26x30-tol{tol}", (30, 26, 34), q, T, [dict(method=1), dict(method=2), dict(method=7), dict(method=7, bins=(16, 8, 2)),
                                                   dict(method=7, no_tma_flush=1)], reps=1, tol=tol)
  run("ext-range", (32, 32, 32), (q * 2.9).astype(np.float32), 2, [dict(method=1), dict(method=7)], reps=1)
  run("tiny-grid", (8, 10, 12), q[:3000], 2, [dict(method=1), dict(method=7)], reps=1)
  run("sparse", (128, 128, 128), q[:2000], 2, [dict(method=1), dict(method=7)], reps=1)
