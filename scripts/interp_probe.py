"""GPU probe: type-2 interpolator variants (stage time + agreement with the first variant)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, pts, T, variants, reps=5, tol=1e-6):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  f = torch.from_numpy(H.random_complex((T, N), 2)).cuda()
  ref = None
  for v in variants:
    c = torch.zeros((T, M), dtype=torch.complex64, device="cuda")
    kw = dict(interp_method=v["method"], profile=1)
    if "bins" in v: kw["bin_dims"] = v["bins"]
    if "msub" in v: kw["max_subproblem_size"] = v["msub"]
    if "no_tma" in v: kw["no_tma"] = v["no_tma"]
    if "nc" in v: kw["coils_per_cta"] = v["nc"]
    if "no_zrange" in v: kw["no_zrange"] = v["no_zrange"]
    try:
      plan = _lib.Plan(2, grid[::-1], -1, T, tol, 0, device=0, **kw)
    except Exception as e:
      print(json.dumps({"case": name, **v, "error": str(e)[:200]}), flush=True)
      continue
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      plan.execute(c.data_ptr(), f.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]:
        best = t
    out = c.cpu().numpy()
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
  which = sys.argv[1:] or ["small", "cfg4", "cfg3", "cfg2", "cfg1"]
  if "small" in which:
    rng = np.random.default_rng(5)
    q = rng.uniform(-np.pi, np.pi, (50000, 2)).astype(np.float32)
    q[:64, 0] = np.float32(np.pi); q[64:128, 1] = -np.float32(np.pi); q[128:160] = 0
    for tol in (1e-6, 1e-3):
      run(f"odd2d-130x94-tol{tol}", (94, 130), q, 3, [dict(method=1), dict(method=2), dict(method=5), dict(method=5, no_tma=1), dict(method=5, msub=8), dict(method=5, bins=(16, 16)), dict(method=4), dict(method=4, no_tma=1), dict(method=4, msub=8)], reps=2, tol=tol)
    q3 = rng.uniform(-np.pi, np.pi, (40000, 3)).astype(np.float32)
    q3[:64, 0] = np.float32(np.pi); q3[64:128, 2] = -np.float32(np.pi); q3[128:160] = 0
    for tol in (1e-6, 1e-2):
      run(f"odd3d-30x44x26-tol{tol}", (26, 44, 30), q3, 3, [dict(method=1), dict(method=2), dict(method=5), dict(method=5, no_tma=1), dict(method=5, bins=(16, 16, 8)), dict(method=5, bins=(16, 8, 4), msub=12), dict(method=4), dict(method=4, no_tma=1), dict(method=4, bins=(16, 8, 4), msub=12)], reps=2, tol=tol)
    run("ext-range3d", (32, 32, 32), (q3 * 2.9).astype(np.float32), 2, [dict(method=1), dict(method=2), dict(method=5)], reps=2)
  if "cfg4" in which:
    p = H.stack_of_stars_points(125, 125, 256)
    run("cfg4-sos-256-T2", (256, 256, 256), p, 2,
        [dict(method=5), dict(method=4), dict(method=4, bins=(16, 16, 4)), dict(method=4, bins=(16, 16, 8)),
         dict(method=4, bins=(16, 8, 4)), dict(method=4, bins=(16, 8, 2)), dict(method=5, bins=(16, 8, 2))])
  if "cfg3" in which:
    p = H.uniform_points(8000000, 3, 3)
    run("cfg3-uniform-128-type2", (128, 128, 128), p, 1,
        [dict(method=5), dict(method=5, bins=(16, 8, 4)), dict(method=4), dict(method=4, bins=(16, 8, 4)), dict(method=5, bins=(16, 8, 2)), dict(method=5, bins=(16, 8, 8))])
  if "cfg2" in which:
    p = H.spiral_points(32, 62500)
    run("cfg2-spiral-512-T8-type2", (512, 512), p, 8,
        [dict(method=5), dict(method=4), dict(method=4, bins=(16, 16))])
  if "cfg1" in which:
    run("cfg1-radial-256", (256, 256), H.radial_points(200, 500), 1, [dict(method=2), dict(method=5), dict(method=5, bins=(16, 16))])
