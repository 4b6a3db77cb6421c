"""Randomised differential test (GPU): the default engine (tile kernels, own pruned FFT passes where
eligible) vs the generic point-driven kernels (spread_method = interp_method = 1) with one full cuFFT
plan (fft_mode = 1) on random ranks, grids (a third of them powers of two, so that the own FFT
runs), tolerances, coil counts, precisions and point distributions (uniform, clustered, on-grid,
fold-boundary values)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200.python.ops import nufft_ops
from tests import helpers as H

def points(rng, kind, M, rank, rdtype):
  if kind == "uniform":
    p = rng.uniform(-np.pi, np.pi, (M, rank))
  elif kind == "cluster":
    c = rng.uniform(-np.pi, np.pi, (1, rank))
    p = c + rng.normal(0, 0.05, (M, rank))
    p = (p + np.pi) % (2 * np.pi) - np.pi
  elif kind == "ongrid":
    n = rng.integers(4, 40)
    p = (rng.integers(0, n, (M, rank)) / n) * 2 * np.pi - np.pi
  elif kind == "planes":   # few distinct values in the last coordinate (stack-of-stars like)
    p = rng.uniform(-np.pi, np.pi, (M, rank))
    p[:, 0] = rng.choice(np.linspace(-np.pi, np.pi, 7, endpoint=False), M)
  else:  # edges
    p = rng.uniform(-np.pi, np.pi, (M, rank))
    vals = np.array([np.pi, -np.pi, 0.0, np.nextafter(np.pi, 0), -np.nextafter(np.pi, 0)])
    k = min(M, 200)
    p[:k] = rng.choice(vals, (k, rank))
  return p.astype(rdtype)

def main(n_cases, seed, gmax2=70, gmax3=28):
  rng = np.random.default_rng(seed)
  worst = 0.0
  for case in range(n_cases):
    rank = int(rng.choice([2, 2, 3]))
    cdtype = np.complex64 if rng.random() < 0.7 else np.complex128
    rdtype = np.float32 if cdtype == np.complex64 else np.float64
    tol = float(rng.choice([1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7] if cdtype == np.complex64 else [1e-4, 1e-6, 1e-8, 1e-10, 1e-12, 1e-13]))
    grid = tuple(int(rng.integers(3, gmax2 if rank == 2 else gmax3)) for _ in range(rank))
    if rng.random() < 0.33:   # power-of-two modes: fine sizes 64 .. 1024, the own FFT passes
      grid = tuple(int(rng.choice([32, 64, 128, 256, 512] if rank == 2 else [32, 32, 64, 128])) for _ in range(rank))
      if rank == 3 and np.prod(grid) > 64 * 64 * 64: grid = (32, 64, 32)
    M = int(rng.choice([1, 7, 33, 500, 5000, 30000]))
    T = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 16, 33]))
    if rank == 3: T = min(T, 5)
    ttype = int(rng.choice([1, 2]))
    kind = str(rng.choice(["uniform", "cluster", "ongrid", "planes", "edges"]))
    pts = points(rng, kind, M, rank, rdtype)
    src = H.random_complex((T, M) if ttype == 1 else (T,) + grid, 1000 + case, cdtype)
    outs = []
    for meth in (1, 0):
      out = nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid, f"type_{ttype}",
                              "backward" if case % 2 else "forward", tol, None, "nufft",
                              engine_kwargs={"spread_method": meth, "interp_method": meth, "fft_mode": meth})
      outs.append(out.cpu().numpy())
    err = H.rel_l2(outs[1], outs[0])
    # two float32 summation orders of up to M terms per cell (clustered points) differ by ~sqrt(M) ulp
    gate = (2e-6 if cdtype == np.complex64 else 1e-12) * max(1.0, (M / 2000.0) ** 0.5)
    worst = max(worst, err / gate)
    status = "ok" if err <= gate and np.isfinite(outs[1]).all() else "FAIL"
    if status != "ok" or case % 20 == 0:
      print(json.dumps({"case": case, "rank": rank, "dtype": np.dtype(cdtype).name, "tol": tol, "grid": grid, "M": M, "T": T,
                        "type": ttype, "points": kind, "rel_l2": err, "status": status}), flush=True)
    assert status == "ok", "kernel disagreement"
  print(json.dumps({"fuzz": "ok", "cases": n_cases, "seed": seed, "worst_err_over_gate": worst}))

if __name__ == "__main__":
  main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0,
       int(sys.argv[3]) if len(sys.argv) > 3 else 70, int(sys.argv[4]) if len(sys.argv) > 4 else 28)
