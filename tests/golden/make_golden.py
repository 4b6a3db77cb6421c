"""Generates tests/golden/*.npz from the reference's OWN CPU plan (oracle/_ref/libref.so, compiled
unmodified from /root/reference by oracle/ref_build/Makefile). Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py
The vectors travel to the GPU box, where the reference tree does not exist. Everything is seeded.
NUM_THREADS is the `options.num_threads` value the reference used for its chunked float
deconvolution factors (nufft_util.cc:94-116); consumers must use the same value.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_THREADS = 8

NUFFT_CASES = [
    # name, rank, grid (x-fastest), M, T, type, dtype, tol, sign, points_range, spread
    ("t2_2d_c64", (16, 20), 400, 2, 2, np.complex64, 1e-6, -1, "extended", np.pi),
    ("t1_2d_c64", (16, 20), 400, 2, 1, np.complex64, 1e-6, +1, "extended", np.pi),
    ("t2_3d_c64", (8, 10, 12), 300, 1, 2, np.complex64, 1e-6, -1, "extended", np.pi),
    ("t1_3d_c64", (8, 10, 12), 300, 2, 1, np.complex64, 1e-6, -1, "extended", np.pi),
    ("t2_1d_c64", (40,), 200, 2, 2, np.complex64, 1e-6, +1, "extended", np.pi),
    ("t1_1d_c64", (40,), 200, 1, 1, np.complex64, 1e-6, -1, "extended", np.pi),
    ("t2_2d_c64_tol1e-3", (24, 24), 300, 1, 2, np.complex64, 1e-3, -1, "extended", np.pi),
    ("t1_2d_c64_tol1e-4", (24, 24), 300, 1, 1, np.complex64, 1e-4, -1, "extended", np.pi),
    ("t2_2d_c128", (16, 20), 300, 1, 2, np.complex128, 1e-12, -1, "extended", np.pi),
    ("t1_2d_c128", (16, 20), 300, 2, 1, np.complex128, 1e-12, +1, "extended", np.pi),
    ("t2_3d_c128", (8, 6, 10), 150, 1, 2, np.complex128, 1e-9, -1, "extended", np.pi),
    ("t1_3d_c128", (8, 6, 10), 150, 1, 1, np.complex128, 1e-9, -1, "extended", np.pi),
    ("t2_2d_c64_extended3pi", (16, 16), 300, 1, 2, np.complex64, 1e-6, -1, "extended", 3 * np.pi),
    ("t2_2d_c64_infinite", (16, 16), 300, 1, 2, np.complex64, 1e-6, -1, "infinite", 10 * np.pi),
    ("t1_2d_c64_strict", (16, 16), 300, 1, 1, np.complex64, 1e-6, -1, "strict", np.pi),
]


def main():
  rng = np.random.default_rng(20261017)
  # ---- parameters ----
  tols = [1e-1, 1e-2, 3e-3, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 5e-8, 1e-8, 1e-9, 1e-10, 1e-12, 1e-14]
  ptab = []
  for cd in (np.complex64, np.complex128):
    for tol in tols:
      p = ref.RefPlan(2, [64, 48], -1, 1, tol, cd, mode="gpuparams", num_threads=NUM_THREADS)
      ptab.append((int(cd == np.complex128), tol, p.kernel_width, p.beta, p.c, p.fine_dims[0], p.fine_dims[1]))
      p.close()
  smooth = np.array([(n, ref.next_smooth_int(n)) for n in
                     [1, 2, 3, 14, 15, 16, 17, 154, 270, 500, 512, 514, 640, 1001, 1024, 1025, 2047, 4100]], np.int64)
  # ---- deconvolution factors ----
  fser = {}
  for dt in (np.float32, np.float64):
    for nf in (32, 512, 540):
      for ns in (2, 4, 7, 14):
        if dt == np.float32 and ns > 9:
          continue
        for nt in (1, 8):
          bon = {2: 2.20, 3: 2.26, 4: 2.38}.get(ns, 2.30)
          beta = float(dt(bon) * dt(ns))
          c = float(dt(4.0 / dt(ns * ns)))
          fser[f"fser_{np.dtype(dt).name}_{nf}_{ns}_{nt}"] = ref.kernel_fseries(nf, ns, beta, c, nt, dt)
  scale = np.array([(int(dt == np.float64), rank, ns,
                     ref.scale_factor(rank, ns, float(dt({2: 2.20, 3: 2.26, 4: 2.38}.get(ns, 2.30)) * dt(ns)),
                                      float(dt(4.0 / dt(ns * ns))), dt))
                    for dt in (np.float32, np.float64) for rank in (1, 2, 3) for ns in (4, 5, 7)])
  np.savez_compressed(os.path.join(HERE, "params.npz"), num_threads=NUM_THREADS,
                      ptab=np.array(ptab, np.float64), smooth=smooth, scale=scale, **fser)

  # ---- fold + CPU bin-sort (deterministic compiled reference code) ----
  sort = {}
  for rank, n in ((2, 48), (3, 20)):
    for rd, cd in ((np.float32, np.complex64), (np.float64, np.complex128)):
      pts = rng.uniform(-np.pi, np.pi, (rank, 1500)).astype(rd)
      pts[:, :4] = np.array([[-np.pi, np.pi, 0.0, np.pi]] * rank, rd)
      for nthr in (1, 4):
        p = ref.RefPlan(2, [n] * rank, -1, 1, 1e-6, cd, mode="gpuparams", num_threads=nthr)
        p.set_points(pts)
        idx, did = p.sort_indices()
        assert did
        key = f"{rank}d_{np.dtype(rd).name}_thr{nthr}"
        sort[key + "_points"] = pts
        sort[key + "_folded"] = p.folded.copy()
        sort[key + "_idx"] = idx
        sort[key + "_nf"] = np.array(p.fine_dims, np.int32)
        p.close()
  np.savez_compressed(os.path.join(HERE, "sort.npz"), **sort)

  # ---- whole transforms ----
  out = {}
  for name, grid, M, T, ttype, cd, tol, sign, prange, spread in NUFFT_CASES:
    rank = len(grid)
    rd = np.float64 if cd == np.complex128 else np.float32
    pts = rng.uniform(-spread, spread, (rank, M)).astype(rd)
    N = int(np.prod(grid))
    shape = (T, M) if ttype == 1 else (T, N)
    src = (rng.uniform(-.5, .5, shape) + 1j * rng.uniform(-.5, .5, shape)).astype(cd)
    for mode in ("gpuparams", "auto"):
      p = ref.RefPlan(ttype, list(grid), sign, T, tol, cd, mode=mode, points_range=prange, num_threads=NUM_THREADS)
      p.set_points(pts)
      res = p.execute(src)
      out[f"{name}__out_{mode}"] = res
      if mode == "gpuparams":
        out[f"{name}__meta"] = np.array([rank, M, T, ttype, sign, p.kernel_width] + list(grid), np.int64)
        out[f"{name}__tol"] = np.array([tol])
        out[f"{name}__range"] = np.array([{"strict": 0, "extended": 1, "infinite": 2}[prange]])
      p.close()
    out[f"{name}__points"] = pts
    out[f"{name}__src"] = src
  np.savez_compressed(os.path.join(HERE, "nufft.npz"), num_threads=NUM_THREADS, **out)
  for f in ("params.npz", "sort.npz", "nufft.npz"):
    print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
  main()
