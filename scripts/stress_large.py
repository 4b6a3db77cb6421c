"""One-off GPU stress: adjointness <A x, y> = <x, A^H y> at sizes well above the BASELINE configs
(index-width / grid-size corner cases)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tensorflow_nufft_b200 as tfft

def check(name, grid, M, T, seed):
  g = torch.Generator(device="cuda").manual_seed(seed)
  rank = len(grid)
  pts = (torch.rand((M, rank), generator=g, device="cuda") * 2 - 1) * np.pi
  x = torch.complex(torch.randn((T,) + grid, generator=g, device="cuda"), torch.randn((T,) + grid, generator=g, device="cuda"))
  y = torch.complex(torch.randn((T, M), generator=g, device="cuda"), torch.randn((T, M), generator=g, device="cuda"))
  torch.cuda.synchronize(); t0 = time.perf_counter()
  Ax = tfft.nufft(x, pts, transform_type="type_2", fft_direction="forward", tol=1e-6)
  AHy = tfft.nufft(y, pts, grid_shape=grid, transform_type="type_1", fft_direction="backward", tol=1e-6)
  torch.cuda.synchronize(); dt = time.perf_counter() - t0
  lhs = torch.sum(Ax.to(torch.complex128) * torch.conj(y.to(torch.complex128)))
  rhs = torch.sum(x.to(torch.complex128) * torch.conj(AHy.to(torch.complex128)))
  rel = abs(lhs - rhs) / abs(lhs)
  print(json.dumps({"case": name, "grid": grid, "M": M, "T": T, "adjoint_rel_err": float(rel), "finite": bool(torch.isfinite(Ax).all() and torch.isfinite(AHy).all()), "seconds": round(dt, 3)}), flush=True)
  tfft.clear_plan_cache()
  assert rel < 2e-5

check("2d-2048-40M", (2048, 2048), 40_000_000, 2, 1)
check("3d-192-30M", (192, 192, 192), 30_000_000, 1, 2)
check("2d-4096-8M-T3", (4096, 4096), 8_000_000, 3, 3)
check("3d-320-6M", (320, 320, 320), 6_000_000, 1, 4)
check("2d-32-20M-dense", (32, 32), 20_000_000, 8, 5)
# power-of-two fine grids: the engine's own FFT passes at their largest sizes (nf = 1024 per axis:
# 8-column bundles, 64-bit strides) and with batches
check("3d-512-10M-ownfft", (512, 512, 512), 10_000_000, 1, 6)
check("2d-512-4M-T40-ownfft", (512, 512), 4_000_000, 40, 7)
check("3d-256-4M-T3-ownfft", (256, 256, 256), 4_000_000, 3, 8)
check("3d-mixed-ownfft", (32, 512, 128), 3_000_000, 2, 9)
print("stress_large: ok")
