"""Text summary of one .ncu-rep: per-kernel raw metrics (duration, DRAM bytes, shared-memory
wavefronts, issue/occupancy) followed by the hot-SASS listing of scripts/ncu_top.py.
Usage: python scripts/ncu_summary.py file.ncu-rep [topN]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]
top = sys.argv[2] if len(sys.argv) > 2 else "22"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
  d = dict(zip(hdr, r))
  u = dict(zip(hdr, units))
  print(f"# {d.get('Kernel Name', '?')[:110]}")
  for k in KEYS:
    if k in d and d[k] != "":
      print(f"{k} {d[k]} {u.get(k, '')}")
  print()
here = os.path.dirname(os.path.abspath(__file__))
sys.stdout.flush()
print(subprocess.run([sys.executable, os.path.join(here, "ncu_top.py"), rep, top], capture_output=True, text=True).stdout)
