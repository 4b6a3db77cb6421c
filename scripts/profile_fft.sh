#!/bin/bash
# Round-2 profiling pass of the own FFT passes (run under gpurun): ncu --set full of the five pass
# kernels of cfg4 (type 2, 512^3) and cfg2 (type 1, 1024^2), summarised on the box, plus the launch
# lists of the bench command on every BASELINE config. Outputs under gpurun_out/.
set -u
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fft_ -c 5 -o $O/r02_fft_passes python scripts/fft_prof.py > $O/prof_fft.log 2>&1
python scripts/ncu_summary.py $O/r02_fft_passes.ncu-rep 14 > $O/r02_ncu_fft_passes_summary.txt 2>&1
rm -f $O/r02_fft_passes.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --only-main > $O/bench_under_ncu.log 2>&1
for c in cfg1 cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --only-main >> $O/bench_under_ncu.log 2>&1
done
tail -2 $O/prof_fft.log
ls -la $O/r02_*
