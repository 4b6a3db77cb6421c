#!/bin/bash
# Round profiling pass (run under gpurun): ncu --set full captures of every engine kernel on the
# BASELINE configs + the launch list of the default bench command. Outputs under gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:spread_ws2 -c 1 -s 2 -o $O/r1_ws2_cfg2 python scripts/prof_one.py cfg2 1 > $O/prof.log 2>&1
$NCU -k regex:spread_tile -c 1 -s 2 -o $O/r1_spread3d_cfg3 python scripts/prof_one.py cfg3 1 >> $O/prof.log 2>&1
$NCU -k regex:interp_qw -c 1 -s 2 -o $O/r1_interp_cfg4 python scripts/prof_one.py cfg4 2 >> $O/prof.log 2>&1
$NCU -k regex:interp_qw -c 1 -s 2 -o $O/r1_interp_cfg2t2 python scripts/prof_one.py cfg2 2 8 >> $O/prof.log 2>&1
$NCU -k regex:"fold_key|radix_scatter|stencil_record8|radix_hist" -c 4 -s 8 -o $O/r1_setpts_cfg3 python scripts/prof_one.py cfg3 1 >> $O/prof.log 2>&1
$NCU -k regex:"deconvolve" -c 1 -s 2 -o $O/r1_deconv_cfg2 python scripts/prof_one.py cfg2 1 >> $O/prof.log 2>&1
$NCU -k regex:"amplify" -c 1 -s 2 -o $O/r1_amplify_cfg4 python scripts/prof_one.py cfg4 2 >> $O/prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
# gpurun copies back at most 64 MiB: summarise on the box, keep only the two headline reports
for r in $O/r1_*.ncu-rep; do python scripts/ncu_summary.py $r 22 > ${r%.ncu-rep}_summary.txt 2>&1; done
rm -f $O/r1_setpts_cfg3.ncu-rep $O/r1_deconv_cfg2.ncu-rep $O/r1_amplify_cfg4.ncu-rep $O/r1_interp_cfg2t2.ncu-rep $O/r1_spread3d_cfg3.ncu-rep
tail -3 $O/prof.log
ls -la $O/r1_*
