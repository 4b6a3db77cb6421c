#!/bin/bash
# Final measurement pass of the round (run under gpurun): bench lines of every BASELINE config, the
# reference arm, the reference's benchmark shapes, and the ncu launch lists of the same commands.
set -u
O=gpurun_out
python bench.py > $O/r02_bench_final.json 2> $O/final.err
python bench.py --impl reference --steps 10 --warmup 3 > $O/r02_bench_final_ref.json 2>> $O/final.err
for c in cfg1 cfg3 cfg4; do python bench.py --config $c --only-main > $O/r02_bench_final_$c.json 2>> $O/final.err; done
for c in ref1 ref2 ref3 ref4 ref5 ref6 ref7 ref8; do python bench.py --config $c --only-main > $O/r02_bench_final_$c.json 2>> $O/final.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --only-main --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for c in cfg1 cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --only-main --no-cpu-baseline >> $O/bench_under_ncu.log 2>&1
done
python scripts/sass_summary.py > $O/r02_sass_summary.txt 2>&1
tail -3 $O/final.err
