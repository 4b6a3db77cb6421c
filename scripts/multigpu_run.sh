#!/bin/bash
# Multi-GPU measurements on ONE box (run under `gpurun --gpus 8`): concurrent pinned-copy bandwidth
# at 1/2/4/8 ranks, bench.py at N = 2/4/8 (weak headline + type2 + strong scaling with gather), and
# the two-GPU device-hygiene test. Writes everything under gpurun_out/.
set -u
OUT=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/r02_mg_gpus.txt 2>&1
python scripts/h2d_probe.py > $OUT/r02_h2d_n1.json 2> $OUT/r02_h2d_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    scripts/h2d_probe.py > $OUT/r02_h2d_n$N.json 2> $OUT/r02_h2d_n$N.err
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02_scale_n1.json 2> $OUT/r02_scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r02_scale_n$N.json 2> $OUT/r02_scale_n$N.err
done
python -m pytest tests/test_gpu_boundary.py -m gpu -q -k "device" > $OUT/r02_mg_pytest.log 2>&1
tail -2 $OUT/r02_mg_pytest.log
cat $OUT/r02_h2d_n*.json
for N in 1 2 4 8; do python - <<PY
import json
try:
  d = json.loads(open("$OUT/r02_scale_n$N.json").read().strip().splitlines()[-1])
  print("N=$N", "value %.3g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.3g" % d["e2e"]["value"],
        "type2 %.3g" % d["type2"]["value"], {k: (round(v["ms_per_step"], 3), round(v["gather_ms"], 3)) for k, v in d["strong"].items()})
except Exception as e:
  print("N=$N failed:", e)
PY
done
