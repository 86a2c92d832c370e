#!/bin/bash
# Round profile capture (run under gpurun, one GPU): plain runs first, then the ncu passes.
set -x
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/ncu_bench.log 2>&1
for spec in "fashion 262144 sfwd_kernel|sbwd_kernel" "cifar10_pde1 65536 sfwd_kernel|sbwd_kernel" "emotion 98304 emo_fwd_tiled|emo_bwd_tiled"; do
  set -- $spec
  python tools/prof_layer.py $1 $2 3 > gpurun_out/plain_$1.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on --kernel-name regex:"$3" --launch-skip 2 --launch-count 2 \
      -o gpurun_out/prof_$1_r1_final -f python tools/prof_layer.py $1 $2 3 > gpurun_out/ncu_$1.log 2>&1
done
