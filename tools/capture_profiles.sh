#!/bin/bash
# Round profile capture (run under gpurun, one GPU): plain runs first, then the ncu passes.
#   gpurun --timeout 2400 -- bash tools/capture_profiles.sh r02
R=${1:-r02}
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err || exit 1
# launch lists (per-launch durations are cold-cache and serialised under ncu: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_fashion_bench.csv \
    python bench.py --steps 2 --warmup 1 --inner 1 --cpu-seconds 1 --all-layers 0 --train 0 > gpurun_out/ncu_bench.log 2>&1
python train.py --model cifar10 --batch 512 --steps 50 --warmup 10 --graph > gpurun_out/plain_train.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 900 --csv --log-file gpurun_out/${R}_launches_train_cifar10.csv \
    python train.py --model cifar10 --batch 512 --steps 6 --warmup 2 > gpurun_out/ncu_train.log 2>&1
for spec in "fashion 262144 sfwd_kernel|sbwd_kernel" "cifar10_pde1 65536 sfwd_kernel|sbwd_kernel" "emotion 98304 emo_fwd_tiled|emo_bwd_tiled" "tiny 16384 tiny_fwd_kernel|tiny_bwd_kernel" "mnist 262144 sfwd_kernel|sbwd_kernel"; do
  set -- $spec
  python tools/prof_layer.py $1 $2 3 > gpurun_out/plain_$1.log 2>&1 || exit 1
  # (gpurun copies back at most 64 MiB: the source pages travel only with the captures whose stalls are read by line)
  SRC=on; case $1 in mnist|tiny) SRC=off;; esac
  ncu --set full --clock-control none --import-source $SRC --kernel-name regex:"$3" --launch-skip 2 --launch-count 2 \
      -o gpurun_out/prof_$1_${R} -f python tools/prof_layer.py $1 $2 3 > gpurun_out/ncu_$1.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
