#!/bin/bash
# Multi-GPU session of a round, run under  gpurun --gpus N -- 'bash scripts/gpu_scale.sh N'  (N = 2, 4 or 8).
# Weak-scaling lines of bench.py at 1, 2, .. N ranks (the launch line the driver uses), the 16-antenna co-add
# (configs[4]), the reference arm under torchrun, the 2-rank NCCL co-add test, and the concurrent H2D ceiling.
set -u
N=${1:-2}
LITE=${2:-}          # "lite": bench lines only (the ubench, the 2-rank test and the reference arm were run earlier)
O=gpurun_out/r02_scale
mkdir -p $O
nvidia-smi -L > $O/box_${N}gpu.txt; nproc >> $O/box_${N}gpu.txt; nvidia-smi topo -m >> $O/box_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
[ -z "$LITE" ] && timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > $O/pytest_multirank_${N}gpu.log 2>&1; tail -2 $O/pytest_multirank_${N}gpu.log
P=29600
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  P=$((P+1))
  if [ $n -eq 1 ]; then
    [ -z "$LITE" ] && timeout 300 python scripts/ubench/h2d_concurrent.py > $O/h2d_concurrent_${n}.log 2>&1
    timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/scale_${n}gpu.log 2>$O/scale_${n}gpu.err
  else
    [ -z "$LITE" ] && timeout 300 $TR --nproc-per-node $n --master-port $P scripts/ubench/h2d_concurrent.py > $O/h2d_concurrent_${n}.log 2>&1
    P=$((P+1))
    timeout 600 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --steps 10 --warmup 3 > $O/scale_${n}gpu.log 2>$O/scale_${n}gpu.err
    P=$((P+1))
    { [ -z "$LITE" ] || [ $n -eq $N ]; } && timeout 600 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --antennas-total 16 --seconds-per-step 4 > $O/coadd16_${n}gpu.log 2>$O/coadd16_${n}gpu.err
  fi
  tail -c 200 $O/scale_${n}gpu.log
done
P=$((P+1))
[ -z "$LITE" ] && timeout 400 $TR --nproc-per-node 2 --master-port $P bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/ref_2gpu.log 2>&1
