# 1 and 2 ranks back to back on one box (run under gpurun --gpus 2): the 2-rank NCCL test and bench lines with the co-add check
O=gpurun_out/s2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q > $O/pytest.log 2>&1; tail -1 $O/pytest.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-legacy --no-e2e > $O/n1.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29821 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/n2.log 2>&1
for f in n1 n2; do tail -1 $O/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['coadd_check_after'])"; done
