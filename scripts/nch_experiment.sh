O=gpurun_out/nch; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
P=29700
for c in default 1 2 4 8; do
  P=$((P+1))
  if [ $c = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$c; fi
  timeout 300 $TR --master-port $P bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e --no-check --no-cpu-baseline > $O/nch_$c.log 2>$O/nch_$c.err
  echo -n "$c: "; tail -1 $O/nch_$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
