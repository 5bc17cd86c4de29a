O=gpurun_out/s2; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29831 bench.py --gpus 2 --steps 5 --warmup 3 --antennas-total 16 --seconds-per-step 4 --no-e2e > $O/c16.log 2>&1
tail -1 $O/c16.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['coadd_check_after'])"
