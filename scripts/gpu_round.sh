#!/bin/bash
# One single-GPU session of a round: tests, bench lines, launch list, ncu captures.  Run under gpurun from the repo root:
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'
# Every output lands in gpurun_out/r02/; summaries are copied into profiles/ by scripts/collect_profiles.py r02.
set -u
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L > $O/box.txt; nproc >> $O/box.txt; nvidia-smi topo -m >> $O/box.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_1ant.log 2>$O/bench_1ant.err
timeout 300 python bench.py --steps 10 --warmup 3 --max-batch 1 --no-cpu-baseline --no-legacy --no-e2e > $O/bench_1ant_nobatch.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --clean --no-cpu-baseline --no-legacy --no-e2e > $O/bench_1ant_clean.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --antennas 8 --seconds-per-step 4 --no-cpu-baseline --no-legacy > $O/bench_8ant.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --nbit 8 --no-cpu-baseline --no-legacy --no-e2e > $O/bench_1ant_nbit8.log 2>&1
# the first process on a fresh box pays for paging in the CUDA libraries inside its observation: warm up, then measure
timeout 300 vlite-fast_b200/bin/process_baseband -S 2 -L 5 -F -w 0 -j > /dev/null 2>&1
timeout 300 vlite-fast_b200/bin/process_baseband -S 10 -L 6 -F -w 0 -j -T -o > $O/exe_60s.log 2>$O/exe_60s.err
python scripts/ubench/h2d_rate.py > $O/h2d_rate.log 2>&1
timeout 120 python scripts/k2_time.py > $O/kernel_times_serialised.log 2>&1
timeout 120 python scripts/k2_trace.py > $O/k2_trace.log 2>&1
# launch list (serialised, cold-cache per-launch times: shares only), then one full capture of each kernel
CMD="python bench.py --steps 1 --warmup 3 --seconds-per-step 4 --no-cpu-baseline --no-legacy --no-e2e --no-check"
$CMD > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vf_k1_pipelined -s 8 -c 1 -o $O/prof_k1 -f $CMD > $O/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vf_k2_normalise -s 8 -c 1 -o $O/prof_k2 -f $CMD > $O/ncu_k2.log 2>&1
timeout 600 ncu --metrics sass__inst_executed_per_opcode,smsp__inst_executed.sum --clock-control none -k regex:vf_k1_pipelined -s 8 -c 1 --csv --log-file $O/k1_opcodes.csv $CMD > $O/ncu_k1_op.log 2>&1
tail -c 400 $O/bench_1ant.log
