#!/bin/bash
# One GPU session of a round: tests, bench lines, launch list, ncu captures.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh'
# Every output lands in gpurun_out/; summaries are copied into profiles/ by scripts/collect_profiles.py afterwards.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_1ant.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --k1-threads 640 --no-cpu-baseline --no-legacy > $O/bench_1ant_mono.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --max-batch 1 --no-cpu-baseline --no-legacy > $O/bench_1ant_nobatch.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --antennas 8 --no-cpu-baseline --no-legacy > $O/bench_8ant.log 2>&1
# the first process on a fresh box pays for paging in the CUDA libraries inside its observation: warm up, then measure
timeout 300 vlite-fast_b200/bin/process_baseband -S 2 -L 5 -F -w 0 -j > /dev/null 2>&1
timeout 300 vlite-fast_b200/bin/process_baseband -S 10 -L 6 -F -w 0 -j > $O/exe_60s.log 2>$O/exe_60s.err
# launch list (serialised, cold-cache per-launch times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legacy --no-e2e > $O/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k vf_k1_pipelined -s 4 -c 1 -o $O/prof_k1_final -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legacy --no-e2e > $O/ncu_k1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vf_k2 -s 4 -c 1 -o $O/prof_k2_final -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legacy --no-e2e > $O/ncu_k2.log 2>&1
scripts/ubench/fp32_rate > $O/fp32_rate.log 2>&1
python scripts/ubench/h2d_rate.py > $O/h2d_rate.log 2>&1
tail -c 600 $O/bench_1ant.log
