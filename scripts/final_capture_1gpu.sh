#!/bin/bash
# Final 1-GPU evidence of a build: tests, bench lines, launch list, ncu captures (each profiled command first runs plainly and must
# exit 0). Usage (on the GPU box, from the repo root): bash scripts/final_capture_1gpu.sh <prefix>     -> gpurun_out/<prefix>_*
set -u
P=${1:-rX}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${P}_tests.log 2>&1; tail -2 $O/${P}_tests.log
python bench.py > $O/${P}_bench_c3_1gpu.json 2> $O/${P}_bench_c3_1gpu.err; tail -c 600 $O/${P}_bench_c3_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_bench_reference_arm_c3.json 2>/dev/null
python bench.py --workload C4 > $O/${P}_bench_c4_1gpu.json 2>/dev/null
python bench.py --workload C2 > $O/${P}_bench_c2_1gpu.json 2>/dev/null
python bench.py --workload C1 > $O/${P}_bench_c1_1gpu.json 2>/dev/null
for w in C2 C3 C4 C5; do python scripts/phase_timers.py $w 592 > $O/${P}_phase_timers_$w.txt 2>&1; done
# launch list of the bench command
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${P}_plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${P}_ncu_bench.log 2>&1
# full captures of the rollout kernel
python scripts/probe.py C3 1184 --reps 1 > $O/${P}_plain_c3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rbo_rollout_kernel -c 1 -o $O/${P}_c3_rollout -f python scripts/probe.py C3 1184 --reps 1 > $O/${P}_ncu_c3.log 2>&1
python scripts/probe.py C5 296 --reps 1 > $O/${P}_plain_c5.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rbo_rollout_kernel -c 1 -o $O/${P}_c5_rollout -f python scripts/probe.py C5 296 --reps 1 > $O/${P}_ncu_c5.log 2>&1
# DRAM / L2 bytes of full-size launches
python scripts/probe.py C3 16384 --reps 1 > $O/${P}_plain_c3_full.log 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:rbo_rollout_kernel -c 1 --csv --log-file $O/${P}_dram_c3_fullM.csv python scripts/probe.py C3 16384 --reps 1 > /dev/null 2>&1
python scripts/probe.py C5 1184 --reps 1 > $O/${P}_plain_c5_b.log 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:rbo_rollout_kernel -c 1 --csv --log-file $O/${P}_dram_c5_M1184.csv python scripts/probe.py C5 1184 --reps 1 > /dev/null 2>&1
cat $O/${P}_plain_c3.log $O/${P}_plain_c5.log $O/${P}_plain_c3_full.log $O/${P}_plain_c5_b.log
