set -x
python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_c5_reference_stepper_properties > gpurun_out/r2_tests3.log 2>&1; tail -25 gpurun_out/r2_tests3.log
python scripts/tune.py c1 --reps 5 --variants "PSI_MIN_BLOCKS=6|128" "PSI_MIN_BLOCKS=8|128" "PSI_MIN_BLOCKS=10|128" "PSI_MIN_BLOCKS=12|128" "PSI_MIN_BLOCKS=8|64" "PSI_MIN_BLOCKS=16|64" > gpurun_out/r2_tune_c1.jsonl 2>&1; cat gpurun_out/r2_tune_c1.jsonl
python scripts/tune.py c3 --reps 3 --variants "PSI_MIN_BLOCKS=6|128" "PSI_MIN_BLOCKS=5|128" "PSI_MIN_BLOCKS=8|128" > gpurun_out/r2_tune_c3.jsonl 2>&1; cat gpurun_out/r2_tune_c3.jsonl
python bench.py --workload c1 --extras none --steps 20 > gpurun_out/r2_bench_c1.json 2>gpurun_out/r2_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c1.json')); print('C1', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e_pageable'])"
M=sm__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for W in c2 c3 c4 c1; do
  python bench.py --workload $W --extras none --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/plain_$W.log 2>&1 && ncu --metrics $M --clock-control none -k regex:psi_entry -c 12 --csv --log-file gpurun_out/r2_opcounts_$W.csv python bench.py --workload $W --extras none --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu_$W.log 2>&1
done
ls -la gpurun_out | tail -12
