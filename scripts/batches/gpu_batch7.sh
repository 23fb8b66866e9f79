set -x
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_c5_reference_stepper_properties > gpurun_out/r2_tests7.log 2>&1; tail -4 gpurun_out/r2_tests7.log
for W in c1 c3 c2; do
python bench.py --workload $W --extras none --no-cpu-baseline --steps 10 > gpurun_out/r2_b7_$W.json 2>gpurun_out/r2_b7_$W.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b7_$W.json')); print('$W', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], (d['e2e_pageable'] or {}).get('ms_per_step'))"
done
python scripts/tune.py c1 --reps 5 --variants "|128" "PSI_MIN_BLOCKS=8|128" "PSI_MIN_BLOCKS=12|128" "PSI_MIN_BLOCKS=10 PSI_PROG_STAGE=1|128" > gpurun_out/r2_tune7_c1.jsonl 2>&1; cut -c1-110 gpurun_out/r2_tune7_c1.jsonl
