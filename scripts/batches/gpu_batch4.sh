set -x
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_c5_reference_stepper_properties > gpurun_out/r2_tests4.log 2>&1; tail -8 gpurun_out/r2_tests4.log
for W in c2 c4 c1 c3; do
python bench.py --workload $W --extras none --no-cpu-baseline --steps 10 > gpurun_out/r2_b4_$W.json 2>gpurun_out/r2_b4_$W.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b4_$W.json')); print('$W', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], (d['e2e_pageable'] or {}).get('ms_per_step'), d['roofline']['device_counters'])"
done
