set -x
python scripts/tune.py c4 --reps 3 --variants "|" > gpurun_out/r2_tune20_c4.jsonl 2>&1; cut -c1-140 gpurun_out/r2_tune20_c4.jsonl
python scripts/tune.py c3 --reps 3 --variants "|" > gpurun_out/r2_tune20_c3.jsonl 2>&1; cut -c1-140 gpurun_out/r2_tune20_c3.jsonl
python scripts/tune.py c1 --reps 5 --variants "|" > gpurun_out/r2_tune20_c1.jsonl 2>&1; cut -c1-140 gpurun_out/r2_tune20_c1.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests20.log 2>&1; tail -4 gpurun_out/r2_tests20.log
