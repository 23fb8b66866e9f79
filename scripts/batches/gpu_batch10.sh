set -x
python scripts/tune.py c2 --reps 5 --variants "|" > gpurun_out/r2_tune10_c2.jsonl 2>&1; cut -c1-240 gpurun_out/r2_tune10_c2.jsonl
python scripts/tune.py c3 --reps 3 --variants "|" > gpurun_out/r2_tune10_c3.jsonl 2>&1; cut -c1-240 gpurun_out/r2_tune10_c3.jsonl
python scripts/tune.py c4 --reps 3 --variants "|" > gpurun_out/r2_tune10_c4.jsonl 2>&1; cut -c1-240 gpurun_out/r2_tune10_c4.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests10.log 2>&1; tail -4 gpurun_out/r2_tests10.log
