set -x
python scripts/tune.py c5 --nsub 200 --nspp 256 --reps 2 --variants "|128" > gpurun_out/r2_tune23_c5.jsonl 2>&1; cut -c1-200 gpurun_out/r2_tune23_c5.jsonl
python -m pytest tests -m gpu -q -x -k "sde or c5 or SDE or particle or kalman" > gpurun_out/r2_tests23.log 2>&1; tail -4 gpurun_out/r2_tests23.log
