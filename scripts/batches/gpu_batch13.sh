set -x
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|128" > gpurun_out/r2_tune13_c5_fp32.jsonl 2>&1; cut -c1-300 gpurun_out/r2_tune13_c5_fp32.jsonl
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 2 --sde-fp64 --variants "|128" > gpurun_out/r2_tune13_c5_fp64.jsonl 2>&1; cut -c1-300 gpurun_out/r2_tune13_c5_fp64.jsonl
python -m pytest tests -m gpu -q -x -k "sde or c5 or SDE or particle or kalman" > gpurun_out/r2_tests13.log 2>&1; tail -4 gpurun_out/r2_tests13.log
