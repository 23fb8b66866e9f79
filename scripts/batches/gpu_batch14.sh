set -x
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "PSI_MIN_BLOCKS=6|128" "PSI_MIN_BLOCKS=7|128" "PSI_MIN_BLOCKS=8|128" "PSI_MIN_BLOCKS=9|128" "PSI_MIN_BLOCKS=10|128" > gpurun_out/r2_tune14_c5_minblocks.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune14_c5_minblocks.jsonl
python -m pytest tests -m gpu -q -x -k "sde or c5 or SDE or particle or kalman" > gpurun_out/r2_tests14.log 2>&1; tail -4 gpurun_out/r2_tests14.log
