set -x
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|128" > gpurun_out/r2_tune16_c5.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune16_c5.jsonl
PHARMSOL_B200_SDE_TICKET=0 python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|128" > gpurun_out/r2_tune16_c5_static.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune16_c5_static.jsonl
python scripts/tune.py c5 --nsub 200 --nspp 1024 --reps 2 --variants "|128" > gpurun_out/r2_tune16_c5_big.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune16_c5_big.jsonl
PHARMSOL_B200_SDE_TICKET=0 python scripts/tune.py c5 --nsub 200 --nspp 1024 --reps 2 --variants "|128" > gpurun_out/r2_tune16_c5_big_static.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune16_c5_big_static.jsonl
python -m pytest tests -m gpu -q -x -k "sde or c5 or SDE or particle or kalman" > gpurun_out/r2_tests16.log 2>&1; tail -4 gpurun_out/r2_tests16.log
