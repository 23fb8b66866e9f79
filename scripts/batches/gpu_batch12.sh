set -x
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|128" > gpurun_out/r2_tune12_c5_fp32.jsonl 2>&1; cut -c1-300 gpurun_out/r2_tune12_c5_fp32.jsonl
python -m pytest tests -m gpu -q -x -k "sde or c5 or SDE or particle or kalman" > gpurun_out/r2_tests12.log 2>&1; tail -4 gpurun_out/r2_tests12.log
ncu --set full --clock-control none --import-source on -k regex:psi_entry -s 1 -c 1 -f -o gpurun_out/r2_full_c5 python scripts/tune.py c5 --nsub 40 --nspp 256 --reps 1 --variants "|128" > gpurun_out/ncu12_c5.log 2>&1; tail -3 gpurun_out/ncu12_c5.log
