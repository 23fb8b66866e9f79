set -x
python scripts/tune.py c5 --nsub 200 --nspp 256 --reps 1 --variants "|128" > gpurun_out/r2_tune22_c5.jsonl 2>&1; cut -c1-200 gpurun_out/r2_tune22_c5.jsonl
ncu --set full --clock-control none --import-source on -k regex:psi_entry -s 2 -c 1 -f -o gpurun_out/r2b_full_c5 python scripts/tune.py c5 --nsub 200 --nspp 256 --reps 1 --variants "|128" > gpurun_out/ncu22_c5.log 2>&1; tail -3 gpurun_out/ncu22_c5.log
