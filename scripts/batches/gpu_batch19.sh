set -x
./scripts/micro/rcp_check > gpurun_out/r2_rcp_check.json 2>&1; cat gpurun_out/r2_rcp_check.json
python scripts/tune.py c4 --reps 3 --variants "|" > gpurun_out/r2_tune19_c4.jsonl 2>&1; cut -c1-140 gpurun_out/r2_tune19_c4.jsonl
