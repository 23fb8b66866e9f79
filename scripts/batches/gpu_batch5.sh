set -x
for W in c2 c4; do
python bench.py --workload $W --extras none --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/r2_b5_$W.json 2>gpurun_out/r2_b5_$W.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b5_$W.json')); print('$W', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
python scripts/tune.py c4 --reps 3 --variants "PSI_MIN_BLOCKS=6|128" "PSI_MIN_BLOCKS=5|128" "PSI_MIN_BLOCKS=4|128" "PSI_MIN_BLOCKS=8|128" "PSI_MIN_BLOCKS=6|64" > gpurun_out/r2_tune_c4.jsonl 2>&1; cut -c1-100 gpurun_out/r2_tune_c4.jsonl
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 2 --variants "|128" > gpurun_out/r2_tune_c5.jsonl 2>&1; cut -c1-160 gpurun_out/r2_tune_c5.jsonl
