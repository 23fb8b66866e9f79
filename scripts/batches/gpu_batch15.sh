set -x
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|64" "|96" "|128" > gpurun_out/r2_tune15_c5_block.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune15_c5_block.jsonl
PHARMSOL_B200_SDE_SMEM=0 python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 3 --variants "|128" > gpurun_out/r2_tune15_c5_nosmem.jsonl 2>&1; cut -c1-120 gpurun_out/r2_tune15_c5_nosmem.jsonl
