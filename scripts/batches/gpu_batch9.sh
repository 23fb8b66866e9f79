set -x
python -m pytest tests/test_gpu_multi_device.py tests/test_c_caller.py tests/test_gpu_parity.py -m gpu -q -x -k "multi or concurrent or chunk or few_columns or c_caller" > gpurun_out/r2_tests9.log 2>&1; tail -4 gpurun_out/r2_tests9.log
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 2 --variants "|128" > gpurun_out/r2_tune9_c5_fp32.jsonl 2>&1; cut -c1-200 gpurun_out/r2_tune9_c5_fp32.jsonl
python scripts/tune.py c5 --nsub 40 --nspp 512 --reps 2 --sde-fp64 --variants "|128" > gpurun_out/r2_tune9_c5_fp64.jsonl 2>&1; cut -c1-200 gpurun_out/r2_tune9_c5_fp64.jsonl
./examples/native_matrix > gpurun_out/r2_native_matrix_c.jsonl 2>&1; cat gpurun_out/r2_native_matrix_c.jsonl | cut -c1-260
./examples/native_matrix --devices 0,0 8192 | cut -c1-300
python benches/native_matrix.py > gpurun_out/r2_native_matrix_py.jsonl 2>&1; cut -c1-220 gpurun_out/r2_native_matrix_py.jsonl
