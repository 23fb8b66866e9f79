set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke21.log 2>&1; tail -2 gpurun_out/r2_smoke21.log
python bench.py > gpurun_out/r2_bench21_n1.json 2> gpurun_out/r2_bench21_n1.err; tail -c 600 gpurun_out/r2_bench21_n1.json; tail -3 gpurun_out/r2_bench21_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench21_ref.json 2> gpurun_out/r2_bench21_ref.err; tail -c 400 gpurun_out/r2_bench21_ref.json
