set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests24.log 2>&1; tail -4 gpurun_out/r2_tests24.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke24.log 2>&1; tail -2 gpurun_out/r2_smoke24.log
python bench.py > gpurun_out/r2_bench24_n1.json 2> gpurun_out/r2_bench24_n1.err; tail -c 300 gpurun_out/r2_bench24_n1.json; tail -3 gpurun_out/r2_bench24_n1.err
