set -x
run() { W=$1; S=$2; python bench.py --workload $W --extras none --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/plain18_$W.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:psi_entry -s $S -c 1 -f -o gpurun_out/r2b_full_$W python bench.py --workload $W --extras none --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > gpurun_out/ncu18_$W.log 2>&1; }
run c4 7
run c3 3
run c2 7
ls -la gpurun_out/r2b*.ncu-rep
