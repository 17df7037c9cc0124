set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gru_tcb_kernel' -s 3 -c 1 -o gpurun_out/prof_c3 python bench.py --steps 3 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency > gpurun_out/prof_c3.log 2>&1
tail -2 gpurun_out/prof_c3.log | cut -c1-200
