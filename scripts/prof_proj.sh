set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'input_proj_kernel' -s 4 -c 1 -o gpurun_out/prof_proj python bench.py --steps 4 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency --no-config3 > gpurun_out/prof_proj.log 2>&1
tail -1 gpurun_out/prof_proj.log | cut -c1-100
