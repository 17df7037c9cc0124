# Round-2 measurement recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
# bench.py primes every stream with 24 untimed ticks and warms up for 3; a tick of the default path at 131072 streams launches three
# kernels (mfcc_tc3_plan_kernel, mfcc_tc3_kernel, gru_mma16_kernel), so the first timed tick starts at launch 27 * 3 (+ the one-time rebuild of the
# projection cache and a fill).
set -x
mkdir -p gpurun_out
TAG=${1:-r2}
# (0) the bench line itself (both arms), not under a profiler
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
# (1) launch list of the bench command: every kernel with its device time
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 36 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 8 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency --no-config3 > gpurun_out/launches_$TAG.log 2>&1
# (2) full capture of the hot kernels: two consecutive steady-state ticks (a one-frame and a two-frame tick)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mfcc_tc3|gru_mma' -s 81 -c 6 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 4 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency --no-config3 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-300
