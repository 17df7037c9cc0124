# Round-1 profiling recipe (run under gpurun, 1 GPU); superseded by scripts/prof_r2.sh (three kernels per tick), kept for the round-1 profiles.  Outputs land in gpurun_out/.
# bench.py primes every stream with 24 untimed ticks (K1, projection, K2 each); the -s counts below skip them and the warm-up so that the
# captured launches are steady-state updates (full 29-frame windows).
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
# (1) launch list of the bench command: every kernel with its device time
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 24 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 8 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency > gpurun_out/launches_$TAG.log 2>&1
# (2) full capture of the hot kernels (one launch each): K1, K2 (default network), K2 (wide network, configs[2])
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mfcc_fast_stream_kernel|gru_mma_kernel|input_proj_kernel' -s 80 -c 6 \
    -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency --no-config3 > gpurun_out/prof_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gru_tcb_kernel' -s 27 -c 1 \
    -o gpurun_out/prof_c3_$TAG python bench.py --steps 3 --warmup 3 --no-small-batch --no-cpu-baseline --no-latency > gpurun_out/prof_c3_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-300
