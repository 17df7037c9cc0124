# Round-1 profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -15
# (1) launch list of the bench command: every kernel with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 8 --warmup 3 --no-small-batch --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
# (2) full capture of the two hot kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:'mfcc_fast_stream_kernel|gru_mma_kernel' -s 8 -c 2 \
    -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-small-batch --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-300
