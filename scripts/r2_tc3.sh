# Bring-up / regression recipe for the tensor-core MFCC kernels (run under gpurun, ONE GPU): parity tests of every K1 variant,
# A/B timing at the bench size, phase timeline of mfcc_tc3 (warps 0 = epilogue, 8 = conversion, 15 = records, 16 = MMA issuer).
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -k "two_stage or alternative_mfcc or large_tick" -x -q 2>&1 | tail -6 | tee gpurun_out/tc3_test.log
timeout 200 python scripts/tc2_time.py 2 5 2>&1 | tee gpurun_out/tc3_time.log
timeout 100 python scripts/tc3_timeline.py 0 8 15 16 2>&1 | grep "^mode" | tee gpurun_out/tc3_timeline.log
