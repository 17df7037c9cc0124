# round-2 first GPU call: experimental-kernel parity, full GPU suite, K1 A/B timing
set -x
mkdir -p gpurun_out
timeout 150 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k experimental_mfcc -x -q -s 2>&1 | tail -25 | tee gpurun_out/tc_k1_test.log
timeout 150 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k tcgen05_scan_over -x -q 2>&1 | tail -8 | tee gpurun_out/tc5_proj_test.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
timeout 200 python scripts/tc2_time.py 3 2 4 2>&1 | tee gpurun_out/tc2_time.log
