set -x
mkdir -p gpurun_out
timeout 120 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k "experimental_mfcc and tensor_core_two_stage" -x -q -s 2>&1 | tail -6 | tee gpurun_out/tc3_test.log
timeout 200 python scripts/tc2_time.py 3 5 2>&1 | tee gpurun_out/tc3_time.log
bash scripts/ncu_k1.sh 5 mfcc_tc3_kernel
