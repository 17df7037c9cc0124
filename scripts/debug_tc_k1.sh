# Bring-up recipe for the opt-in kernels written after round 1's GPU budget ran out: the tensor-core MFCC tick
# (csrc/mfcc_tc.cuh, k1 mode 1), the lean fast MFCC kernel (k1 mode 2) and the tcgen05 GRU scan over cached projections
# (gru mode 8); run under gpurun, ONE GPU.
# Everything is wrapped in `timeout`: a wrong mbarrier phase hangs the kernel, and a hung box is a strike.
set -x
mkdir -p gpurun_out
# 1. parity of the windows against the default kernels (small batch: 300 streams, 40 ticks)
timeout 120 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k experimental_mfcc -x -q 2>&1 | tail -15 | tee gpurun_out/tc_k1_test.log
# 1b. the tcgen05 GRU scan over cached projections (pb_debug_gru_mode 8)
timeout 120 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k tcgen05_scan_over -x -q 2>&1 | tail -8 | tee gpurun_out/tc5_proj_test.log
timeout 200 python bench.py --gru-mode 8 --no-cpu-baseline --no-latency --no-config3 --no-small-batch > gpurun_out/bench_mode8.json 2> gpurun_out/bench_mode8.err; tail -c 300 gpurun_out/bench_mode8.json
# 2. memory checker on a tiny run (only if step 1 did not hang)
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 env PB_TEST_TC_K1=1 python -m pytest tests/test_gpu_parity.py -m gpu -k experimental_mfcc -x -q 2>&1 | tail -20 | tee gpurun_out/tc_k1_memcheck.log
# 3. timing against the default kernel at the bench size
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/tc_k1_time.log
import numpy as np, torch, mycroft_precise_b200 as m
S = 131072
model = m.GruModel.random(13, 20, seed=0, scale=0.1)
pcm = torch.from_numpy((np.random.RandomState(0).randn(S, 1024) * 3000).astype(np.int16)).cuda()
for mode in (0, 2, 1):
    sb = m.StreamBatch(model, S, chunk_samples=1024)
    sb.core.k1_mode(mode)
    for _ in range(30):
        sb.update(pcm)
    torch.cuda.synchronize()
    sb.core.profile(True)
    for _ in range(20):
        sb.update(pcm)
    ms, n = sb.core.profile_read()
    print('k1_mode', mode, 'K1 %.1f us per tick' % (1e3 * ms[0] / n[0]))
    sb.core.close()
PY
