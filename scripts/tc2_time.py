"""K1 A/B at the bench size (pb_debug_k1_mode): 2 = FFT kernel on the CUDA cores, 3 = the same with 64-bit set-up, 4 = mfcc_tc2 (stage 2 on
tcgen05), 5 = mfcc_tc3 (both DFT stages on tcgen05; the default at this size), 6 = mfcc_tc3 with the shuffle-gather epilogue tail."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, mycroft_precise_b200 as m
S = 131072
modes = [int(a) for a in sys.argv[1:]] or [2, 5]
model = m.GruModel.random(13, 20, seed=0, scale=0.1)
pcm = torch.from_numpy((np.random.RandomState(0).randn(S, 1024) * 3000).astype(np.int16)).cuda()
for mode in modes:
    sb = m.StreamBatch(model, S, chunk_samples=1024)
    sb.core.k1_mode(mode)
    for _ in range(30):
        sb.update(pcm)
    torch.cuda.synchronize()
    sb.core.profile(True)
    for _ in range(20):
        sb.update(pcm)
    ms, n = sb.core.profile_read()
    print('k1_mode', mode, 'K1 %.1f us per tick' % (1e3 * ms[0] / n[0]), 'K2 %.1f' % (1e3 * ms[1] / n[1]), flush=True)
    sb.core.close()
