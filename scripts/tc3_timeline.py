"""Phase timeline of mfcc_tc3_kernel (CTA 0): cycles in INT, waiting at sync A, in P, waiting at sync B, for an EPI warp (k1 mode 5) and a CONV warp (mode 6)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, mycroft_precise_b200 as m
S = 131072
model = m.GruModel.random(13, 20, seed=0, scale=0.1)
pcm = torch.from_numpy((np.random.RandomState(0).randn(S, 1024) * 3000).astype(np.int16)).cuda()
for mode in [100 + int(a) for a in sys.argv[1:]] or [100, 108]:
    sb = m.StreamBatch(model, S, chunk_samples=1024)
    sb.core.k1_mode(mode)
    sb.core.debug_counters()                 # arms the buffer
    for t in range(12):
        sb.update(pcm)
        torch.cuda.synchronize()
        c = sb.core.debug_counters()
        if t == 10:
            tiles = (S * (2 if sum(c) > 1.5e6 else 1) / 32 + 147) // 148
            print('mode', mode, 'tick', t, c, flush=True)
    sb.core.close()
