# One --set full capture of the K1 kernel selected by $1 (k1 mode) at the bench size; report lands in gpurun_out/.
# usage (under gpurun, ONE GPU): bash scripts/ncu_k1.sh 4 mfcc_tc2
MODE=${1:-4}; KNAME=${2:-mfcc_tc2}
mkdir -p gpurun_out
cat > /tmp/ncu_drv.py <<PY
import sys, os
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import numpy as np, torch, mycroft_precise_b200 as m
S = 131072
model = m.GruModel.random(13, 20, seed=0, scale=0.1)
pcm = torch.from_numpy((np.random.RandomState(0).randn(S, 1024) * 3000).astype(np.int16)).cuda()
sb = m.StreamBatch(model, S, chunk_samples=1024)
sb.core.k1_mode($MODE)
for _ in range(8):
    sb.update(pcm)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$KNAME --launch-skip 5 --launch-count 2 -o gpurun_out/ncu_k1_mode$MODE -f python /tmp/ncu_drv.py > gpurun_out/ncu_k1_mode$MODE.log 2>&1
tail -3 gpurun_out/ncu_k1_mode$MODE.log
