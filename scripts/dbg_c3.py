import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import mycroft_precise_b200 as m
pr=m.ListenerParams(n_filt=40,n_mfcc=40)
S=18944  # 148 CTAs x 128
model=m.GruModel.random(40,128,seed=1,scale=0.1/np.sqrt(128/20.0))
c=m.PreciseB200(pr,hidden=128,max_streams=S)
c.load_weights(model.kernel,model.recurrent,model.bias,model.dense_w,model.dense_b)
c.debug_counters()
x=torch.randn(S,29,40,device='cuda')
for _ in range(3): c.predict(x)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record(); c.predict(x); b.record(); torch.cuda.synchronize()
print('one wave ms', a.elapsed_time(b)); print('ready,full,issue,total clk', c.debug_counters())
