import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, synth
from chainer_maskrcnn_b200 import _engine, _lib
rng=np.random.RandomState(0)
N,H,W=2,50,84
rois=torch.from_numpy(synth.make_rois(rng,N,512,H*16,W*16,size_range=(32.,400.))).cuda()
for C in (488,490,492):
    x=torch.randn(N,C,H,W,device='cuda').contiguous(memory_format=torch.channels_last)
    gy=torch.rand(rois.shape[0],C,7,7,device='cuda').contiguous(memory_format=torch.channels_last)
    def step():
        outs,plan=_engine.forward([x],rois,None,[1/16.],[7],sampling_ratio=1)
        _engine.backward(plan,[gy])
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    print(C, "ms per fwd+bwd", e0.elapsed_time(e1)/20)
