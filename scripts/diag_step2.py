import os, sys, argparse
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import cmlpl_oracle as O
from cmlpl_b200.tools.models import BaseNet2
dev=torch.device("cuda")
def rel(a,b):
    a=np.asarray(a,dtype=np.float64); b=np.asarray(b,dtype=np.float64); return np.abs(a-b).max()/np.abs(b).max()
torch.manual_seed(5)
for trial in range(3):
    sd=O.basenet2_init(103,9)
    net=BaseNet2(103,0,9); net.load_state_dict(sd); net=net.to(dev).train()
    x=torch.randn(256,60,20,20); y=torch.randn(256,103)
    w=torch.randn(256,9); wf=torch.randn(256,1024)
    sdr={k:v.clone().requires_grad_(True) for k,v in sd.items()}
    lo,fe=O.basenet2_forward(sdr,x,y); ((lo*w).sum()+(fe*wf).sum()).backward()
    lo2,fe2=net(x.to(dev),y.to(dev)); ((lo2*w.to(dev)).sum()+(fe2*wf.to(dev)).sum()).backward()
    print("trial",trial, {k: "%.1e"%rel(p.grad.cpu(), sdr[k].grad) for k,p in net.named_parameters() if k in O.LIVE_KEYS and 'conv' in k})
    # second backward on a fresh forward with other tensors allocated in between
    junk=[torch.randn(1000,1000,device=dev) for _ in range(3)]
