import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import cmlpl_oracle as O
from cmlpl_b200 import ops
from cmlpl_b200.tools.models import BaseNet2
dev=torch.device("cuda")
log=[]
def wrap(name):
    f=getattr(ops,name)
    def g(*a,**k):
        out=f(*a,**k)
        torch.cuda.synchronize()
        out2=f(*a,**k)
        torch.cuda.synchronize()
        o1=out if isinstance(out,tuple) else (out,); o2=out2 if isinstance(out2,tuple) else (out2,)
        for i,(p,q) in enumerate(zip(o1,o2)):
            d=float((p-q).abs().max()/max(float(q.abs().max()),1e-30))
            if d>1e-5: log.append((name,i,d,[tuple(t.shape) for t in a if torch.is_tensor(t)]))
        return out
    setattr(ops,name,g)
for n in ["conv2d","conv2d_dgrad","conv2d_wgrad","avgpool2","avgpool2_bwd","relu_bwd","sgemm","colsum","l2norm","l2norm_bwd"]:
    wrap(n)
def rel(a,b):
    a=np.asarray(a,dtype=np.float64); b=np.asarray(b,dtype=np.float64); return np.abs(a-b).max()/np.abs(b).max()
torch.manual_seed(5)
for trial in range(6):
    sd=O.basenet2_init(103,9)
    net=BaseNet2(103,0,9); net.load_state_dict(sd); net=net.to(dev).train()
    x=torch.randn(256,60,20,20); y=torch.randn(256,103)
    w=torch.randn(256,9); wf=torch.randn(256,1024)
    sdr={k:v.clone().requires_grad_(True) for k,v in sd.items()}
    lo,fe=O.basenet2_forward(sdr,x,y); ((lo*w).sum()+(fe*wf).sum()).backward()
    lo2,fe2=net(x.to(dev),y.to(dev)); ((lo2*w.to(dev)).sum()+(fe2*wf.to(dev)).sum()).backward()
    print("trial",trial, {k: "%.1e"%rel(p.grad.cpu(), sdr[k].grad) for k,p in net.named_parameters() if k in ('conv1.weight','conv2.weight')}, log)
    log.clear()
    junk=[torch.randn(1000,1000,device=dev) for _ in range(3)]
