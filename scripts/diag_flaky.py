import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from cmlpl_b200 import ops
dev=torch.device("cuda")
torch.manual_seed(0)
b=256
dz2=torch.randn(b,64,10,10,device=dev); w2=torch.randn(64,64,3,3,device=dev)*0.05
a1=torch.randn(b,64,20,20,device=dev); a0=torch.randn(b,64,20,20,device=dev); x=torch.randn(b,60,20,20,device=dev)
w1=torch.randn(64,64,3,3,device=dev)*0.05
def chain():
    dp1=ops.conv2d_dgrad(dz2,w2,res=dz2)
    da1=ops.avgpool2_bwd(dp1,20,20)
    dz1=ops.relu_bwd(a1,da1)
    dw1,db1=ops.conv2d_wgrad(a0,dz1,3)
    da0=ops.conv2d_dgrad(dz1,w1,res=dz1)
    dw0,db0=ops.conv2d_wgrad(x,da0,1)
    return dict(dp1=dp1,da1=da1,dz1=dz1,dw1=dw1,db1=db1,da0=da0,dw0=dw0,db0=db0)
ref=chain(); torch.cuda.synchronize()
for it in range(12):
    junk=[torch.randn(np.random.randint(100,3000),1000,device=dev) for _ in range(np.random.randint(1,4))]
    del junk
    out=chain(); torch.cuda.synchronize()
    bad={k: float((out[k]-ref[k]).abs().max()/ref[k].abs().max()) for k in out if not torch.equal(out[k],ref[k])}
    print(it, {k:"%.1e"%v for k,v in bad.items()})
