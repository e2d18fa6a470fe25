import sys, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from oracle import cmlpl_oracle as O
from cmlpl_b200 import ops
dev=torch.device("cuda")
torch.manual_seed(5)
for trial in range(6):
    sd=O.basenet2_init(103,9)
    x=torch.randn(256,60,20,20); y=torch.randn(256,103); w=torch.randn(256,9); wf=torch.randn(256,1024)
    a0=F.conv2d(x,sd["conv0.weight"],sd["conv0.bias"]); a1=F.relu(F.conv2d(a0,sd["conv1.weight"],sd["conv1.bias"],padding=1)+a0)
    p1=F.avg_pool2d(a1,2,2); a2=F.relu(F.conv2d(p1,sd["conv2.weight"],sd["conv2.bias"],padding=1)+p1)
    d=lambda t:t.to(dev).contiguous()
    m0=ops.conv2d(d(x),d(sd["conv0.weight"]),d(sd["conv0.bias"])); m1=ops.conv2d(m0,d(sd["conv1.weight"]),d(sd["conv1.bias"]),res=m0,relu=True)
    mp1=ops.avgpool2(m1); m2=ops.conv2d(mp1,d(sd["conv2.weight"]),d(sd["conv2.bias"]),res=mp1,relu=True)
    f1=int(((m1.cpu()>0)!=(a1>0)).sum()); f2=int(((m2.cpu()>0)!=(a2>0)).sum())
    print("trial",trial,"relu1 flips",f1,"relu2 flips",f2,"max|a1 diff| %.1e"%float((m1.cpu()-a1).abs().max()))
    junk=torch.randn(256,9); junk=torch.randn(256,1024)
