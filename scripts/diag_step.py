import os, sys, argparse
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import cmlpl_oracle as O
from test_oracle_cpu import replay_step
from cmlpl_b200 import train as T
g="/root/repo/tests/golden"
z=np.load(g+"/step.npz"); ti=np.load(g+"/train_infer.npz")
r, ost = replay_step(z, ti)
inp = ost.extras["inputs"]; nz, a = inp["noise"], inp["args"]
dev=torch.device("cuda")
args = argparse.Namespace(temperature=a.temperature, thr=a.thr, num_epochs=a.num_epochs, queue_batch=a.queue_batch, alpha=a.alpha, lr=a.lr, labeled_batch_size=128, dropout=0, noise=a.noise)
st = T.make_state(103, 9, args, dev)
st.Base.load_state_dict(inp["sd"]); st.Base1.load_state_dict(inp["sd1"])
for dst, src in zip((st.queue_feats, st.queue_probs, st.queue_feats1, st.queue_probs1), inp["queues"]): dst.copy_(src)
st.extras["keep_grads"]=True
d=lambda t:t.to(dev)
XP_b = torch.cat([inp["XP_l"] + nz["xp_l1"] * a.noise, inp["XP_u"] + nz["xp_u1"] * a.noise], 0)
X_b = torch.cat([inp["X_l"] + nz["x_l1"] * a.noise, inp["X_u"] + nz["x_u1"] * a.noise], 0)
XP_e = torch.cat([inp["XP_l"] + nz["xp_l2"] * a.noise, inp["XP_u"] + nz["xp_u2"] * a.noise], 0)
X_e = torch.cat([inp["X_l"] + nz["x_l2"] * a.noise, inp["X_u"] + nz["x_u2"] * a.noise], 0)
hist, aux = T.mutual_step(st, d(XP_b), d(X_b), d(XP_e), d(X_e), d(inp["Y_l"]), 1, 0, args)
# float64 oracle
sd64={k:v.double() for k,v in inp["sd"].items()}; sd164={k:v.double() for k,v in inp["sd1"].items()}
st64=O.make_state(sd64, sd164, 9, a)
for dst, src in zip((st64.queue_feats, st64.queue_probs, st64.queue_feats1, st64.queue_probs1), inp["queues"]):
    dst.data=src.double()
nz64={k:v.double() for k,v in nz.items()}
r64=O.ref_step(st64, inp["XP_l"].double(), inp["X_l"].double(), inp["Y_l"], inp["XP_u"].double(), inp["X_u"].double(), nz64, 1, 0, a)
def rel(a,b): 
    a=np.asarray(a,dtype=np.float64); b=np.asarray(b,dtype=np.float64); return np.abs(a-b).max()/np.abs(b).max()
for net,(gm,go,g64) in {"net0":(aux["grads"], r["grads"], r64["grads"]), "net1":(aux["grads1"], r["grads1"], r64["grads1"])}.items():
    for k in O.LIVE_KEYS:
        print(net, k, "mine-vs-f32oracle %.2e  mine-vs-f64 %.2e  f32oracle-vs-f64 %.2e" % (rel(gm[k].cpu(), go[k]), rel(gm[k].cpu(), g64[k]), rel(go[k], g64[k])))
