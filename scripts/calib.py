"""Prints measured deviations (max-abs-relative and L2-relative) of the CUDA path from the float64 CPU oracle for
whole networks, float32 and bf16 modes -- used to set and justify the test tolerances."""
import sys, random, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import combat_oracle as O
from combat_b200.nets import Classifier, Generator
from combat_b200 import ops
def rm(a,b):
    a=a.detach().float().cpu().double(); b=b.detach().double(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
def r2(a,b):
    a=a.detach().float().cpu().double(); b=b.detach().double(); return float((a-b).norm()/b.norm().clamp_min(1e-30))
def todbl(d): return {k:(v.clone().double() if v.is_floating_point() else v.clone()) for k,v in d.items()}
for arch,size,ncls in [("preact_resnet18",32,10),("resnet18",64,8)]:
    gen = torch.Generator().manual_seed(11)
    init = O.init_preact_resnet18_state if arch=="preact_resnet18" else O.init_resnet18_state
    p,b = init(gen, num_classes=ncls, scaler={32:1,64:4}[size])
    B = 8 if size==32 else 4
    x = torch.rand(B,3,size,size,generator=gen)*2-1; t = torch.randint(0,ncls,(B,),generator=gen)
    pr = {k:v.requires_grad_(True) for k,v in todbl(p).items()}; br = todbl(b); xr = x.double().requires_grad_(True)
    lo = O.CLASSIFIERS[arch](pr, br, xr, True); F.cross_entropy(lo,t).backward()
    for dtype in (torch.float32, torch.bfloat16):
        net = Classifier(arch,ncls,3,size,device="cuda",dtype=dtype); net.load_state_dict({**p,**b})
        logits, ctx = net.forward(x.cuda(), train=True, save=True)
        _, dl, _ = ops.cross_entropy(logits, t.cuda(), 1.0, True)
        net.zero_grad(); dx = net.backward(ctx, dl, True, True)
        gm = max(rm(net.store.g(k), pr[k].grad) for k in p); g2 = max(r2(net.store.g(k), pr[k].grad) for k in p)
        print("%s %s: logits max %.2e l2 %.2e | dx max %.2e l2 %.2e | worst param-grad max %.2e l2 %.2e" % (arch, str(dtype)[6:], rm(logits,lo), r2(logits,lo), rm(dx,xr.grad), r2(dx,xr.grad), gm, g2))
for size,cond in [(32,0),(64,8)]:
    gen = torch.Generator().manual_seed(12+size)
    p = O.init_unet_state(gen, num_classes=cond); B = 4 if size==32 else 2
    x = torch.rand(B,3,size,size,generator=gen)*2-1
    lab = torch.randint(0,max(cond,1),(B,),generator=gen) if cond else None
    w = torch.rand(B,3,size,size,generator=gen)-0.5
    pr = {k:v.requires_grad_(True) for k,v in todbl(p).items()}
    y = O.unet_forward(pr, x.double(), lab, cond if cond else None); (y*w.double()).sum().backward()
    live = [k for k in p if not (k.endswith(".bias") and k not in ("conv0_0.bias","upconv0_0.bias"))]
    for dtype in (torch.float32, torch.bfloat16):
        net = Generator(3,64,cond,device="cuda",dtype=dtype); net.load_state_dict(p)
        yd, ctx = net.forward(x.cuda(), lab.cuda() if cond else None, save=True)
        net.zero_grad(); net.backward(ctx, w.cuda())
        gm = max(rm(net.store.g(k), pr[k].grad) for k in live); g2 = max(r2(net.store.g(k), pr[k].grad) for k in live)
        dead = max(float(net.store.g(k).abs().max()) for k in p if k not in live)
        print("unet%d cond%d %s: y max %.2e l2 %.2e | worst live param-grad max %.2e l2 %.2e | dead-bias |g| %.1e (w-grad scale %.1e)" % (size,cond,str(dtype)[6:], rm(yd,y), r2(yd,y), gm, g2, dead, float(pr["conv1_1.weight"].grad.abs().max())))
