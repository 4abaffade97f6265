import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import combat_oracle as O
from combat_b200.nets import Classifier
from combat_b200 import ops
def r2(a,b):
    a=a.detach().float().cpu().double(); b=b.detach().double(); return float((a-b).norm()/b.norm().clamp_min(1e-30))
gen = torch.Generator().manual_seed(11)
p,b = O.init_preact_resnet18_state(gen)
x = torch.rand(8,3,32,32,generator=gen)*2-1; t = torch.randint(0,10,(8,),generator=gen)
pr = {k:v.clone().double().requires_grad_(True) for k,v in p.items()}
br = {k:(v.clone().double() if v.is_floating_point() else v.clone()) for k,v in b.items()}
xr = x.double().requires_grad_(True)
lo = O.preact_resnet18_forward(pr, br, xr, True); F.cross_entropy(lo,t).backward()
for use_tc in (False, True):
    net = Classifier("preact_resnet18",10,3,32,device="cuda",dtype=torch.bfloat16, use_tc=use_tc)
    net.load_state_dict({**p,**b})
    logits, ctx = net.forward(x.cuda(), train=True, save=True)
    _, dl, _ = ops.cross_entropy(logits, t.cuda(), 1.0, True)
    net.zero_grad(); dx = net.backward(ctx, dl, True, True)
    print("use_tc", use_tc, "logits", r2(logits, lo), "dx", r2(dx, xr.grad))
    for k in p:
        if "conv" in k or "shortcut" in k or k.startswith("linear"): print("   %-30s %.2e" % (k, r2(net.store.g(k), pr[k].grad)))
