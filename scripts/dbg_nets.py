import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import combat_oracle as O
from combat_b200.nets import Classifier
from combat_b200 import ops
def rel(a,b): return float((a.detach().float().cpu().double()-b.detach().double()).abs().max()/b.detach().double().abs().max())
gen = torch.Generator().manual_seed(4)
p,b = O.init_preact_resnet18_state(gen)
x = torch.rand(8,3,32,32,generator=gen)*2-1; t = torch.randint(0,10,(8,),generator=gen)
pr = {k:v.clone().double().requires_grad_(True) for k,v in p.items()}
br = {k:(v.clone().double() if v.is_floating_point() else v.clone()) for k,v in b.items()}
xr = x.clone().double().requires_grad_(True)
lo = O.preact_resnet18_forward(pr, br, xr, True); F.cross_entropy(lo,t).backward()
net = Classifier("preact_resnet18",10,3,32,device="cuda",dtype=torch.float32)
net.load_state_dict({**p,**b})
logits, ctx = net.forward(x.cuda(), train=True, save=True)
_, dl, _ = ops.cross_entropy(logits, t.cuda(), 1.0, True)
net.zero_grad(); dx = net.backward(ctx, dl, True, True)
print("logits", rel(logits, lo), "dx", rel(dx, xr.grad))
for k in p: print("%-32s %.2e" % (k, rel(net.store.g(k), pr[k].grad)))
