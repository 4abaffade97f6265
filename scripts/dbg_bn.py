import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from combat_b200 import ops
def rel(a,b): return float((a.detach().float().cpu().double()-b.detach().double()).abs().max()/b.detach().double().abs().max())
g = torch.Generator().manual_seed(0)
for (N,C,H) in [(8,128,16),(8,64,32),(8,256,8),(6,64,8)]:
    x = torch.randn(N,C,H,H,generator=g).double().requires_grad_(True)
    gamma = torch.ones(C).double().requires_grad_(True); beta = torch.zeros(C).double().requires_grad_(True)
    y = F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5))
    dy = (torch.randn(N,C,H,H,generator=g)*1e-3).double()
    y.backward(dy)
    nh = lambda t: t.permute(0,2,3,1).contiguous().float().cuda()
    xd, dyd = nh(x.detach()), nh(dy)
    rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
    sc, sh, mean, invstd = ops.bn_train_prepare(xd, N*H*H, C, gamma.detach().float().cuda(), beta.detach().float().cuda(), rm, rv, 0.1, 1e-5)
    yd = ops.affine_act(xd, sc, sh, True)
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dx, _ = ops.bn_bwd_train(dyd, xd, yd, gamma.detach().float().cuda(), mean, invstd, True, dg, db)
    print((N,C,H), "y", rel(yd.permute(0,3,1,2), y), "dx", rel(dx.permute(0,3,1,2), x.grad), "dgamma", rel(dg, gamma.grad), "dbeta", rel(db, beta.grad))
