import torch, sys, os
sys.path.insert(0, os.getcwd())
from unsamflow_b200 import ops
g = torch.Generator().manual_seed(7)
shape = (8, 32, 64, 208)
x1, x2 = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
go = torch.randn(8, 81, 64, 208, generator=g).cuda()
r1, r2 = ops.corr_bwd(x1, x2, go)
r1, r2 = r1.clone(), r2.clone()
bad = 0
for i in range(300):
    a, b = ops.corr_bwd(x1, x2, go)
    if not (torch.equal(a, r1) and torch.equal(b, r2)):
        bad += 1
o = ops.corr_fwd(x1, x2).clone()
badf = sum(0 if torch.equal(ops.corr_fwd(x1, x2), o) else 1 for _ in range(300))
print("bwd mismatching runs:", bad, "/300   fwd:", badf, "/300")
