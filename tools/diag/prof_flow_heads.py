"""Driver for `ncu -k regex:conv3_|deconv4_ ...`: two forward + backward passes of predict_flow2 (565 -> 2, 128^2) and upfeat3 (597 -> 2,
64^2) at the bench batch (profiles/r02_ncu_flow_heads.txt)."""
import sys, os
sys.path.insert(0, "/root/repo")
import torch, ofr_b200
B=24
for kind,C,h in (("c",565,128),("d",597,64)):
    m = (ofr_b200.FlowConv2d(C,2,3,1,1) if kind=="c" else ofr_b200.FlowUpsample2d(C,2,4,2,1)).cuda()
    x = torch.randn(B,C,h,h,device="cuda",requires_grad=True)
    for _ in range(2):
        out = m(x)
        torch.autograd.grad(out,[x,m.weight,m.bias],torch.ones_like(out))
torch.cuda.synchronize()
