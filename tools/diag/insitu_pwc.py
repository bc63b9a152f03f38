"""Diagnostic: where do the in-situ PWCDCNet gradient differences (CUDA layers vs oracle layers) come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import oracle, ofr_b200
from ofr_b200.harness import reference_env as env, synthetic_pairs
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = False
if len(sys.argv) > 1 and sys.argv[1] == "det":
    torch.backends.cudnn.deterministic = True

def rel(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)

with env.ReferenceEnv("ofr") as e:
    torch.manual_seed(6); net = e.pwcnet().to(DEV)
with env.ReferenceEnv(oracle.MODULES) as e:
    net_o = e.pwcnet().to(DEV)
net_o.load_state_dict(net.state_dict())
x = synthetic_pairs(2, 256, seed=6).to(DEV)
g = torch.Generator(device=DEV).manual_seed(7)

def run(n, cot=None):
    n.zero_grad(set_to_none=True)
    cap = []
    def hook(mod, inp, out):
        a, b = inp
        cap.append((a.detach().clone(), b.detach().clone(), out))
        out.retain_grad()
    h = n.corr.register_forward_hook(hook)
    flows = n(x)
    h.remove()
    if cot is None:
        cot = [torch.randn(f.shape, device=DEV, generator=g) for f in flows]
    sum((f * c).sum() for f, c in zip(flows, cot)).backward()
    grads = {k: p.grad.clone() for k, p in n.named_parameters() if p.grad is not None}
    return flows, grads, cot, [(a, b, o.detach(), o.grad.clone()) for a, b, o in cap]

f1, g1, cot, c1 = run(net)
f2, g2, _, c2 = run(net_o, cot)
f3, g3, _, c3 = run(net_o, cot)
print("flows ofr-vs-oracle:", [f"{rel(a,b):.1e}" for a, b in zip(f1, f2)])
print("flows oracle-vs-oracle:", [f"{rel(a,b):.1e}" for a, b in zip(f3, f2)])
worst = sorted(((rel(g1[k], g2[k]), k) for k in g1), reverse=True)[:8]
print("grads ofr-vs-oracle worst:", [(f"{r:.1e}", k) for r, k in worst])
worst = sorted(((rel(g3[k], g2[k]), k) for k in g3), reverse=True)[:4]
print("grads oracle-vs-oracle worst:", [(f"{r:.1e}", k) for r, k in worst])
# each Correlation call in isolation: same captured inputs and upstream gradient through both layers
for i, ((a, b, o, go), (a2, b2, o2, go2)) in enumerate(zip(c1, c2)):
    A, Bt = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = ofr_b200.Correlation(4, 1, 4, 1, 1, 1)(A, Bt)
    ga, gb = torch.autograd.grad(out, (A, Bt), go)
    Ad, Bd = a.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = oracle.correlation(Ad, Bd, 4, 1, 4, 1, 1)
    ra, rb = torch.autograd.grad(ref, (Ad, Bd), go.double())
    print(f"corr call {i} C{a.shape[1]} {a.shape[2]}x{a.shape[3]}: fwd {(out.double()-ref).abs().max().item():.1e} "
          f"dI1 {rel(ga, ra):.1e} dI2 {rel(gb, rb):.1e} | in-situ: inputs differ {rel(a, a2):.1e}/{rel(b, b2):.1e} "
          f"out {rel(o, o2):.1e} gO {rel(go, go2):.1e}")
