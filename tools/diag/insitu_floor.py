"""Diagnostic: distribution of the in-situ noise floor (several perturbed oracles) vs the CUDA layers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import oracle, ofr_b200
from ofr_b200.harness import reference_env as env, synthetic_pairs
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = False
torch.backends.cudnn.deterministic = True

class Corr64(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__(); self.inner = oracle.MODULES["Correlation"](*a, **k)
    def forward(self, a, b): return self.inner(a.double(), b.double()).float()
class CorrFlip(torch.nn.Module):   # fp32, channels summed in the opposite order
    def __init__(self, *a, **k):
        super().__init__(); self.inner = oracle.MODULES["Correlation"](*a, **k)
    def forward(self, a, b): return self.inner(a.flip(1), b.flip(1))
class CorrNoise(torch.nn.Module):  # fp64 then multiplicative last-bit noise
    def __init__(self, *a, **k):
        super().__init__(); self.inner = oracle.MODULES["Correlation"](*a, **k); self.g = torch.Generator(device=DEV).manual_seed(3)
    def forward(self, a, b):
        o = self.inner(a.double(), b.double())
        return (o * (1 + 6e-8 * torch.randn(o.shape, device=o.device, generator=self.g, dtype=torch.float64))).float()
class CorrOfrFwdOracleBwd(torch.nn.Module):  # CUDA forward values, oracle backward (which side carries the difference?)
    def __init__(self, *a, **k):
        super().__init__(); self.inner = oracle.MODULES["Correlation"](*a, **k); self.ofr = ofr_b200.Correlation(*a, **k)
    def forward(self, a, b):
        o = self.inner(a, b)
        with torch.no_grad():
            d = self.ofr(a, b) - o
        return o + d
class CorrOracleFwdOfrBwd(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__(); self.inner = oracle.MODULES["Correlation"](*a, **k); self.ofr = ofr_b200.Correlation(*a, **k)
    def forward(self, a, b):
        o = self.ofr(a, b)
        with torch.no_grad():
            d = self.inner(a, b) - o
        return o + d

def mods(c): return {"Correlation": c, "Resample2d": oracle.MODULES["Resample2d"], "ChannelNorm": oracle.MODULES["ChannelNorm"]}
def l2(ga, gb):
    num = sum(float((ga[n].double() - gb[n].double()).pow(2).sum()) for n in gb)
    den = sum(float(gb[n].double().pow(2).sum()) for n in gb)
    return (num / den) ** 0.5

with env.ReferenceEnv("ofr") as e:
    torch.manual_seed(6); net = e.pwcnet().to(DEV)
nets = {"ofr": net}
for name, c in [("o64", Corr64), ("o32", oracle.MODULES["Correlation"]), ("flip", CorrFlip), ("noise", CorrNoise),
                ("ofrF_oraB", CorrOfrFwdOracleBwd), ("oraF_ofrB", CorrOracleFwdOfrBwd)]:
    with env.ReferenceEnv(mods(c)) as e:
        n = e.pwcnet().to(DEV)
    n.load_state_dict(net.state_dict()); nets[name] = n
for seed in (6, 11):
    x = synthetic_pairs(2, 256, seed=seed).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(7)
    cot = None; res = {}
    for name, n in nets.items():
        n.zero_grad(set_to_none=True)
        flows = n(x)
        if cot is None: cot = [torch.randn(f.shape, device=DEV, generator=g) for f in flows]
        sum((f * c).sum() for f, c in zip(flows, cot)).backward()
        res[name] = {k: p.grad.clone() for k, p in n.named_parameters() if p.grad is not None}
    print(f"seed {seed}: L2 distance of weight gradients to the float64-oracle network:",
          {k: f"{l2(v, res['o64']):.2e}" for k, v in res.items() if k != "o64"})
