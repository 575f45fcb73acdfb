"""Which GELU does cuBLASLt's epilogue (torch._addmm_activation(use_gelu=True)) apply?  Compared in fp32 against the exact
erf form and the tanh approximation evaluated on the fp32 pre-activation (tools/diag, round-2 note in DESIGN.md section 7)."""
import torch
torch.manual_seed(0)
a = torch.randn(4096, 384, device="cuda", dtype=torch.bfloat16)
w = torch.randn(1536, 384, device="cuda", dtype=torch.bfloat16) * 0.05
b = torch.randn(1536, device="cuda", dtype=torch.bfloat16)
y = torch._addmm_activation(b, a, w.t(), use_gelu=True).float()
lin32 = a.float() @ w.float().t() + b.float()
erf = torch.nn.functional.gelu(lin32)
tanh = torch.nn.functional.gelu(lin32, approximate="tanh")
# where the two forms differ by more than a bf16 ulp the epilogue's choice is visible
d = (erf - tanh).abs()
sel = d > 2e-3 * erf.abs().clamp_min(0.05)
print("elements where erf and tanh forms differ visibly:", int(sel.sum()))
print("mean |y - erf | there:", (y - erf)[sel].abs().mean().item())
print("mean |y - tanh| there:", (y - tanh)[sel].abs().mean().item())
two = torch.nn.functional.gelu(torch.nn.functional.linear(a, w, b)).float()
print("eager two-rounding sequence vs epilogue: mismatch frac", (two != y).float().mean().item(),
      " max rel", ((two - y).abs() / y.abs().clamp_min(1e-3)).max().item())
