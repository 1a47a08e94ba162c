"""Dev tool: PyTorch's GroupNorm(4, 32) on [B,32,11,14] and LayerNorm(32) on [154*B, 32] against equivalent
formulations built from better-behaved kernels (forward + backward, bf16 autocast like the trainer)."""
import torch, torch.nn as nn, torch.nn.functional as F

dev = "cuda"
B = 8192

def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def gn_fast(x, gn):
    Bn, C, H, W = x.shape
    G = gn.num_groups
    y = F.layer_norm(x.reshape(Bn, G, -1), (C // G * H * W,), None, None, gn.eps)
    return torch.addcmul(gn.bias.view(1, C, 1, 1), y.view(Bn, C, H, W), gn.weight.view(1, C, 1, 1))

def ln_small(x, ln):
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    xc = xf - mu
    var = (xc * xc).mean(-1, keepdim=True)
    return (xc * torch.rsqrt(var + ln.eps)) * ln.weight + ln.bias

def ln_var_mean(x, ln):
    xf = x.float()
    var, mu = torch.var_mean(xf, dim=-1, correction=0, keepdim=True)
    return torch.addcmul(ln.bias, (xf - mu) * torch.rsqrt(var + ln.eps), ln.weight)

for amp in (False, True):
    gn = nn.GroupNorm(4, 32).to(dev)
    x = torch.randn(B, 32, 11, 14, device=dev, requires_grad=True)
    ln = nn.LayerNorm(32).to(dev)
    z = torch.randn(154, 4096, 32, device=dev, requires_grad=True)
    def run(f, inp):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            xi = inp.to(torch.bfloat16) if amp else inp
            y = f(xi)
        y.float().sum().backward()
    print(f"amp={amp}")
    print("  GroupNorm   torch %.3f ms   fast %.3f ms" % (timed(lambda: run(gn, x)), timed(lambda: run(lambda t: gn_fast(t, gn), x))))
    print("  LayerNorm32 torch %.3f ms   explicit %.3f ms   var_mean %.3f ms" % (
        timed(lambda: run(ln, z)), timed(lambda: run(lambda t: ln_small(t, ln), z)), timed(lambda: run(lambda t: ln_var_mean(t, ln), z))))
    with torch.no_grad():
        print("  max diff GN %.2e  LN %.2e %.2e" % ((gn(x) - gn_fast(x, gn)).abs().max().item(), (ln(z) - ln_small(z, ln)).abs().max().item(),
                                                  (ln(z) - ln_var_mean(z, ln)).abs().max().item()))
