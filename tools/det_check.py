"""Debug helper: run-to-run and path-to-path determinism of the gradients, with poisoned free memory."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unetca_b200
from unetca_b200 import _lib
from oracle import unet_ca_port as port
lib = _lib.load()
def poison(val):
    t = torch.full((1 << 28,), val, dtype=torch.float32, device="cuda")   # 1 GiB
    del t
def run(use_se, prec, fused, B=2, H=32, W=48, seed=1, pval=None):
    sd = port.make_state_dict(seed=seed, use_se=use_se)
    x, y = port.make_batch(seed, B, H, W)
    m = unetca_b200.UNet(3, 2, use_se=use_se).cuda().set_precision(prec)
    m.load_state_dict(sd); m.train()
    xc, yc = x.cuda(), y.cuda()
    if pval is not None:
        poison(pval)
    if fused:
        loss = m.loss(xc, yc)
    else:
        loss = torch.nn.CrossEntropyLoss(ignore_index=255)(m(xc), yc)
    loss.backward()
    torch.cuda.synchronize()
    return {n: p.grad.clone() for n, p in m.named_parameters()}, loss.item()
def diff(a, b):
    worst = max(a, key=lambda n: ((a[n] - b[n]).abs().max() / (a[n].abs().max() + 1e-30)).item())
    return ((a[worst] - b[worst]).abs().max() / (a[worst].abs().max() + 1e-30)).item(), worst
for narrow in (0, 1):
    lib.unetca_tc_force_wgrad_narrow(narrow)
    for prec in ("bf16", "fp32"):
        a, la = run(False, prec, True, pval=0.0)
        b, lb = run(False, prec, True, pval=float("nan"))
        c, lc = run(False, prec, True, pval=1e30)
        d, ld = run(False, prec, False, pval=0.0)
        print(f"wgrad_narrow={narrow} {prec}: zero-vs-nan {diff(a, b)} zero-vs-1e30 {diff(a, c)} fused-vs-plain {diff(a, d)} losses {la} {lb} {lc} {ld}")
