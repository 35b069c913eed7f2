"""Host-side logic of unetca_b200.optim.Adam that needs no GPU: constructor contract, the
torch.optim.Adam state_dict layout, the conv filters it claims for the packing kernel, and the refusal to step on the
CPU (there is no CPU path)."""
import pytest
import torch

import unetca_b200
from unetca_b200 import optim as uoptim


def test_constructor_contract():
    p = [torch.nn.Parameter(torch.zeros(4))]
    with pytest.raises(NotImplementedError):
        uoptim.Adam(p, amsgrad=True)
    with pytest.raises(NotImplementedError):
        uoptim.Adam(p, maximize=True)
    with pytest.raises(ValueError):
        uoptim.Adam(p, lr=-1.0)
    with pytest.raises(ValueError):
        uoptim.Adam(p, betas=(0.9, 1.0))
    opt = uoptim.Adam(p, lr=3e-4, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.01)
    g = opt.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (3e-4, (0.8, 0.95), 1e-6, 0.01)
    # same param_groups keys as torch.optim.Adam needs to load our state_dict (it fills the rest with its defaults)
    t = torch.optim.Adam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    t.load_state_dict(opt.state_dict())
    assert t.param_groups[0]["lr"] == 3e-4 and t.param_groups[0]["betas"] == (0.8, 0.95)


def test_no_cpu_step():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uoptim.Adam([p]).step()
    q = torch.nn.Parameter(torch.zeros(4))          # no gradient: nothing to do, no error
    uoptim.Adam([q]).step()


def test_conv_filters_claimed_for_the_packing_kernel():
    m = unetca_b200.UNet(3, 2, use_se=True)
    opt = uoptim.Adam(m.parameters(), lr=1e-4, model=m)
    claimed = {id(c.weight) for c in opt._conv_of.values()}
    names = {n for n, p in m.named_parameters() if id(p) in claimed}
    # every 3x3 filter except the first conv (3 input channels: packed in its own pixel-pair layout by the forward)
    want = {n for n, p in m.named_parameters() if p.dim() == 4 and p.shape[2:] == (3, 3) and p.shape[1] % 32 == 0}
    assert names == want and len(names) == 17
    assert "inc.double_conv.0.weight" not in names
    total = sum(p.numel() for p in m.parameters())
    assert sum(p.numel() for n, p in m.named_parameters() if n in names) / total > 0.9


def test_load_state_dict_rejects_mixed_step_counts():
    ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    t = torch.optim.Adam(ps, lr=1e-3)
    ps[0].grad = torch.ones(4)
    t.step()                                          # only the first parameter has stepped
    ps[1].grad = torch.ones(3)
    t.step()
    opt = uoptim.Adam([torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))], lr=1e-3)
    with pytest.raises(ValueError, match="one step count per parameter group"):
        opt.load_state_dict(t.state_dict())
