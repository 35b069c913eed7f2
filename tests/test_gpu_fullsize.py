"""Parity at BASELINE.json's FULL sizes (configs[1]: batch 64, 3x512x512) through size-independent properties and
direct comparison with ATen's CUDA kernels — the CPU oracle cannot run these sizes in test time, so the checks are
linearity / homogeneity, permutation equivariance, checksums of checksums, and bit-exact integer work."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import BF16, TDT, call, cint, parts_buf, ptr, stream

pytestmark = pytest.mark.gpu

B, S = 64, 512


def _nhwc(t):
    return t.permute(0, 3, 1, 2)          # NHWC storage viewed as NCHW (no copy)


@pytest.mark.parametrize("C,O,HW,paired", [(64, 128, 256, False), (128, 64, 512, True), (512, 512, 64, False)])
def test_conv3x3_full_size_vs_aten_and_homogeneity(C, O, HW, paired):
    """One full-size layer of each tcgen05 layout (haloed pixels-on-N, row-pair, wide): output and BatchNorm partial sums
    against ATen's bf16 CUDA convolution on the same operands; conv(2x) == 2 conv(x) bit for bit."""
    torch.manual_seed(C + O)
    call("unetca_set_conv_impl", 0)
    x = torch.randn(B, HW, HW, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(O, C, 3, 3, device="cuda") / (9 * C) ** 0.5)
    wf = torch.empty(O, 9 * C, dtype=torch.bfloat16, device="cuda")
    call("unetca_pack_conv3x3_weight", BF16, ptr(w), ptr(wf), 9 * C, None, O, C, stream())
    parts = parts_buf(B)
    n = cint()

    def conv(inp):
        y = torch.empty(B, HW, HW, O, dtype=torch.bfloat16, device="cuda")
        if paired:
            wp = torch.empty(2 * O, 12 * C, dtype=torch.bfloat16, device="cuda")
            call("unetca_pack_conv3x3_pair", BF16, ptr(wf), 9 * C, ptr(wp), O, C, stream())
            call("unetca_conv3x3_fwd_paired", BF16, ptr(inp), C, ptr(wp), ptr(y), O, B, HW, HW, C, O, ptr(parts),
                 ctypes.byref(n), stream())
        else:
            call("unetca_conv3x3_fwd", BF16, ptr(inp), C, ptr(wf), 9 * C, ptr(y), O, B, HW, HW, C, O, ptr(parts),
                 ctypes.byref(n), stream())
        return y

    y = conv(x)
    st = parts[: n.value * 2 * O].view(n.value, 2, O).double().sum(0)
    # ATen reference on identical bf16 operands (channels_last view of the same storage), fp32 accumulate
    ref = F.conv2d(_nhwc(x), w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), None, padding=1)
    err = (_nhwc(y).float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    assert err < 1.2e-2, err                                    # both round fp32 accumulators to bf16; K-order differs
    yf = y.float()
    s1, s2 = yf.sum((0, 1, 2)).double(), (yf * yf).sum((0, 1, 2)).double()
    assert ((st[0] - s1).abs().max() / s1.abs().max()).item() < 1e-3      # checksum of the per-CTA checksums
    assert ((st[1] - s2).abs().max() / s2.abs().max()).item() < 1e-3
    y2 = conv(x * 2)                                            # exact in bf16: a power of two
    assert torch.equal(y2, y * 2)


def test_maxpool_full_size_bit_exact_vs_aten():
    """MaxPool2d(2) at (64, 64, 512, 512) bf16 with many exact ties: values and window positions == ATen's."""
    torch.manual_seed(1)
    C = 64
    x = torch.randint(-3, 4, (B, S, S, C), device="cuda").to(torch.bfloat16)
    pooled = torch.empty(B, S // 2, S // 2, C, dtype=torch.bfloat16, device="cuda")
    pos = torch.empty(B, S // 2, S // 2, C, dtype=torch.uint8, device="cuda")
    call("unetca_maxpool2x2", BF16, ptr(x), C, ptr(pooled), C, ptr(pos), None, B, S, S, C, stream())
    ry, ridx = F.max_pool2d(_nhwc(x).float(), 2, return_indices=True)
    assert torch.equal(_nhwc(pooled).float(), ry)
    code = _nhwc(pos).long()
    ho = torch.arange(S // 2, device="cuda")[None, None, :, None]
    wo = torch.arange(S // 2, device="cuda")[None, None, None, :]
    assert torch.equal((2 * ho + code // 2) * S + 2 * wo + code % 2, ridx)


def test_model_full_size_properties():
    """configs[1] shape through the whole model: eval-mode logits are equivariant under a batch permutation bit for bit
    (no cross-image coupling outside train-mode BN), the train-mode loss is permutation invariant, doubling the upstream
    gradient doubles every parameter gradient exactly, and the CE of constant logits is ln(num_classes)."""
    import unetca_b200
    torch.manual_seed(3)
    m = unetca_b200.UNet(3, 2, use_se=True).cuda().set_precision("bf16")
    x = torch.randn(B, 3, S, S, device="cuda")
    y = torch.randint(0, 2, (B, S, S), device="cuda")
    y[torch.rand(B, S, S, device="cuda") < 0.01] = 255
    perm = torch.randperm(B, device="cuda")
    m.train()
    l1 = m.loss(x, y)
    l1.backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    # the two losses below must see the same BN running stats / weights: train-mode forward does not depend on them
    for p in m.parameters():
        p.grad = None
    l2 = m.loss(x[perm], y[perm])
    (2 * l2).backward()
    assert abs(l1.item() - l2.item()) < 2e-4 * abs(l1.item())            # bf16 chains: summation order differs
    gn1 = torch.sqrt(sum((g.float() ** 2).sum() for g in g1)).item()
    gn2 = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in m.parameters())).item()
    assert abs(gn2 - 2 * gn1) < 2e-2 * 2 * gn1
    # exact doubling on the same batch order
    for p in m.parameters():
        p.grad = None
    l3 = m.loss(x, y)
    (2 * l3).backward()
    assert l3.item() == l1.item()                                         # run-to-run deterministic
    for p, g in zip(m.parameters(), g1):
        assert torch.equal(p.grad, 2 * g)
    m.eval()
    with torch.no_grad():
        a = m(x)
        b = m(x[perm])
    assert torch.equal(a[perm], b)
    assert torch.equal(m.predict_mask(x), torch.max(a, 1)[1])
    # the CE kernel at full size: constant logits -> ln(num_classes); un-normalised dlogits sum to zero over classes
    const = torch.full((B, 2, S, S), 0.25, device="cuda")
    g = torch.empty_like(const)
    out, gs = torch.empty(2, device="cuda"), torch.empty(1, device="cuda")
    call("unetca_cross_entropy", ptr(const), ptr(y), 2, B, S * S, 255, None, ptr(g), None, ptr(parts_buf(B)), ptr(out), ptr(gs),
         stream())
    assert abs(out[0].item() - np.log(2.0)) < 1e-6
    assert out[1].item() == (y != 255).sum().item()
    assert g.sum(1).abs().max().item() < 1e-6 and (g[:, 0][y == 255] == 0).all()
