"""oracle/unet_ca_port.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU port of the reference hot path (`/root/reference/Unet-ChannalAttention.py:45-163` + the
`nn.CrossEntropyLoss(ignore_index=255)` of `:465`) written against `torch.nn.functional`, driven by a
plain `state_dict`-keyed dict of tensors instead of `nn.Module`s.  It calls the same ATen CPU kernels
the reference's modules dispatch to (conv2d / batch_norm / relu / max_pool2d / conv_transpose2d /
linear / sigmoid / cross_entropy), so it is both

  * the checker for the CUDA path at sizes where the numpy restatement (`oracle/np_ops.py`) is too
    slow (configs[0]: B=4, 3x256x256), and
  * the `cpu_baseline` / `--impl reference` arm of `bench.py` on the GPU box, where
    `/root/reference` does not exist (kind = "port").

Pinned against the unmodified reference classes by `oracle/make_golden.py` (run in the build
container, where the reference is importable) -> `tests/golden/*.npz`, checked in
`tests/test_oracle_cpu.py`.

Also holds the seeded fixture generators (`make_state_dict`, `make_batch`) shared by the golden
generator, the tests and the bench; they use `numpy.random.RandomState`, whose streams are frozen
across numpy versions, so the fixtures can be regenerated anywhere instead of being committed
(a full state_dict is 125 MB).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

ENC = (('inc', 'inc.double_conv'), ('down1', 'down1.1.double_conv'), ('down2', 'down2.1.double_conv'),
       ('down3', 'down3.1.double_conv'), ('down4', 'down4.1.double_conv'))
DEC = (('up1', 'conv1'), ('up2', 'conv2'), ('up3', 'conv3'), ('up4', 'conv4'))
WIDTHS = (64, 128, 256, 512, 1024)


def state_dict_spec(in_channels=3, num_classes=2, use_se=True, reduction=16):
    """Ordered (key, shape) list identical to `UNet(in_channels, num_classes, use_se).state_dict()`
    of the reference (UCA:101-125; 154 keys with SE, 136 without — SURVEY.md §8b)."""
    spec = []

    def dc(pre, cin, cout):
        for i, ci in ((0, cin), (3, cout)):
            spec.append((f'{pre}.{i}.weight', (cout, ci, 3, 3)))
            spec.append((f'{pre}.{i}.bias', (cout,)))
            j = i + 1
            spec.append((f'{pre}.{j}.weight', (cout,)))
            spec.append((f'{pre}.{j}.bias', (cout,)))
            spec.append((f'{pre}.{j}.running_mean', (cout,)))
            spec.append((f'{pre}.{j}.running_var', (cout,)))
            spec.append((f'{pre}.{j}.num_batches_tracked', ()))
        if use_se:
            spec.append((f'{pre}.6.fc.0.weight', (cout // reduction, cout)))
            spec.append((f'{pre}.6.fc.2.weight', (cout, cout // reduction)))

    cin = in_channels
    for (name, pre), c in zip(ENC, WIDTHS):
        dc(pre, cin, c)
        cin = c
    for (up, conv), c in zip(DEC, (512, 256, 128, 64)):
        spec.append((f'{up}.weight', (2 * c, c, 2, 2)))
        spec.append((f'{up}.bias', (c,)))
        dc(f'{conv}.double_conv', 2 * c, c)
    spec.append(('outc.weight', (num_classes, 64, 1, 1)))
    spec.append(('outc.bias', (num_classes,)))
    return spec


def make_state_dict(seed=0, in_channels=3, num_classes=2, use_se=True, dtype=torch.float32):
    """Deterministic weights with the reference's default-init *scales* (uniform +-1/sqrt(fan_in)),
    non-trivial BN affine (gamma ~ 1 + 0.1 N, beta ~ 0.1 N) and fresh running stats."""
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for key, shape in state_dict_spec(in_channels, num_classes, use_se):
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith('running_mean'):
            sd[key] = torch.zeros(shape, dtype=dtype)
        elif key.endswith('running_var'):
            sd[key] = torch.ones(shape, dtype=dtype)
        elif len(shape) == 1 and ('.1.' in key or '.4.' in key):           # BN affine
            if key.endswith('weight'):
                v = 1.0 + 0.1 * rs.standard_normal(shape)
            else:
                v = 0.1 * rs.standard_normal(shape)
            sd[key] = torch.from_numpy(v).to(dtype)
        else:
            if len(shape) == 4 and key.startswith('up'):                   # ConvTranspose2d (Cin,Cout,2,2)
                fan_in = shape[1] * 4
            elif len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
            elif len(shape) == 2:
                fan_in = shape[1]
            else:                                                         # conv / convT / outc bias
                fan_in = None
            if fan_in is None:
                v = rs.uniform(-0.05, 0.05, shape)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                v = rs.uniform(-bound, bound, shape)
            sd[key] = torch.from_numpy(v).to(dtype)
    return sd


def make_batch(seed, B, H, W, in_channels=3, num_classes=2, ignore_frac=0.01):
    """x ~ N(0,1) (B,Cin,H,W) fp32; y ~ randint{0..num_classes-1} with ~1 % pixels set to 255."""
    rs = np.random.RandomState(1000 + seed)
    x = rs.standard_normal((B, in_channels, H, W)).astype(np.float32)
    y = rs.randint(0, num_classes, (B, H, W)).astype(np.int64)
    if ignore_frac > 0:
        y[rs.uniform(size=(B, H, W)) < ignore_frac] = 255
    return torch.from_numpy(x), torch.from_numpy(y)


def make_structured_batch(seed, B, H, W, ignore_frac=0.01):
    """make_batch with LEARNABLE labels: class = sign of the 9x9 box mean of input channel 0 (the ~1 % ignored pixels
    of make_batch are kept).  Training on fresh batches of this kind is well conditioned — the unmodified reference
    run in fp32 stays within 1e-3 of its own fp64 run over 100 Adam steps — whereas memorising the random labels of
    make_batch is chaotic (the reference departs from its own fp64 / other-thread-count run by 5 % after ~60 steps)."""
    x, y = make_batch(seed, B, H, W, ignore_frac=ignore_frac)
    m = F.avg_pool2d(x[:, :1], 9, 1, 4)[:, 0]
    y2 = (m > 0).long()
    y2[y == 255] = 255
    return x, y2


def trajectory_batch(mode, step, B, H, W):
    """Batch of train step `step` of the two trajectory fixtures (tests/golden/trajectory_*.npz)."""
    if mode == "random4":                      # four fixed batches with random labels, cycled
        return make_batch(1000 + step % 4, B, H, W)
    if mode == "struct":                       # a fresh batch with learnable labels every step
        return make_structured_batch(1000 + step, B, H, W)
    raise ValueError(mode)


# ------------------------------------------------------------------------------------------------

def _double_conv(x, p, pre, use_se, train, bufs):
    """DoubleConv.forward (UCA:80-97): [conv3x3 -> BN -> ReLU] x2 [-> SE]."""
    for i in (0, 3):
        x = F.conv2d(x, p[f'{pre}.{i}.weight'], p[f'{pre}.{i}.bias'], padding=1)           # UCA:81,84
        j = i + 1
        x = F.batch_norm(x, bufs[f'{pre}.{j}.running_mean'], bufs[f'{pre}.{j}.running_var'],
                         p[f'{pre}.{j}.weight'], p[f'{pre}.{j}.bias'], training=train,
                         momentum=0.1, eps=1e-5)                                             # UCA:82,85
        if train:
            bufs[f'{pre}.{j}.num_batches_tracked'] += 1
        x = F.relu(x)                                                                        # UCA:83,86
    if use_se:
        x = se_layer(x, p[f'{pre}.6.fc.0.weight'], p[f'{pre}.6.fc.2.weight'])
    return x


def se_layer(x, w1, w2):
    """SELayer.forward (UCA:61-72) on its own: x * sigmoid(W2 relu(W1 mean_hw(x)))."""
    b, c = x.shape[:2]
    y = F.adaptive_avg_pool2d(x, 1).view(b, c)                                               # UCA:65
    y = F.relu(F.linear(y, w1))                                                              # UCA:55-56
    y = torch.sigmoid(F.linear(y, w2))                                                       # UCA:57-58
    return x * y.view(b, c, 1, 1)                                                            # UCA:72


def double_conv(x, p, pre, use_se, train):
    """DoubleConv.forward (UCA:96-97) on its own, parameters / buffers keyed `pre.<index>.<name>` as in the state_dict."""
    return _double_conv(x, p, pre, use_se, train, p)


def unet_forward(x, p, bufs=None, use_se=True, train=True, return_aux=False):
    """UNet.forward (UCA:127-163) on a dict of tensors.  `bufs` (running stats) defaults to `p`
    and is updated in place in train mode, exactly like the modules' buffers."""
    if bufs is None:
        bufs = p
    skips, pool_idx = [], []
    h = x
    for li, (name, pre) in enumerate(ENC):
        if li > 0:
            h, idx = F.max_pool2d(h, 2, return_indices=True)                                 # UCA:106-109
            pool_idx.append(idx)
        h = _double_conv(h, p, pre, use_se, train, bufs)
        skips.append(h)
    for di, (up, conv) in enumerate(DEC):
        u = F.conv_transpose2d(h, p[f'{up}.weight'], p[f'{up}.bias'], stride=2)              # UCA:136..155
        skip = skips[3 - di]
        if u.shape[2:] != skip.shape[2:]:
            # resize guard (UCA:138-139 ...): torchvision F_T.resize(tensor, size, BILINEAR) is
            # interpolate(mode='bilinear', align_corners=False, antialias=True); only taken when H or W is not in 16*N
            u = F.interpolate(u, size=skip.shape[2:], mode='bilinear', align_corners=False, antialias=True)
        h = torch.cat([skip, u], dim=1)                                                      # UCA:140..158
        h = _double_conv(h, p, f'{conv}.double_conv', use_se, train, bufs)
    logits = F.conv2d(h, p['outc.weight'], p['outc.bias'])                                   # UCA:162
    if return_aux:
        return logits, {'pool_idx': pool_idx, 'skips': skips}
    return logits


def loss_fn(logits, target):
    """nn.CrossEntropyLoss(ignore_index=255) (UCA:465, called UCA:344)."""
    return F.cross_entropy(logits, target, ignore_index=255)


def train_step_grads(sd, x, y, use_se=True, dtype=torch.float32):
    """One reference train-step forward+backward (UCA:343-345) -> (logits, loss, grads dict, new bufs)."""
    p = {k: (v.detach().clone().to(dtype).requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k
             else v.detach().clone()) for k, v in sd.items()}
    for k in p:
        if 'running' in k:
            p[k] = p[k].to(dtype)
    logits, aux = unet_forward(x.to(dtype), p, use_se=use_se, train=True, return_aux=True)
    loss = loss_fn(logits, y)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items() if v.requires_grad}
    bufs = {k: v.detach() for k, v in p.items() if not v.requires_grad}
    return logits.detach(), loss.detach(), grads, bufs, aux
