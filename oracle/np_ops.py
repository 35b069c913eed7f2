"""oracle/np_ops.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

A from-scratch numpy restatement, forward and backward, of every arithmetic op on the U-Net-CA hot
path of the reference (`/root/reference/Unet-ChannalAttention.py`, "UCA" below).  The arithmetic of
the reference lives in a third-party dependency, `torch.nn` (no version pinned by the reference;
torch 2.11.0 is what is installed), so every function restates the *published*
semantics of the torch op at the reference call site it cites, including the tie / NaN / ignore
behaviours probed in SURVEY.md §8(a).

Pinning: `tests/test_oracle_cpu.py` checks these functions (a) against the committed golden vectors in
`tests/golden/` that `oracle/make_golden.py` produced by running the UNMODIFIED reference classes in
the build container, and (b) against the torch port in `oracle/unet_ca_port.py`.

All tensors are NCHW numpy arrays, like the reference; compute dtype follows the inputs (use float64
inputs for the arbiter).  Loops are vectorised with im2col so the whole model runs in seconds at
B=2, 32x32.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# conv 3x3, pad 1, stride 1, bias                                            UCA:81,84 (nn.Conv2d)
# ----------------------------------------------------------------------------------------------

def _im2col3x3(x):
    """(B,C,H,W) -> (B,H,W,9*C), tap-major (kh,kw) then channel; zero padding of 1."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((B, H, W, 9, C), dtype=x.dtype)
    for kh in range(3):
        for kw in range(3):
            cols[:, :, :, kh * 3 + kw, :] = xp[:, :, kh:kh + H, kw:kw + W].transpose(0, 2, 3, 1)
    return cols.reshape(B, H, W, 9 * C)


def conv3x3_fwd(x, w, b):
    """y[b,o,h,w] = sum_{c,kh,kw} x[b,c,h+kh-1,w+kw-1] * w[o,c,kh,kw] + b[o]   (UCA:81,84)."""
    B, C, H, W = x.shape
    O = w.shape[0]
    cols = _im2col3x3(x)                                   # (B,H,W,9C)
    wm = w.transpose(0, 2, 3, 1).reshape(O, 9 * C)         # (O, tap, c)
    y = cols.reshape(-1, 9 * C) @ wm.T
    if b is not None:
        y = y + b[None, :]
    return y.reshape(B, H, W, O).transpose(0, 3, 1, 2).copy()


def conv3x3_bwd(x, w, dy):
    """Returns (dx, dw, db) of conv3x3_fwd — what autograd runs at UCA:345 for each Conv2d."""
    B, C, H, W = x.shape
    O = w.shape[0]
    cols = _im2col3x3(x).reshape(-1, 9 * C)
    dym = dy.transpose(0, 2, 3, 1).reshape(-1, O)
    dw = (dym.T @ cols).reshape(O, 3, 3, C).transpose(0, 3, 1, 2).copy()
    db = dym.sum(0)
    # dgrad = conv of dy with the 180-degree rotated, channel-transposed filter
    wr = w[:, :, ::-1, ::-1].transpose(1, 0, 2, 3)          # (C,O,3,3)
    dx = conv3x3_fwd(dy, np.ascontiguousarray(wr), None)
    return dx, dw, db


def conv1x1_fwd(x, w, b):
    """outc: 1x1 conv -> class logits (UCA:125,162)."""
    O, C = w.shape[0], w.shape[1]
    y = np.einsum('bchw,oc->bohw', x, w.reshape(O, C))
    return y + b[None, :, None, None]


def conv1x1_bwd(x, w, dy):
    O, C = w.shape[0], w.shape[1]
    dx = np.einsum('bohw,oc->bchw', dy, w.reshape(O, C))
    dw = np.einsum('bohw,bchw->oc', dy, x).reshape(O, C, 1, 1)
    db = dy.sum((0, 2, 3))
    return dx, dw, db

# ----------------------------------------------------------------------------------------------
# BatchNorm2d                                                                 UCA:82,85
# ----------------------------------------------------------------------------------------------

def bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    """Train mode: normalise with the *biased* batch variance, update running stats with the
    *unbiased* one.  Returns (y, cache, new_running_mean, new_running_var)."""
    B, C, H, W = x.shape
    n = B * H * W
    if n <= 1:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {x.shape}")
    mean = x.mean((0, 2, 3))
    var = x.var((0, 2, 3))                                  # biased
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
    y = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
    new_rm = (1 - momentum) * running_mean + momentum * mean
    new_rv = (1 - momentum) * running_var + momentum * var * (n / (n - 1))
    return y, (xhat, invstd, gamma), new_rm, new_rv


def bn_train_bwd(dy, cache):
    xhat, invstd, gamma = cache
    n = dy.shape[0] * dy.shape[2] * dy.shape[3]
    dbeta = dy.sum((0, 2, 3))
    dgamma = (dy * xhat).sum((0, 2, 3))
    dx = (gamma * invstd)[None, :, None, None] * (
        dy - dbeta[None, :, None, None] / n - xhat * dgamma[None, :, None, None] / n)
    return dx, dgamma, dbeta


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, eps=1e-5):
    """Eval mode (UCA:276): running statistics."""
    scale = gamma / np.sqrt(running_var + eps)
    return x * scale[None, :, None, None] + (beta - running_mean * scale)[None, :, None, None]

# ----------------------------------------------------------------------------------------------
# ReLU                                                                         UCA:83,86
# ----------------------------------------------------------------------------------------------

def relu_fwd(x):
    return np.maximum(x, 0)


def relu_bwd(y, dy):
    return dy * (y > 0)

# ----------------------------------------------------------------------------------------------
# MaxPool2d(2)                                                                 UCA:106-109
# ----------------------------------------------------------------------------------------------

def maxpool2x2_fwd(x):
    """2x2 / stride 2, floor.  First maximum in row-major window order wins; NaN propagates
    (`v > best or isnan(v)`), scanning (0,0),(0,1),(1,0),(1,1) — probed in SURVEY.md §7.3.
    Returns (y, idx) with idx the int64 flat offset h*W+w into the input plane, as torch does."""
    B, C, H, W = x.shape
    Ho, Wo = H // 2, W // 2
    best = x[:, :, 0:2 * Ho:2, 0:2 * Wo:2].copy()
    hh = np.arange(Ho)[:, None] * 2
    ww = np.arange(Wo)[None, :] * 2
    idx = np.broadcast_to((hh * W + ww)[None, None], best.shape).astype(np.int64).copy()
    for dh, dw in ((0, 1), (1, 0), (1, 1)):
        v = x[:, :, dh:2 * Ho:2, dw:2 * Wo:2]
        take = (v > best) | np.isnan(v)
        best = np.where(take, v, best)
        cand = ((hh + dh) * W + (ww + dw))[None, None]
        idx = np.where(take, cand, idx)
    return best, idx


def maxpool2x2_bwd(dy, idx, in_shape):
    B, C, H, W = in_shape
    dx = np.zeros((B, C, H * W), dtype=dy.dtype)
    np.put_along_axis(dx, idx.reshape(B, C, -1), dy.reshape(B, C, -1), axis=2)
    return dx.reshape(B, C, H, W)

# ----------------------------------------------------------------------------------------------
# SELayer                                                                      UCA:45-72
# ----------------------------------------------------------------------------------------------

def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def se_fwd(x, w1, w2):
    """s = sigmoid(W2 relu(W1 mean_hw(x))), y = x*s; W1 (C/r,C), W2 (C,C/r), no biases (UCA:54-59)."""
    p = x.mean((2, 3))                                      # (B,C)    UCA:65
    z = np.maximum(p @ w1.T, 0)                             # (B,C/r)  UCA:55-56
    s = _sigmoid(z @ w2.T)                                  # (B,C)    UCA:57-58
    return x * s[:, :, None, None], (x, p, z, s)            # UCA:72


def se_bwd(dy, cache, w1, w2):
    x, p, z, s = cache
    hw = x.shape[2] * x.shape[3]
    ds = (dy * x).sum((2, 3))                               # (B,C)
    dpre2 = ds * s * (1 - s)
    dw2 = dpre2.T @ z
    dz = (dpre2 @ w2) * (z > 0)
    dw1 = dz.T @ p
    dp = dz @ w1
    dx = dy * s[:, :, None, None] + dp[:, :, None, None] / hw
    return dx, dw1, dw2

# ----------------------------------------------------------------------------------------------
# ConvTranspose2d k=2 s=2                                                      UCA:112,115,118,121
# ----------------------------------------------------------------------------------------------

def convT2x2_fwd(x, w, b):
    """out[b,o,2i+d,2j+e] = sum_c x[b,c,i,j] * w[c,o,d,e] + b[o]; weight layout (Cin,Cout,2,2)."""
    B, C, H, W = x.shape
    O = w.shape[1]
    t = np.einsum('bcij,code->boidje', x, w)                # (B,O,H,2,W,2)
    return t.reshape(B, O, 2 * H, 2 * W) + b[None, :, None, None]


def convT2x2_bwd(x, w, dy):
    B, C, H, W = x.shape
    O = w.shape[1]
    g = dy.reshape(B, O, H, 2, W, 2)
    dx = np.einsum('boidje,code->bcij', g, w)
    dw = np.einsum('boidje,bcij->code', g, x)
    db = dy.sum((0, 2, 3))
    return dx, dw, db

# ----------------------------------------------------------------------------------------------
# CrossEntropyLoss(ignore_index=255), reduction='mean'                         UCA:465, 344
# ----------------------------------------------------------------------------------------------

def bilinear_taps(n_in, n_out):
    """(i0, i1, lam) per output index of a 1-D bilinear resize, align_corners=False (the resize guard, UCA:138-157).
    torchvision's F_T.resize(antialias=True) only up-samples here (2*floor(h/2) -> h), where the antialias filter
    reduces to these two taps (weights renormalised at the borders == index clamping)."""
    scale = n_in / n_out
    src = np.maximum((np.arange(n_out) + 0.5) * scale - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    return i0, i1, src - i0


def resize_bilinear_fwd(x, size):
    """(B,C,h,w) -> (B,C,H,W)"""
    h0, h1, lh = bilinear_taps(x.shape[2], size[0])
    w0, w1, lw = bilinear_taps(x.shape[3], size[1])
    rows = x[:, :, h0, :] * (1 - lh)[None, None, :, None] + x[:, :, h1, :] * lh[None, None, :, None]
    return rows[:, :, :, w0] * (1 - lw) + rows[:, :, :, w1] * lw


def resize_bilinear_bwd(dy, in_size):
    """adjoint of resize_bilinear_fwd: (B,C,H,W) -> (B,C,h,w)"""
    B, C, H, W = dy.shape
    h0, h1, lh = bilinear_taps(in_size[0], H)
    w0, w1, lw = bilinear_taps(in_size[1], W)
    tmp = np.zeros((B, C, H, in_size[1]), dtype=dy.dtype)
    np.add.at(tmp, (slice(None), slice(None), slice(None), w0), dy * (1 - lw))
    np.add.at(tmp, (slice(None), slice(None), slice(None), w1), dy * lw)
    dx = np.zeros((B, C, in_size[0], in_size[1]), dtype=dy.dtype)
    np.add.at(dx, (slice(None), slice(None), h0, slice(None)), tmp * (1 - lh)[None, None, :, None])
    np.add.at(dx, (slice(None), slice(None), h1, slice(None)), tmp * lh[None, None, :, None])
    return dx


def cross_entropy_fwd(logits, target, ignore_index=255):
    """-mean_{t != ignore} log_softmax(logits)[t].  All-ignored -> NaN (0/0), like torch."""
    m = logits.max(1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(1, keepdims=True))
    logp = logits - lse
    valid = target != ignore_index
    t = np.where(valid, target, 0)
    picked = np.take_along_axis(logp, t[:, None], axis=1)[:, 0]
    nvalid = valid.sum()
    with np.errstate(invalid='ignore', divide='ignore'):
        loss = -(picked * valid).sum() / logits.dtype.type(nvalid)
    return loss, (logp, t, valid, nvalid)


def cross_entropy_bwd(cache, gout=1.0):
    logp, t, valid, nvalid = cache
    p = np.exp(logp)
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, t[:, None], 1.0, axis=1)
    with np.errstate(invalid='ignore', divide='ignore'):
        return (p - onehot) * valid[:, None] * (gout / logp.dtype.type(nvalid))


def argmax_mask(logits):
    """torch.max(outputs, 1)[1] (UCA:220): ties -> lowest class index."""
    return np.argmax(logits, axis=1).astype(np.int64)

# ----------------------------------------------------------------------------------------------
# Whole model, forward + backward, keyed by the reference's state_dict names   UCA:75-163
# ----------------------------------------------------------------------------------------------

def _dc_fwd(x, sd, pre, use_se, train, new_stats):
    c = {}
    def bn(v, i):
        g, b_ = sd[f'{pre}.{i}.weight'], sd[f'{pre}.{i}.bias']
        rm, rv = sd[f'{pre}.{i}.running_mean'], sd[f'{pre}.{i}.running_var']
        if train:
            y, cache, nrm, nrv = bn_train_fwd(v, g, b_, rm, rv)
            new_stats[f'{pre}.{i}.running_mean'] = nrm
            new_stats[f'{pre}.{i}.running_var'] = nrv
            return y, cache
        return bn_eval_fwd(v, g, b_, rm, rv), None
    c['x0'] = x
    y1 = conv3x3_fwd(x, sd[f'{pre}.0.weight'], sd[f'{pre}.0.bias'])
    n1, c['bn1'] = bn(y1, 1)
    a1 = relu_fwd(n1); c['a1'] = a1
    y2 = conv3x3_fwd(a1, sd[f'{pre}.3.weight'], sd[f'{pre}.3.bias'])
    n2, c['bn2'] = bn(y2, 4)
    a2 = relu_fwd(n2); c['a2'] = a2
    if use_se:
        out, c['se'] = se_fwd(a2, sd[f'{pre}.6.fc.0.weight'], sd[f'{pre}.6.fc.2.weight'])
    else:
        out = a2
    return out, c


def _dc_bwd(g, c, sd, pre, use_se, grads):
    if use_se:
        g, dw1, dw2 = se_bwd(g, c['se'], sd[f'{pre}.6.fc.0.weight'], sd[f'{pre}.6.fc.2.weight'])
        grads[f'{pre}.6.fc.0.weight'] = dw1
        grads[f'{pre}.6.fc.2.weight'] = dw2
    g = relu_bwd(c['a2'], g)
    g, grads[f'{pre}.4.weight'], grads[f'{pre}.4.bias'] = bn_train_bwd(g, c['bn2'])
    g, grads[f'{pre}.3.weight'], grads[f'{pre}.3.bias'] = conv3x3_bwd(c['a1'], sd[f'{pre}.3.weight'], g)
    g = relu_bwd(c['a1'], g)
    g, grads[f'{pre}.1.weight'], grads[f'{pre}.1.bias'] = bn_train_bwd(g, c['bn1'])
    g, grads[f'{pre}.0.weight'], grads[f'{pre}.0.bias'] = conv3x3_bwd(c['x0'], sd[f'{pre}.0.weight'], g)
    return g


ENC = (('inc', 'inc.double_conv'), ('down1', 'down1.1.double_conv'), ('down2', 'down2.1.double_conv'),
       ('down3', 'down3.1.double_conv'), ('down4', 'down4.1.double_conv'))
DEC = (('up1', 'conv1'), ('up2', 'conv2'), ('up3', 'conv3'), ('up4', 'conv4'))


def unet_forward(x, sd, use_se=True, train=True):
    """UNet.forward (UCA:127-163).  Returns (logits, cache, new_running_stats, pool_indices)."""
    if x.shape[2] % 16 or x.shape[3] % 16:
        raise ValueError("oracle restates the fast path only: H, W must be multiples of 16 (UCA:138-157 resize guard not taken)")
    cache, new_stats, pool_idx = {}, {}, []
    skips = []
    h = x
    for li, (name, pre) in enumerate(ENC):
        if li > 0:
            cache[f'{name}.pool_in'] = h.shape
            h, idx = maxpool2x2_fwd(h)                      # UCA:106-109
            cache[f'{name}.pool_idx'] = idx
            pool_idx.append(idx)
        h, cache[name] = _dc_fwd(h, sd, pre, use_se, train, new_stats)
        skips.append(h)
    for di, (up, conv) in enumerate(DEC):
        skip = skips[3 - di]
        cache[f'{up}.x'] = h
        u = convT2x2_fwd(h, sd[f'{up}.weight'], sd[f'{up}.bias'])       # UCA:136,143,149,155
        h = np.concatenate([skip, u], axis=1)               # skip first (UCA:140,146,152,158)
        h, cache[conv] = _dc_fwd(h, sd, f'{conv}.double_conv', use_se, train, new_stats)
    cache['outc.x'] = h
    logits = conv1x1_fwd(h, sd['outc.weight'], sd['outc.bias'])         # UCA:162
    return logits, cache, new_stats, pool_idx


def unet_backward(dlogits, cache, sd, use_se=True):
    """Gradients of every parameter (dict keyed like state_dict) — autograd at UCA:345."""
    grads = {}
    g, grads['outc.weight'], grads['outc.bias'] = conv1x1_bwd(cache['outc.x'], sd['outc.weight'], dlogits)
    skip_g = [None] * 4
    for di in (3, 2, 1, 0):
        up, conv = DEC[di]
        g = _dc_bwd(g, cache[conv], sd, f'{conv}.double_conv', use_se, grads)
        C = g.shape[1] // 2
        skip_g[3 - di] = g[:, :C]
        g, grads[f'{up}.weight'], grads[f'{up}.bias'] = convT2x2_bwd(cache[f'{up}.x'], sd[f'{up}.weight'], g[:, C:])
    for li in (4, 3, 2, 1, 0):
        name, pre = ENC[li]
        if li < 4:
            g = g + skip_g[li]
        g = _dc_bwd(g, cache[name], sd, pre, use_se, grads)
        if li > 0:
            g = maxpool2x2_bwd(g, cache[f'{name}.pool_idx'], cache[f'{name}.pool_in'])
    grads['input'] = g
    return grads
