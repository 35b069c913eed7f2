"""Drop-in `UNet` / `DoubleConv` / `SELayer` for the reference's U-Net-CA (Unet-ChannalAttention.py:45-163).

Same constructor signatures, same attribute tree and therefore the same 154 (136 without SE) `state_dict` keys and
OIHW fp32 parameter layout as the reference, same NCHW float input -> NCHW float class-logit output, same
train()/eval() BatchNorm semantics.  The child modules (nn.Conv2d, nn.BatchNorm2d, ...) are only parameter
containers: `UNet.forward` never calls them.  Forward and backward run as hand-written sm_100a CUDA kernels behind
the C ABI in include/unetca_b200.h (NHWC activations; tcgen05 implicit-GEMM convolutions in bf16 mode, FFMA in fp32
parity mode), driven by one `torch.autograd.Function`.  PyTorch only owns memory and the stream.

There is no CPU path and no ATen fallback: CPU tensors, a missing shared library or an unsupported shape raise.
`DoubleConv` and `SELayer` are also usable on their own, like the reference's (UCA:61-72, 96-97): their `forward`
runs the same kernels behind a small autograd Function each.  The whole autograd surface of the reference is kept:
gradients w.r.t. the parameters and, when `images.requires_grad`, w.r.t. the input; backward also works under
`model.eval()` (running-statistics BatchNorm).

Extra, beyond the reference API:
  * `model.precision` / `model.set_precision('bf16' | 'fp32')` — storage type of activations and contraction
    operands ('fp32' is the bit-exact-mask parity mode);
  * `model.loss(images, masks, ignore_index=255)` — fused forward + softmax cross-entropy (UCA:343-344); the
    logits of that call stay available as `model.last_logits`;
  * `model.predict_mask(images)` — argmax class map of UCA:220.
"""
from __future__ import annotations

import ctypes
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib

_WIDTHS = (64, 128, 256, 512, 1024)
_WS_FLOATS = 48 * 1024 * 1024          # split-K scratch for the weight gradients (192 MB)


class SELayer(nn.Module):
    """Squeeze-and-excitation channel attention; parameters as in the reference (UCA:45-59)."""

    def __init__(self, channel: int, reduction: int = 16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x * sigmoid(W2 relu(W1 mean_hw(x))) for an NCHW tensor of any sign (UCA:61-72).  Inside UNet the layer runs
        fused with its neighbours instead (unetca_se_squeeze / unetca_se_fc3 / unetca_se_scale_pool)."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("SELayer expects a 4-D (B, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("unetca_b200.SELayer runs only on CUDA tensors (sm_100a); there is no CPU fallback")
        w1, w2 = self.fc[0].weight, self.fc[2].weight
        if x.shape[1] != w1.shape[1]:
            raise RuntimeError(f"expected input with {w1.shape[1]} channels, got {x.shape[1]}")
        with torch.cuda.device(x.device):
            out = _SELayerFn.apply(x, w1, w2)
        return out if x.dtype == torch.float32 else out.to(x.dtype)


class DoubleConv(nn.Module):
    """(conv3x3 -> BN -> ReLU) x 2 [-> SE]; parameters as in the reference (UCA:75-94)."""

    def __init__(self, in_channels: int, out_channels: int, use_se: bool = False):
        super().__init__()
        layers = [
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        ]
        if use_se:
            layers.append(SELayer(out_channels))
        self.double_conv = nn.Sequential(*layers)

        object.__setattr__(self, "precision", "bf16")
        object.__setattr__(self, "_eng", None)
        object.__setattr__(self, "_blk", None)

    def set_precision(self, precision: str) -> "DoubleConv":
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        object.__setattr__(self, "precision", precision)
        return self

    def _dt(self):
        return (_lib.BF16, torch.bfloat16) if self.precision == "bf16" else (_lib.F32, torch.float32)

    def _engine(self):
        if self._eng is None:
            object.__setattr__(self, "_eng", _Engine(self))
            conv1 = self.double_conv[0]
            object.__setattr__(self, "_blk", _Block("double_conv", self, conv1.in_channels, conv1.out_channels, 0,
                                                    first=conv1.in_channels <= 5))
        return self._eng

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """The block on its own (UCA:96-97): NCHW in, NCHW out, train()/eval() BatchNorm semantics, differentiable
        w.r.t. the parameters and the input.  In bf16 mode the input channels must be <= 5 or a multiple of 64 (the
        shapes UNet uses); 'fp32' mode (FFMA) takes any channel count that is <= 5 or a multiple of 4."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("DoubleConv expects a 4-D (B, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("unetca_b200.DoubleConv runs only on CUDA tensors (sm_100a); there is no CPU fallback")
        conv1 = self.double_conv[0]
        if x.shape[1] != conv1.in_channels:
            raise RuntimeError(f"expected input with {conv1.in_channels} channels, got {x.shape[1]}")
        self._engine()
        keep = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        with torch.cuda.device(x.device):
            out = _DoubleConvFn.apply(self, x, keep, *[p for _, p in self.named_parameters()])
        return out if x.dtype == torch.float32 else out.to(x.dtype)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Block:
    """Static description of one DoubleConv: where its parameters live and its shapes."""

    def __init__(self, name, mod: DoubleConv, cin, cout, level, first=False):
        self.name = name
        seq = mod.double_conv
        self.conv1, self.bn1, self.conv2, self.bn2 = seq[0], seq[1], seq[3], seq[4]
        self.se = seq[6] if len(seq) > 6 else None
        self.cin, self.cout, self.level, self.first = cin, cout, level, first
        self.prefix = name


class _Engine:
    """Per-model execution state: packed operand copies of the weights and scratch buffers."""

    def __init__(self, model: "UNet"):
        self.model = model
        self._packed = {}
        self.epoch = 0            # bumped by every forward that may see stepped parameters (see _cached)
        self.last_train = False
        self._scratch = {}

    # ---- scratch -------------------------------------------------------------------------------------
    def scratch(self, key, numel, dtype, device):
        t = self._scratch.get(key)
        if t is None or t.numel() < numel or t.device != device or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=device)
            self._scratch[key] = t
        return t

    def parts(self, B, device):
        rows = _lib.load().unetca_max_parts(B)
        return self.scratch("parts", rows * 4 * 1024, torch.float32, device)

    def ws(self, device):
        return self.scratch("ws", _WS_FLOATS, torch.float32, device)

    def identity_affine(self, C, device):
        """(ones, zeros) of length C: the BatchNorm arguments of the SE kernels when they run on an activation"""
        key = ("ident", C)
        t = self._scratch.get(key)
        if t is None or t[0].device != device:
            t = (torch.ones(C, dtype=torch.float32, device=device), torch.zeros(C, dtype=torch.float32, device=device))
            self._scratch[key] = t
        return t

    # ---- packed weights ------------------------------------------------------------------------------
    def _cached(self, key, param, tdt, build):
        """Packed operand copy of `param`.  The copies are rebuilt by every train-mode forward and by the first eval
        forward after one (`self.epoch`; the matching backward reuses them): fused optimizers such as
        torch.optim.Adam(fused=True) update parameters WITHOUT bumping `Tensor._version`, so the version alone cannot
        be trusted to see a step.  Between consecutive eval-mode forwards (inference: weights only change through
        load_state_dict / copy_, which do bump it) the version and data pointer key the cache."""
        ent = self._packed.get(key)
        ver = (param._version, param.data_ptr(), tdt, self.epoch)
        if ent is None or ent[0] != ver:
            # same storage type as before: rebuild INTO the existing buffers, so their addresses never change (a CUDA
            # graph of the train step and an optimizer that writes the packed filters keep pointing at live memory)
            old = ent[1] if ent is not None and ent[0][2] == tdt and ent[0][1] == param.data_ptr() else None
            ent = (ver, build(old))
            self._packed[key] = ent
        return ent[1]

    def invalidate(self):
        """Forget that the packed operand copies are current (the next forward rebuilds them in place).  For writes
        that bypass the version counter: `p.data.copy_()`, raw-pointer optimizers, dist.broadcast(t.data)."""
        self.epoch += 1

    def _derive(self, w, dt, tdt, first, wf, wd, ldk, wfp=None, wdp=None):
        """The layouts derived from the base packed filters wf / wd (allocated on first use, then rewritten in place):
        narrow outputs (a multiple of 64 but not of 128 channels) run through the row-pair layout of the tcgen05 path,
        which wants the filter re-expressed over 4x3 virtual taps."""
        O, C = w.shape[0], w.shape[1]
        st = _stream()
        if dt == _lib.BF16 and first and C <= 5:
            # first conv: pixel-pair layout (unetca_im2col_pairs), filter over the 4x3 patch a row pair shares
            if wfp is None:
                wfp = torch.empty(2 * O, 64, dtype=tdt, device=w.device)              # (a plain tensor: first conv only)
            _lib.call("unetca_pack_first_pairs", dt, _ptr(w), _ptr(wfp), O, C, st)
        if dt == _lib.BF16 and not first:
            # 64 output channels: the row-pair layout ("pair"), except from 128 input channels where the kw-stacked
            # layout ("kw", filter resident in shared memory) wins: its epilogue reads three accumulator columns per
            # output from TMEM (64 B/cycle/SM) and only hides behind a mainloop of >= 2 channel chunks
            if O % 128:
                if O == 64 and C == 64:
                    wfp = ("rp64", None)            # resident-filter row-pair kernel: reads the ordinary packed filter
                elif O == 64 and C == 128:
                    wfp = wfp or ("kw", torch.empty(9 * C, 64, dtype=tdt, device=w.device))
                    _lib.call("unetca_pack_conv3x3_kw", dt, _ptr(wf), ldk, _ptr(wfp[1]), C, st)
                else:
                    wfp = wfp or ("pair", torch.empty(2 * O, 12 * C, dtype=tdt, device=w.device))
                    _lib.call("unetca_pack_conv3x3_pair", dt, _ptr(wf), ldk, _ptr(wfp[1]), O, C, st)
            if C % 128:
                if C == 64 and O == 64:
                    wdp = ("rp64", None)
                elif C == 64 and O == 128:
                    wdp = wdp or ("kw", torch.empty(9 * O, 64, dtype=tdt, device=w.device))
                    _lib.call("unetca_pack_conv3x3_kw", dt, _ptr(wd), 9 * O, _ptr(wdp[1]), O, st)
                else:
                    wdp = wdp or ("pair", torch.empty(2 * C, 12 * O, dtype=tdt, device=w.device))
                    _lib.call("unetca_pack_conv3x3_pair", dt, _ptr(wd), 9 * O, _ptr(wdp[1]), C, O, st)
        return wfp, wdp

    def conv_w(self, conv: nn.Conv2d, dt, tdt, first):
        w = conv.weight
        O, C = w.shape[0], w.shape[1]

        def build(old):
            ldk = ((9 * C + 63) // 64) * 64 if first else 9 * C
            if old is not None:
                wf, wd, _, wfp, wdp = old
            else:
                wf = torch.empty(O, ldk, dtype=tdt, device=w.device)
                wd = None if first else torch.empty(C, 9 * O, dtype=tdt, device=w.device)
                wfp = wdp = None
            _lib.call("unetca_pack_conv3x3_weight", dt, _ptr(w), _ptr(wf), ldk, _ptr(wd), O, C, _stream())
            wfp, wdp = self._derive(w, dt, tdt, first, wf, wd, ldk, wfp, wdp)
            return wf, wd, ldk, wfp, wdp
        return self._cached(("c", id(conv)), w, tdt, build)

    def conv_w_first_dgrad(self, conv: nn.Conv2d, dt, tdt):
        """dgrad operand [Cin][9*O] of the first conv — only needed for the gradient w.r.t. the input image."""
        w = conv.weight
        O, C = w.shape[0], w.shape[1]

        def build(old):
            ldk = ((9 * C + 63) // 64) * 64
            wf, wd = old if old is not None else (torch.empty(O, ldk, dtype=tdt, device=w.device),
                                                  torch.empty(C, 9 * O, dtype=tdt, device=w.device))
            _lib.call("unetca_pack_conv3x3_weight", dt, _ptr(w), _ptr(wf), ldk, _ptr(wd), O, C, _stream())
            return wf, wd
        return self._cached(("cd", id(conv)), w, tdt, build)[1]

    def conv_w_adopt(self, conv: nn.Conv2d, dt, tdt):
        """For an optimizer that has just written the stepped weights into this layer's base packed filters
        (optim.Adam -> unetca_adam_step_conv3x3): refresh the derived layouts in place and mark the entry as the one the
        next forward (which bumps `epoch`) will accept, so that forward repacks nothing."""
        w = conv.weight
        key = ("c", id(conv))
        _, (wf, wd, ldk, wfp, wdp) = self._packed[key]
        wfp, wdp = self._derive(w, dt, tdt, False, wf, wd, ldk, wfp, wdp)
        self._packed[key] = ((w._version, w.data_ptr(), tdt, self.epoch + 1), (wf, wd, ldk, wfp, wdp))

    def convT_w(self, up: nn.ConvTranspose2d, dt, tdt):
        w = up.weight
        Cin, Cout = w.shape[0], w.shape[1]

        def build(old):
            wf, wd = old if old is not None else (torch.empty(4 * Cout, Cin, dtype=tdt, device=w.device),
                                                  torch.empty(Cin, 4 * Cout, dtype=tdt, device=w.device))
            _lib.call("unetca_pack_convT_weight", dt, _ptr(w), _ptr(wf), _ptr(wd), Cin, Cout, _stream())
            return wf, wd
        return self._cached(("t", id(up)), w, tdt, build)


def _conv3x3(dt, x, ldx, w, ldk, w_pair, y, ldy, B, H, W, C, O, sp, nparts, st):
    """conv3x3 forward (or dgrad with the dgrad-packed filter): row-pair tcgen05 layout when a pair-packed filter exists."""
    tc = _lib.load().unetca_get_conv_impl() == 0
    if isinstance(w_pair, tuple) and w_pair[0] == "rp64" and H % 2 == 0 and tc:
        _lib.call("unetca_conv3x3_fwd_rp64", dt, _ptr(x), ldx, _ptr(w), ldk, _ptr(y), ldy, B, H, W, sp, nparts, st)
    elif isinstance(w_pair, tuple) and w_pair[0] == "kw" and tc:
        _lib.call("unetca_conv3x3_fwd_kw", dt, _ptr(x), ldx, _ptr(w_pair[1]), _ptr(y), ldy, B, H, W, C, sp, nparts, st)
    elif isinstance(w_pair, tuple) and w_pair[0] == "pair" and H % 2 == 0 and tc:
        _lib.call("unetca_conv3x3_fwd_paired", dt, _ptr(x), ldx, _ptr(w_pair[1]), _ptr(y), ldy, B, H, W, C, O, sp, nparts, st)
    else:
        _lib.call("unetca_conv3x3_fwd", dt, _ptr(x), ldx, _ptr(w), ldk, _ptr(y), ldy, B, H, W, C, O, sp, nparts, st)


def _conv3x3_cat(dt, xin, wf, w_pair, y, ldy, B, H, W, C, O, sp, nparts, st, scale=None, shift=None, sq_parts=None):
    """conv3x3 of torch.cat([skip, up], 1) (UCA:140) with the two halves as dense tensors xin = (skip, up): haloed kernel for
    O % 128 == 0, kw-stacked kernel for 128 -> 64; optional eval-mode BatchNorm + ReLU / SE squeeze epilogue."""
    x1, x2 = xin
    w = w_pair[1] if (O % 128 and isinstance(w_pair, tuple) and w_pair[0] == "kw") else wf
    if O % 128 and w is wf:
        raise RuntimeError(f"no two-source conv kernel for {C} -> {O} channels")
    _lib.call("unetca_conv3x3_fwd_cat", dt, _ptr(x1), x1.stride(2), _ptr(x2), x2.stride(2), x1.shape[3], _ptr(w), _ptr(y), ldy,
              B, H, W, C, O, sp, _ptr(scale), _ptr(shift), _ptr(sq_parts), nparts, st)


def sv_pairs(col, B, H, W):
    """True when `col` is the pixel-pair im2col buffer (one row per pair of rows) rather than the per-pixel one."""
    return col is not None and col.shape[0] == B * (H // 2) * W and H % 2 == 0


FUSE_SQUEEZE = True      # inference: take the SE squeeze in the second conv's epilogue (False: separate read-only pass)
FUSE_HEAD = True          # outc (UCA:162) fused with the last block's SE-scale / BN2-backward passes: its input and input-gradient never exist
PLANAR_CAT = os.environ.get("UNETCA_PLANAR_CAT", "0") == "1"   # decoder conv1 input as two dense tensors (two-source conv operand) instead of one [.., 2C] concat buffer; bit-identical, measured neutral -> off
SPLIT_DCAT = True         # decoder conv1 dgrad writes d(skip) and d(up) as two dense tensors (False: one [.., 2C] buffer)
FUSE_BN_BWD_STATS = True  # training: BN1-backward statistics in the epilogue of the dgrad conv that writes dA1 (False: reduce pass)


def _conv3x3_bnrelu(dt, x, ldx, w, w_pair, y, ldy, B, H, W, C, O, scale, shift, st, sq_parts=None, nparts=None):
    """Inference: conv3x3 + folded eval-mode BatchNorm + ReLU in one tcgen05 kernel (UCA:81-86 under model.eval()).
    Returns False when this shape has no fused kernel (the caller then runs conv and bn_relu separately)."""
    if dt != _lib.BF16 or _lib.load().unetca_get_conv_impl() != 0:
        return False
    if isinstance(w_pair, tuple) and w_pair[0] == "rp64" and H % 2 == 0:
        layout, wt = 3, w
    elif isinstance(w_pair, tuple) and w_pair[0] == "kw":
        layout, wt = 2, w_pair[1]
    elif isinstance(w_pair, tuple) and w_pair[0] == "pair" and H % 2 == 0:
        layout, wt = 1, w_pair[1]
    elif O % 128 == 0:
        layout, wt = 0, w
    else:
        return False
    if layout == 2:
        sq_parts = None                     # the kw-stacked kernel has no squeeze sums (it never is a block's second conv)
    _lib.call("unetca_conv3x3_bnrelu_fwd", dt, _ptr(x), ldx, _ptr(wt), layout, _ptr(y), ldy, B, H, W, C, O, _ptr(scale),
              _ptr(shift), _ptr(sq_parts), nparts, st)
    return True


def _check_input(model, x):
    if not isinstance(x, torch.Tensor) or x.dim() != 4:
        raise ValueError("UNet expects a 4-D (B, C, H, W) tensor")
    if not x.is_cuda:
        raise RuntimeError("unetca_b200.UNet runs only on CUDA tensors (sm_100a); there is no CPU fallback")
    if x.shape[1] != model.in_channels:
        raise RuntimeError(f"expected input with {model.in_channels} channels, got {x.shape[1]}")
    H, W = x.shape[2], x.shape[3]
    if H < 16 or W < 16:
        # four MaxPool2d(2) need at least 16 pixels per side; torch raises from max_pool2d in the reference
        raise RuntimeError(f"input {H}x{W} is too small: four 2x2 max-pools need H, W >= 16 "
                           "(Unet-ChannalAttention.py:106-109)")


# =======================================================================================================
# forward
# =======================================================================================================
def _head_ok(dt, O, nc, hw):
    """Shapes the fused output-head kernels take (unetca_se_scale_outc_fwd and friends)."""
    return FUSE_HEAD and dt == _lib.BF16 and O == 64 and 1 <= nc <= 2 and hw % 128 == 0


def _double_conv_fwd(eng, blk, xin, col, B, Hl, Wl, out_view, pooled, pos, train, dt, tdt, keep, head=None):
    """One DoubleConv (+SE, + the following MaxPool when `pooled` is given).  Returns what backward needs.
    head = (outc module, logits): the last block — its SE-scale pass also applies the 1x1 outc conv and writes the logits;
    the block output itself is never stored (out_view is None)."""
    dev = (out_view if out_view is not None else head[1]).device
    st = _stream()
    C, O = blk.cin, blk.cout
    npix = B * Hl * Wl
    parts = eng.parts(B, dev)
    nparts = ctypes.c_int(0)
    sv = SimpleNamespace(blk=blk, xin=xin if keep else None, col=col if keep else None, B=B, H=Hl, W=Wl, train=train)

    def bn_params(bn, conv, tag):
        mean = torch.empty(O, dtype=torch.float32, device=dev)
        invstd = torch.empty_like(mean)
        scale = torch.empty_like(mean)
        shift = torch.empty_like(mean)
        if bn.running_mean is None or bn.running_var is None:
            raise NotImplementedError("BatchNorm2d(track_running_stats=False) is not supported by the CUDA path "
                                      "(the reference uses the default, UCA:82,85)")
        if train:
            # momentum=None is torch's cumulative moving average: factor 1 / (batches seen so far, this one included)
            mom = float(bn.momentum) if bn.momentum is not None else 1.0 / (int(bn.num_batches_tracked) + 1)
            _lib.call("unetca_bn_finalize_train", _ptr(parts), nparts.value, O, npix, _ptr(conv.bias), _ptr(bn.weight),
                      _ptr(bn.bias), _ptr(bn.running_mean), _ptr(bn.running_var), mom, float(bn.eps),
                      _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), st)
            bn.num_batches_tracked += 1
        else:
            _lib.call("unetca_bn_fold_eval", O, _ptr(conv.bias), _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean),
                      _ptr(bn.running_var), float(bn.eps), _ptr(scale), _ptr(shift), st)
            if keep:
                # backward under eval(): BatchNorm is the fixed affine of its running statistics.  The stored conv
                # output carries no bias, so "mean" is running_mean - bias (two O-element host-side formulas)
                torch.rsqrt(bn.running_var.detach().float() + float(bn.eps), out=invstd)
                torch.sub(bn.running_mean.detach().float(), conv.bias.detach().float(), out=mean)
        setattr(sv, "mean" + tag, mean); setattr(sv, "invstd" + tag, invstd)
        setattr(sv, "scale" + tag, scale); setattr(sv, "shift" + tag, shift)
        return scale, shift

    sp = _ptr(parts) if train else None
    if not train and not keep and dt == _lib.BF16:
        done = _double_conv_eval_fused(eng, blk, xin, col, B, Hl, Wl, out_view, pooled, pos, dt, tdt, bn_params, parts, nparts,
                                       head)
        if done:
            return sv
    # ---- conv1 -> BN -> ReLU
    wf1, _, ldk1, wfp1, _ = eng.conv_w(blk.conv1, dt, tdt, blk.first)
    y1 = torch.empty(B, Hl, Wl, O, dtype=tdt, device=dev)
    if blk.first and sv_pairs(col, B, Hl, Wl):
        _lib.call("unetca_first_pairs_fwd", dt, _ptr(col), _ptr(wfp1), _ptr(y1), O, B, Hl, Wl, O, sp, ctypes.byref(nparts), st)
    elif blk.first:
        _lib.call("unetca_gemm_nt", dt, _ptr(col), col.shape[1], _ptr(wf1), ldk1, _ptr(y1), O, npix, O, col.shape[1],
                  sp, ctypes.byref(nparts), st)
    elif isinstance(xin, tuple):
        _conv3x3_cat(dt, xin, wf1, wfp1, y1, O, B, Hl, Wl, C, O, sp, ctypes.byref(nparts), st)
    else:
        _conv3x3(dt, xin, xin.stride(2), wf1, ldk1, wfp1, y1, O, B, Hl, Wl, C, O, sp, ctypes.byref(nparts), st)
    scale1, shift1 = bn_params(blk.bn1, blk.conv1, "1")
    a1 = torch.empty_like(y1)
    _lib.call("unetca_bn_relu", dt, _ptr(y1), O, _ptr(a1), O, B, Hl * Wl, O, _ptr(scale1), _ptr(shift1), None, None, st)
    # ---- conv2 -> BN -> ReLU [-> SE] [-> MaxPool]
    wf2, _, ldk2, wfp2, _ = eng.conv_w(blk.conv2, dt, tdt, False)
    y2 = torch.empty_like(y1)
    _conv3x3(dt, a1, O, wf2, ldk2, wfp2, y2, O, B, Hl, Wl, O, O, sp, ctypes.byref(nparts), st)
    scale2, shift2 = bn_params(blk.bn2, blk.conv2, "2")
    s = None
    if blk.se is not None:
        w1, w2 = blk.se.fc[0].weight, blk.se.fc[2].weight
        Cr = w1.shape[0]
        _lib.call("unetca_se_squeeze", dt, _ptr(y2), O, B, Hl * Wl, O, _ptr(scale2), _ptr(shift2), _ptr(parts),
                  ctypes.byref(nparts), st)
        p = torch.empty(B, O, dtype=torch.float32, device=dev)
        z = torch.empty(B, Cr, dtype=torch.float32, device=dev)
        s = torch.empty(B, O, dtype=torch.float32, device=dev)
        sums34 = torch.empty(B, 2, O, dtype=torch.float32, device=dev) if keep else None
        _lib.call("unetca_se_fc3", _ptr(parts), nparts.value, B, O, Cr, Hl * Wl, _ptr(w1), _ptr(w2), _ptr(scale2),
                  _ptr(shift2), _ptr(sv.mean2) if (train or keep) else None, _ptr(p), _ptr(z), _ptr(s), _ptr(sums34), st)
        sv.p, sv.z, sv.s, sv.sums34 = p, z, s, sums34
    if head is not None:
        outc, logits = head
        _lib.call("unetca_se_scale_outc_fwd", dt, _ptr(y2), O, B, Hl * Wl, O, _ptr(scale2), _ptr(shift2), _ptr(s),
                  _ptr(outc.weight), _ptr(outc.bias), outc.weight.shape[0], _ptr(logits), st)
    elif pooled is not None and (Hl % 2 or Wl % 2):
        # odd extent: MaxPool2d(2) floors (UCA:106-109) -> scale pass, then the standalone pool over the stored values
        _lib.call("unetca_se_scale_pool", dt, _ptr(y2), O, _ptr(out_view), out_view.stride(2), None, 0, None, B, Hl, Wl,
                  O, _ptr(scale2), _ptr(shift2), _ptr(s), st)
        _lib.call("unetca_maxpool2x2", dt, _ptr(out_view), out_view.stride(2), _ptr(pooled), pooled.stride(2), _ptr(pos),
                  None, B, Hl, Wl, O, st)
    else:
        _lib.call("unetca_se_scale_pool", dt, _ptr(y2), O, _ptr(out_view), out_view.stride(2), _ptr(pooled),
                  pooled.stride(2) if pooled is not None else 0, _ptr(pos), B, Hl, Wl, O, _ptr(scale2), _ptr(shift2),
                  _ptr(s), st)
    if keep:
        sv.y1, sv.a1, sv.y2 = y1, a1, y2
    return sv


def _double_conv_eval_fused(eng, blk, xin, col, B, Hl, Wl, out_view, pooled, pos, dt, tdt, bn_params, parts, nparts, head=None):
    """Inference form of one DoubleConv: BatchNorm uses running statistics, so both conv + BN + ReLU pairs run as single
    tcgen05 kernels (the activation is what gets written; the pre-BN tensors never exist), followed by the SE squeeze /
    scale (+ max-pool) passes on the activation with an identity affine.  Returns False if a conv of this block has
    no fused kernel for its shape."""
    dev = (out_view if out_view is not None else head[1]).device
    st = _stream()
    C, O = blk.cin, blk.cout
    wf1, _, ldk1, wfp1, _ = eng.conv_w(blk.conv1, dt, tdt, blk.first)
    wf2, _, ldk2, wfp2, _ = eng.conv_w(blk.conv2, dt, tdt, False)
    # can both convolutions be fused?  (decide before launching anything)
    tc = _lib.load().unetca_get_conv_impl() == 0

    def fusable(w_pair, Cin, first):
        if not tc:
            return False
        if first:
            return sv_pairs(col, B, Hl, Wl)
        if isinstance(w_pair, tuple) and (w_pair[0] == "kw" or Hl % 2 == 0):
            return True
        return O % 128 == 0
    if not (fusable(wfp1, C, blk.first) and fusable(wfp2, O, False)):
        return False
    scale1, shift1 = bn_params(blk.bn1, blk.conv1, "1")
    a1 = torch.empty(B, Hl, Wl, O, dtype=tdt, device=dev)
    if blk.first:
        _lib.call("unetca_first_pairs_bnrelu_fwd", dt, _ptr(col), _ptr(wfp1), _ptr(a1), O, B, Hl, Wl, O, _ptr(scale1),
                  _ptr(shift1), st)
    elif isinstance(xin, tuple):
        _conv3x3_cat(dt, xin, wf1, wfp1, a1, O, B, Hl, Wl, C, O, None, None, st, scale1, shift1)
    else:
        ok = _conv3x3_bnrelu(dt, xin, xin.stride(2), wf1, wfp1, a1, O, B, Hl, Wl, C, O, scale1, shift1, st)
        assert ok
    scale2, shift2 = bn_params(blk.bn2, blk.conv2, "2")
    a2 = torch.empty(B, Hl, Wl, O, dtype=tdt, device=dev)
    one, zero = eng.identity_affine(O, dev)
    s = None
    if blk.se is not None:
        w1, w2 = blk.se.fc[0].weight, blk.se.fc[2].weight
        Cr = w1.shape[0]
        p = torch.empty(B, O, dtype=torch.float32, device=dev)
        z = torch.empty(B, Cr, dtype=torch.float32, device=dev)
        s = torch.empty(B, O, dtype=torch.float32, device=dev)
        kw2 = isinstance(wfp2, tuple) and wfp2[0] == "kw"
        if FUSE_SQUEEZE and not kw2:
            # the second conv's epilogue also leaves the SE squeeze: per-image partial channel sums of the activation
            nsq = B * _lib.load().unetca_num_sms() * O
            sq = eng.scratch("sq_parts", nsq, torch.float32, dev)
            sq[:nsq].zero_()
            ok = _conv3x3_bnrelu(dt, a1, O, wf2, wfp2, a2, O, B, Hl, Wl, O, O, scale2, shift2, st, sq, ctypes.byref(nparts))
            assert ok
            _lib.call("unetca_se_fc", _ptr(sq), nparts.value, B, O, Cr, Hl * Wl, _ptr(w1), _ptr(w2), _ptr(p), _ptr(z),
                      _ptr(s), st)
        else:
            ok = _conv3x3_bnrelu(dt, a1, O, wf2, wfp2, a2, O, B, Hl, Wl, O, O, scale2, shift2, st)
            assert ok
            _lib.call("unetca_se_squeeze", dt, _ptr(a2), O, B, Hl * Wl, O, _ptr(one), _ptr(zero), _ptr(parts),
                      ctypes.byref(nparts), st)
            _lib.call("unetca_se_fc3", _ptr(parts), nparts.value, B, O, Cr, Hl * Wl, _ptr(w1), _ptr(w2), _ptr(one),
                      _ptr(zero), None, _ptr(p), _ptr(z), _ptr(s), None, st)
    else:
        ok = _conv3x3_bnrelu(dt, a1, O, wf2, wfp2, a2, O, B, Hl, Wl, O, O, scale2, shift2, st)
        assert ok
    if head is not None:
        outc, logits = head
        _lib.call("unetca_se_scale_outc_fwd", dt, _ptr(a2), O, B, Hl * Wl, O, _ptr(one), _ptr(zero), _ptr(s), _ptr(outc.weight),
                  _ptr(outc.bias), outc.weight.shape[0], _ptr(logits), st)
    elif pooled is not None and (Hl % 2 or Wl % 2):
        _lib.call("unetca_se_scale_pool", dt, _ptr(a2), O, _ptr(out_view), out_view.stride(2), None, 0, None, B, Hl, Wl,
                  O, _ptr(one), _ptr(zero), _ptr(s), st)
        _lib.call("unetca_maxpool2x2", dt, _ptr(out_view), out_view.stride(2), _ptr(pooled), pooled.stride(2), _ptr(pos),
                  None, B, Hl, Wl, O, st)
    else:
        _lib.call("unetca_se_scale_pool", dt, _ptr(a2), O, _ptr(out_view), out_view.stride(2), _ptr(pooled),
                  pooled.stride(2) if pooled is not None else 0, _ptr(pos), B, Hl, Wl, O, _ptr(one), _ptr(zero), _ptr(s), st)
    return True


def _first_conv_rows(dt, tdt, xf, B, Cin, H, W):
    """im2col rows of the network input for the first convolution (K = 9*Cin): one row per pixel pair on the bf16
    tensor-core path with an even H (unetca_im2col_pairs), one per pixel otherwise."""
    dev, st = xf.device, _stream()
    if dt == _lib.BF16 and Cin <= 5 and H % 2 == 0 and _lib.load().unetca_get_conv_impl() == 0:
        col = torch.empty(B * (H // 2) * W, 64, dtype=tdt, device=dev)           # one row per pixel pair (rows 2i, 2i+1)
        _lib.call("unetca_im2col_pairs", dt, _ptr(xf), _ptr(col), B, Cin, H, W, st)
    else:
        Kpad = ((9 * Cin + 63) // 64) * 64
        col = torch.empty(B * H * W, Kpad, dtype=tdt, device=dev)
        _lib.call("unetca_im2col3x3_nchw", dt, _ptr(xf), _ptr(col), B, Cin, H, W, Kpad, st)
    return col


def _forward(model: "UNet", x: torch.Tensor, keep: bool):
    """UNet.forward (UCA:127-163).  Returns (logits NCHW fp32, saved-state or None)."""
    _check_input(model, x)
    eng = model._engine()
    dt, tdt = model._dt()
    train = model.training
    if train or eng.last_train:
        eng.epoch += 1            # the parameters may have been stepped since the last train-mode forward: repack them
    eng.last_train = train
    dev = x.device
    st = _stream()
    B, Cin, H, W = x.shape
    Hs = [H >> l for l in range(5)]                  # MaxPool2d(2) floors: repeated halving == shift
    Ws = [W >> l for l in range(5)]
    if train and B * (H // 16) * (W // 16) <= 1:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                         f"{[B, 1024, H // 16, W // 16]}")
    xf = x.detach()
    if xf.dtype != torch.float32 or not xf.is_contiguous():
        xf = xf.float().contiguous()
    col = _first_conv_rows(dt, tdt, xf, B, Cin, H, W)

    sv = SimpleNamespace(enc=[], dec=[], up_in=[], cat=[], pos=[], B=B, H=H, W=W, Hs=Hs, Ws=Ws)
    # torch.cat([skip, up], 1) (UCA:140..158) never happens: the producers write the operand of the decoder's first conv in
    # place.  bf16 tensor-core path: two DENSE tensors per level (the conv kernels take a two-source operand), so nobody
    # writes or reads half-pixels at a doubled stride; otherwise the two channel halves of one [B,H,W,2C] buffer.
    planar = PLANAR_CAT and dt == _lib.BF16 and _lib.load().unetca_get_conv_impl() == 0
    if planar:
        cat = [(torch.empty(B, Hs[l], Ws[l], _WIDTHS[l], dtype=tdt, device=dev),
                torch.empty(B, Hs[l], Ws[l], _WIDTHS[l], dtype=tdt, device=dev)) for l in range(4)]
    else:
        cat = [torch.empty(B, Hs[l], Ws[l], 2 * _WIDTHS[l], dtype=tdt, device=dev) for l in range(4)]
    # ---- encoder
    xin = None
    for l, blk in enumerate(model._enc_blocks):
        Hl, Wl, Cl = Hs[l], Ws[l], _WIDTHS[l]
        if l < 4:
            out_view = cat[l][0] if planar else cat[l][..., :Cl]
            pooled = torch.empty(B, Hl // 2, Wl // 2, Cl, dtype=tdt, device=dev)
            pos = torch.empty(B, Hl // 2, Wl // 2, Cl, dtype=torch.uint8, device=dev)
        else:
            out_view = torch.empty(B, Hl, Wl, Cl, dtype=tdt, device=dev)
            pooled = pos = None
        s = _double_conv_fwd(eng, blk, xin, col if l == 0 else None, B, Hl, Wl, out_view, pooled, pos, train, dt, tdt,
                             keep)
        sv.enc.append(s)
        sv.pos.append(pos)
        xin = pooled if l < 4 else out_view
    # ---- decoder
    h = xin
    for i, (up, blk) in enumerate(zip(model._ups, model._dec_blocks)):
        l = 3 - i
        Hl, Wl, Cl = Hs[l], Ws[l], _WIDTHS[l]
        hi, wi = Hs[l + 1], Ws[l + 1]
        wf, _ = eng.convT_w(up, dt, tdt)
        up_view = cat[l][1] if planar else cat[l][..., Cl:]
        if (2 * hi, 2 * wi) == (Hl, Wl):
            _lib.call("unetca_convT2x2_fwd", dt, _ptr(h), h.stride(2), _ptr(wf), _ptr(up.bias), _ptr(up_view),
                      up_view.stride(2), B, hi, wi, 2 * Cl, Cl, st)
        else:
            # resize guard (UCA:138-157): the transposed conv gives 2*floor(H/2) rows / columns, the skip has H
            u = torch.empty(B, 2 * hi, 2 * wi, Cl, dtype=tdt, device=dev)
            _lib.call("unetca_convT2x2_fwd", dt, _ptr(h), h.stride(2), _ptr(wf), _ptr(up.bias), _ptr(u), Cl, B, hi, wi,
                      2 * Cl, Cl, st)
            _lib.call("unetca_resize_bilinear_fwd", dt, _ptr(u), Cl, 2 * hi, 2 * wi, _ptr(up_view), up_view.stride(2), Hl, Wl,
                      B, Cl, st)
        sv.up_in.append(h if keep else None)
        head = None
        if l == 0 and _head_ok(dt, Cl, model.num_classes, Hl * Wl):
            # the last block: outc rides in its SE-scale pass, the block output is never written
            logits = torch.empty(B, model.num_classes, H, W, dtype=torch.float32, device=dev)
            head, out = (model.outc, logits), None
        else:
            out = torch.empty(B, Hl, Wl, Cl, dtype=tdt, device=dev)
        s = _double_conv_fwd(eng, blk, cat[l], None, B, Hl, Wl, out, None, None, train, dt, tdt, keep, head)
        sv.dec.append(s)
        h = out
    # ---- outc
    nc = model.num_classes
    if h is not None:
        logits = torch.empty(B, nc, H, W, dtype=torch.float32, device=dev)
        _lib.call("unetca_outc_fwd", dt, _ptr(h), h.stride(2), 64, _ptr(model.outc.weight), _ptr(model.outc.bias), nc,
                  _ptr(logits), B, H * W, st)
    if not keep:
        return logits, None
    sv.cat = cat
    sv.dec_out = h                       # None when the head was fused
    return logits, sv


# =======================================================================================================
# backward
# =======================================================================================================
def _block_grad_order(pre, use_se):
    """Order in which _double_conv_bwd completes the parameter gradients of one block (DP buckets follow it)."""
    out = [pre + ".6.fc.0.weight", pre + ".6.fc.2.weight"] if use_se else []
    return out + [pre + ".4.weight", pre + ".4.bias", pre + ".3.bias", pre + ".3.weight",
                  pre + ".1.weight", pre + ".1.bias", pre + ".0.bias", pre + ".0.weight"]


def _double_conv_bwd(eng, sv, dout, G, dt, tdt, need_dx, lazy=None, dx_stats=None, head=None, split=0):
    """Gradient of one DoubleConv block; dout is d(loss)/d(block output) (NHWC view).  Returns d/d(block input).
    G: gradient sink with alloc(name, like) -> tensor to write into and put(name) once it is complete.
    lazy = (skip_grad, dpooled, pos) instead of dout: the output gradient of an encoder block, skip gradient plus the
    max-pool routing of the pooled gradient, is rebuilt inside the two kernels that consume it (SE blocks, even H, W).
    dx_stats = (parts, nparts): have the dgrad convolution that writes the returned gradient leave its per-CTA channel
    sums there ([n][2][C], the BatchNorm-statistics epilogue) — the decoder takes the ConvTranspose bias gradient from them.
    split = C_skip > 0 (decoder blocks on the bf16 tensor-core path): return (d skip, d upsampled) as two dense tensors instead of
    one [.., 2C] buffer — every consumer of either half would otherwise read 128-byte rows at a 256-byte stride.
    head = (g, gscale, outc, dw, db) instead of dout (last block, fused output head): dout = W_outc^T g is rebuilt per pixel inside
    the reduction and the apply pass from the dlogits g; the outc weight / bias gradients come out of the reduction."""
    blk = sv.blk
    B, Hl, Wl = sv.B, sv.H, sv.W
    C, O = blk.cin, blk.cout
    dev = dout.device if dout is not None else (lazy[0] if lazy is not None else head[0]).device
    st = _stream()
    npix = B * Hl * Wl
    parts = eng.parts(B, dev)
    ws = eng.ws(dev)
    nparts = ctypes.c_int(0)
    pre = blk.prefix
    s = dp = sums = None
    if blk.se is not None:
        # one pass over (dO, Y2) yields the four per-image sums that both the SE excitation gradient and the BN2
        # backward statistics are linear in (csrc/elementwise.cu: se_bn_bwd_reduce_kernel)
        w1, w2 = blk.se.fc[0].weight, blk.se.fc[2].weight
        Cr = w1.shape[0]
        if head is not None:
            hg, hgs, outc, hdw, hdb = head
            _lib.call("unetca_outc_bn_bwd_reduce", dt, _ptr(hg), _ptr(hgs), _ptr(outc.weight), outc.weight.shape[0], _ptr(sv.y2), O, B,
                      Hl * Wl, O, _ptr(sv.scale2), _ptr(sv.shift2), _ptr(sv.mean2), _ptr(sv.s), _ptr(parts), ctypes.byref(nparts),
                      _ptr(ws), ws.numel(), _ptr(hdw), _ptr(hdb), st)
            G.put("outc.weight")
            G.put("outc.bias")
        elif lazy is not None:
            sg, dpl, pos = lazy
            _lib.call("unetca_se_bn_bwd_reduce_pool", dt, _ptr(sg), sg.stride(2), _ptr(dpl), dpl.stride(2), _ptr(pos),
                      _ptr(sv.y2), O, B, Hl, Wl, O, _ptr(sv.scale2), _ptr(sv.shift2), _ptr(sv.mean2), _ptr(parts),
                      ctypes.byref(nparts), st)
        else:
            _lib.call("unetca_se_bn_bwd_reduce", dt, _ptr(dout), dout.stride(2), _ptr(sv.y2), O, B, Hl * Wl, O,
                      _ptr(sv.scale2), _ptr(sv.shift2), _ptr(sv.mean2), _ptr(parts), ctypes.byref(nparts), st)
        dpre2 = torch.empty(B, O, dtype=torch.float32, device=dev)
        dz = torch.empty(B, Cr, dtype=torch.float32, device=dev)
        dp = torch.empty(B, O, dtype=torch.float32, device=dev)
        sums = torch.empty(B, 4, O, dtype=torch.float32, device=dev)
        dw1 = G.alloc(pre + ".6.fc.0.weight", w1)
        dw2 = G.alloc(pre + ".6.fc.2.weight", w2)
        _lib.call("unetca_se_fc_bwd_fused", _ptr(parts), nparts.value, B, O, Cr, _ptr(w1), _ptr(w2), _ptr(sv.p), _ptr(sv.z),
                  _ptr(sv.s), _ptr(sv.scale2), _ptr(sv.shift2), _ptr(sv.mean2), _ptr(sv.sums34), _ptr(sums), _ptr(dpre2), _ptr(dz),
                  _ptr(dp),
                  _ptr(dw1), _ptr(dw2), st)
        G.put(pre + ".6.fc.0.weight")
        G.put(pre + ".6.fc.2.weight")
        s = sv.s

    def bn_relu_bwd(d_in, ld_in, y, tag, s_, dp_, bn_idx, sums_=None, have_parts=False):
        scale, shift = getattr(sv, "scale" + tag), getattr(sv, "shift" + tag)
        mean, invstd = getattr(sv, "mean" + tag), getattr(sv, "invstd" + tag)
        bn = blk.bn1 if tag == "1" else blk.bn2
        dgamma = G.alloc(f"{pre}.{bn_idx}.weight", bn.weight)
        dbeta = G.alloc(f"{pre}.{bn_idx}.bias", bn.bias)
        coef = torch.empty(3, O, dtype=torch.float32, device=dev)
        if sums_ is not None:
            _lib.call("unetca_bn_bwd_finalize_se", _ptr(sums_), B, O, npix, Hl * Wl, _ptr(bn.weight), _ptr(invstd), _ptr(s_),
                      _ptr(dp_), _ptr(dgamma), _ptr(dbeta), _ptr(coef), st)
        else:
            if head is not None and tag == "2":          # plain U-Net: fused output head without SE
                hg, hgs, outc, hdw, hdb = head
                _lib.call("unetca_outc_bn_bwd_reduce", dt, _ptr(hg), _ptr(hgs), _ptr(outc.weight), outc.weight.shape[0], _ptr(y), O, B,
                          Hl * Wl, O, _ptr(scale), _ptr(shift), _ptr(mean), None, _ptr(parts), ctypes.byref(nparts), _ptr(ws),
                          ws.numel(), _ptr(hdw), _ptr(hdb), st)
                G.put("outc.weight")
                G.put("outc.bias")
                nparts.value *= B
            elif not have_parts:             # (else: the dgrad conv that wrote d_in left the statistics in `parts` already)
                _lib.call("unetca_bn_bwd_reduce", dt, _ptr(d_in), ld_in, _ptr(y), O, B, Hl * Wl, O, _ptr(scale), _ptr(shift),
                          _ptr(mean), _ptr(invstd), _ptr(s_), _ptr(dp_), _ptr(parts), ctypes.byref(nparts), st)
            _lib.call("unetca_bn_bwd_finalize", _ptr(parts), nparts.value, O, npix, _ptr(bn.weight), _ptr(invstd),
                      _ptr(dgamma), _ptr(dbeta), _ptr(coef), st)
        if not sv.train:
            # eval(): no batch statistics to differentiate through -> dY = gamma*invstd_running * dz (c1 = c2 = 0), and the
            # conv bias in front (folded into the shift) gets  sum dY = gamma*invstd * sum dz
            coef[1:].zero_()
            conv = blk.conv1 if tag == "1" else blk.conv2
            torch.mul(coef[0], dbeta, out=G.alloc(f"{pre}.{bn_idx - 1}.bias", conv.bias))
        dy = torch.empty(B, Hl, Wl, O, dtype=tdt, device=dev)
        if head is not None and tag == "2":
            hg, hgs, outc, _, _ = head
            _lib.call("unetca_outc_bn_bwd_apply", dt, _ptr(hg), _ptr(hgs), _ptr(outc.weight), outc.weight.shape[0], _ptr(y), O, _ptr(dy),
                      O, B, Hl * Wl, O, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(s_), _ptr(dp_), _ptr(coef), st)
        elif sums_ is not None and lazy is not None:
            sg, dpl, pos = lazy
            _lib.call("unetca_bn_bwd_apply_pool", dt, _ptr(sg), sg.stride(2), _ptr(dpl), dpl.stride(2), _ptr(pos), _ptr(y), O,
                      _ptr(dy), O, B, Hl, Wl, O, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(s_), _ptr(dp_),
                      _ptr(coef), st)
        else:
            _lib.call("unetca_bn_bwd_apply", dt, _ptr(d_in), ld_in, _ptr(y), O, _ptr(dy), O, B, Hl * Wl, O, _ptr(scale),
                      _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(s_), _ptr(dp_), _ptr(coef), st)
        G.put(f"{pre}.{bn_idx}.weight")
        G.put(f"{pre}.{bn_idx}.bias")
        return dy

    # ---- [SE ->] ReLU -> BN2 backward
    dy2 = bn_relu_bwd(dout, dout.stride(2) if dout is not None else 0, sv.y2, "2", s, dp, 4, sums)
    # a conv bias in front of a train-mode BatchNorm has an analytically zero gradient (BN removes the mean); under
    # eval() bn_relu_bwd has just written the real one for .3.bias, and writes .0.bias below
    if sv.train:
        G.alloc(pre + ".3.bias", blk.conv2.bias).zero_()
    G.put(pre + ".3.bias")
    # ---- conv2 wgrad + dgrad
    dw = G.alloc(pre + ".3.weight", blk.conv2.weight)
    _lib.call("unetca_conv3x3_wgrad", dt, _ptr(dy2), O, _ptr(sv.a1), O, _ptr(ws), ws.numel(), B, Hl, Wl, O, O, _ptr(dw), st)
    G.put(pre + ".3.weight")
    _, wd2, _, _, wdp2 = eng.conv_w(blk.conv2, dt, tdt, False)
    da1 = torch.empty(B, Hl, Wl, O, dtype=tdt, device=dev)
    # bf16 tensor-core path: the dgrad's epilogue also takes the statistics of the ReLU + BN1 backward that follows (it holds
    # every dA1 value it writes; the saved Y1 of the same pixels is read from L2 / HBM under the MMAs) — no reduce pass
    fused_stats = (FUSE_BN_BWD_STATS and dt == _lib.BF16 and _lib.load().unetca_get_conv_impl() == 0 and
                   (O % 128 == 0 or (O == 64 and Hl % 2 == 0)))
    if fused_stats:
        _lib.call("unetca_conv3x3_dgrad_bnstats", dt, _ptr(dy2), O, _ptr(wd2), 9 * O, _ptr(da1), O, B, Hl, Wl, O, O, _ptr(sv.y1), O,
                  _ptr(sv.scale1), _ptr(sv.shift1), _ptr(sv.mean1), _ptr(parts), ctypes.byref(nparts), st)
    else:
        _conv3x3(dt, dy2, O, wd2, 9 * O, wdp2, da1, O, B, Hl, Wl, O, O, None, None, st)
    del dy2
    # ---- ReLU -> BN1 backward
    dy1 = bn_relu_bwd(da1, O, sv.y1, "1", None, None, 1, have_parts=fused_stats)
    del da1
    if sv.train:
        G.alloc(pre + ".0.bias", blk.conv1.bias).zero_()
    G.put(pre + ".0.bias")
    # ---- conv1 wgrad (+ dgrad)
    dw = G.alloc(pre + ".0.weight", blk.conv1.weight)
    if blk.first and sv_pairs(sv.col, B, Hl, Wl):
        _lib.call("unetca_first_pairs_wgrad", dt, _ptr(dy1), O, _ptr(sv.col), _ptr(ws), ws.numel(), B, Hl, Wl, C, O,
                  _ptr(dw), st)
    elif blk.first:
        _lib.call("unetca_im2col_wgrad", dt, _ptr(dy1), O, _ptr(sv.col), sv.col.shape[1], _ptr(ws), ws.numel(), npix, C, O,
                  _ptr(dw), st)
    elif isinstance(sv.xin, tuple):
        x1, x2 = sv.xin
        _lib.call("unetca_conv3x3_wgrad_cat", dt, _ptr(dy1), O, _ptr(x1), x1.stride(2), _ptr(x2), x2.stride(2), x1.shape[3],
                  _ptr(ws), ws.numel(), B, Hl, Wl, C, O, _ptr(dw), st)
    else:
        _lib.call("unetca_conv3x3_wgrad", dt, _ptr(dy1), O, _ptr(sv.xin), sv.xin.stride(2), _ptr(ws), ws.numel(), B, Hl,
                  Wl, C, O, _ptr(dw), st)
    G.put(pre + ".0.weight")
    if not need_dx:
        return None
    if blk.first:
        # gradient w.r.t. the network input (only when the caller asked for it: images.requires_grad): the dgrad of the
        # K = 9*Cin first conv is 0.9 GFLOP/img with N = Cin <= 5 output channels — FFMA kernel, NHWC rows of Cin values
        wd0 = eng.conv_w_first_dgrad(blk.conv1, dt, tdt)
        dx = torch.empty(B, Hl, Wl, C, dtype=tdt, device=dev)
        _lib.call("unetca_simt_conv3x3_fwd", dt, _ptr(dy1), O, _ptr(wd0), 9 * O, _ptr(dx), C, B, Hl, Wl, O, C, st)
        return dx
    _, wd1, _, _, wdp1 = eng.conv_w(blk.conv1, dt, tdt, False)
    sp, npp = (dx_stats[0], ctypes.byref(dx_stats[1])) if dx_stats is not None else (None, None)
    if split and C % 128 == 0 and dt == _lib.BF16 and _lib.load().unetca_get_conv_impl() == 0:
        dskip = torch.empty(B, Hl, Wl, split, dtype=tdt, device=dev)
        dup = torch.empty(B, Hl, Wl, C - split, dtype=tdt, device=dev)
        _lib.call("unetca_conv3x3_fwd_split", dt, _ptr(dy1), O, _ptr(wd1), 9 * O, _ptr(dskip), split, _ptr(dup), C - split, split,
                  B, Hl, Wl, O, C, _ptr(sp) if sp is not None else None, npp, st)
        return dskip, dup
    dx = torch.empty(B, Hl, Wl, C, dtype=tdt, device=dev)
    _conv3x3(dt, dy1, O, wd1, 9 * O, wdp1, dx, C, B, Hl, Wl, O, C, _ptr(sp) if sp is not None else None, npp, st)
    return dx


class _GradSink:
    """Default gradient sink: fresh tensors, nothing to do when one completes.  parallel.GradBuckets replaces it with
    views into flat all-reduce buckets and launches the collective as each bucket fills."""

    def __init__(self):
        self.grads = {}

    def alloc(self, name, like):
        t = torch.empty_like(like)
        self.grads[name] = t
        return t

    def put(self, name):
        pass

    def finish(self):
        return self.grads


def grad_order(model: "UNet"):
    """Parameter names in the order _backward completes their gradients."""
    names = ["outc.weight", "outc.bias"]
    for l in range(4):
        i = 3 - l
        names += _block_grad_order(model._dec_blocks[i].prefix, model.use_se)
        names += [model._up_names[i] + ".bias", model._up_names[i] + ".weight"]
    for l in range(4, -1, -1):
        names += _block_grad_order(model._enc_blocks[l].prefix, model.use_se)
    return names


def _backward(model: "UNet", sv, g: torch.Tensor, gscale: torch.Tensor, need_dx: bool = False):
    """loss.backward() (UCA:345) for every parameter.  g: (B,nc,H,W) fp32 dlogits up to the factor *gscale.
    Returns ({name: gradient} or None when the sink owns .grad, d(loss)/d(input) NCHW fp32 or None)."""
    G = model._grad_sink_factory() if model._grad_sink_factory is not None else _GradSink()
    eng = model._engine()
    dt, tdt = model._dt()
    dev = g.device
    st = _stream()
    B, H, W = sv.B, sv.H, sv.W
    nc = model.num_classes
    parts = eng.parts(B, dev)
    ws = eng.ws(dev)
    # ---- outc
    h = sv.dec_out
    dw = G.alloc("outc.weight", model.outc.weight)
    db = G.alloc("outc.bias", model.outc.bias)
    head = None
    if h is None:
        # fused output head: the last block's backward passes rebuild d(block output) from the dlogits and leave dw / db
        head, dcur = (g, gscale, model.outc, dw, db), None
    else:
        dcur = torch.empty(B, H, W, 64, dtype=tdt, device=dev)
        _lib.call("unetca_outc_bwd", dt, _ptr(g), _ptr(gscale), _ptr(h), h.stride(2), _ptr(dcur), 64, 64,
                  _ptr(model.outc.weight), nc, B, H * W, _ptr(parts), _ptr(dw), _ptr(db), st)
        G.put("outc.weight")
        G.put("outc.bias")
    # ---- decoder, shallow to deep
    skip_grads = [None] * 4
    for l in range(4):
        i = 3 - l
        up, name = model._ups[i], model._up_names[i]
        Hl, Wl, Cl = sv.Hs[l], sv.Ws[l], _WIDTHS[l]
        hi, wi = sv.Hs[l + 1], sv.Ws[l + 1]
        # bf16 tensor-core path, no resize guard: the dgrad conv that writes dcat also leaves its per-CTA channel sums,
        # whose upper half is the ConvTranspose bias gradient (one less pass over dcat)
        fused_db = (dt == _lib.BF16 and _lib.load().unetca_get_conv_impl() == 0 and (2 * hi, 2 * wi) == (Hl, Wl))
        nst = ctypes.c_int(0)
        dcat = _double_conv_bwd(eng, sv.dec[i], dcur, G, dt, tdt, True,
                                dx_stats=(parts, nst) if fused_db else None,        # (B,Hl,Wl,2Cl), or its two halves
                                head=head if l == 0 else None, split=Cl if SPLIT_DCAT else 0)
        if isinstance(dcat, tuple):
            skip_grads[l], du = dcat
            ldu = Cl
        else:
            du, ldu = dcat[..., Cl:], 2 * Cl
            skip_grads[l] = dcat[..., :Cl]
        if (2 * hi, 2 * wi) != (Hl, Wl):                                         # adjoint of the resize guard
            du_s = torch.empty(B, 2 * hi, 2 * wi, Cl, dtype=tdt, device=dev)
            _lib.call("unetca_resize_bilinear_bwd", dt, _ptr(du), ldu, Hl, Wl, _ptr(du_s), Cl, 2 * hi, 2 * wi, B, Cl, st)
            du, ldu = du_s, Cl
        dbias = G.alloc(name + ".bias", up.bias)
        if fused_db and nst.value > 0:
            _lib.call("unetca_sum_rows", parts.data_ptr() + 4 * Cl, nst.value, 4 * Cl, Cl, _ptr(dbias), st)
        else:
            _lib.call("unetca_chan_sum", dt, _ptr(du), ldu, Cl, B * 4 * hi * wi, _ptr(parts), _ptr(dbias), st)
        G.put(name + ".bias")
        up_in = sv.up_in[i]
        dwu = G.alloc(name + ".weight", up.weight)
        _lib.call("unetca_convT2x2_wgrad", dt, _ptr(up_in), up_in.stride(2), _ptr(du), ldu, _ptr(ws), ws.numel(), B,
                  hi, wi, 2 * Cl, Cl, _ptr(dwu), st)
        G.put(name + ".weight")
        _, wd = eng.convT_w(up, dt, tdt)
        dcur = torch.empty(B, hi, wi, 2 * Cl, dtype=tdt, device=dev)
        _lib.call("unetca_convT2x2_dgrad", dt, _ptr(du), ldu, _ptr(wd), _ptr(dcur), 2 * Cl, B, hi, wi, 2 * Cl, Cl, st)
    # ---- encoder, deep to shallow
    lazy = None
    for l in range(4, -1, -1):
        dpooled = _double_conv_bwd(eng, sv.enc[l], dcur, G, dt, tdt, l > 0 or need_dx, lazy)
        if l == 0:
            break
        Hp, Wp, Cp = sv.Hs[l - 1], sv.Ws[l - 1], _WIDTHS[l - 1]
        sg = skip_grads[l - 1]
        if model.use_se and Hp % 2 == 0 and Wp % 2 == 0:
            # the next block rebuilds dO = skip gradient + unpool(dpooled) inside its own reduction / apply kernels
            lazy, dcur = (sg, dpooled, sv.pos[l - 1]), None
            continue
        lazy = None
        dcur = torch.empty(B, Hp, Wp, Cp, dtype=tdt, device=dev)
        _lib.call("unetca_pool_bwd_add", dt, _ptr(sg), sg.stride(2), _ptr(dpooled), Cp, _ptr(sv.pos[l - 1]), _ptr(dcur),
                  Cp, B, Hp, Wp, Cp, st)
    dx = None
    if need_dx:
        Cin = model.in_channels
        dx = torch.empty(B, Cin, H, W, dtype=torch.float32, device=dev)
        _lib.call("unetca_nhwc_to_nchw", dt, _ptr(dpooled), Cin, _ptr(dx), B, Cin, H, W, st)
    return G.finish(), dx


def _param_grads(model, grads):
    """Gradients in parameter order for autograd; a sink that owns `.grad` itself (parallel.GradBuckets) hands None."""
    if grads is None:
        return (None,) * len(model._param_names)
    return tuple(grads[n] for n in model._param_names)


class _UNetFn(torch.autograd.Function):
    """logits = UNet(x); backward receives dlogits from whatever criterion follows (UCA:343-345)."""

    @staticmethod
    def forward(ctx, model, x, keep, *params):
        with torch.cuda.device(x.device):
            logits, sv = _forward(model, x, keep)
        ctx.model, ctx.sv = model, sv
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model, sv = ctx.model, ctx.sv
        if sv is None:
            raise RuntimeError("backward through a forward that ran without grad")
        g = dlogits.contiguous().float()
        one = torch.ones(1, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            grads, dx = _backward(model, sv, g, one, ctx.needs_input_grad[1])
        ctx.sv = None
        return (None, dx, None) + _param_grads(model, grads)


class _UNetLossFn(torch.autograd.Function):
    """Fused model + nn.CrossEntropyLoss(ignore_index) (UCA:343-344, 465)."""

    @staticmethod
    def forward(ctx, model, x, target, ignore_index, keep, *params):
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        if target.shape != (B, H, W) or target.dtype != torch.int64 or not target.is_cuda:
            raise ValueError(f"target must be a CUDA int64 tensor of shape {(B, H, W)}")
        with torch.cuda.device(x.device):
            logits, sv = _forward(model, x, keep)
            nc = logits.shape[1]
            dev = logits.device
            target = target.contiguous()
            eng = model._engine()
            parts = eng.parts(B, dev)
            g = torch.empty_like(logits) if keep else None
            out = torch.empty(2, dtype=torch.float32, device=dev)
            gscale = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.call("unetca_cross_entropy", _ptr(logits), _ptr(target), nc, B, H * W, int(ignore_index), None, _ptr(g),
                      None, _ptr(parts), _ptr(out), _ptr(gscale), _stream())
        model.last_logits = logits
        ctx.model, ctx.sv, ctx.g, ctx.gscale = model, sv, g, gscale
        return out[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        model, sv = ctx.model, ctx.sv
        if sv is None:
            raise RuntimeError("backward through a forward that ran without grad")
        gscale = ctx.gscale * dloss.reshape(1).float()
        with torch.cuda.device(gscale.device):
            grads, dx = _backward(model, sv, ctx.g, gscale, ctx.needs_input_grad[1])
        ctx.sv = ctx.g = None
        return (None, dx, None, None, None) + _param_grads(model, grads)


class _SELayerFn(torch.autograd.Function):
    """SELayer.forward on its own (UCA:61-72): NCHW fp32 plane passes around the SE FC kernels."""

    @staticmethod
    def forward(ctx, x, w1, w2):
        B, C, H, W = x.shape
        Cr, hw, st, dev = w1.shape[0], H * W, _stream(), x.device
        xf = x.detach().float().contiguous()
        w1f, w2f = w1.detach().float().contiguous(), w2.detach().float().contiguous()
        sums = torch.empty(B, C, dtype=torch.float32, device=dev)
        _lib.call("unetca_plane_dot", _ptr(xf), None, B * C, hw, _ptr(sums), st)
        p = torch.empty(B, C, dtype=torch.float32, device=dev)
        z = torch.empty(B, Cr, dtype=torch.float32, device=dev)
        s = torch.empty(B, C, dtype=torch.float32, device=dev)
        _lib.call("unetca_se_fc", _ptr(sums), 1, B, C, Cr, hw, _ptr(w1f), _ptr(w2f), _ptr(p), _ptr(z), _ptr(s), st)
        out = torch.empty_like(xf)
        _lib.call("unetca_plane_scale_add", _ptr(xf), _ptr(s), None, 0.0, B * C, hw, _ptr(out), st)
        ctx.save_for_backward(xf, w1f, w2f, p, z, s)
        return out

    @staticmethod
    def backward(ctx, dy):
        xf, w1f, w2f, p, z, s = ctx.saved_tensors
        B, C, H, W = xf.shape
        Cr, hw, dev = w1f.shape[0], H * W, xf.device
        with torch.cuda.device(dev):
            st = _stream()
            dyf = dy.float().contiguous()
            ds = torch.empty(B, C, dtype=torch.float32, device=dev)
            _lib.call("unetca_plane_dot", _ptr(dyf), _ptr(xf), B * C, hw, _ptr(ds), st)
            dpre2 = torch.empty(B, C, dtype=torch.float32, device=dev)
            dz = torch.empty(B, Cr, dtype=torch.float32, device=dev)
            dp = torch.empty(B, C, dtype=torch.float32, device=dev)
            dw1, dw2 = torch.empty_like(w1f), torch.empty_like(w2f)
            _lib.call("unetca_se_fc_bwd", _ptr(ds), 1, B, C, Cr, _ptr(w1f), _ptr(w2f), _ptr(p), _ptr(z), _ptr(s), _ptr(dpre2),
                      _ptr(dz), _ptr(dp), _ptr(dw1), _ptr(dw2), st)
            dx = torch.empty_like(xf)
            _lib.call("unetca_plane_scale_add", _ptr(dyf), _ptr(s), _ptr(dp), 1.0 / hw, B * C, hw, _ptr(dx), st)
        return dx, dw1, dw2


class _DoubleConvFn(torch.autograd.Function):
    """DoubleConv.forward on its own (UCA:96-97): the block kernels of UNet between two NCHW <-> NHWC boundary passes."""

    @staticmethod
    def forward(ctx, mod, x, keep, *params):
        eng, blk = mod._engine(), mod._blk
        dt, tdt = mod._dt()
        train = mod.training
        if train or eng.last_train:
            eng.epoch += 1
        eng.last_train = train
        B, C, H, W = x.shape
        O, dev = blk.cout, x.device
        if train and B * H * W <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {[B, O, H, W]}")
        with torch.cuda.device(dev):
            st = _stream()
            xf = x.detach().float().contiguous()
            if blk.first:
                xin, col = None, _first_conv_rows(dt, tdt, xf, B, C, H, W)
            else:
                xin, col = torch.empty(B, H, W, C, dtype=tdt, device=dev), None
                _lib.call("unetca_nchw_to_nhwc", dt, _ptr(xf), _ptr(xin), C, B, C, H, W, st)
            out = torch.empty(B, H, W, O, dtype=tdt, device=dev)
            sv = _double_conv_fwd(eng, blk, xin, col, B, H, W, out, None, None, train, dt, tdt, keep)
            y = torch.empty(B, O, H, W, dtype=torch.float32, device=dev)
            _lib.call("unetca_nhwc_to_nchw", dt, _ptr(out), O, _ptr(y), B, O, H, W, st)
        ctx.mod, ctx.sv = mod, sv if keep else None
        return y

    @staticmethod
    def backward(ctx, dy):
        mod, sv = ctx.mod, ctx.sv
        if sv is None:
            raise RuntimeError("backward through a forward that ran without grad")
        eng, blk = mod._engine(), mod._blk
        dt, tdt = mod._dt()
        B, H, W, C, O, dev = sv.B, sv.H, sv.W, blk.cin, blk.cout, dy.device
        with torch.cuda.device(dev):
            st = _stream()
            g = dy.float().contiguous()
            dout = torch.empty(B, H, W, O, dtype=tdt, device=dev)
            _lib.call("unetca_nchw_to_nhwc", dt, _ptr(g), _ptr(dout), O, B, O, H, W, st)
            G = _GradSink()
            need_dx = ctx.needs_input_grad[1]
            dxn = _double_conv_bwd(eng, sv, dout, G, dt, tdt, need_dx)
            dx = None
            if need_dx:
                dx = torch.empty(B, C, H, W, dtype=torch.float32, device=dev)
                _lib.call("unetca_nhwc_to_nchw", dt, _ptr(dxn), C, _ptr(dx), B, C, H, W, st)
        ctx.sv = None
        grads = G.finish()
        return (None, dx, None) + tuple(grads[n] for n, _ in mod.named_parameters())


class UNet(nn.Module):
    """U-Net with optional channel attention; constructor and attribute tree of the reference (UCA:100-125)."""

    def __init__(self, in_channels: int = 1, num_classes: int = 2, use_se: bool = False):
        super().__init__()
        self.inc = DoubleConv(in_channels, 64, use_se=use_se)
        self.down1 = nn.Sequential(nn.MaxPool2d(2), DoubleConv(64, 128, use_se=use_se))
        self.down2 = nn.Sequential(nn.MaxPool2d(2), DoubleConv(128, 256, use_se=use_se))
        self.down3 = nn.Sequential(nn.MaxPool2d(2), DoubleConv(256, 512, use_se=use_se))
        self.down4 = nn.Sequential(nn.MaxPool2d(2), DoubleConv(512, 1024, use_se=use_se))
        self.up1 = nn.ConvTranspose2d(1024, 512, kernel_size=2, stride=2)
        self.conv1 = DoubleConv(1024, 512, use_se=use_se)
        self.up2 = nn.ConvTranspose2d(512, 256, kernel_size=2, stride=2)
        self.conv2 = DoubleConv(512, 256, use_se=use_se)
        self.up3 = nn.ConvTranspose2d(256, 128, kernel_size=2, stride=2)
        self.conv3 = DoubleConv(256, 128, use_se=use_se)
        self.up4 = nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2)
        self.conv4 = DoubleConv(128, 64, use_se=use_se)
        self.outc = nn.Conv2d(64, num_classes, kernel_size=1)

        object.__setattr__(self, "in_channels", in_channels)
        object.__setattr__(self, "num_classes", num_classes)
        object.__setattr__(self, "use_se", use_se)
        object.__setattr__(self, "precision", "bf16")
        object.__setattr__(self, "last_logits", None)
        object.__setattr__(self, "_eng", None)
        object.__setattr__(self, "_grad_sink_factory", None)     # set by parallel.GradBuckets (data parallel)
        enc = [("inc.double_conv", self.inc), ("down1.1.double_conv", self.down1[1]),
               ("down2.1.double_conv", self.down2[1]), ("down3.1.double_conv", self.down3[1]),
               ("down4.1.double_conv", self.down4[1])]
        cin = in_channels
        blocks = []
        for l, (name, mod) in enumerate(enc):
            blocks.append(_Block(name, mod, cin, _WIDTHS[l], l, first=(l == 0)))
            cin = _WIDTHS[l]
        object.__setattr__(self, "_enc_blocks", blocks)
        dec = [("conv1.double_conv", self.conv1, 3), ("conv2.double_conv", self.conv2, 2),
               ("conv3.double_conv", self.conv3, 1), ("conv4.double_conv", self.conv4, 0)]
        object.__setattr__(self, "_dec_blocks", [_Block(n, m, 2 * _WIDTHS[l], _WIDTHS[l], l) for n, m, l in dec])
        object.__setattr__(self, "_ups", [self.up1, self.up2, self.up3, self.up4])
        object.__setattr__(self, "_up_names", ["up1", "up2", "up3", "up4"])
        object.__setattr__(self, "_param_names", [n for n, _ in self.named_parameters()])

    # ---- configuration --------------------------------------------------------------------------------
    def set_precision(self, precision: str) -> "UNet":
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        object.__setattr__(self, "precision", precision)
        return self

    def _dt(self):
        return (_lib.BF16, torch.bfloat16) if self.precision == "bf16" else (_lib.F32, torch.float32)

    def _engine(self) -> _Engine:
        if self._eng is None:
            object.__setattr__(self, "_eng", _Engine(self))
        return self._eng

    def _keep(self, x=None) -> bool:
        """Will this forward be differentiated?  (decided here: autograd.Function.forward always runs in no-grad mode)"""
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                            (x is not None and x.requires_grad))

    def invalidate_packed(self) -> None:
        """Call after writing parameters behind autograd's back (`p.data.copy_()`, EMA/SWA swaps, raw-pointer optimizers):
        the next forward re-derives the packed bf16 operand copies from the fp32 parameters."""
        self._engine().invalidate()

    def _params(self):
        return [p for _, p in self.named_parameters()]

    # ---- reference API --------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B,Cin,H,W) float -> class logits (B,num_classes,H,W), like the reference's UNet.forward (UCA:127-163)."""
        _check_input(self, x)
        out = _UNetFn.apply(self, x, self._keep(x), *self._params())
        return out if x.dtype == torch.float32 else out.to(x.dtype)

    # ---- fused extras ---------------------------------------------------------------------------------
    def loss(self, images: torch.Tensor, masks: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
        """criterion(model(images), masks) with nn.CrossEntropyLoss(ignore_index) semantics, fused."""
        _check_input(self, images)
        return _UNetLossFn.apply(self, images, masks, ignore_index, self._keep(images), *self._params())

    @torch.no_grad()
    def predict_mask(self, images: torch.Tensor) -> torch.Tensor:
        """torch.max(model(images), 1)[1] (UCA:220): int64 class map, ties -> lowest class."""
        logits = self.forward(images).float().contiguous()
        B, nc, H, W = logits.shape
        mask = torch.empty(B, H, W, dtype=torch.int64, device=logits.device)
        with torch.cuda.device(logits.device):
            _lib.call("unetca_cross_entropy", _ptr(logits), None, nc, B, H * W, -1, None, None, _ptr(mask), None, None, None,
                      _stream())
        return mask
