"""oracle/make_golden.py — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference classes (`UNet`, `nn.CrossEntropyLoss`) imported
by path from /root/reference/Unet-ChannalAttention.py on CPU, fp32, on the seeded fixtures of
oracle/unet_ca_port.py (`make_state_dict`, `make_batch`).  Run in the build container only (the GPU box has no
/root/reference):  python oracle/make_golden.py [case ...]      (UNETCA_GOLDEN_OUT=/tmp/g re-generates elsewhere;
oracle/check_golden.py diffs such a directory against tests/golden)

The reference ships no tests or golden vectors of its own, so these files are what pins the oracle.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_ca_port as port  # noqa: E402

REF = "/root/reference/Unet-ChannalAttention.py"
OUT = os.environ.get("UNETCA_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))   # override to re-generate elsewhere and diff


def load_reference():
    spec = importlib.util.spec_from_file_location("unet_ca_reference", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_case(ref, name, seed, B, H, W, use_se, full):
    torch.manual_seed(0)
    sd = port.make_state_dict(seed=seed, use_se=use_se)
    x, y = port.make_batch(seed, B, H, W)
    model = ref.UNet(in_channels=3, num_classes=2, use_se=use_se)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    crit = torch.nn.CrossEntropyLoss(ignore_index=255)
    # pool indices as autograd would save them: hook the MaxPool2d inputs
    pool_in = []
    hooks = [m.register_forward_hook(lambda m_, i, o: pool_in.append(i[0].detach().clone()))
             for m in model.modules() if isinstance(m, torch.nn.MaxPool2d)]
    logits = model(x)
    loss = crit(logits, y)
    loss.backward()
    for h in hooks:
        h.remove()
    out = {
        "keys": np.array(list(model.state_dict().keys())),
        "loss": np.float32(loss.item()),
        "argmax_packed": np.packbits(torch.max(logits, 1)[1].numpy().astype(np.uint8)),
        "grad_norms": np.array([p.grad.norm().item() for _, p in model.named_parameters()], dtype=np.float64),
        "param_names": np.array([n for n, _ in model.named_parameters()]),
    }
    if full:
        out["logits"] = logits.detach().numpy()
        for i, t in enumerate(pool_in):
            _, idx = torch.nn.functional.max_pool2d(t, 2, return_indices=True)
            out[f"pool_idx{i}"] = idx.numpy().astype(np.int32)
        named = dict(model.named_parameters())
        for k in ("outc.weight", "outc.bias", "inc.double_conv.1.weight", "inc.double_conv.1.bias",
                  "inc.double_conv.0.weight", "conv4.double_conv.4.weight", "up4.bias"):
            out["grad:" + k] = named[k].grad.numpy()
        if use_se:
            for k in ("inc.double_conv.6.fc.0.weight", "inc.double_conv.6.fc.2.weight"):
                out["grad:" + k] = named[k].grad.numpy()
        bufs = dict(model.named_buffers())
        for k in ("inc.double_conv.1.running_mean", "inc.double_conv.1.running_var",
                  "down4.1.double_conv.4.running_mean", "down4.1.double_conv.4.running_var"):
            out["buf:" + k] = bufs[k].numpy()
    else:
        out["logits_sub"] = logits.detach().numpy()[:, :, ::8, ::8].copy()
    # eval-mode forward with the updated running stats
    model.eval()
    with torch.no_grad():
        ev = model(x)
    out["eval_logits_sub" if not full else "eval_logits"] = ev.numpy() if full else ev.numpy()[:, :, ::8, ::8].copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", loss.item(), "->", path, os.path.getsize(path) // 1024, "KiB")


def run_metrics(ref):
    """compute_metrics (UCA:214-269) of the unmodified reference on the seeded cases of oracle/metrics_port.py."""
    from oracle import metrics_port
    out = {}
    for name, logits, masks, nc in metrics_port.metric_cases():
        m = ref.compute_metrics(logits, masks, nc)
        out[name] = np.array([m["acc"], m["miou"], m["mpa"], m["mf1"]], dtype=np.float64)
        print("metrics", name, out[name])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


def run_dataset(ref):
    """A miniature VOC tree (tests/golden/voc_mini) and what the reference's VOCSegDataset + transforms
    (UCA:166-212, 428-433) return for it."""
    import torchvision.transforms as T
    from PIL import Image
    root = os.path.join(OUT, "voc_mini")
    for d in ("JPEGImages", "SegmentationClass", os.path.join("ImageSets", "Segmentation")):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    rs = np.random.RandomState(5)
    ids = {"train": ["t0", "t1", "t2", "t3"], "val": ["v0", "v1"]}
    for split, names in ids.items():
        with open(os.path.join(root, "ImageSets", "Segmentation", f"{split}.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        for n in names:
            h, w = (24, 20) if n != "t3" else (32, 32)                   # t3 already has the target size
            yy, xx = np.mgrid[0:h, 0:w]
            img = (127 + 90 * np.sin(xx / 3.0 + rs.rand() * 6) * np.cos(yy / 4.0) + rs.randint(-20, 20, (h, w))).clip(0, 255)
            Image.fromarray(img.astype(np.uint8), "L").save(os.path.join(root, "JPEGImages", f"{n}.jpg"), quality=92)
            m = np.where(rs.rand(h, w) < 0.45, 255, 0).astype(np.uint8)
            m[rs.rand(h, w) < 0.05] = 128                                # a grey value: .long() truncates it to 0
            Image.fromarray(m, "L").save(os.path.join(root, "SegmentationClass", f"{n}.png"))
    S = 32
    tf = T.Compose([T.Resize((S, S)), T.ToTensor(), T.Normalize(mean=[0.5], std=[0.5])])
    out = {}
    for split in ids:
        ds = ref.VOCSegDataset(voc_root=root, image_size=S, image_set=split, transforms=tf)
        for i, n in enumerate(ds.ids):
            img, mask = ds[i]
            out[f"img:{n}"] = img.numpy()
            out[f"mask:{n}"] = mask.numpy()
    np.savez_compressed(os.path.join(OUT, "voc_mini_expected.npz"), **out)
    print("dataset fixture:", root, len(out) // 2, "items")


def run_trajectory(ref, name, mode, seed, B, H, W, steps):
    """100 steps of the reference's own train loop body (UCA:338-346: zero_grad, forward, criterion, backward,
    Adam(lr=1e-4).step) with the UNMODIFIED classes on the seeded batches of port.trajectory_batch(mode, step): the loss
    and the global gradient norm of every step, in fp32 (the golden values) AND in fp64 — the difference is the
    reference's own sensitivity to rounding, i.e. how tight a per-step bound on another implementation can be.
    B = 4, 128x128: the bottleneck BatchNorm normalises over B*(H/16)*(W/16) = 256 values."""
    out = {}
    for dtype, tag in ((torch.float32, ""), (torch.float64, "ref_fp64_")):
        sd = port.make_state_dict(seed=seed)
        model = ref.UNet(in_channels=3, num_classes=2, use_se=True)
        model.load_state_dict(sd, strict=True)
        model = model.to(dtype)
        model.train()
        crit = torch.nn.CrossEntropyLoss(ignore_index=255)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        losses, gnorms = [], []
        for s in range(steps):
            x, y = port.trajectory_batch(mode, s, B, H, W)
            opt.zero_grad()
            loss = crit(model(x.to(dtype)), y)
            loss.backward()
            gnorms.append(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())).item())
            losses.append(loss.item())
            opt.step()
        out[tag + "loss"] = np.array(losses, dtype=np.float64)
        out[tag + "grad_norm"] = np.array(gnorms, dtype=np.float64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, cfg=np.array([seed, B, H, W, steps]), mode=np.array(mode), **out)
    sens = np.abs(out["grad_norm"] - out["ref_fp64_grad_norm"]) / out["ref_fp64_grad_norm"]
    print(name, "loss", out["loss"][0], "->", out["loss"][-1], "gnorm", out["grad_norm"][0], "->", out["grad_norm"][-1],
          "reference fp32 vs its own fp64: worst gnorm rel", sens.max(), "first step > 1e-2:",
          int(np.argmax(sens > 1e-2)) if (sens > 1e-2).any() else None, path)


def main():
    """python oracle/make_golden.py [case ...] — no arguments: regenerate everything."""
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    cases = [
        ("metrics", lambda: run_metrics(ref)),
        ("voc_mini", lambda: run_dataset(ref)),
        ("unetca_se_b2_32", lambda: run_case(ref, "unetca_se_b2_32", seed=0, B=2, H=32, W=32, use_se=True, full=True)),
        ("unet_plain_b2_32", lambda: run_case(ref, "unet_plain_b2_32", seed=1, B=2, H=32, W=48, use_se=False, full=True)),
        # BASELINE configs[0]
        ("unetca_se_b4_256", lambda: run_case(ref, "unetca_se_b4_256", seed=0, B=4, H=256, W=256, use_se=True, full=False)),
        # H, W not multiples of 16: floor max-pools (5 -> 2, 13 -> 6) and the bilinear resize guard of UCA:138-157
        ("unetca_se_b2_40x52", lambda: run_case(ref, "unetca_se_b2_40x52", seed=3, B=2, H=40, W=52, use_se=True, full=True)),
        # the benchmarked tile size (BASELINE configs[1] is 64 of these): 8 images of 3x512x512 — what host RAM and a
        # CPU oracle run in test time; every kernel family the B=64 bench dispatches is on this path
        ("unetca_se_b8_512", lambda: run_case(ref, "unetca_se_b8_512", seed=4, B=8, H=512, W=512, use_se=True, full=False)),
        # well-conditioned trajectory (fresh learnable batches): the fixture the 1e-2-per-step bound is asserted on
        ("trajectory_struct_b4_128", lambda: run_trajectory(ref, "trajectory_struct_b4_128", "struct", seed=7, B=4, H=128, W=128, steps=100)),
        # chaotic trajectory (four random-label batches memorised): the reference leaves its own fp64 run by 5 % after ~60 steps
        ("trajectory_b4_128", lambda: run_trajectory(ref, "trajectory_b4_128", "random4", seed=7, B=4, H=128, W=128, steps=100)),
    ]
    want = set(sys.argv[1:])
    unknown = want - {n for n, _ in cases}
    assert not unknown, f"unknown case(s) {sorted(unknown)}"
    for name, fn in cases:
        if not want or name in want:
            fn()


if __name__ == "__main__":
    main()
