"""Resume checkpoints: round trip of model + Adam state, reference-file compatibility."""
import os

import torch

import unetca_b200
from unetca_b200 import checkpoint
from oracle import unet_ca_port as port


def _fake_step(model, opt, seed):
    g = torch.Generator().manual_seed(seed)
    for p in model.parameters():
        p.grad = torch.randn(p.shape, generator=g) * 1e-2
    opt.step()


def test_resume_round_trip(tmp_path):
    m = unetca_b200.UNet(3, 2, use_se=True)
    m.load_state_dict(port.make_state_dict(seed=4))
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    _fake_step(m, opt, 1)
    _fake_step(m, opt, 2)
    hist = [{"epoch": 1, "train_loss": 0.7, "val_miou": 0.31}, {"epoch": 2, "train_loss": 0.6, "val_miou": 0.35}]
    path = os.path.join(tmp_path, "ckpt", "resume.pt")
    checkpoint.save_resume(path, m, opt, epoch=2, best_m_iou=0.35, history=hist)
    assert not [f for f in os.listdir(os.path.dirname(path)) if ".tmp." in f]       # atomic rename, nothing left over
    m2 = unetca_b200.UNet(3, 2, use_se=True)
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-4)
    epoch, best, h = checkpoint.load_resume(path, m2, opt2)
    assert (epoch, best, h) == (2, 0.35, hist)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    # the restored optimizer continues bit-identically
    _fake_step(m, opt, 3)
    _fake_step(m2, opt2, 3)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_best_model_file_is_the_reference_format(tmp_path):
    """save_best writes what UCA:386 writes: a bare state_dict with the reference's keys, loadable either way."""
    m = unetca_b200.UNet(3, 2, use_se=True)
    sd = port.make_state_dict(seed=5)
    m.load_state_dict(sd)
    path = os.path.join(tmp_path, "best.pth")
    checkpoint.save_best(m, path)
    raw = torch.load(path, weights_only=True)
    assert list(raw.keys()) == list(sd.keys()) and all(torch.equal(raw[k], sd[k]) for k in sd)
    m2 = unetca_b200.UNet(3, 2, use_se=True)
    assert checkpoint.load_resume(path, m2) == (0, -1.0, [])                       # a bare reference checkpoint
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
