"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; the nn.Module mirrors the reference's constructor / state_dict; and nothing falls back to the CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_ca_port as port


def test_library_exports_every_declared_symbol(built_lib):
    import unetca_b200
    decls = unetca_b200._lib.parse_header()
    assert len(decls) >= 45
    # the development knobs live in their own header: the drop-in ABI declares no process-global switch but conv_impl
    tuning = unetca_b200._lib.parse_header(unetca_b200._lib.TUNING_HEADER)
    assert tuning and not (set(tuning) & set(decls))
    assert not [n for n in decls if "force" in n or "set_tuning" in n or n.startswith("unetca_tc_set")]
    for name in list(decls) + list(tuning):
        assert hasattr(built_lib, name), name
    assert built_lib.unetca_abi_version() == 1
    assert built_lib.unetca_max_parts(4) > 4


def test_sass_is_blackwell_native(built_lib):
    """The tensor-core path must be tcgen05/TMA (UTCHMMA / UTMALDG / UTMASTG / LDTM in SASS), not mma.sync (HMMA)."""
    import subprocess
    import unetca_b200
    sass = subprocess.run(["cuobjdump", "-sass", unetca_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnem in sass, mnem
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_module_mirrors_reference_state_dict(golden_dir):
    import unetca_b200
    g = np.load(os.path.join(golden_dir, "unetca_se_b2_32.npz"))
    m = unetca_b200.UNet(in_channels=3, num_classes=2, use_se=True)
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert sum(p.numel() for p in m.parameters()) == 31_261_698
    sd = port.make_state_dict(seed=0)
    m.load_state_dict(sd, strict=True)                               # reference checkpoints load unchanged
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and v.dtype == sd[k].dtype, k
    g2 = np.load(os.path.join(golden_dir, "unet_plain_b2_32.npz"))
    m2 = unetca_b200.UNet(3, 2)                                      # use_se defaults to False like the reference
    assert list(m2.state_dict().keys()) == [str(k) for k in g2["keys"]]
    m3 = unetca_b200.UNet()                                          # defaults: in_channels=1, num_classes=2
    assert m3.inc.double_conv[0].weight.shape == (64, 1, 3, 3)
    assert m3.outc.weight.shape == (2, 64, 1, 1)


def test_package_directory_is_importable():
    """`insar-unet-ca_b200/` is a real package (importlib: the hyphen cannot appear in an import statement); `unetca_b200` is
    its canonical alias."""
    import importlib
    m = importlib.import_module("insar-unet-ca_b200")
    assert m.UNet.__name__ == "UNet" and hasattr(m, "DoubleConv") and hasattr(m, "SELayer")


def test_no_cpu_fallback():
    import unetca_b200
    m = unetca_b200.UNet(3, 2, True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.loss(torch.zeros(1, 3, 32, 32), torch.zeros(1, 32, 32, dtype=torch.int64))


def test_standalone_modules_have_no_cpu_fallback_either():
    import unetca_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        unetca_b200.SELayer(64)(torch.zeros(1, 64, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        unetca_b200.DoubleConv(3, 64, use_se=True)(torch.zeros(1, 3, 8, 8))


def test_reference_arm_is_the_unmodified_reference():
    """build() stages a byte-identical copy of the reference's hot-path script in the git-ignored oracle/_ref/ whenever
    /root/reference is present (bench.py --impl reference and cpu_baseline import UNet / CrossEntropyLoss from it)."""
    import __graft_entry__ as g
    if not os.path.exists(g.REFERENCE_SRC):
        pytest.skip("no /root/reference on this box")
    assert g.stage_reference()
    assert open(g.REFERENCE_COPY, "rb").read() == open(g.REFERENCE_SRC, "rb").read()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert "oracle/_ref/" in open(os.path.join(root, ".gitignore")).read()


def test_product_does_not_import_oracle():
    """The product path may not route through oracle/ (it is test infrastructure)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in ("insar-unet-ca_b200", "unetca_b200"):
        for dirpath, _, files in os.walk(os.path.join(root, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
