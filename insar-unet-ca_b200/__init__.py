"""B200-native U-Net-CA hot path (see ../README.md).

The directory name carries a hyphen (it is the repository's package name), so it cannot appear in an `import`
statement; `import unetca_b200` (the alias package next to this directory) is the canonical way in, and
`importlib.import_module("insar-unet-ca_b200")` works as well when the repository root is on `sys.path`.
"""
from .model import UNet, DoubleConv, SELayer  # noqa: F401
from . import _lib  # noqa: F401
