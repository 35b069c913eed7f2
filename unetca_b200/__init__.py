"""Importable alias of the `insar-unet-ca_b200/` package directory (a hyphen is not a legal module name).

`import unetca_b200` exposes everything in `/insar-unet-ca_b200/` (model, ops, _lib, parallel, tiling): this file
only points the package search path there and re-exports the public names.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "insar-unet-ca_b200"))

from .model import UNet, DoubleConv, SELayer  # noqa: E402,F401
from . import _lib  # noqa: E402,F401
