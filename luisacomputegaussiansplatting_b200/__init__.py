"""lcgs-b200: B200-native forward splat-render path of LuisaComputeGaussianSplatting.

`lcgs` (host mirror of the reference's classes over the C ABI) needs torch + a CUDA device and is
imported on demand; `scenes` (synthetic inputs) is numpy-only.
"""
from . import scenes  # noqa: F401

__all__ = ["scenes", "lcgs", "build"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "lcgs":
        import importlib

        return importlib.import_module(".lcgs", __name__)
    raise AttributeError(name)
