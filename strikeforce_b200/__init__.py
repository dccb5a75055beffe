"""strikeforce_b200 -- B200-native batched StrikeForce tick engine.

The simulation lives in ``libstrikeforce_b200.so`` (CUDA, sm_100a) behind the C ABI of
``include/strikeforce_b200.h``; this package is the thin host side: data loaders, the ctypes
binding, the batched driver and the mirror of the reference's bot plugin surface.  There is no
CPU implementation."""
from . import config, data  # noqa: F401

__all__ = ["config", "data", "lib", "sim", "bots", "dist"]
