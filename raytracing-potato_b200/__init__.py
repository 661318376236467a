"""raytracing-potato_b200 — B200-native rendering core for raytracing-potato's hot path.

The product is `lib/librtp_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/rtp.h).
This package is the thin Python host layer over that ABI: `api` mirrors the reference crate's scene
and render interface, `scenes` restates its example scenes and the BASELINE configs, `assets`
provides their inputs. The directory name contains a hyphen; import it with
`importlib.import_module("raytracing-potato_b200")` (see `rtp_b200.py` at the repo root).
"""
from . import _abi  # noqa: F401
from . import api, assets, dist, scenes  # noqa: F401
from .api import *  # noqa: F401,F403

__all__ = ["api", "assets", "dist", "scenes", "_abi"]
