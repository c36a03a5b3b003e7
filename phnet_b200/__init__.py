"""phnet_b200 -- B200 (sm_100a) drop-in for PHNet's lane-NMS op (`libs/ops`).

    from phnet_b200.ops import nms            # same signature as libs/ops/nms.py:32
    phnet_b200.install_as_libs_ops()          # makes `from libs.ops import nms` resolve to this package
"""
import sys
import types

__version__ = "0.1.0"


def install_as_libs_ops() -> None:
    """Register this package's op as `libs.ops` / `libs.ops.nms` so PHNet's `from libs.ops import nms`
    (libs/models/Router4OL.py:10 and siblings) picks it up without editing the reference tree."""
    from . import ops
    libs = sys.modules.get("libs")
    if libs is None:
        libs = types.ModuleType("libs")
        libs.__path__ = []  # namespace-like
        sys.modules["libs"] = libs
    sys.modules["libs.ops"] = ops
    sys.modules["libs.ops.nms"] = sys.modules["phnet_b200.ops.nms"]
    setattr(libs, "ops", ops)
