"""Mirror of PHNet `libs/ops/__init__.py:1-3`: `from phnet_b200.ops import nms`."""
from .nms import nms, nms_batched, sort_order, plan

__all__ = ["nms", "nms_batched", "sort_order", "plan"]
