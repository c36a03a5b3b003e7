"""Mirror of PHNet `libs/ops/__init__.py:1-3`: `from phnet_b200.ops import nms`."""
from .nms import nms, nms_batched, sort_order, plan
from .pipeline import HostLaneNMS, nms_host
from .get_lanes import get_lanes, decode_lanes
from .line_iou import line_iou
from .assign import dynamic_k_assign, dynamic_k_assign_batched
from .graphed import GraphedNMS

__all__ = ["nms", "nms_batched", "sort_order", "plan", "HostLaneNMS", "nms_host", "get_lanes", "decode_lanes", "line_iou",
           "dynamic_k_assign", "dynamic_k_assign_batched", "GraphedNMS"]
