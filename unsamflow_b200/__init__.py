"""unsamflow_b200 -- B200-native (sm_100a) correlation cost volume + flow_warp for PWC-Lite.

Drop-in for the hot path of turswiming/UnSAMFlow:
    from unsamflow_b200 import Correlation, flow_warp
replaces `models/correlation_package/correlation.Correlation` (or correlation_native) and
`utils.warp_utils.flow_warp`; `get_occu_mask_backward` / `get_occu_mask_bidirection` replace the two
occlusion masks of `utils.warp_utils` that the photometric loss builds from the flows.  The math runs in hand-written CUDA kernels behind the C ABI in
include/unsamflow_b200.h; there is no CPU or PyTorch fallback.
"""
from .correlation import Correlation, CorrelationFunction  # noqa: F401
from .occlusion import get_occu_mask_backward, get_occu_mask_bidirection  # noqa: F401
from .segpool import segment_max_pool  # noqa: F401
from .warp import FlowWarpFunction, flow_warp  # noqa: F401

__all__ = ["Correlation", "CorrelationFunction", "FlowWarpFunction", "flow_warp",
           "get_occu_mask_backward", "get_occu_mask_bidirection", "segment_max_pool"]
__version__ = "0.1.0"
