"""Drop-in for the UPFlow operators on the hot path: correlation_package.correlation.CorrelationFunction,
pwc_modules.upsample2d_flow_as, pwc_modules.WarpingLayer_no_div, utils.pytorch_correlation.Corr_pyTorch; the flow network itself
(UPFlow_net.forward_2_frame_v3, occlusion check) is `upflow.net.UPFlowNet` / `upflow.net.occ_check`."""
from .correlation import CorrelationFunction, correlation_cuda      # noqa: F401
from .pwc_modules import WarpingLayer_no_div, upsample2d_flow_as    # noqa: F401
from .utils.pytorch_correlation import Corr_pyTorch                # noqa: F401
