"""vknet -- B200-native (sm_100a) implementation of Video K-Net's KernelUpdateHead hot path.

The package is the host-side mirror of the reference's plugin interface: `KernelUpdateHead`,
`VideoKernelUpdateHead` and `KernelUpdator` keep the reference's registry names, constructor
kwargs, state_dict keys and forward contracts, and call hand-written CUDA through the C ABI of
libvknet.so (include/vknet.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib
from ._lib import VknError, kernel_names
from .registry import HEADS, TRANSFORMER_LAYER, build_head, build_transformer_layer
from .kernel_updator import KernelUpdator
from .kernel_update_head import KernelUpdateHead
from .video_kernel_update_head import VideoKernelUpdateHead
from .tracker_kernel_update_head import KernelUpdateHeadVideo
from .iter_loop import FramesInFlight, KernelIterLoop

__all__ = ['KernelUpdator', 'KernelUpdateHead', 'VideoKernelUpdateHead', 'KernelUpdateHeadVideo', 'KernelIterLoop', 'FramesInFlight', 'HEADS',
           'TRANSFORMER_LAYER', 'build_head', 'build_transformer_layer', 'VknError', 'kernel_names', '_lib']
