from vknet.tracker_kernel_update_head import KernelUpdateHeadVideo  # noqa: F401  (knet_vis/tracker/kernel_update_head.py:19)
