from vknet.kernel_updator import KernelUpdator  # noqa: F401  (knet_vis/kernel_updator.py:7)
