from vknet.kernel_updator import KernelUpdator  # noqa: F401  (knet/kernel_updator.py:7)
