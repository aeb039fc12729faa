from vknet.kernel_update_head import KernelUpdateHead  # noqa: F401  (knet/det/kernel_update_head.py:16)
