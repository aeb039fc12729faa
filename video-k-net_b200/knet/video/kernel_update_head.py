from vknet.video_kernel_update_head import VideoKernelUpdateHead  # noqa: F401  (knet/video/kernel_update_head.py:17)
