"""Alias package: `custom_imports` of the reference configs name `knet.kernel_updator`,
`knet.det.kernel_update_head`, `knet.video.kernel_update_head`
(configs/det/_base_/models/knet_kitti_step_s3_r50_fpn.py:179-220).  With `video-k-net_b200/` ahead of the
reference tree on PYTHONPATH those imports land here and register the B200 implementations."""
from _overlay import extend_path as _ext
__path__ = _ext(__path__, __name__)
