"""Alias package for configs/video_knet_vis (`knet_vis.kernel_updator`, `knet_vis.det.kernel_update_head`;
configs/video_knet_vis/_base_/models/knet_track_r50.py:223-244)."""
from _overlay import extend_path as _ext
__path__ = _ext(__path__, __name__)
