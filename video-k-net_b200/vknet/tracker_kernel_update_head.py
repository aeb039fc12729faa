"""KernelUpdateHeadVideo -- drop-in for knet_vis/tracker/kernel_update_head.py:19-374 (the clip-level
stage of configs/video_knet_vis): x [B,F,C,H,W], mask_preds [B,F,N,H,W].

  gathered mode  (proposal_feat [B,N,C,1,1], with_cls=True): the pooled feature is the MEAN over the F
                 frames of a clip (:245-246, query_merge_method='mean'); ONE kernel set per clip is updated
                 and convolves all F frames (:329-339)            -> (cls, masks [B,F,N,H,W], obj [B,N,C,1,1])
  per-frame mode (proposal_feat [B,F,N,C,1,1], with_cls=False): every frame is an independent stage
                 (:268-278, :340-352)                              -> (None, masks, obj [B,F,N,C,1,1])
'attention' / 'attention_pos' query merging is selected by no shipped config and raises.
"""
import torch

from . import _lib
from .kernel_update_head import _HeadBase
from .registry import HEADS


@HEADS.register_module(force=True)
class KernelUpdateHeadVideo(_HeadBase):

    def __init__(self, with_cls=True, num_proposals=100, num_classes=80, num_ffn_fcs=2, num_heads=8, num_cls_fcs=1,
                 num_mask_fcs=3, feedforward_channels=2048, in_channels=256, out_channels=256, dropout=0.0,
                 mask_thr=0.5, act_cfg=dict(type='ReLU', inplace=True), ffn_act_cfg=dict(type='ReLU', inplace=True),
                 conv_kernel_size=3, feat_transform_cfg=None, hard_mask_thr=0.5, kernel_init=False, with_ffn=True,
                 mask_out_stride=4, relative_coors=False, relative_coors_off=False, feat_gather_stride=1,
                 mask_transform_stride=1, mask_upsample_stride=1, num_thing_classes=80, num_stuff_classes=53,
                 mask_assign_stride=4, ignore_label=255, thing_label_in_seg=0, query_merge_method='mean',
                 kernel_updator_cfg=dict(type='DynamicConv', in_channels=256, feat_channels=64, out_channels=256,
                                         input_feat_shape=1, act_cfg=dict(type='ReLU', inplace=True),
                                         norm_cfg=dict(type='LN')),
                 loss_rank=None, loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0),
                 loss_dice=dict(type='DiceLoss', loss_weight=3.0),
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0)):
        super().__init__()
        self._build_common(num_classes, num_ffn_fcs, num_heads, num_cls_fcs if with_cls else 0, num_mask_fcs,
                           feedforward_channels, in_channels, out_channels, dropout, mask_thr, act_cfg, ffn_act_cfg,
                           conv_kernel_size, feat_transform_cfg, hard_mask_thr, kernel_init, with_ffn, mask_out_stride,
                           relative_coors, relative_coors_off, feat_gather_stride, mask_transform_stride,
                           mask_upsample_stride, num_thing_classes, num_stuff_classes, mask_assign_stride,
                           ignore_label, thing_label_in_seg, kernel_updator_cfg, loss_rank, loss_mask, loss_dice,
                           loss_cls)
        self.with_cls = with_cls
        self.num_proposals = num_proposals
        self.query_merge_method = query_merge_method
        self._num_cls_out = self.fc_cls.out_features
        if not with_cls:                       # the reference builds no cls branch at all (:137-150)
            del self.cls_fcs
            del self.fc_cls
            self.cls_fcs = torch.nn.ModuleList()
            self.fc_cls = None

    _REF_FILE = 'knet_vis/tracker/kernel_update_head.py'
    _REF_CLASS = 'KernelUpdateHeadVideo'

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                torch.nn.init.xavier_uniform_(p)
        if self.with_cls and getattr(self.loss_cls, 'use_sigmoid', True):
            from .bricks import bias_init_with_prob
            torch.nn.init.constant_(self.fc_cls.bias, bias_init_with_prob(0.01))
        if self.kernel_init:
            torch.nn.init.normal_(self.fc_mask.weight, mean=0, std=0.01)

    def check_supported(self):
        super().check_supported()
        if self.with_cls and self.query_merge_method != 'mean':
            raise NotImplementedError("query_merge_method=%r: only 'mean' (the default every shipped config uses) is "
                                      'on the CUDA path' % self.query_merge_method)

    def _shape(self, B, N, H, W, x_dtype, w_dtype, frames_per_set=1):
        from .kernel_update_head import thr_logit
        return _lib.make_shape(B, N, self.in_channels, H, W, self.feedforward_channels, self._num_cls_out,
                               self.num_heads, x_dtype, w_dtype, self.with_ffn, self.engine,
                               thr_logit(self.hard_mask_thr), frames_per_set)

    @torch.no_grad()
    def forward(self, x, proposal_feat, mask_preds, prev_cls_score=None, mask_shape=None, img_metas=None, pos=None):
        self.check_supported()
        if x.dim() != 5:
            raise _lib.VknError('KernelUpdateHeadVideo expects x [B, F, C, H, W]')
        Bc, Fr, Cc, H, W = x.shape
        gathered = proposal_feat.dim() != 6                                  # :217-225
        if gathered and not self.with_cls:
            raise _lib.VknError('5-D proposal_feat (gathered queries) needs with_cls=True (reference asserts, :223)')
        if not gathered and self.with_cls:
            raise _lib.VknError('6-D proposal_feat (per-frame queries) needs with_cls=False (reference asserts, :219)')
        N = proposal_feat.shape[1] if gathered else proposal_feat.shape[2]
        if N != self.num_proposals:
            raise _lib.VknError('num_proposals mismatch: %d vs %d (reference asserts, :226)' % (N, self.num_proposals))
        if mask_preds.shape[-2:] != (H, W):
            raise NotImplementedError('mask_preds at a different resolution than x is not on the shipped path')
        if mask_shape is not None and mask_shape[0] != H:
            raise NotImplementedError('mask_shape resize raises in the reference as well (:365-366)')
        xf = x.reshape(Bc * Fr, Cc, H, W)
        mf = mask_preds.reshape(Bc * Fr, N, H, W)
        sets = Bc if gathered else Bc * Fr
        pf = proposal_feat.reshape(sets, N, Cc, -1)
        xf, pf, mf, _, _, _, _, xd = self._prepare(xf, pf, mf, frames_per_set=Fr if gathered else 1)
        w, _, wd = self.packed_weights(x.device)
        shape = self._shape(sets, N, H, W, xd, wd, Fr if gathered else 1)
        dev = x.device
        cls = torch.empty(sets, N, self._num_cls_out, dtype=torch.float32, device=dev) if self.with_cls else None
        new_mask = torch.empty(Bc * Fr, N, H, W, dtype=xf.dtype, device=dev)
        obj = torch.empty(sets, N, Cc, dtype=torch.float32, device=dev)
        ws, wsb = self._ws.get(shape, dev)
        _lib.check(_lib.lib().vkn_stage_forward(shape, w, _lib.ptr(xf), _lib.ptr(pf), _lib.ptr(mf), None,
                                                _lib.ptr(cls), _lib.ptr(new_mask), _lib.ptr(obj), None, ws, wsb,
                                                _lib.stream_ptr()))
        new_mask = new_mask.reshape(Bc, Fr, N, H, W)
        if gathered:
            return cls, new_mask, obj.reshape(Bc, N, Cc, 1, 1)
        return None, new_mask, obj.reshape(Bc, Fr, N, Cc, 1, 1)
