"""KernelUpdateHead -- drop-in for knet/det/kernel_update_head.py:16-277 (and the knet_vis copy,
knet_vis/det/kernel_update_head.py): same registry key, constructor kwargs, state_dict keys
 and forward contract.  One forward = one `vkn_stage_forward` call into
libvknet.so; inference only (no autograd graph is built), no PyTorch fallback.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, pack
from .bricks import Conv1x1Params, FFNParams, MultiheadAttentionParams, bias_init_with_prob, make_ln
from .registry import HEADS, build_loss, build_transformer_layer


def thr_logit(thr):
    """sigmoid(m) > thr  <=>  m > logit(thr); 0.0 for the default hard_mask_thr = 0.5
    (knet/det/kernel_update_head.py:190-192).  fp32 sigmoid rounds 0 < m < ~6e-8 to exactly 0.5, a
    measure-zero sliver where the two predicates differ (documented in DESIGN.md)."""
    if thr <= 0.0:
        return -float('inf')
    if thr >= 1.0:
        return float('inf')
    return math.log(thr / (1.0 - thr))


class _HeadBase(nn.Module):
    """Shared construction + the C-ABI plumbing of the det and video heads."""

    def _build_common(self, num_classes, num_ffn_fcs, num_heads, num_cls_fcs, num_mask_fcs, feedforward_channels,
                      in_channels, out_channels, dropout, mask_thr, act_cfg, ffn_act_cfg, conv_kernel_size,
                      feat_transform_cfg, hard_mask_thr, kernel_init, with_ffn, mask_out_stride, relative_coors,
                      relative_coors_off, feat_gather_stride, mask_transform_stride, mask_upsample_stride,
                      num_thing_classes, num_stuff_classes, mask_assign_stride, ignore_label, thing_label_in_seg,
                      kernel_updator_cfg, loss_rank, loss_mask, loss_dice, loss_cls):
        self.num_classes = num_classes
        self.loss_cls = build_loss(loss_cls)
        self.loss_mask = build_loss(loss_mask)
        self.loss_dice = build_loss(loss_dice)
        self.loss_rank = build_loss(loss_rank) if loss_rank is not None else None
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.mask_thr = mask_thr
        self.fp16_enabled = False
        self.dropout = dropout
        self.num_heads = num_heads
        self.hard_mask_thr = hard_mask_thr
        self.kernel_init = kernel_init
        self.with_ffn = with_ffn
        self.mask_out_stride = mask_out_stride
        self.relative_coors = relative_coors
        self.relative_coors_off = relative_coors_off
        self.conv_kernel_size = conv_kernel_size
        self.feat_gather_stride = feat_gather_stride
        self.mask_transform_stride = mask_transform_stride
        self.mask_upsample_stride = mask_upsample_stride
        self.num_thing_classes = num_thing_classes
        self.num_stuff_classes = num_stuff_classes
        self.mask_assign_stride = mask_assign_stride
        self.ignore_label = ignore_label
        self.thing_label_in_seg = thing_label_in_seg
        self.feedforward_channels = feedforward_channels
        if act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError('only ReLU is on the shipped path')
        E = in_channels * conv_kernel_size ** 2
        self.attention = MultiheadAttentionParams(E, num_heads, dropout)
        self.attention_norm = make_ln(dict(type='LN'), E)
        self.kernel_update_conv = build_transformer_layer(kernel_updator_cfg)
        if feat_transform_cfg is not None:
            cfg = dict(feat_transform_cfg)
            kernel_size = cfg.pop('kernel_size', 1)
            self.feat_transform = Conv1x1Params(in_channels, in_channels, kernel_size, stride=feat_gather_stride,
                                                padding=int(feat_gather_stride // 2), **cfg)
        else:
            self.feat_transform = None
        if self.with_ffn:
            self.ffn = FFNParams(in_channels, feedforward_channels, num_ffn_fcs, act_cfg=ffn_act_cfg, dropout=dropout)
            self.ffn_norm = make_ln(dict(type='LN'), in_channels)
        self.cls_fcs = nn.ModuleList()
        for _ in range(num_cls_fcs):
            self.cls_fcs.append(nn.Linear(in_channels, in_channels, bias=False))
            self.cls_fcs.append(make_ln(dict(type='LN'), in_channels))
            self.cls_fcs.append(nn.ReLU(inplace=True))
        use_sigmoid = getattr(self.loss_cls, 'use_sigmoid', True)
        self.fc_cls = nn.Linear(in_channels, self.num_classes if use_sigmoid else self.num_classes + 1)
        self.mask_fcs = nn.ModuleList()
        for _ in range(num_mask_fcs):
            self.mask_fcs.append(nn.Linear(in_channels, in_channels, bias=False))
            self.mask_fcs.append(make_ln(dict(type='LN'), in_channels))
            self.mask_fcs.append(nn.ReLU(inplace=True))
        self.fc_mask = nn.Linear(in_channels, out_channels)
        self.engine = _lib.ENGINE_AUTO
        self._ws = _lib.Workspace()
        self._packed = None
        self._packed_key = None

    def init_weights(self):
        """knet/det/kernel_update_head.py:151-168."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if getattr(self.loss_cls, 'use_sigmoid', True):
            nn.init.constant_(self.fc_cls.bias, bias_init_with_prob(0.01))
        if self.kernel_init:
            nn.init.normal_(self.fc_mask.weight, mean=0, std=0.01)

    # ---- C-ABI plumbing ---------------------------------------------------------------------------
    def check_supported(self):
        if self.conv_kernel_size != 1:
            raise NotImplementedError('conv_kernel_size != 1: every shipped config uses 1x1 dynamic kernels '
                                      '(configs/det/_base_/models/knet_kitti_step_s3_r50_fpn.py:3)')
        if self.out_channels != self.in_channels:
            raise NotImplementedError('out_channels must equal in_channels')
        if self.feat_gather_stride != 1:
            raise NotImplementedError('feat_gather_stride != 1 is not on the shipped path')
        self.kernel_update_conv.check_supported()

    def invalidate_weight_cache(self):
        self._packed = None
        self.__dict__['_param_list'] = None

    def _apply(self, fn, *args, **kwargs):          # .to() / .bfloat16() / .cuda(): parameters may be re-created
        self.__dict__['_param_list'] = None
        return super()._apply(fn, *args, **kwargs)

    def _weights(self):
        """The module's parameters as a list, walked once: `self.parameters()` traverses the module tree on every call (86 us
        for the 44 tensors of a stage -- more than the stage's device time at one frame per call).  Re-derived after `_apply`
        and `invalidate_weight_cache()`; code that swaps Parameter OBJECTS of a built head must call the latter."""
        pl = self.__dict__.get('_param_list')
        if pl is not None and next(self.parameters(), None) is not pl[0]:
            pl = None                              # a replica / re-built module: its first parameter is another object
        if pl is None:
            pl = list(self.parameters())
            self.__dict__['_param_list'] = pl
        return pl

    def _pack_extra(self, pk):
        return None

    def packed_weights(self, device):
        """(VknHeadW, extra, w_dtype) for `device`; rebuilt when any parameter changed in place."""
        # _version catches in-place edits through the parameter, data_ptr re-assignment of p.data (EMA / weight surgery);
        # edits through p.data.copy_() bump neither: call invalidate_weight_cache() after those
        key = (str(device),) + tuple((p._version, p.data_ptr(), p.dtype) for p in self._weights())
        if self._packed is None or self._packed_key != key:
            wd = pack.weight_dtype_of(self._weights())
            pk = pack.Packer(device, wd)
            w = pack.pack_head(pk, self)
            pack.attach_frame_chain_pack(pk, w, self._shape(1, 1, 1, 1, _lib.VKN_BF16, wd))
            self._packed = (w, self._pack_extra(pk), wd, pk)
            self._packed_key = key
            # the derived operands (ft_wt_ext, fc_pack, dtype copies) were produced on the CURRENT stream; a caller that drives
            # several streams may use them from another one next: finish them once, here (weights change rarely)
            if torch.device(device).type == 'cuda' and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(device).synchronize()
        return self._packed[0], self._packed[1], self._packed[2]

    def _shape(self, B, N, H, W, x_dtype, w_dtype):
        return _lib.make_shape(B, N, self.in_channels, H, W, self.feedforward_channels, self.fc_cls.out_features,
                               self.num_heads, x_dtype, w_dtype, self.with_ffn, self.engine,
                               thr_logit(self.hard_mask_thr))

    def _prepare(self, x, proposal_feat, mask_preds, frames_per_set=1):
        self.check_supported()
        if not x.is_cuda:
            raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
        self._refuse_autograd(x, proposal_feat, mask_preds)
        B, N = proposal_feat.shape[:2]
        Cc, H, W = x.shape[-3:]
        if Cc != self.in_channels:
            raise _lib.VknError('x has %d channels, head expects %d' % (Cc, self.in_channels))
        # the C ABI receives raw pointers + (B, N, H, W): a mismatched batch / kernel count would be an out-of-bounds device
        # read where the reference's einsum raises -- check every extent here
        if x.dim() != 4 or x.shape[0] != B * frames_per_set:
            raise _lib.VknError('x is %s but proposal_feat describes %d kernel set(s) x %d frame(s)' % (
                tuple(x.shape), B, frames_per_set))
        if proposal_feat.numel() != B * N * self.in_channels * self.conv_kernel_size ** 2:
            raise _lib.VknError('proposal_feat %s does not hold [B=%d, N=%d, C=%d] kernels' % (
                tuple(proposal_feat.shape), B, N, self.in_channels))
        if mask_preds is not None and (mask_preds.dim() != 4 or mask_preds.shape[0] != B * frames_per_set or
                                       mask_preds.shape[1] != N):
            raise _lib.VknError('mask_preds is %s, expected [%d, %d, h, w]' % (tuple(mask_preds.shape), B * frames_per_set, N))
        x = x.contiguous()
        xd = _lib.dtype_code(x.dtype)
        if mask_preds is not None:
            if mask_preds.shape[-2:] != (H, W):   # knet/det/kernel_update_head.py:183-188 (not hit by shipped configs)
                mask_preds = F.interpolate(mask_preds.float(), (H, W), align_corners=False, mode='bilinear')
            mask_preds = mask_preds.to(x.dtype).contiguous()
        pf = proposal_feat.reshape(B, N, self.in_channels, -1)
        if pf.shape[-1] != 1:
            raise NotImplementedError('conv_kernel_size != 1 is not on the shipped path')
        pf = pf.reshape(B, N, self.in_channels).to(torch.float32).contiguous()
        return x, pf, mask_preds, B, N, H, W, xd

    def _post_masks(self, new_mask_preds, mask_shape, H):
        if self.mask_transform_stride == 2:       # :230-236, 261-266 -- never set by the shipped configs
            raise NotImplementedError('mask_transform_stride == 2 is not on the shipped path')
        if mask_shape is not None and mask_shape[0] != H:   # :268-273
            new_mask_preds = F.interpolate(new_mask_preds.float(), mask_shape, align_corners=False,
                                           mode='bilinear').to(new_mask_preds.dtype)
        return new_mask_preds

    # helpers the callers of the reference use -----------------------------------------------------------------------
    def rescale_masks(self, masks_per_img, img_meta):
        """knet/det/kernel_update_head.py:443-458 -- one fused launch (vkn_rescale_masks) instead of three full-resolution
        temporaries.  `masks_per_img` are the (already x mask_upsample_stride) scaled_mask_preds of the selected kernels."""
        from . import ops
        return ops.rescale_masks(masks_per_img, img_meta, 1, None)[0]

    def get_seg_masks(self, masks_per_img, labels_per_img, scores_per_img, test_cfg, img_meta):
        """:460-467: rescale + threshold on the device, then the reference's list packing (segm2result)."""
        from . import ops
        thr = test_cfg['mask_thr'] if isinstance(test_cfg, dict) else test_cfg.mask_thr
        seg_masks = ops.rescale_masks(masks_per_img, img_meta, 1, thr, probs=False)[1]
        return self.segm2result(seg_masks, labels_per_img, scores_per_img)

    def segm2result(self, mask_preds, det_labels, cls_scores):
        """Result packing of knet/det/kernel_update_head.py:469-483: per class, the score-only pseudo boxes [n_c, 5] and the
        list of that class's masks (in input order)."""
        import numpy as np
        masks = mask_preds.cpu().numpy()
        labels = det_labels.cpu().numpy()
        boxes = np.zeros((labels.shape[0], 5), dtype=np.float32)
        boxes[:, 4] = cls_scores.cpu().numpy()
        per_class = [np.flatnonzero(labels == c) for c in range(self.num_classes)]
        return [boxes[idx] for idx in per_class], [[masks[i] for i in idx] for idx in per_class]

    @staticmethod
    def _refuse_autograd(*tensors):
        """The CUDA path is forward only.  The reference's forward is differentiable, so silently detaching would train a
        frozen head: refuse instead."""
        if torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
            raise NotImplementedError('vknet heads are inference-only: an input requires grad and autograd is enabled; wrap '
                                      'the call in torch.no_grad() (or use the reference modules for training)')

    # ---- result packing helpers of the knet_vis heads (knet_vis/det/kernel_update_head.py:484-500,
    #      knet_vis/tracker/kernel_update_head.py:581-600) -------------------------------------------------------------
    def get_seg_masks_tracking(self, masks_per_img, labels_per_img, scores_per_img, ids_per_img, test_cfg, img_meta):
        """rescale + threshold on the device (one launch), then mmtrack's outs2results packing: per class the rows
        [id, 0, 0, 0, 0, score] of the tracked instances (ids > -1) and the list of their masks."""
        from . import ops
        thr = test_cfg['mask_thr'] if isinstance(test_cfg, dict) else test_cfg.mask_thr
        seg_masks = ops.rescale_masks(masks_per_img, img_meta, 1, thr, probs=False)[1]
        return self.tracks2result(seg_masks, labels_per_img, scores_per_img, ids_per_img)

    def tracks2result(self, seg_masks, labels, scores, ids):
        """mmtrack.transform.outs2results (mmtrack/transform.py:6-74) for the call the reference makes: fake boxes
        [0,0,0,0,score], instances with id > -1 only."""
        import numpy as np
        ids = ids.detach().cpu().numpy()
        valid = ids > -1
        ids = ids[valid]
        labels = labels.detach().cpu().numpy()[valid]
        scores = scores.detach().cpu().numpy().astype(np.float32)[valid]
        masks = seg_masks.detach().cpu().numpy()[valid]
        boxes = np.zeros((ids.shape[0], 5), dtype=np.float32)
        boxes[:, 4] = scores
        if ids.shape[0] == 0:
            bbox_results = [np.zeros((0, 6), dtype=np.float32) for _ in range(self.num_classes)]
        else:
            bbox_results = [np.concatenate((ids[labels == c, None], boxes[labels == c, :]), axis=1)
                            for c in range(self.num_classes)]
        mask_results = [[] for _ in range(self.num_classes)]
        for i in range(ids.shape[0]):
            mask_results[labels[i]].append(masks[i])
        return bbox_results, mask_results

    # ---- training-side methods: pass-throughs to the reference's own pure-torch code -----------------
    _REF_FILE = 'knet/det/kernel_update_head.py'
    _REF_CLASS = 'KernelUpdateHead'

    def _reference_method(self, name):
        from ._refpass import reference_function
        return reference_function(self._REF_FILE, self._REF_CLASS, name)

    def loss(self, *args, **kwargs):
        """knet/det/kernel_update_head.py:279-... unchanged: the reference's own function, bound to this module (same
        attribute names).  Needs the reference tree and mmdet importable; raises NotImplementedError otherwise."""
        return self._reference_method('loss')(self, *args, **kwargs)

    def get_targets(self, *args, **kwargs):
        return self._reference_method('get_targets')(self, *args, **kwargs)

    def _get_target_single(self, *args, **kwargs):
        return self._reference_method('_get_target_single')(self, *args, **kwargs)


@HEADS.register_module(force=True)
class KernelUpdateHead(_HeadBase):

    def __init__(self, num_classes=80, num_ffn_fcs=2, num_heads=8, num_cls_fcs=1, num_mask_fcs=3,
                 feedforward_channels=2048, in_channels=256, out_channels=256, dropout=0.0, mask_thr=0.5,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_act_cfg=dict(type='ReLU', inplace=True),
                 conv_kernel_size=3, feat_transform_cfg=None, hard_mask_thr=0.5, kernel_init=False,
                 with_ffn=True, mask_out_stride=4, relative_coors=False, relative_coors_off=False,
                 feat_gather_stride=1, mask_transform_stride=1, mask_upsample_stride=1, num_thing_classes=80,
                 num_stuff_classes=53, mask_assign_stride=4, ignore_label=255, thing_label_in_seg=0,
                 kernel_updator_cfg=dict(type='DynamicConv', in_channels=256, feat_channels=64, out_channels=256,
                                         input_feat_shape=1, act_cfg=dict(type='ReLU', inplace=True),
                                         norm_cfg=dict(type='LN')),
                 loss_rank=None, loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0),
                 loss_dice=dict(type='DiceLoss', loss_weight=3.0),
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                 ffn_drop=None):
        super().__init__()
        self._build_common(num_classes, num_ffn_fcs, num_heads, num_cls_fcs, num_mask_fcs, feedforward_channels,
                           in_channels, out_channels, dropout, mask_thr, act_cfg, ffn_act_cfg, conv_kernel_size,
                           feat_transform_cfg, hard_mask_thr, kernel_init, with_ffn, mask_out_stride,
                           relative_coors, relative_coors_off, feat_gather_stride, mask_transform_stride,
                           mask_upsample_stride, num_thing_classes, num_stuff_classes, mask_assign_stride,
                           ignore_label, thing_label_in_seg, kernel_updator_cfg, loss_rank, loss_mask, loss_dice,
                           loss_cls)

    @torch.no_grad()
    def forward(self, x, proposal_feat, mask_preds, prev_cls_score=None, mask_shape=None, img_metas=None):
        """-> (cls_score [B,N,ncls] fp32, new_mask_preds [B,N,H,W] x.dtype, obj_feat [B,N,C,1,1] fp32)
        (knet/det/kernel_update_head.py:170-176, 275-277)."""
        x, pf, mask_preds, B, N, H, W, xd = self._prepare(x, proposal_feat, mask_preds)
        w, _, wd = self.packed_weights(x.device)
        shape = self._shape(B, N, H, W, xd, wd)
        dev = x.device
        cls = torch.empty(B, N, self.fc_cls.out_features, dtype=torch.float32, device=dev)
        new_mask = torch.empty(B, N, H, W, dtype=x.dtype, device=dev)
        obj = torch.empty(B, N, self.in_channels, dtype=torch.float32, device=dev)
        ws, wsb = self._ws.get(shape, dev)
        _lib.check(_lib.lib().vkn_stage_forward(shape, w, _lib.ptr(x), _lib.ptr(pf), _lib.ptr(mask_preds), None,
                                                _lib.ptr(cls), _lib.ptr(new_mask), _lib.ptr(obj), None, ws, wsb,
                                                _lib.stream_ptr()))
        new_mask = self._post_masks(new_mask, mask_shape, H)
        return cls, new_mask, obj.reshape(B, N, self.in_channels, 1, 1)
