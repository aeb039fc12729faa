"""VideoKernelUpdateHead -- drop-in for knet/video/kernel_update_head.py:17-541: the stage of
KernelUpdateHead plus the cross-frame link blocks the shipped configs select
  previous_type='ffn'                  (:394-415)  tracking-kernel link after the update
  previous_type='update'               (:417-444)  same, previous kernels first re-updated
  previous_link='update_dynamic_cov'   (:324-348)  kernel fusion before the update
('link_atten' and 'update_obj' are selected by no shipped config and raise NotImplementedError).
Same registry key, constructor kwargs, state_dict keys and 5-tuple return.
"""
import torch

from . import _lib, pack
from .bricks import FFNParams, MultiheadAttentionParams, make_ln
from .kernel_update_head import _HeadBase
from .registry import HEADS, build_transformer_layer


@HEADS.register_module(force=True)
class VideoKernelUpdateHead(_HeadBase):

    def __init__(self, num_classes=80, num_ffn_fcs=2, num_heads=8, num_cls_fcs=1, num_mask_fcs=3,
                 feedforward_channels=2048, in_channels=256, out_channels=256, dropout=0.0, mask_thr=0.5,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_act_cfg=dict(type='ReLU', inplace=True),
                 conv_kernel_size=3, feat_transform_cfg=None, hard_mask_thr=0.5, kernel_init=False,
                 with_ffn=True, mask_out_stride=4, relative_coors=False, relative_coors_off=False,
                 feat_gather_stride=1, mask_transform_stride=1, mask_upsample_stride=1, num_thing_classes=80,
                 num_stuff_classes=53, mask_assign_stride=4, ignore_label=255, thing_label_in_seg=0,
                 previous=None, previous_x_feat=None, previous_link=None, previous_type=None,
                 previous_detach=False, previous_detach_link=False, previous_link_detach=False,
                 kernel_updator_cfg=dict(type='DynamicConv', in_channels=256, feat_channels=64, out_channels=256,
                                         input_feat_shape=1, act_cfg=dict(type='ReLU', inplace=True),
                                         norm_cfg=dict(type='LN')),
                 loss_rank=None, loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0),
                 loss_dice=dict(type='DiceLoss', loss_weight=3.0),
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0)):
        super().__init__()
        self._build_common(num_classes, num_ffn_fcs, num_heads, num_cls_fcs, num_mask_fcs, feedforward_channels,
                           in_channels, out_channels, dropout, mask_thr, act_cfg, ffn_act_cfg, conv_kernel_size,
                           feat_transform_cfg, hard_mask_thr, kernel_init, with_ffn, mask_out_stride,
                           relative_coors, relative_coors_off, feat_gather_stride, mask_transform_stride,
                           mask_upsample_stride, num_thing_classes, num_stuff_classes, mask_assign_stride,
                           ignore_label, thing_label_in_seg, kernel_updator_cfg, loss_rank, loss_mask, loss_dice,
                           loss_cls)
        self.previous = previous
        self.previous_type = previous_type
        self.previous_link = previous_link
        self.previous_x_feat = previous_x_feat
        self.previous_detach = previous_detach
        self.previous_detach_link = previous_detach_link
        self.previous_link_detach = previous_link_detach
        if self.previous is not None:                      # knet/video/kernel_update_head.py:167-260
            E = in_channels * conv_kernel_size ** 2

            def link_ffn():
                return FFNParams(in_channels, feedforward_channels, num_ffn_fcs, act_cfg=ffn_act_cfg, dropout=dropout)

            if previous_type == 'ffn':
                self.attention_previous = MultiheadAttentionParams(E, 8, 0.0)
                self.attention_previous_norm = make_ln(dict(type='LN'), E)
                self.link_ffn = link_ffn()
                self.link_ffn_norm = make_ln(dict(type='LN'), in_channels)
            elif previous_type in ('update', 'update_obj'):
                self.attention_previous_update_track = build_transformer_layer(kernel_updator_cfg)
                self.attention_previous_track = MultiheadAttentionParams(E, 8, 0.0)
                self.attention_previous_norm_track = make_ln(dict(type='LN'), E)
                self.link_ffn_track = link_ffn()
                self.link_ffn_norm_track = make_ln(dict(type='LN'), in_channels)
            if previous_link == 'update_dynamic_cov':
                self.attention_previous_update_link = build_transformer_layer(kernel_updator_cfg)
                self.attention_previous_link = MultiheadAttentionParams(E, 8, 0.0)
                self.attention_previous_norm_link = make_ln(dict(type='LN'), E)
                self.link_ffn_link = link_ffn()
                self.link_ffn_norm_link = make_ln(dict(type='LN'), in_channels)
            elif previous_link == 'link_atten':
                self.attention_previous_link = MultiheadAttentionParams(E, 8, 0.0)
                self.attention_previous_norm_link = make_ln(dict(type='LN'), E)
                self.link_ffn_link = link_ffn()
                self.link_ffn_norm_link = make_ln(dict(type='LN'), in_channels)

    _REF_FILE = 'knet/video/kernel_update_head.py'
    _REF_CLASS = 'VideoKernelUpdateHead'

    # ---- result helpers with the VIDEO head's contract (knet/video/kernel_update_head.py:725-748): 3-tuples, boxes from
    #      the masks; called by VideoKernelIterHead.get_panoptic / simple_test (knet/video/kernel_iter_head.py:605-609) ----
    def get_seg_masks(self, masks_per_img, labels_per_img, scores_per_img, test_cfg, img_meta):
        from . import ops
        thr = test_cfg['mask_thr'] if isinstance(test_cfg, dict) else test_cfg.mask_thr
        seg_masks = ops.rescale_masks(masks_per_img, img_meta, 1, thr, probs=False)[1]
        return self.segm2result(seg_masks, labels_per_img, scores_per_img)

    def segm2result(self, mask_preds, det_labels, cls_scores):
        """-> (bboxes [n,5] float32 ndarray: extent of each mask's non-zero pixels clipped at 0 + score,
               segm_result: per class the list of that class's mask tensors (input order), mask_preds)   (:734-748)"""
        import numpy as np
        from . import ops
        labels = det_labels.detach().cpu().numpy()
        n = mask_preds.shape[0]
        bboxes = np.zeros((n, 5), dtype=np.float32)
        bboxes[:, 4] = cls_scores.detach().cpu().numpy()
        if n:
            boxes = ops.mask_boxes(mask_preds) if mask_preds.is_cuda else self.mask_boxes_torch(mask_preds)
            bboxes[:, :4] = boxes.cpu().numpy().clip(min=0)
        segm_result = [[] for _ in range(self.num_classes)]
        for idx in range(n):
            segm_result[labels[idx]].append(mask_preds[idx])
        return bboxes, segm_result, mask_preds

    @staticmethod
    def mask_boxes_torch(masks):
        """torch formulation of the mask -> box reduction (host-side packing of CPU tensors; the CUDA kernel's checker)"""
        nz = masks != 0
        K, H, W = nz.shape
        rows, cols = nz.any(2), nz.any(1)                                  # [K,H], [K,W]
        ys = torch.arange(H, device=masks.device).expand(K, H)
        xs = torch.arange(W, device=masks.device).expand(K, W)
        big = max(H, W) + 1
        y0 = torch.where(rows, ys, torch.full_like(ys, big)).min(1).values
        y1 = torch.where(rows, ys, torch.full_like(ys, -1)).max(1).values
        x0 = torch.where(cols, xs, torch.full_like(xs, big)).min(1).values
        x1 = torch.where(cols, xs, torch.full_like(xs, -1)).max(1).values
        out = torch.stack([x0, y0, x1, y1], 1).float()
        empty = ~rows.any(1)
        out[empty] = torch.tensor([-1.0, -1.0, 10.0, 10.0], device=masks.device)
        return out

    def check_supported(self):
        super().check_supported()
        if self.previous is not None:
            if self.previous_type == 'update_obj' or self.previous_link == 'link_atten':
                raise NotImplementedError("previous_type='update_obj' / previous_link='link_atten' are selected by "
                                          'no shipped config and are outside the CUDA path')
            if self.num_heads != 8:
                raise NotImplementedError('link attention uses 8 heads (knet/video/kernel_update_head.py:170)')

    def _pack_extra(self, pk):
        links = {}
        if self.previous is None:
            return links
        if self.previous_type == 'ffn':
            links['track'] = pack.pack_link(pk, None, self.attention_previous, self.attention_previous_norm,
                                            self.link_ffn, self.link_ffn_norm)
        elif self.previous_type == 'update':
            links['track'] = pack.pack_link(pk, self.attention_previous_update_track, self.attention_previous_track,
                                            self.attention_previous_norm_track, self.link_ffn_track,
                                            self.link_ffn_norm_track)
        if self.previous_link == 'update_dynamic_cov':
            links['link'] = pack.pack_link(pk, self.attention_previous_update_link, self.attention_previous_link,
                                           self.attention_previous_norm_link, self.link_ffn_link,
                                           self.link_ffn_norm_link)
        return links

    def _link(self, shape, lw, cur, prev, x_feat, ws, wsb):
        out = torch.empty_like(cur)
        _lib.check(_lib.lib().vkn_link_attend(shape, lw, _lib.ptr(cur), _lib.ptr(prev), _lib.ptr(x_feat),
                                              _lib.ptr(out), ws, wsb, _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def forward(self, x, proposal_feat, mask_preds, prev_cls_score=None, mask_shape=None, img_metas=None,
                previous_obj_feats=None, previous_mask_preds=None, previous_x_feats=None):
        """-> (cls_score, new_mask_preds, obj_feat [B,N,C,1,1], x_feat [B,N,C], obj_feat_track | None)
        (knet/video/kernel_update_head.py:281-291, 534-541)."""
        x, pf, mask_preds, B, N, H, W, xd = self._prepare(x, proposal_feat, mask_preds)
        w, links, wd = self.packed_weights(x.device)
        shape = self._shape(B, N, H, W, xd, wd)
        dev, Cc = x.device, self.in_channels
        L = _lib.lib()
        ws, wsb = self._ws.get(shape, dev)
        st = _lib.stream_ptr()
        x_feat = torch.empty(B, N, Cc, dtype=torch.float32, device=dev)
        _lib.check(L.vkn_mask_pool(shape, w, _lib.ptr(x), _lib.ptr(mask_preds), _lib.ptr(x_feat), ws, wsb, st))
        prev = None
        if previous_obj_feats is not None:
            self._refuse_autograd(previous_obj_feats)
            if previous_obj_feats.numel() != B * N * Cc:
                raise _lib.VknError('previous_obj_feats %s does not hold [B=%d, N=%d, C=%d] kernels' % (
                    tuple(previous_obj_feats.shape), B, N, Cc))
            prev = previous_obj_feats.reshape(B, N, Cc).to(torch.float32).contiguous()
        if prev is not None and 'link' in links:                                    # :324-348
            pf = self._link(shape, links['link'], pf, prev, x_feat, ws, wsb)
        cls = torch.empty(B, N, self.fc_cls.out_features, dtype=torch.float32, device=dev)
        new_mask = torch.empty(B, N, H, W, dtype=x.dtype, device=dev)
        obj = torch.empty(B, N, Cc, dtype=torch.float32, device=dev)
        _lib.check(L.vkn_stage_forward(shape, w, _lib.ptr(x), _lib.ptr(pf), None, _lib.ptr(x_feat), _lib.ptr(cls),
                                       _lib.ptr(new_mask), _lib.ptr(obj), None, ws, wsb, st))
        track = None
        if prev is not None and 'track' in links:                                   # :394-444
            track = self._link(shape, links['track'], obj, prev, x_feat, ws, wsb).reshape(B, N, Cc, 1, 1)
        new_mask = self._post_masks(new_mask, mask_shape, H)
        return cls, new_mask, obj.reshape(B, N, Cc, 1, 1), x_feat, track
