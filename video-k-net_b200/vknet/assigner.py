"""Drop-in for the reference's training-side matcher: `MaskHungarianAssigner`, `DiceCost`, `MaskCost`
(knet/det/mask_hungarian_assigner.py) with the cost matrix computed by ONE fused CUDA pass (ops.match_cost -> vkn_match_cost)
instead of two full-size activations and three einsums; the Hungarian solve stays scipy's on the host, as in the reference
(:246-251).  Same constructor kwargs, same `assign` signature and result fields.  Registered into mmdet's BBOX_ASSIGNERS /
MATCH_COST registries when mmdet is importable.
"""
import numpy as np
import torch

from . import _lib, ops

try:
    from scipy.optimize import linear_sum_assignment
except ImportError:                                      # pragma: no cover
    linear_sum_assignment = None

try:                                                     # the real result class / registries when mmdet is installed
    from mmdet.core import AssignResult                  # type: ignore
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS   # type: ignore
    from mmdet.core.bbox.match_costs.builder import MATCH_COST  # type: ignore
    HAVE_MMDET = True
except Exception:  # noqa: BLE001
    HAVE_MMDET = False

    class AssignResult:                                  # the fields the reference's callers read (mmdet v2.18)
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


def _supported(cost, what):
    if not cost.pred_act or cost.act_mode != 'sigmoid':
        raise NotImplementedError('%s: only pred_act=True, act_mode="sigmoid" (what every shipped config sets, e.g. '
                                  'configs/det/_base_/models/knet_s3_r50_fpn.py:116-118) is on the CUDA path' % what)


class DiceCost:
    """knet/det/mask_hungarian_assigner.py:14-75."""

    def __init__(self, weight=1., pred_act=False, act_mode='sigmoid', eps=1e-3):
        self.weight, self.pred_act, self.act_mode, self.eps = weight, pred_act, act_mode, eps

    def __call__(self, mask_preds, gt_masks):
        _supported(self, 'DiceCost')
        return ops.match_cost(mask_preds, None, gt_masks, None, w_cls=0.0, w_mask=0.0, w_dice=float(self.weight), dice_eps=self.eps)


class MaskCost:
    """knet/det/mask_hungarian_assigner.py:78-110."""

    def __init__(self, weight=1., pred_act=False, act_mode='sigmoid'):
        self.weight, self.pred_act, self.act_mode = weight, pred_act, act_mode

    def __call__(self, cls_pred, target):
        _supported(self, 'MaskCost')
        return ops.match_cost(cls_pred, None, target, None, w_cls=0.0, w_mask=float(self.weight), w_dice=0.0)


class FocalLossCost:
    """mmdet.core.bbox.match_costs.FocalLossCost (v2.18) -- only its parameters; the arithmetic runs inside vkn_match_cost."""

    def __init__(self, weight=1., alpha=0.25, gamma=2, eps=1e-12):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps


_COSTS = {'DiceCost': DiceCost, 'MaskCost': MaskCost, 'FocalLossCost': FocalLossCost}


def _build_cost(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ not in _COSTS:
        raise NotImplementedError('match cost %r is not on the CUDA path (shipped configs use FocalLossCost / MaskCost / DiceCost)' % typ)
    return _COSTS[typ](**cfg)


class MaskHungarianAssigner:
    """knet/det/mask_hungarian_assigner.py:113-266.  `assign(bbox_pred=mask logits [N,H,W], cls_pred [N,ncls], gt_bboxes=gt masks
    [M,H,W], gt_labels [M])` -> AssignResult(num_gts, gt_inds (0 = background, k = gt k-1), None, labels)."""

    def __init__(self, cls_cost=dict(type='FocalLossCost', weight=1.), mask_cost=dict(type='MaskCost', weight=1.0, pred_act=True),
                 dice_cost=dict(type='DiceCost', weight=1.0, pred_act=True), boundary_cost=None, topk=1):
        self.cls_cost, self.mask_cost, self.dice_cost = _build_cost(cls_cost), _build_cost(mask_cost), _build_cost(dice_cost)
        if boundary_cost is not None:
            raise NotImplementedError('boundary_cost: no shipped config sets it')
        self.boundary_cost = None
        self.topk = topk
        if not isinstance(self.cls_cost, FocalLossCost) or not isinstance(self.mask_cost, MaskCost) or \
                not isinstance(self.dice_cost, DiceCost):
            raise NotImplementedError('cls_cost / mask_cost / dice_cost must be FocalLossCost / MaskCost / DiceCost')

    def cost_matrix(self, bbox_pred, cls_pred, gt_bboxes, gt_labels):
        if self.mask_cost.weight != 0:
            _supported(self.mask_cost, 'MaskCost')
        if self.dice_cost.weight != 0:
            _supported(self.dice_cost, 'DiceCost')
        c = self.cls_cost
        return ops.match_cost(bbox_pred, cls_pred if c.weight != 0 else None, gt_bboxes, gt_labels, w_cls=float(c.weight),
                              w_mask=float(self.mask_cost.weight), w_dice=float(self.dice_cost.weight), dice_eps=self.dice_cost.eps,
                              focal_alpha=c.alpha, focal_gamma=c.gamma, focal_eps=c.eps)

    @torch.no_grad()
    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta=None, gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, 'Only case when gt_bboxes_ignore is None is supported.'
        num_gts, num_bboxes = gt_bboxes.size(0), bbox_pred.size(0)
        assigned_gt_inds = bbox_pred.new_full((num_bboxes,), -1, dtype=torch.long)
        assigned_labels = bbox_pred.new_full((num_bboxes,), -1, dtype=torch.long)
        if num_gts == 0 or num_bboxes == 0:                                   # :218-224
            if num_gts == 0:
                assigned_gt_inds[:] = 0
            return AssignResult(num_gts, assigned_gt_inds, None, labels=assigned_labels)
        cost = self.cost_matrix(bbox_pred, cls_pred, gt_bboxes, gt_labels).cpu()      # :228-247
        if linear_sum_assignment is None:
            raise ImportError('Please run "pip install scipy" to install scipy first.')
        if self.topk == 1:                                                    # :251-264
            rows, cols = linear_sum_assignment(cost)
        else:
            rr, cc = [], []
            for _ in range(self.topk):
                r, c = linear_sum_assignment(cost)
                rr.append(r)
                cc.append(c)
                cost[r] = 1e10
            rows, cols = np.concatenate(rr), np.concatenate(cc)
        rows = torch.from_numpy(rows).to(bbox_pred.device)
        cols = torch.from_numpy(cols).to(bbox_pred.device)
        assigned_gt_inds[:] = 0                                               # :271-276
        assigned_gt_inds[rows] = cols + 1
        assigned_labels[rows] = gt_labels.to(bbox_pred.device)[cols]
        return AssignResult(num_gts, assigned_gt_inds, None, labels=assigned_labels)


if HAVE_MMDET:                                           # pragma: no cover  (mmdet is not installable in the build image)
    BBOX_ASSIGNERS.register_module(name='MaskHungarianAssigner', force=True, module=MaskHungarianAssigner)
    MATCH_COST.register_module(name='DiceCost', force=True, module=DiceCost)
    MATCH_COST.register_module(name='MaskCost', force=True, module=MaskCost)
