"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

mmcv / mmdet stand-ins that let the reference's hot-path files import UNMODIFIED
from /root/reference (they `import mmcv`, `import mmdet` at module top;
neither wheel exists in this image and there is no network).

Used only by oracle/make_golden.py and by the CPU tests that cross-check the
restatement in oracle/knet_oracle.py against the live reference.  The GPU box
has no /root/reference, so nothing under `-m gpu`, smoke() or bench.py touches
this file.

What is restated here is third-party behaviour:
  * mmcv.cnn.ConvModule(norm_cfg=None, act_cfg=None)  -> nn.Conv2d(bias=True) as `.conv`
  * mmcv.cnn.build_norm_layer(dict(type='LN'), n)    -> ('ln', nn.LayerNorm(n))
  * mmcv.cnn.build_activation_layer(dict(type='ReLU', inplace=True))
  * mmcv.cnn.bias_init_with_prob(p) = -log((1-p)/p)
  * mmcv.cnn.bricks.transformer.FFN / MultiheadAttention / registries
  * mmdet registries (HEADS), build_loss (only `.use_sigmoid` is read on the path)
Reference call sites: knet/det/kernel_update_head.py:100-126,
knet/kernel_updator.py:3-4, knet/video/kernel_update_head.py:108-260.
"""
import importlib
import math
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = '/root/reference'


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg, **default):
        cfg = dict(cfg)
        typ = cfg.pop('type')
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        for k, v in default.items():
            cfg.setdefault(k, v)
        return cls(**cfg)


class ConvModule(nn.Module):
    """mmcv ConvModule restricted to what the hot path uses (no norm, no act)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), bias='auto', **kw):
        super().__init__()
        assert norm_cfg is None, 'shim: only the norm-free ConvModule is on the hot path'
        assert act_cfg is None, 'shim: only the activation-free ConvModule is on the hot path'
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, bias=True)

    def forward(self, x):
        return self.conv(x)


def build_norm_layer(cfg, num_features, postfix=''):
    assert cfg['type'] == 'LN'
    return 'ln' + str(postfix), nn.LayerNorm(num_features)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    assert typ == 'ReLU'
    return nn.ReLU(**cfg)


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


class FFN(nn.Module):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        if 'dropout' in kwargs:  # deprecated alias kept by mmcv
            ffn_drop = kwargs['dropout']
        self.embed_dims = embed_dims
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels),
                                        build_activation_layer(act_cfg), nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return out
        if identity is None:
            identity = x
        return identity + out


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_pos=None, attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class _FakeLoss(nn.Module):
    def __init__(self, use_sigmoid=False, **kw):
        super().__init__()
        self.use_sigmoid = use_sigmoid

    def forward(self, *a, **k):
        raise NotImplementedError('losses are outside the hot path')


def _mod(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Put the stand-ins into sys.modules and /root/reference on sys.path."""
    global _installed
    if _installed:
        return
    names = ['mmcv', 'mmcv.cnn', 'mmcv.cnn.bricks', 'mmcv.cnn.bricks.transformer', 'mmcv.runner',
             'mmdet', 'mmdet.core', 'mmdet.models', 'mmdet.models.builder',
             'mmdet.models.dense_heads', 'mmdet.models.dense_heads.atss_head',
             'mmdet.models.losses', 'mmdet.utils', 'unitrack', 'unitrack.mask', 'mmtrack', 'mmtrack.transform']
    mods = {n: _mod(n) for n in names}
    for n, m in mods.items():
        if '.' in n:
            parent, child = n.rsplit('.', 1)
            setattr(mods[parent], child, m)
    cnn = mods['mmcv.cnn']
    cnn.ConvModule = ConvModule
    cnn.bias_init_with_prob = bias_init_with_prob
    cnn.build_activation_layer = build_activation_layer
    cnn.build_norm_layer = build_norm_layer
    cnn.normal_init = lambda module, mean=0, std=1, bias=0: (nn.init.normal_(module.weight, mean, std),
                                                            module.bias is not None and nn.init.constant_(module.bias, bias))
    tr = mods['mmcv.cnn.bricks.transformer']
    tr.TRANSFORMER_LAYER = Registry('transformer_layer')
    tr.FFN = FFN
    tr.MultiheadAttention = MultiheadAttention
    tr.build_transformer_layer = lambda cfg, default_args=None: tr.TRANSFORMER_LAYER.build(cfg)
    mods['mmcv.runner'].force_fp32 = lambda *a, **k: (lambda f: f)
    mods['mmcv.runner'].auto_fp16 = lambda *a, **k: (lambda f: f)
    core = mods['mmdet.core']
    core.multi_apply = lambda func, *args, **kw: tuple(map(list, zip(*map(
        lambda *a: func(*a, **kw), *args))))
    core.bbox2result = lambda *a, **k: None
    core.mask_matrix_nms = lambda *a, **k: None
    core.build_assigner = lambda *a, **k: None
    core.build_sampler = lambda *a, **k: None
    core.reduce_mean = lambda t: t
    bld = mods['mmdet.models.builder']
    bld.HEADS = Registry('head')
    bld.build_loss = lambda cfg: _FakeLoss(**{k: v for k, v in cfg.items() if k == 'use_sigmoid'})
    bld.build_head = lambda cfg: bld.HEADS.build(cfg)

    class _PassthroughNeck(nn.Module):      # stands in for SemanticFPNWrapper: hands the given feature maps through
        def forward(self, feats):
            return list(feats)
    bld.build_neck = lambda cfg: _PassthroughNeck()
    mods['mmdet.models.dense_heads.atss_head'].reduce_mean = lambda t: t
    mods['mmdet.models.losses'].accuracy = lambda *a, **k: None
    mods['mmdet.utils'].get_root_logger = lambda *a, **k: __import__('logging').getLogger('ref')
    mods['unitrack.mask'].tensor_mask2box = lambda *a, **k: None
    mods['unitrack.mask'].mask2box = lambda *a, **k: None
    mods['mmtrack.transform'].outs2results = lambda *a, **k: None
    _installed = True
    # The result-packing helpers of the video / VIS heads call two small pure-python utilities that ship INSIDE the reference
    # tree (unitrack/utils/mask.py: tensor_mask2box; mmtrack/transform.py: outs2results).  Load the real ones by path when
    # their own imports can be satisfied (pycocotools is not installed here and is not used by these two functions).
    try:
        if 'pycocotools' not in sys.modules:
            pc = types.ModuleType('pycocotools')
            pc.mask = types.ModuleType('pycocotools.mask')
            sys.modules['pycocotools'], sys.modules['pycocotools.mask'] = pc, pc.mask
        um = _load_file('_ref_unitrack.utils.mask', 'unitrack/utils/mask.py')
        mods['unitrack.mask'].tensor_mask2box = um.tensor_mask2box
        mods['unitrack.mask'].mask2box = um.mask2box
    except Exception:       # noqa: BLE001 -- keep the stand-ins
        pass
    try:
        core.bbox2result = lambda bboxes, labels, num_classes: [
            (bboxes.cpu().numpy() if hasattr(bboxes, 'cpu') else bboxes)[
                (labels.cpu().numpy() if hasattr(labels, 'cpu') else labels) == i, :] for i in range(num_classes)]
        mt = _load_file('_ref_mmtrack.transform', 'mmtrack/transform.py')
        mods['mmtrack.transform'].outs2results = mt.outs2results
    except Exception:       # noqa: BLE001
        pass


def available():
    import os
    return os.path.isdir(REFERENCE_ROOT + '/knet')


def _load_file(modname, relpath):
    """Import one reference file by PATH under a private module name.  (The product ships alias
    packages named `knet` / `knet_vis` for the configs' custom_imports; importing by path guarantees
    the oracle side really is the reference's source, whatever sys.path says.)"""
    import importlib.util
    import os
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith(REFERENCE_ROOT)
    return mod


def load_video_iter_head():
    """knet/video/kernel_iter_head.py (VideoKernelIterHead: get_panoptic and the merge_* post-processing) by path.
    Its two extra imports are stubbed for the duration of the import only: mmdet's BaseRoIHead (an nn.Module base class
    here) and knet.det.mask_pseudo_sampler (training-side, unused by the post-processing)."""
    install()
    saved = {k: sys.modules.get(k) for k in ('mmdet.models.roi_heads', 'knet', 'knet.det', 'knet.det.mask_pseudo_sampler')}
    rh = types.ModuleType('mmdet.models.roi_heads')
    rh.BaseRoIHead = nn.Module
    sys.modules['mmdet.models.roi_heads'] = rh
    for name in ('knet', 'knet.det', 'knet.det.mask_pseudo_sampler'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules['knet.det.mask_pseudo_sampler'].MaskPseudoSampler = type('MaskPseudoSampler', (), {})
    try:
        mod = _load_file('_ref_knet.video.kernel_iter_head', 'knet/video/kernel_iter_head.py')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_tracker():
    """knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py by path (QuasiDenseEmbedTracker: match / update_memo).  Its
    relative import `..builder` is satisfied by a stand-in package holding a TRACKERS registry; mmdet's bbox_overlaps (IoU,
    mmdet v2.18 semantics: no +1, union clamped at 1e-6) is the oracle's restatement -- third-party arithmetic, see
    knet_oracle.bbox_iou."""
    install()
    import importlib.util
    import os
    import knet_oracle as ko
    sys.modules['mmdet.core'].bbox_overlaps = lambda a, b, mode='iou', is_aligned=False, eps=1e-6: ko.bbox_iou(a, b, eps)
    for name in ('_ref_qd', '_ref_qd.trackers'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    if '_ref_qd.builder' not in sys.modules:
        b = types.ModuleType('_ref_qd.builder')
        b.TRACKERS = Registry('tracker')
        sys.modules['_ref_qd.builder'] = b
    path = os.path.join(REFERENCE_ROOT, 'knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py')
    spec = importlib.util.spec_from_file_location('_ref_qd.trackers.quasi_dense_embed_tracker', path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = '_ref_qd.trackers'
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_assigner():
    """knet/det/mask_hungarian_assigner.py by path (DiceCost, MaskCost, MaskHungarianAssigner).  mmdet's side of it is stubbed:
    AssignResult (a record), BaseAssigner (an empty base), the two registries, and FocalLossCost -- third-party arithmetic,
    restated in knet_oracle.focal_loss_cost (mmdet v2.18)."""
    install()
    import knet_oracle as ko
    core = sys.modules['mmdet.core']

    class AssignResult:
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    core.AssignResult = AssignResult
    core.BaseAssigner = type('BaseAssigner', (), {})
    core.reduce_mean = lambda t: t
    for name in ('mmdet.core.bbox', 'mmdet.core.bbox.builder', 'mmdet.core.bbox.match_costs', 'mmdet.core.bbox.match_costs.builder'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    b = sys.modules['mmdet.core.bbox.builder']
    if not hasattr(b, 'BBOX_ASSIGNERS'):
        b.BBOX_ASSIGNERS = Registry('bbox_assigner')
    mc = sys.modules['mmdet.core.bbox.match_costs.builder']
    if not hasattr(mc, 'MATCH_COST'):
        mc.MATCH_COST = Registry('match_cost')

        class FocalLossCost:
            def __init__(self, weight=1., alpha=0.25, gamma=2, eps=1e-12):
                self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

            def __call__(self, cls_pred, gt_labels):
                return ko.focal_loss_cost(cls_pred, gt_labels, self.weight, self.alpha, self.gamma, self.eps)

        mc.MATCH_COST.register_module()(FocalLossCost)

        def build_match_cost(cfg, default_args=None):
            return mc.MATCH_COST.build(cfg)
        mc.build_match_cost = build_match_cost
    return _load_file('_ref_knet.det.mask_hungarian_assigner', 'knet/det/mask_hungarian_assigner.py')


def load(tree='knet'):
    """Import the reference hot-path modules verbatim.  `tree` is 'knet' or 'knet_vis'
    (they register the same registry keys, so use one per process)."""
    install()
    out = types.SimpleNamespace()
    if tree == 'knet':
        out.kernel_updator = _load_file('_ref_knet.kernel_updator', 'knet/kernel_updator.py')
        out.det_head = _load_file('_ref_knet.det.kernel_update_head', 'knet/det/kernel_update_head.py')
        out.video_head = _load_file('_ref_knet.video.kernel_update_head', 'knet/video/kernel_update_head.py')
        out.KernelUpdator = out.kernel_updator.KernelUpdator
        out.KernelUpdateHead = out.det_head.KernelUpdateHead
        out.VideoKernelUpdateHead = out.video_head.VideoKernelUpdateHead
        out.kernel_head = _load_file('_ref_knet.det.kernel_head', 'knet/det/kernel_head.py')
        out.ConvKernelHead = out.kernel_head.ConvKernelHead
    else:
        out.kernel_updator = _load_file('_ref_knet_vis.kernel_updator', 'knet_vis/kernel_updator.py')
        out.det_head = _load_file('_ref_knet_vis.det.kernel_update_head', 'knet_vis/det/kernel_update_head.py')
        out.tracker_head = _load_file('_ref_knet_vis.tracker.kernel_update_head', 'knet_vis/tracker/kernel_update_head.py')
        out.KernelUpdator = out.kernel_updator.KernelUpdator
        out.KernelUpdateHead = out.det_head.KernelUpdateHead
        out.KernelUpdateHeadVideo = out.tracker_head.KernelUpdateHeadVideo
    return out
