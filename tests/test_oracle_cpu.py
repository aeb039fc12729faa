"""CPU: pins the oracle (oracle/knet_oracle.py) against the reference.
 (1) golden fixtures written by the unmodified reference (tests/golden, made by oracle/make_golden.py);
 (2) when /root/reference is present (build container only), the live reference on fresh seeds."""
import copy

import numpy as np
import pytest
import torch

import knet_oracle as ko
from helpers import golden_files, load_golden, maxabs


@pytest.mark.parametrize('path', golden_files('det_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_det(path):
    g = load_golden(path)
    t = g['t']
    outs = ko.iter_forward(g['sds'], [g['cfg']] * g['S'], t['x'], t['proposal_feat'], t['mask_preds'])
    for s, (cls, m, obj) in enumerate(outs):
        assert maxabs(cls, t['s%d.cls_score' % s]) < 2e-5
        assert maxabs(obj, t['s%d.obj_feat' % s]) < 2e-5
        assert maxabs(m, t['s%d.mask_preds' % s]) < 2e-4
        assert torch.equal(m.argmax(1), t['s%d.mask_preds' % s].argmax(1))


@pytest.mark.parametrize('path', golden_files('video_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_video(path):
    g = load_golden(path)
    t = g['t']
    out = ko.video_kernel_update_head_forward(g['sds'][0], g['cfg'], t['x'], t['proposal_feat'], t['mask_preds'],
                                              t['previous_obj_feats'])
    for name, got in zip(('cls_score', 'mask_preds', 'obj_feat', 'x_feat', 'obj_feat_track'), out):
        ref = t['s0.' + name]
        assert maxabs(got, ref) < 2e-5 * max(1.0, ref.abs().max().item()), name
    out = ko.video_kernel_update_head_forward(g['sds'][0], g['cfg'], t['x'], t['proposal_feat'], t['mask_preds'], None)
    assert out[4] is None
    assert maxabs(out[2], t['noprev.obj_feat']) < 2e-5


@pytest.mark.parametrize('path', golden_files('clip_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_clip_head(path):
    g = load_golden(path)
    t = g['t']
    cls, nm, obj = ko.kernel_update_head_video_forward(g['sds'][0], g['cfg'], t['x'], t['proposal_feat'], t['mask_preds'])
    assert maxabs(nm, t['s0.mask_preds']) < 2e-4 and maxabs(obj, t['s0.obj_feat']) < 2e-5
    assert torch.equal(nm.argmax(2), t['s0.mask_preds'].argmax(2))
    if 's0.cls_score' in t:
        assert maxabs(cls, t['s0.cls_score']) < 2e-5
    else:
        assert cls is None


def test_fixture_generator_is_committed_and_fixtures_exist():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert os.path.exists(os.path.join(root, 'oracle', 'make_golden.py'))
    assert len(golden_files('det_')) >= 3 and len(golden_files('video_')) >= 3 and len(golden_files('clip_')) >= 2


def _live_reference():
    import ref_shim
    if not ref_shim.available():
        pytest.skip('/root/reference not present (GPU box): fixtures cover this')
    return ref_shim.load('knet')


@pytest.mark.parametrize('B,N,C,H,W', [(1, 100, 256, 25, 11), (2, 17, 64, 7, 9)])
def test_oracle_matches_live_reference_det(B, N, C, H, W):
    ref = _live_reference()
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256)
    sd = ko.random_state_dict(cfg, seed=42)
    head = ref.KernelUpdateHead(**copy.deepcopy(cfg))
    head.load_state_dict(sd, strict=True)
    head.eval()
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=11)
    with torch.no_grad():
        want = head(x, pf, mask)
    got = ko.kernel_update_head_forward(sd, cfg, x, pf, mask)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert maxabs(a, b) < 2e-5 * max(1.0, b.abs().max().item())


def test_oracle_init_matches_reference_init():
    """init_weights semantics (kernel_update_head.py:151-168): fc_cls.bias = -log(99), LN defaults."""
    ref = _live_reference()
    cfg = ko.default_cfg(num_classes=19)
    torch.manual_seed(0)
    head = ref.KernelUpdateHead(**copy.deepcopy(cfg))
    head.init_weights()
    import vknet
    torch.manual_seed(0)
    mine = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    mine.init_weights()
    sd_r, sd_m = head.state_dict(), mine.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys())
    for k in sd_r:
        assert sd_r[k].shape == sd_m[k].shape, k
    assert torch.allclose(sd_m['fc_cls.bias'], sd_r['fc_cls.bias'])
    assert torch.equal(sd_m['attention_norm.weight'], sd_r['attention_norm.weight'])


def test_threshold_sliver_is_documented_behaviour():
    """sigmoid(m) > 0.5 vs m > 0 differ only for 0 < m < ~6e-8 in fp32."""
    m = torch.tensor([-1.0, -1e-30, 0.0, 1e-30, 1e-8, 1e-6, 1.0])
    ref = (torch.sigmoid(m) > 0.5)
    ours = m > 0.0
    assert torch.equal(ref[[0, 1, 2, 5, 6]], ours[[0, 1, 2, 5, 6]])
    assert not ref[3] and ours[3]      # the sliver


def test_oracle_matches_reference_fixture_init_proposals():
    """tail of ConvKernelHead._decode_init_proposals, fixture produced by the real class."""
    import numpy as np
    z = np.load(golden_files('init_')[0])
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    prop, mask = ko.init_proposals(t['init_w'], None, t['loc_feats'], t['x_feats'])
    assert maxabs(prop, t['proposal_feats']) < 2e-5 * t['proposal_feats'].abs().max().item()
    assert maxabs(mask, t['mask_preds']) < 2e-5 * t['mask_preds'].abs().max().item()


@pytest.mark.parametrize('path', golden_files('rescale_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_rescale_masks(path):
    """Post-loop mask path: oracle restatement == the reference's rescale_masks output."""
    import numpy as np
    z = np.load(path)
    K, H, W, up, Hb, Wb, h, w, Ho, Wo = (int(v) for v in z['meta'])
    meta = dict(img_shape=(h, w, 3), batch_input_shape=(Hb, Wb), ori_shape=(Ho, Wo, 3))
    got = ko.rescale_masks(torch.from_numpy(z['masks']), meta, up)
    assert got.shape == (K, Ho, Wo)
    assert (got - torch.from_numpy(z['seg'])).abs().max().item() < 1e-6


@pytest.mark.parametrize('path', golden_files('panoptic_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_panoptic_merge(path):
    """Next part of row 8f-2 (not yet on the GPU): the joint score-weighted argmax merge.  Oracle restatement == the
    reference's merge_stuff_thing_stuff_joint outputs (id map, segment table, kept thing indices)."""
    import numpy as np
    z = np.load(path)
    K, M, H, W, nthing = (int(v) for v in z['meta'])
    t = {k: torch.from_numpy(z[k]) for k in ('thing_masks', 'stuff_masks', 'thing_scores', 'stuff_scores', 'thing_labels',
                                             'stuff_labels')}
    seg, info, kept = ko.panoptic_merge_joint(t['thing_masks'], t['thing_labels'], t['thing_scores'], t['stuff_masks'],
                                              t['stuff_labels'], t['stuff_scores'], nthing, float(z['thr'][0]), float(z['thr'][1]))
    assert np.array_equal(seg.numpy(), z['seg'])
    rows = np.array([[d['id'], int(d['isthing']), d['category_id'], d.get('instance_id', -1), d.get('area', -1)] for d in info],
                    dtype=np.int64).reshape(-1, 5)
    assert np.array_equal(rows, z['info'])
    assert kept == z['kept'].tolist()
    assert len(info) >= 10 and any(not d['isthing'] for d in info) and len(info) < K + M      # fixtures exercise every branch


def test_no_cpu_path_for_the_post_loop_ops(built_lib):
    from vknet import _lib, ops
    meta = dict(img_shape=(8, 8, 3), batch_input_shape=(8, 8), ori_shape=(8, 8, 3))
    with pytest.raises(_lib.VknError):
        ops.rescale_masks(torch.zeros(2, 4, 4), meta)


@pytest.mark.parametrize('path', golden_files('track_match_'), ids=lambda p: p.split('/')[-1][:-4])
def test_oracle_matches_reference_fixtures_tracker_match(path):
    """Row 8f-3: the association restatement == the reference's QuasiDenseEmbedTracker.match outputs, frame by frame,
    against the memory the reference held at that frame (fixtures written by oracle/make_golden.py track)."""
    import numpy as np
    z = np.load(path)
    frames = int(z['meta'][0])
    thr = [float(v) for v in z['cfg']]
    matched = news = 0
    for f in range(frames):
        t = {k: torch.from_numpy(z['f%d.%s' % (f, k)]) for k in ('bboxes', 'labels', 'feats', 'memo_labels', 'memo_embeds', 'memo_ids',
                                                               'out_bboxes', 'out_labels', 'out_ids')}
        n0, n1 = (int(v) for v in z['f%d.num_tracklets' % f])
        sel, ids, nnew = ko.tracker_match(t['bboxes'], t['labels'], t['feats'], t['memo_labels'], t['memo_embeds'], t['memo_ids'], n0,
                                          *thr, with_cats=True)
        assert torch.equal(t['bboxes'][sel], t['out_bboxes']) and torch.equal(t['labels'][sel], t['out_labels'])
        assert torch.equal(ids, t['out_ids']) and n0 + nnew == n1
        matched += int(((ids > -1) & (ids < n0)).sum())
        news += nnew
    assert matched >= 10 and news >= 10          # the fixtures exercise matching, births and the duplicate rule


# ---- row f4: MaskHungarianAssigner cost matrix ---------------------------------------------------------------------------
@pytest.mark.parametrize('path', golden_files('assign_'), ids=lambda p: p.split('/')[-1][:-4])
def test_match_cost_oracle_matches_reference_fixtures(path):
    """The restated DiceCost / MaskCost / FocalLossCost against the outputs of the reference's own cost objects, and the
    Hungarian assignment on the restated cost against the reference's `assign`."""
    from scipy.optimize import linear_sum_assignment
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    pred, cls, gt, lab = t['mask_logits'], t['cls_logits'], t['gt_masks'], t['gt_labels']
    assert (ko.focal_loss_cost(cls, lab, 2.0) - t['cost_cls']).abs().max() < 1e-6
    assert (ko.mask_cost(pred, gt, 1.0) - t['cost_mask']).abs().max() < 1e-6
    assert (ko.dice_cost(pred, gt, 4.0) - t['cost_dice']).abs().max() < 1e-6
    cost = ko.match_cost(pred, cls, gt, lab)
    assert (cost - t['cost']).abs().max() < 2e-6
    rows, cols = linear_sum_assignment(cost)
    inds = torch.zeros(pred.shape[0], dtype=torch.long)
    inds[torch.from_numpy(rows)] = torch.from_numpy(cols) + 1
    assert torch.equal(inds, t['gt_inds'])


def test_match_cost_oracle_matches_live_reference():
    import ref_shim
    if not ref_shim.available():
        pytest.skip('needs /root/reference')
    mod = ref_shim.load_assigner()
    g = torch.Generator().manual_seed(12)
    pred, gt = 2 * torch.randn(9, 7, 11, generator=g), torch.rand(4, 7, 11, generator=g)
    assert (mod.DiceCost(weight=4.0, pred_act=True)(pred, gt) - ko.dice_cost(pred, gt, 4.0)).abs().max() < 1e-6
    assert (mod.MaskCost(weight=1.0, pred_act=True)(pred, gt) - ko.mask_cost(pred, gt, 1.0)).abs().max() < 1e-6
