"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
modules (imported from /root/reference through oracle/ref_shim.py) on seeded inputs.

Run in the build container (the only place /root/reference exists):
    python oracle/make_golden.py
The fixtures travel to the GPU box; tests compare both the oracle restatement (CPU) and the
CUDA path (GPU) against them.  Each .npz stores the inputs, the full state_dict (so the fixture
does not depend on RNG reproducibility) and the reference outputs.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import knet_oracle as ko  # noqa: E402
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _pack(prefix, sd):
    return {prefix + k: v.numpy() for k, v in sd.items()}


def case_det(name, B, N, C, H, W, S, Fh, ncls, seed):
    ref = ref_shim.load('knet')
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=Fh)
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=seed)
    blob = dict(x=x.numpy(), proposal_feat=pf.numpy(), mask_preds=mask.numpy(),
                meta=np.array([B, N, C, H, W, S, Fh, ncls], dtype=np.int64))
    obj = pf
    with torch.no_grad():
        for s in range(S):
            sd = ko.random_state_dict(cfg, seed=100 * seed + s)
            head = ref.KernelUpdateHead(**copy.deepcopy(cfg))
            head.load_state_dict(sd, strict=True)
            head.eval()
            cls, mask, obj = head(x, obj, mask)     # knet/det/kernel_iter_head.py:246-253 chaining
            blob.update(_pack('s%d.w.' % s, sd))
            blob['s%d.cls_score' % s] = cls.numpy()
            blob['s%d.mask_preds' % s] = mask.numpy()
            blob['s%d.obj_feat' % s] = obj.numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print(name, 'ok', {k: v.shape for k, v in blob.items() if not k.split('.')[-2:][0] == 'w'} if False else '')


def case_video(name, B, N, C, H, W, Fh, ncls, seed, previous_type, previous_link):
    ref = ref_shim.load('knet')
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=Fh,
                         previous='placeholder', previous_type=previous_type, previous_link=previous_link)
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    prev = torch.randn(B, N, C, 1, 1, generator=g)
    sd = ko.random_state_dict(cfg, seed=100 * seed)
    head = ref.VideoKernelUpdateHead(**copy.deepcopy(cfg))
    head.load_state_dict(sd, strict=True)
    head.eval()
    blob = dict(x=x.numpy(), proposal_feat=pf.numpy(), mask_preds=mask.numpy(), previous_obj_feats=prev.numpy(),
                meta=np.array([B, N, C, H, W, 1, Fh, ncls], dtype=np.int64))
    blob.update(_pack('s0.w.', sd))
    with torch.no_grad():
        cls, nm, obj, x_feat, track = head(x, pf, mask, previous_obj_feats=prev)
        cls0, nm0, obj0, x_feat0, track0 = head(x, pf, mask)       # no previous -> 5th is None
    assert track0 is None
    blob.update({'s0.cls_score': cls.numpy(), 's0.mask_preds': nm.numpy(), 's0.obj_feat': obj.numpy(),
                 's0.x_feat': x_feat.numpy(), 's0.obj_feat_track': track.numpy(),
                 'noprev.cls_score': cls0.numpy(), 'noprev.mask_preds': nm0.numpy(), 'noprev.obj_feat': obj0.numpy()})
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print(name, 'ok')


def case_init(name, B, N, C, H, W, seed):
    """ConvKernelHead._decode_init_proposals through the real reference class (neck stubbed to a passthrough,
    no loc/seg convs): pins the init-kernel conv, the thresholded pooling and the proposal assembly."""
    ref = ref_shim.load('knet')
    g = torch.Generator().manual_seed(seed)
    head = ref.ConvKernelHead(num_proposals=N, in_channels=C, out_channels=C, num_loc_convs=0, num_seg_convs=0,
                              localization_fpn=dict(type='Passthrough'), semantic_fpn=True, num_classes=7,
                              use_binary=True, proposal_feats_with_obj=True, feat_downsample_stride=1,
                              num_thing_classes=3, num_stuff_classes=4, cat_stuff_mask=False)
    with torch.no_grad():
        head.init_kernels.weight.copy_(torch.randn(N, C, 1, 1, generator=g) * 0.3)
    head.eval()
    loc = torch.randn(B, C, H, W, generator=g)
    sem = torch.randn(B, C, H, W, generator=g)
    with torch.no_grad():
        prop, x_feats, mask_preds, cls_scores, seg_preds = head._decode_init_proposals((loc, sem), [dict()] * B)
    assert torch.equal(x_feats, sem + loc)
    blob = dict(loc_feats=loc.numpy(), x_feats=x_feats.numpy(), init_w=head.init_kernels.weight.detach().numpy(),
                proposal_feats=prop.numpy(), mask_preds=mask_preds.numpy(),
                meta=np.array([B, N, C, H, W, 0, 0, 0], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print(name, 'ok')


def case_clip(name, B, Fr, N, C, H, W, Fh, ncls, seed, with_cls):
    """KernelUpdateHeadVideo (knet_vis tree: must run in a process where `knet` was not loaded)."""
    ref = ref_shim.load('knet_vis')
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=Fh)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Fr, C, H, W, generator=g)
    if with_cls:
        pf = torch.randn(B, N, C, 1, 1, generator=g)
        mask = torch.einsum('bnc,bfchw->bfnhw', pf.view(B, N, C), x)
    else:
        pf = torch.randn(B, Fr, N, C, 1, 1, generator=g)
        mask = torch.einsum('bfnc,bfchw->bfnhw', pf.view(B, Fr, N, C), x)
    sd = ko.random_state_dict(cfg, seed=100 * seed)
    if not with_cls:
        sd = {k: v for k, v in sd.items() if not (k.startswith('cls_fcs') or k.startswith('fc_cls'))}
    head = ref.KernelUpdateHeadVideo(with_cls=with_cls, num_proposals=N, **copy.deepcopy(cfg))
    head.load_state_dict(sd, strict=True)
    head.eval()
    with torch.no_grad():
        cls, nm, obj = head(x, pf, mask)
    blob = dict(x=x.numpy(), proposal_feat=pf.numpy(), mask_preds=mask.numpy(),
                meta=np.array([B, N, C, H, W, 1, Fh, ncls], dtype=np.int64), frames=np.array([Fr], dtype=np.int64))
    blob.update(_pack('s0.w.', sd))
    blob['s0.mask_preds'] = nm.numpy()
    blob['s0.obj_feat'] = obj.numpy()
    if cls is not None:
        blob['s0.cls_score'] = cls.numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print(name, 'ok')


def main_clip():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    case_clip('clip_gathered_b2_f3_n10_c64_8x12', 2, 3, 10, 64, 8, 12, 128, 5, 8, True)
    case_clip('clip_perframe_b1_f3_n10_c64_8x12', 1, 3, 10, 64, 8, 12, 128, 5, 9, False)


def case_rescale(name, K, H, W, up, batch_shape, img_shape, ori_shape, seed):
    """Post-loop mask path: the x`up` upsample of _mask_forward (knet/det/kernel_iter_head.py:122-128, restated as the same
    F.interpolate call -- the iter head itself needs the whole detector) followed by the UNMODIFIED reference
    KernelUpdateHead.rescale_masks (knet/det/kernel_update_head.py:443-458)."""
    import torch.nn.functional as F
    ref = ref_shim.load('knet')
    g = torch.Generator().manual_seed(seed)
    masks = torch.randn(K, H, W, generator=g) * 3.0
    img_meta = dict(img_shape=tuple(img_shape) + (3,), batch_input_shape=tuple(batch_shape), ori_shape=tuple(ori_shape) + (3,))
    with torch.no_grad():
        scaled = F.interpolate(masks.unsqueeze(0), scale_factor=up, align_corners=False, mode='bilinear').squeeze(0) if up > 1 else masks
        seg = ref.KernelUpdateHead.rescale_masks(None, scaled, img_meta)
    blob = dict(masks=masks.numpy(), seg=seg.numpy(),
                meta=np.array([K, H, W, up, batch_shape[0], batch_shape[1], img_shape[0], img_shape[1], ori_shape[0], ori_shape[1]],
                              dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **blob)
    print(name, 'ok', tuple(seg.shape))


def case_panoptic(name, K, M, H, W, num_thing, seed):
    """Joint panoptic merge (merge_joint=True): the UNMODIFIED VideoKernelIterHead.merge_stuff_thing_stuff_joint
    (knet/video/kernel_iter_head.py:818-882) on smooth random probability maps."""
    import types
    import torch.nn.functional as F
    mod = ref_shim.load_video_iter_head()
    g = torch.Generator().manual_seed(seed)

    # a random partition of the image into K + M regions (argmax of smooth noise), each mask = a soft indicator of its region;
    # a few masks are then replaced by near-duplicates of a neighbour so that the overlap rejection (:846-848) triggers
    noise = F.interpolate(torch.randn(1, K + M, H // 4, W // 4, generator=g), size=(H, W), mode='bilinear', align_corners=False)[0]
    region = noise.argmax(0)
    soft = torch.stack([(region == i).float() for i in range(K + M)])
    soft = F.avg_pool2d(soft.unsqueeze(0), 3, 1, 1)[0]
    probs = torch.sigmoid(8.0 * (soft - 0.5)) * (0.6 + 0.4 * torch.rand(K + M, 1, 1, generator=g))
    for dup in range(0, K + M, 5):
        probs[dup] = 0.9 * probs[(dup + 1) % (K + M)]
    thing_masks, stuff_masks = probs[:K].contiguous(), probs[K:].contiguous()
    thing_scores, stuff_scores = torch.rand(K, generator=g), torch.rand(M, generator=g)
    thing_labels = torch.randint(0, num_thing, (K,), generator=g)
    stuff_labels = torch.arange(M) + num_thing                      # get_panoptic: stuff_inds + num_thing_classes (:625)
    cfg = types.SimpleNamespace(instance_score_thr=0.25, overlap_thr=0.6, iou_thr=0.5, stuff_max_area=4096)
    fake_self = types.SimpleNamespace(num_thing_classes=num_thing)
    obj_t, obj_s = torch.arange(K).float().view(K, 1), torch.arange(K, K + M).float().view(M, 1)
    (seg, info), kept_obj = mod.VideoKernelIterHead.merge_stuff_thing_stuff_joint(
        fake_self, thing_masks, thing_labels, thing_scores, stuff_masks, stuff_labels, stuff_scores, cfg, obj_t, obj_s)
    rows = np.array([[d['id'], int(d['isthing']), d['category_id'], d.get('instance_id', -1), d.get('area', -1)] for d in info],
                    dtype=np.int64).reshape(-1, 5)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), thing_masks=thing_masks.numpy(), stuff_masks=stuff_masks.numpy(),
                        thing_scores=thing_scores.numpy(), stuff_scores=stuff_scores.numpy(), thing_labels=thing_labels.numpy(),
                        stuff_labels=stuff_labels.numpy(), seg=seg, info=rows, kept=kept_obj.view(-1).long().numpy(),
                        meta=np.array([K, M, H, W, num_thing], dtype=np.int64), thr=np.array([0.25, 0.6]))
    print(name, 'ok', seg.shape, len(info), 'segments')


def main_panoptic():
    os.makedirs(OUT, exist_ok=True)
    case_panoptic('panoptic_joint_k12_m5_40x64', 12, 5, 40, 64, 2, 21)
    case_panoptic('panoptic_joint_k20_m9_32x48', 20, 9, 32, 48, 2, 22)


def main_rescale():
    os.makedirs(OUT, exist_ok=True)
    case_rescale('rescale_k5_12x20_up2_96x160_crop90x150_to47x83', 5, 12, 20, 2, (96, 160), (90, 150), (47, 83), 11)
    case_rescale('rescale_k3_9x13_up1_36x52_crop33x50_to70x101', 3, 9, 13, 1, (36, 52), (33, 50), (70, 101), 12)


def main():
    assert ref_shim.available(), '/root/reference is required to (re)generate the fixtures'
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    # BASELINE.json configs[0]: single 64x64 frame, N=10, C=64, S=1 (reference CPU case)
    case_det('det_cfg0_b1_n10_c64_64x64_s1', 1, 10, 64, 64, 64, 1, 2048, 19, seed=1)
    # ragged: odd H x W (HW not a multiple of any tile), B=2, two chained stages
    case_det('det_b2_n12_c64_13x17_s2', 2, 12, 64, 13, 17, 2, 128, 5, seed=2)
    # N above one 128-row tile is exercised at C=64 as well (N=130), 3 chained stages
    case_det('det_b1_n130_c64_9x31_s3', 1, 130, 64, 9, 31, 3, 64, 3, seed=3)
    case_video('video_link_ffn_b2_n12_c64_9x11', 2, 12, 64, 9, 11, 128, 5, 4, 'ffn', None)
    case_video('video_link_update_b2_n12_c64_9x11', 2, 12, 64, 9, 11, 128, 5, 5, 'update', 'update_dynamic_cov')
    case_video('video_link_cov_ffn_b1_n12_c64_9x11', 1, 12, 64, 9, 11, 128, 5, 6, 'ffn', 'update_dynamic_cov')
    case_init('init_b2_n20_c64_12x16', 2, 20, 64, 12, 16, 10)
    main_rescale()
    main_panoptic()


def case_track(name, frames, nmax, D, ncls, seed):
    """QuasiDenseEmbedTracker.match of the UNMODIFIED reference class over a short sequence (KITTI-STEP tracker settings,
    configs/det/video_knet_kitti_step/...joint_train_8e.py:61-73): per frame the inputs, the memory it matched against and
    its outputs.  Instances persist across frames (drifting boxes / embeddings) with births, deaths and duplicates."""
    mod = ref_shim.load_tracker()
    cfg = dict(init_score_thr=0.35, obj_score_thr=0.3, match_score_thr=0.5, memo_tracklet_frames=5, memo_backdrop_frames=1,
               memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3, nms_class_iou_thr=0.7, with_cats=True,
               match_metric='bisoftmax')
    trk = mod.QuasiDenseEmbedTracker(**cfg)
    g = torch.Generator().manual_seed(seed)
    n_obj = nmax
    base_emb = torch.randn(n_obj, D, generator=g) * 1.2
    base_box = torch.rand(n_obj, 2, generator=g) * 300
    size = 20 + torch.rand(n_obj, 2, generator=g) * 80
    lab = torch.randint(0, ncls, (n_obj,), generator=g)
    out = dict(meta=np.array([frames, D, ncls], dtype=np.int64),
               cfg=np.array([cfg['obj_score_thr'], cfg['match_score_thr'], cfg['init_score_thr'], cfg['nms_conf_thr'],
                             cfg['nms_backdrop_iou_thr'], cfg['nms_class_iou_thr']], dtype=np.float32))
    for f in range(frames):
        alive = torch.rand(n_obj, generator=g) > 0.25
        idx = torch.nonzero(alive).squeeze(1)
        if f == 2:                                   # a duplicate detection of one object (must be removed by the IoU rule)
            idx = torch.cat([idx, idx[:1]])
        xy = base_box[idx] + f * 4.0 + torch.randn(len(idx), 2, generator=g)
        wh = size[idx]
        score = 0.15 + 0.85 * torch.rand(len(idx), generator=g)
        bboxes = torch.cat([xy, xy + wh, score[:, None]], dim=1)
        feats = base_emb[idx] + 0.15 * torch.randn(len(idx), D, generator=g)
        labels = lab[idx].clone()
        if trk.empty:
            memo = (torch.zeros(0, dtype=torch.long), torch.zeros(0, D), torch.zeros(0, dtype=torch.long))
        else:
            mb, ml, me, mi, mv = trk.memo
            memo = (ml.clone(), me.clone(), mi.clone())
        ntr = int(trk.num_tracklets)
        ob, ol, oi = trk.match(bboxes=bboxes.clone(), labels=labels.clone(), track_feats=feats.clone(), frame_id=f)
        for k, v in (('bboxes', bboxes), ('labels', labels), ('feats', feats), ('memo_labels', memo[0]), ('memo_embeds', memo[1]),
                     ('memo_ids', memo[2]), ('out_bboxes', ob), ('out_labels', ol), ('out_ids', oi)):
            out['f%d.%s' % (f, k)] = v.numpy()
        out['f%d.num_tracklets' % f] = np.array([ntr, int(trk.num_tracklets)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, {k: v.shape for k, v in out.items() if k.startswith('f1.')})


def main_track():
    ref_shim.install()
    case_track('track_match_kitti_6f_n24_d256', frames=6, nmax=24, D=256, ncls=2, seed=5)
    case_track('track_match_4f_n60_d64', frames=4, nmax=60, D=64, ncls=3, seed=9)


def case_assign(name, N, M, H, W, ncls, seed):
    """MaskHungarianAssigner of the UNMODIFIED reference (shipped weights: FocalLossCost 2, MaskCost 1, DiceCost 4, pred_act=True;
    configs/det/_base_/models/knet_s3_r50_fpn.py:114-118): the three cost terms of its own cost objects, their sum, and the
    assignment `assign` returns.  Targets are bilinearly down-sampled blobs (values in [0, 1], as in training)."""
    mod = ref_shim.load_assigner()
    asg = mod.MaskHungarianAssigner(cls_cost=dict(type='FocalLossCost', weight=2.0), dice_cost=dict(type='DiceCost', weight=4.0, pred_act=True),
                                    mask_cost=dict(type='MaskCost', weight=1.0, pred_act=True))
    g = torch.Generator().manual_seed(seed)
    big = (torch.rand(M, 1, 4 * H // 8 + 2, 4 * W // 8 + 2, generator=g) > 0.6).float()
    gt = torch.nn.functional.interpolate(big, (H, W), mode='bilinear', align_corners=False)[:, 0]
    pred = 3.0 * torch.randn(N, H, W, generator=g)
    for m in range(min(M, N)):                       # some predictions resemble a target (a meaningful matching)
        pred[(m * 7) % N] = (gt[m] - 0.5) * 12 + torch.randn(H, W, generator=g)
    cls = 2.0 * torch.randn(N, ncls, generator=g) - 2.0
    labels = torch.randint(0, ncls, (M,), generator=g)
    with torch.no_grad():
        c_cls, c_mask, c_dice = asg.cls_cost(cls, labels), asg.mask_cost(pred, gt), asg.dice_cost(pred, gt)
        res = asg.assign(pred, cls, gt, labels)
    out = dict(meta=np.array([N, M, H, W, ncls], dtype=np.int64), mask_logits=pred.numpy(), cls_logits=cls.numpy(), gt_masks=gt.numpy(),
               gt_labels=labels.numpy(), cost_cls=c_cls.numpy(), cost_mask=c_mask.numpy(), cost_dice=c_dice.numpy(),
               cost=(c_cls + c_mask + c_dice).numpy(), gt_inds=res.gt_inds.numpy(), labels=res.labels.numpy())
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'matched', int((res.gt_inds > 0).sum()))


def main_assign():
    ref_shim.install()
    case_assign('assign_n100_m17_40x56_c19', 100, 17, 40, 56, 19, seed=3)
    case_assign('assign_n20_m3_13x17_c5', 20, 3, 13, 17, 5, seed=4)
    case_assign('assign_n150_m40_24x32_c124', 150, 40, 24, 32, 124, seed=6)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'assign':
        main_assign()
    elif len(sys.argv) > 1 and sys.argv[1] == 'track':
        main_track()
    elif len(sys.argv) > 1 and sys.argv[1] == 'clip':      # knet_vis registers the same keys as knet: own process
        main_clip()
    elif len(sys.argv) > 1 and sys.argv[1] == 'rescale':
        main_rescale()
    elif len(sys.argv) > 1 and sys.argv[1] == 'panoptic':
        main_panoptic()
    else:
        main()
        import subprocess
        subprocess.run([sys.executable, os.path.abspath(__file__), 'clip'], check=True)
