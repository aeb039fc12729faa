"""Functional wrappers over the individual C-ABI operators (rows a3-a9 of the scope table).
Weights come from a KernelUpdateHead-like module; tensors are torch CUDA tensors.  Used by the
per-operator parity tests and by anyone who wants one piece of the stage.
"""
import torch

from . import _lib

_ws = _lib.Workspace()


def _ctx(head, x_like, B, N, H, W, x_dtype=None):
    dev = x_like.device
    if not x_like.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    w, extra, wd = head.packed_weights(dev)
    xd = _lib.dtype_code(x_like.dtype) if x_dtype is None else x_dtype
    shape = head._shape(B, N, H, W, xd, wd)
    ws, wsb = _ws.get(shape, dev)
    return w, extra, shape, ws, wsb


def mask_pool(head, x, mask_preds):
    """x_feat [B,N,C] = sum_p 1[sigmoid(mask) > thr] * feat_transform(x)   (kernel_update_head.py:179-195)."""
    B, Cc, H, W = x.shape
    N = mask_preds.shape[1]
    if mask_preds.shape != (B, N, H, W) or Cc != head.in_channels:
        raise _lib.VknError('mask_pool: x %s / mask_preds %s do not describe [B,C=%d,H,W] / [B,N,H,W]' % (
            tuple(x.shape), tuple(mask_preds.shape), head.in_channels))
    x = x.contiguous()
    mask_preds = mask_preds.to(x.dtype).contiguous()
    w, _, shape, ws, wsb = _ctx(head, x, B, N, H, W)
    out = torch.empty(B, N, Cc, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().vkn_mask_pool(shape, w, _lib.ptr(x), _lib.ptr(mask_preds), _lib.ptr(out), ws, wsb,
                                        _lib.stream_ptr()))
    return out


def kernel_update(head, x_feat, proposal_feat):
    """KernelUpdator.forward on [B,N,C] rows (kernel_updator.py:56-94) -> [B,N,C]."""
    B, N, Cc = x_feat.shape
    xf = x_feat.float().contiguous()
    pf = proposal_feat.reshape(B, N, Cc).float().contiguous()
    w, _, shape, ws, wsb = _ctx(head, xf, B, N, 1, 1, _lib.VKN_F32)
    out = torch.empty_like(xf)
    _lib.check(_lib.lib().vkn_kernel_update(shape, w.upd, _lib.ptr(xf), _lib.ptr(pf), _lib.ptr(out), ws, wsb,
                                            _lib.stream_ptr()))
    return out


def mhsa_ln(head, q_in, kv_in=None, attn_w=None):
    """LN(q + MHA(q, kv, kv)) across the N kernels of each frame (kernel_update_head.py:204-208)."""
    B, N, Cc = q_in.shape
    q = q_in.float().contiguous()
    kv = None if kv_in is None else kv_in.float().contiguous()
    w, _, shape, ws, wsb = _ctx(head, q, B, N, 1, 1, _lib.VKN_F32)
    out = torch.empty_like(q)
    _lib.check(_lib.lib().vkn_mhsa_ln(shape, attn_w if attn_w is not None else w.attn, _lib.ptr(q), _lib.ptr(kv),
                                      _lib.ptr(out), ws, wsb, _lib.stream_ptr()))
    return out


def ffn_ln(head, inp):
    """LN(x + FFN(x)) (kernel_update_head.py:214-215)."""
    B, N, Cc = inp.shape
    a = inp.float().contiguous()
    w, _, shape, ws, wsb = _ctx(head, a, B, N, 1, 1, _lib.VKN_F32)
    out = torch.empty_like(a)
    _lib.check(_lib.lib().vkn_ffn_ln(shape, w.ffn, _lib.ptr(a), _lib.ptr(out), ws, wsb, _lib.stream_ptr()))
    return out


def heads(head, obj_feat):
    """cls_score [B,N,ncls], mask_kernel [B,N,C] (kernel_update_head.py:217-227)."""
    B, N, Cc = obj_feat.shape
    a = obj_feat.float().contiguous()
    w, _, shape, ws, wsb = _ctx(head, a, B, N, 1, 1, _lib.VKN_F32)
    cls = torch.empty(B, N, head.fc_cls.out_features, dtype=torch.float32, device=a.device)
    mk = torch.empty_like(a)
    _lib.check(_lib.lib().vkn_heads(shape, w, _lib.ptr(a), _lib.ptr(cls), _lib.ptr(mk), ws, wsb, _lib.stream_ptr()))
    return cls, mk


def mask_gemm(head, x, mask_kernel):
    """new_mask [B,N,H,W] = mask_kernel . feat_transform(x)   (kernel_update_head.py:179-180, 247-260)."""
    B, Cc, H, W = x.shape
    N = mask_kernel.shape[1]
    if mask_kernel.shape[0] != B or mask_kernel.numel() != B * N * Cc or Cc != head.in_channels:
        raise _lib.VknError('mask_gemm: mask_kernel %s does not hold [B=%d, N, C=%d] kernels for x %s' % (
            tuple(mask_kernel.shape), B, head.in_channels, tuple(x.shape)))
    x = x.contiguous()
    mk = mask_kernel.reshape(B, N, Cc).float().contiguous()
    w, _, shape, ws, wsb = _ctx(head, x, B, N, H, W)
    out = torch.empty(B, N, H, W, dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().vkn_mask_gemm(shape, w, _lib.ptr(x), _lib.ptr(mk), _lib.ptr(out), ws, wsb,
                                        _lib.stream_ptr()))
    return out


def init_proposals(init_kernels, loc_feats, x_feats, engine=_lib.ENGINE_AUTO):
    """Tail of ConvKernelHead._decode_init_proposals (knet/det/kernel_head.py:212, 234-254) with the
    shipped settings proposal_feats_with_obj=True, use_binary=True:
        mask_preds = init_kernels(loc_feats);  obj = einsum(1[sigmoid(mask_preds) > 0.5], x_feats);
        proposal_feats = init_kernels.weight + obj
    init_kernels: the nn.Conv2d(C, N, 1) of the head.  Returns (proposal_feats [B,N,C,1,1] fp32, mask_preds [B,N,H,W])."""
    w = init_kernels.weight
    if w.shape[-1] != 1 or w.shape[-2] != 1:
        raise NotImplementedError('conv_kernel_size != 1 is not on the shipped path')
    if not x_feats.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    B, Cc, H, W = x_feats.shape
    N = w.shape[0]
    x_feats = x_feats.contiguous()
    loc_feats = loc_feats.to(x_feats.dtype).contiguous()
    xd = _lib.dtype_code(x_feats.dtype)
    wf = w.detach().reshape(N, Cc).to(device=x_feats.device, dtype=torch.float32).contiguous()
    bf = None if init_kernels.bias is None else init_kernels.bias.detach().to(device=x_feats.device, dtype=torch.float32).contiguous()
    shape = _lib.make_shape(B, N, Cc, H, W, 32, 1, 8 if Cc % 8 == 0 and Cc // 8 <= 32 else Cc // 32, xd, _lib.VKN_F32,
                            True, engine, 0.0)
    ws, wsb = _ws.get(shape, x_feats.device)
    mask = torch.empty(B, N, H, W, dtype=x_feats.dtype, device=x_feats.device)
    prop = torch.empty(B, N, Cc, dtype=torch.float32, device=x_feats.device)
    _lib.check(_lib.lib().vkn_init_proposals(shape, _lib.ptr(wf), _lib.ptr(bf), _lib.ptr(loc_feats), _lib.ptr(x_feats),
                                             _lib.ptr(mask), _lib.ptr(prop), ws, wsb, _lib.stream_ptr()))
    return prop.reshape(B, N, Cc, 1, 1), mask


def rescale_masks(masks, img_meta, mask_upsample_stride=1, mask_thr=None, probs=True):
    """Post-loop mask path in one launch:
        F.interpolate(masks, scale_factor=mask_upsample_stride)            knet/det/kernel_iter_head.py:122-128
        KernelUpdateHead.rescale_masks(., img_meta)                        knet/det/kernel_update_head.py:443-458
        (. > mask_thr)                                                     :460-462 (get_seg_masks)
    masks [K,H,W] logits (fp32 / bf16, CUDA).  Returns (seg_probs [K,ori_h,ori_w] fp32 or None,
    seg_masks bool [K,ori_h,ori_w] or None when mask_thr is None)."""
    if not masks.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    if masks.dim() != 3:
        raise ValueError('masks must be [K,H,W]')
    K, H, W = masks.shape
    h, w = img_meta['img_shape'][:2]
    Hb, Wb = img_meta['batch_input_shape'][:2]
    Ho, Wo = img_meta['ori_shape'][:2]
    masks = masks.contiguous()
    out_p = torch.empty(K, Ho, Wo, dtype=torch.float32, device=masks.device) if probs else None
    out_b = torch.empty(K, Ho, Wo, dtype=torch.bool, device=masks.device) if mask_thr is not None else None   # kernel writes 0 / 1 bytes
    if out_p is None and out_b is None:
        raise ValueError('nothing to compute: probs=False and mask_thr=None')
    if K > 0:
        _lib.check(_lib.lib().vkn_rescale_masks(_lib.ptr(masks), _lib.dtype_code(masks.dtype), K, H, W, int(mask_upsample_stride),
                                                int(Hb), int(Wb), int(h), int(w), int(Ho), int(Wo),
                                                float(mask_thr if mask_thr is not None else 0.5), _lib.ptr(out_p), _lib.ptr(out_b),
                                                _lib.stream_ptr()))
    return out_p, out_b


def mask_boxes(masks):
    """boxes [K,4] fp32 (x_min, y_min, x_max, y_max) of the non-zero pixels of each mask, (-1,-1,10,10) for an empty one:
    the mask -> box step of VideoKernelUpdateHead.segm2result (knet/video/kernel_update_head.py:734-744, unitrack
    tensor_mask2box) as one launch instead of a `.nonzero()` + four `.item()` per mask.  masks [K,H,W] bool / uint8 /
    float32 CUDA tensor."""
    if not masks.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    if masks.dim() != 3:
        raise ValueError('masks must be [K,H,W]')
    if masks.dtype in (torch.bool, torch.uint8):
        m, eb = masks.contiguous(), 1
    else:
        m, eb = masks.float().contiguous(), 4
    K, H, W = m.shape
    out = torch.empty(K, 4, dtype=torch.float32, device=m.device)
    _lib.check(_lib.lib().vkn_mask_boxes(_lib.ptr(m), eb, K, H, W, _lib.ptr(out), _lib.stream_ptr(m.device)))
    return out


def panoptic_merge(thing_masks, thing_labels, thing_scores, stuff_masks, stuff_labels, stuff_scores, num_thing_classes,
                   instance_score_thr, overlap_thr):
    """VideoKernelIterHead.merge_stuff_thing_stuff_joint (knet/video/kernel_iter_head.py:832-895) on the device: three
    launches and ONE device->host read of the small result tables instead of 2-3 host synchronisations per segment.
    masks [K,H,W] / [M,H,W] fp32 probabilities, labels integer, scores fp32 (CUDA).
    Returns (panoptic_seg int32 [H,W] CUDA tensor, segments_info list of dicts as the reference builds it, kept thing
    indices -- what the reference uses to gather `thing_obj_feat`)."""
    if not thing_masks.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    dev = thing_masks.device
    masks = torch.cat([thing_masks, stuff_masks], 0).float().contiguous()
    scores = torch.cat([thing_scores, stuff_scores], 0).float().contiguous()
    labels = torch.cat([thing_labels, stuff_labels], 0).to(torch.int32).contiguous()
    T, H, W = masks.shape
    seg = torch.empty(H, W, dtype=torch.int32, device=dev)
    small = torch.zeros(T * 5 + T + 2, dtype=torch.int32, device=dev)          # segment rows | kept | counts
    seg_scores = torch.empty(T, dtype=torch.float32, device=dev)
    ws = torch.empty(H * W + 3 * T, dtype=torch.int32, device=dev)
    table, kept, counts = small[:T * 5], small[T * 5:T * 6], small[T * 6:]
    _lib.check(_lib.lib().vkn_panoptic_merge(_lib.ptr(masks), _lib.ptr(scores), _lib.ptr(labels), T, H, W, int(num_thing_classes),
                                             float(instance_score_thr), float(overlap_thr), _lib.ptr(seg), _lib.ptr(table),
                                             _lib.ptr(seg_scores), _lib.ptr(kept), _lib.ptr(counts), _lib.ptr(ws), ws.numel() * 4,
                                             _lib.stream_ptr(dev)))
    host = small.cpu()                                                          # the one synchronising read
    sc = seg_scores.cpu()
    nseg, nkept = int(host[T * 6]), int(host[T * 6 + 1])
    info = []
    for i in range(nseg):
        sid, isthing, cat, inst, area = (int(v) for v in host[i * 5:i * 5 + 5])
        if isthing:
            info.append(dict(id=sid, isthing=True, score=float(sc[i]), category_id=cat, instance_id=inst))
        else:
            info.append(dict(id=sid, isthing=False, category_id=cat, area=area))
    return seg, info, [int(v) for v in host[T * 5:T * 5 + nkept]]


def mlp(layers, x):
    """A stack of Linear layers on [rows, features] fp32 rows (vkn_mlp).  `layers`: list of
    (linear: nn.Linear, norm: nn.LayerNorm | None, relu: bool) applied as  y = relu?(norm?(linear(x))).
    The tracking-embedding path of the video detectors is such a stack:
        embed_fcs + fc_embed            knet/video/knet_quansi_dense_embed_fc_joint_train.py:113-126, 572-580
        track_head fcs + fc_embed       knet/video/track_heads.py:632-642"""
    import ctypes as C
    if not x.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    dev = x.device
    rows = x.reshape(-1, x.shape[-1]).float().contiguous()
    dts = {lin.weight.dtype for lin, _, _ in layers}
    if len(dts) != 1 or dts.pop() not in (torch.float32, torch.bfloat16):
        raise _lib.VknError('mlp: weights must be uniformly float32 or bfloat16')
    wd = _lib.dtype_code(layers[0][0].weight.dtype)
    arr = (_lib.VknMlpLayer * len(layers))()
    keep = []

    def vec(t):
        t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep.append(t)
        return C.c_void_p(t.data_ptr())
    maxd = 0
    for i, (lin, norm, relu) in enumerate(layers):
        w = lin.weight.detach().to(dev).contiguous()
        keep.append(w)
        arr[i].w = C.c_void_p(w.data_ptr())
        arr[i].b = vec(lin.bias) if lin.bias is not None else None
        arr[i].ln_g = vec(norm.weight) if norm is not None else None
        arr[i].ln_b = vec(norm.bias) if norm is not None else None
        arr[i].in_dim, arr[i].out_dim, arr[i].relu = lin.in_features, lin.out_features, int(bool(relu))
        maxd = max(maxd, lin.out_features)
    out = torch.empty(rows.shape[0], layers[-1][0].out_features, dtype=torch.float32, device=dev)
    buf = (rows.shape[0] * maxd * 4 + 255) // 256 * 256
    ws = torch.empty(2 * buf + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    _lib.check(_lib.lib().vkn_mlp(arr, len(layers), wd, _lib.ptr(rows), _lib.ptr(out), rows.shape[0],
                                  C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _lib.stream_ptr(dev)))
    return out.reshape(tuple(x.shape[:-1]) + (out.shape[-1],))


def track_match(bboxes, labels, track_feats, memo_labels, memo_embeds, memo_ids, num_tracklets, obj_score_thr=0.5,
                match_score_thr=0.5, init_score_thr=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3, nms_class_iou_thr=0.7,
                with_cats=True):
    """QuasiDenseEmbedTracker.match up to the memory update (knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py:
    137-204, match_metric='bisoftmax') as ONE launch and one device->host read, instead of ~4 host synchronisations per
    detection.  Returns (selected: indices of the kept detections in score order, ids: their track ids int64, number of new
    tracks); bboxes[selected], labels[selected], track_feats[selected] are the tensors the reference returns / memorises."""
    if not bboxes.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    dev = bboxes.device
    n, m = bboxes.shape[0], (0 if memo_embeds is None else memo_embeds.shape[0])
    D = track_feats.shape[1] if n else (memo_embeds.shape[1] if m else 1)
    bb = bboxes.float().contiguous()
    lab = labels.to(device=dev, dtype=torch.int64).contiguous()
    emb = track_feats.float().contiguous()
    if m:
        ml, me, mi = (memo_labels.to(device=dev, dtype=torch.int64).contiguous(), memo_embeds.to(dev).float().contiguous(),
                      memo_ids.to(device=dev, dtype=torch.int64).contiguous())
    else:
        ml = me = mi = None
    thr = torch.tensor([obj_score_thr, match_score_thr, init_score_thr, nms_conf_thr, nms_backdrop_iou_thr, nms_class_iou_thr],
                       dtype=torch.float32)
    import ctypes as C
    thr_arr = (C.c_float * 6)(*thr.tolist())
    sel = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    ids = torch.full((max(n, 1),), -1, dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = torch.empty(max(2 * n * m, 2), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().vkn_track_match(_lib.ptr(bb), _lib.ptr(lab), _lib.ptr(emb), n, D, _lib.ptr(ml), _lib.ptr(me), _lib.ptr(mi), m,
                                          thr_arr, int(bool(with_cats)), int(num_tracklets), _lib.ptr(sel), _lib.ptr(ids),
                                          _lib.ptr(counts), _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr(dev)))
    nk, nnew = (int(v) for v in counts.cpu())
    return sel[:nk].long(), ids[:nk], nnew


def match_cost(mask_logits, cls_logits, gt_masks, gt_labels, w_cls=2.0, w_mask=1.0, w_dice=4.0, dice_eps=1e-3, focal_alpha=0.25,
               focal_gamma=2.0, focal_eps=1e-12):
    """Cost matrix [N, M] of MaskHungarianAssigner.assign for one image (knet/det/mask_hungarian_assigner.py:228-247; DiceCost
    :43-75, MaskCost :93-110 with pred_act=True / act_mode='sigmoid', mmdet FocalLossCost) in one pass over the masks
    (vkn_match_cost).  mask_logits [N, H, W], cls_logits [N, ncls] or None, gt_masks [M, H, W] (float), gt_labels [M].
    The default weights are the shipped configs' (cls 2, mask 1, dice 4); a zero weight skips the term."""
    import ctypes as C
    if not mask_logits.is_cuda:
        raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
    dev = mask_logits.device
    N, M = mask_logits.shape[0], gt_masks.shape[0]
    if N == 0 or M == 0:
        return torch.zeros(N, M, dtype=torch.float32, device=dev)
    if tuple(mask_logits.shape[1:]) != tuple(gt_masks.shape[1:]):
        raise _lib.VknError('mask_logits %s and gt_masks %s differ in size' % (tuple(mask_logits.shape), tuple(gt_masks.shape)))
    HW = int(mask_logits[0].numel())
    ml = mask_logits.reshape(N, HW).float().contiguous()
    gm = gt_masks.to(dev).reshape(M, HW).float().contiguous()
    use_cls = cls_logits is not None and w_cls != 0
    cl = cls_logits.float().contiguous() if use_cls else None
    ncls = cl.shape[1] if use_cls else 0
    gl = gt_labels.to(device=dev, dtype=torch.int64).contiguous() if use_cls else None
    if use_cls and (cl.shape[0] != N or gl.numel() != M):
        raise _lib.VknError('cls_logits / gt_labels do not match the mask counts')
    L = _lib.lib()
    nb = C.c_size_t(0)
    _lib.check(L.vkn_match_cost_workspace_bytes(N, M, HW, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    cost = torch.empty(N, M, dtype=torch.float32, device=dev)
    params = (C.c_float * 7)(w_cls if use_cls else 0.0, w_mask, w_dice, dice_eps, focal_alpha, focal_gamma, focal_eps)
    with torch.cuda.device(dev):
        _lib.check(L.vkn_match_cost(_lib.ptr(ml), _lib.ptr(cl), _lib.ptr(gm), _lib.ptr(gl), N, M, HW, ncls, params, _lib.ptr(cost),
                                    _lib.ptr(ws), nb.value, _lib.stream_ptr(dev)))
    return cost
