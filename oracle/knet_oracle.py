"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the KernelUpdateHead hot path.

A plain-torch (CPU) restatement of the reference algorithm, written as explicit
math on a `state_dict`, with every function citing the reference lines it
follows.  It is the CHECKER for the CUDA path: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
package (video-k-net_b200/vknet) never imports anything from oracle/.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c), so this restatement is pinned against the reference
code ITSELF, imported verbatim through oracle/ref_shim.py:
  * tests/test_oracle_vs_reference.py (CPU, this container) runs both on the same
    seeded inputs and demands agreement to fp32 round-off;
  * oracle/make_golden.py stores reference outputs as fixtures under
    tests/golden/, which travel to the GPU box where /root/reference is absent.

All tensors keep the reference layouts: x [B,C,H,W], proposal_feat [B,N,C,K,K],
mask_preds [B,N,H,W].  Computation dtype follows the input dtype (fp32 or fp64;
fp64 gives a "truth" to separate our error from the reference's own round-off).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# primitives (third-party semantics restated)
# ----------------------------------------------------------------------------------------
def linear(x, w, b=None):
    y = x.matmul(w.t().to(x.dtype))
    return y if b is None else y + b.to(x.dtype)


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm(C), eps 1e-5, biased variance (mmcv build_norm_layer(dict(type='LN')))."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w.to(x.dtype) + b.to(x.dtype)


def mha_core(query, key, value, sd, prefix, num_heads):
    """torch nn.MultiheadAttention forward, seq-first [L,B,E], dropout 0, no masks.
    Packed in_proj [3E,E]; q scaled by 1/sqrt(head_dim); softmax over keys; out_proj.
    (the math mmcv MultiheadAttention delegates to: knet/det/kernel_update_head.py:100-101, 206)"""
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    w = sd[prefix + 'attn.in_proj_weight']
    b = sd[prefix + 'attn.in_proj_bias']
    q = linear(query, w[:E], b[:E])
    k = linear(key, w[E:2 * E], b[E:2 * E])
    v = linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(L, B * num_heads, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(S, B * num_heads, hd).transpose(0, 1)
    v = v.reshape(S, B * num_heads, hd).transpose(0, 1)
    att = torch.softmax(q.bmm(k.transpose(1, 2)), dim=-1)
    out = att.bmm(v).transpose(0, 1).reshape(L, B, E)
    return linear(out, sd[prefix + 'attn.out_proj.weight'], sd[prefix + 'attn.out_proj.bias'])


def mha_block(query, key, value, identity, sd, prefix, num_heads):
    """mmcv MultiheadAttention wrapper: identity + attn(q,k,v) (Appendix B)."""
    return identity + mha_core(query, key, value, sd, prefix, num_heads)


def ffn_block(x, sd, prefix):
    """mmcv FFN(num_fcs=2, ReLU, add_identity=True): x + W2 relu(W1 x + b1) + b2
    (knet/det/kernel_update_head.py:119-126, 214-215)."""
    h = torch.relu(linear(x, sd[prefix + 'layers.0.0.weight'], sd[prefix + 'layers.0.0.bias']))
    return x + linear(h, sd[prefix + 'layers.1.weight'], sd[prefix + 'layers.1.bias'])


# ----------------------------------------------------------------------------------------
# KernelUpdator  (knet/kernel_updator.py:56-94)
# ----------------------------------------------------------------------------------------
def kernel_updator(sd, prefix, update_feature, input_feature, in_channels, feat_channels,
                   gate_sigmoid=True, gate_norm_act=False, activate_out=False):
    p = prefix
    update_feature = update_feature.reshape(-1, in_channels)                       # :57
    num_proposals = update_feature.shape[0]
    parameters = linear(update_feature, sd[p + 'dynamic_layer.weight'], sd[p + 'dynamic_layer.bias'])  # :59
    param_in = parameters[:, :feat_channels]                                       # :60-61
    param_out = parameters[:, -feat_channels:]                                     # :62-63
    input_feats = linear(input_feature.reshape(num_proposals, -1, feat_channels),  # :65-66
                         sd[p + 'input_layer.weight'], sd[p + 'input_layer.bias'])
    input_in = input_feats[..., :feat_channels]                                    # :67
    input_out = input_feats[..., -feat_channels:]                                  # :68
    gate_feats = input_in * param_in.unsqueeze(-2)                                 # :70
    if gate_norm_act:                                                              # :71-72
        gate_feats = torch.relu(layer_norm(gate_feats, sd[p + 'gate_norm.weight'], sd[p + 'gate_norm.bias']))
    input_gate = layer_norm(linear(gate_feats, sd[p + 'input_gate.weight'], sd[p + 'input_gate.bias']),
                            sd[p + 'input_norm_in.weight'], sd[p + 'input_norm_in.bias'])   # :74
    update_gate = layer_norm(linear(gate_feats, sd[p + 'update_gate.weight'], sd[p + 'update_gate.bias']),
                             sd[p + 'norm_in.weight'], sd[p + 'norm_in.bias'])               # :75
    if gate_sigmoid:                                                               # :76-78
        input_gate = torch.sigmoid(input_gate)
        update_gate = torch.sigmoid(update_gate)
    param_out = layer_norm(param_out, sd[p + 'norm_out.weight'], sd[p + 'norm_out.bias'])            # :79
    input_out = layer_norm(input_out, sd[p + 'input_norm_out.weight'], sd[p + 'input_norm_out.bias'])  # :80
    if activate_out:                                                               # :82-84
        param_out = torch.relu(param_out)
        input_out = torch.relu(input_out)
    features = update_gate * param_out.unsqueeze(-2) + input_gate * input_out      # :87-88
    features = linear(features, sd[p + 'fc_layer.weight'], sd[p + 'fc_layer.bias'])  # :90
    features = layer_norm(features, sd[p + 'fc_norm.weight'], sd[p + 'fc_norm.bias'])  # :91
    return torch.relu(features)                                                    # :92


# ----------------------------------------------------------------------------------------
# shared pieces of the stage
# ----------------------------------------------------------------------------------------
def hard_mask(mask_preds, thr=0.5):
    """knet/det/kernel_update_head.py:190-192 -- 1[sigmoid(m) > thr] as float."""
    return (torch.sigmoid(mask_preds) > thr).to(mask_preds.dtype)


def _feat_transform(sd, x):
    """ConvModule 1x1 + bias, no norm/act (knet/det/kernel_update_head.py:107-117, 179-180)."""
    if 'feat_transform.conv.weight' not in sd:
        return x
    w = sd['feat_transform.conv.weight'].to(x.dtype)
    assert w.shape[-1] == 1 and w.shape[-2] == 1, 'only the 1x1 feat_transform is on the shipped path'
    return F.conv2d(x, w, sd['feat_transform.conv.bias'].to(x.dtype))   # the op nn.Conv2d dispatches to


def _heads_and_conv(sd, cfg, obj_feat, x, B, N):
    """cls/mask FC stacks + dynamic mask conv (knet/det/kernel_update_head.py:217-260)."""
    C = cfg['in_channels']
    K = cfg.get('conv_kernel_size', 1)
    cls_feat = obj_feat.sum(-2)                                                    # :217
    mask_feat = obj_feat                                                           # :218
    for i in range(cfg.get('num_cls_fcs', 1)):                                     # :220-221
        cls_feat = torch.relu(layer_norm(linear(cls_feat, sd['cls_fcs.%d.weight' % (3 * i)]),
                                         sd['cls_fcs.%d.weight' % (3 * i + 1)], sd['cls_fcs.%d.bias' % (3 * i + 1)]))
    for i in range(cfg.get('num_mask_fcs', 1)):                                    # :222-223
        mask_feat = torch.relu(layer_norm(linear(mask_feat, sd['mask_fcs.%d.weight' % (3 * i)]),
                                          sd['mask_fcs.%d.weight' % (3 * i + 1)], sd['mask_fcs.%d.bias' % (3 * i + 1)]))
    cls_score = linear(cls_feat, sd['fc_cls.weight'], sd['fc_cls.bias']).view(B, N, -1)  # :225
    mask_feat = linear(mask_feat, sd['fc_mask.weight'], sd['fc_mask.bias']).permute(0, 1, 3, 2)  # :227
    H, W = x.shape[-2:]
    mask_feat = mask_feat.reshape(B, N, C, K, K)                                   # :244-246
    new_mask_preds = torch.cat([F.conv2d(x[i:i + 1], mask_feat[i], padding=K // 2)  # :251-259
                                for i in range(B)], dim=0).reshape(B, N, H, W)
    return cls_score, new_mask_preds


def _pool(sd, cfg, x, proposal_feat, mask_preds):
    """feat_transform, hard mask, einsum pooling, proposal reshape
    (knet/det/kernel_update_head.py:178-200)."""
    B, N = proposal_feat.shape[:2]
    C = cfg['in_channels']
    x = _feat_transform(sd, x)
    H, W = x.shape[-2:]
    if mask_preds.shape[-2:] != (H, W):                                            # :183-188
        gather_mask = F.interpolate(mask_preds, (H, W), align_corners=False, mode='bilinear')
    else:
        gather_mask = mask_preds
    m = hard_mask(gather_mask, cfg.get('hard_mask_thr', 0.5))
    x_feat = torch.einsum('bnhw,bchw->bnc', m.to(x.dtype), x)                      # :195
    proposal_feat = proposal_feat.reshape(B, N, C, -1).permute(0, 1, 3, 2)         # :198-200
    return x, x_feat, proposal_feat


def _update_attend_ffn(sd, cfg, x_feat, proposal_feat, B, N):
    """KernelUpdator -> MHSA+LN -> FFN+LN (knet/det/kernel_update_head.py:201-215)."""
    C = cfg['in_channels']
    ku = cfg['kernel_updator_cfg']
    obj_feat = kernel_updator(sd, 'kernel_update_conv.', x_feat, proposal_feat,
                              ku.get('in_channels', 256), ku.get('feat_channels', 64),
                              ku.get('gate_sigmoid', True), ku.get('gate_norm_act', False),
                              ku.get('activate_out', False))                       # :201
    obj_feat = obj_feat.reshape(B, N, -1).permute(1, 0, 2)                         # :204-205
    obj_feat = layer_norm(mha_block(obj_feat, obj_feat, obj_feat, obj_feat, sd, 'attention.',
                                    cfg.get('num_heads', 8)),
                          sd['attention_norm.weight'], sd['attention_norm.bias'])  # :206
    obj_feat = obj_feat.permute(1, 0, 2).reshape(B, N, -1, C)                      # :208-211
    if cfg.get('with_ffn', True):                                                  # :214-215
        obj_feat = layer_norm(ffn_block(obj_feat, sd, 'ffn.'), sd['ffn_norm.weight'], sd['ffn_norm.bias'])
    return obj_feat


# ----------------------------------------------------------------------------------------
# KernelUpdateHead.forward  (knet/det/kernel_update_head.py:170-277)
# ----------------------------------------------------------------------------------------
def kernel_update_head_forward(sd, cfg, x, proposal_feat, mask_preds):
    B, N = proposal_feat.shape[:2]
    C = cfg['in_channels']
    K = cfg.get('conv_kernel_size', 1)
    x, x_feat, pf = _pool(sd, cfg, x, proposal_feat, mask_preds)
    obj_feat = _update_attend_ffn(sd, cfg, x_feat, pf, B, N)
    cls_score, new_mask_preds = _heads_and_conv(sd, cfg, obj_feat, x, B, N)
    return cls_score, new_mask_preds, obj_feat.permute(0, 1, 3, 2).reshape(B, N, C, K, K)  # :275-277


# ----------------------------------------------------------------------------------------
# VideoKernelUpdateHead.forward  (knet/video/kernel_update_head.py:281-541)
# ----------------------------------------------------------------------------------------
def _cross_link(sd, cfg, cur, prev, attn_prefix, norm_prefix, ffn_prefix, ffn_norm_prefix, B, N):
    """LN(cur + MHA(q=cur, k=v=prev)) then LN(FFN(.)) -- the block shared by the three link
    variants (knet/video/kernel_update_head.py:337-348, 404-415, 432-444)."""
    C = cfg['in_channels']
    q = cur.reshape(B, N, -1).permute(1, 0, 2)
    kv = prev.reshape(B, N, -1).permute(1, 0, 2)
    t = layer_norm(mha_block(q, kv, kv, q, sd, attn_prefix, 8), sd[norm_prefix + 'weight'], sd[norm_prefix + 'bias'])
    t = t.permute(1, 0, 2).reshape(B, N, -1, C)
    return layer_norm(ffn_block(t, sd, ffn_prefix), sd[ffn_norm_prefix + 'weight'], sd[ffn_norm_prefix + 'bias'])


def video_kernel_update_head_forward(sd, cfg, x, proposal_feat, mask_preds, previous_obj_feats=None):
    """Returns the reference 5-tuple (cls_score, new_mask_preds, obj_feat, x_feat, obj_feat_track|None).
    Link variants restated: previous_link='update_dynamic_cov' (:324-348),
    previous_type='ffn' (:394-415), previous_type='update' (:417-444)."""
    B, N = proposal_feat.shape[:2]
    C = cfg['in_channels']
    K = cfg.get('conv_kernel_size', 1)
    ku = cfg['kernel_updator_cfg']
    ku_args = (ku.get('in_channels', 256), ku.get('feat_channels', 64))
    x, x_feat, pf = _pool(sd, cfg, x, proposal_feat, mask_preds)                   # :293-317
    prev = None
    if previous_obj_feats is not None:
        prev = previous_obj_feats.reshape(B, N, C, -1).permute(0, 1, 3, 2)
    if prev is not None and cfg.get('previous_link') == 'update_dynamic_cov':      # :324-348
        prev_upd = kernel_updator(sd, 'attention_previous_update_link.', x_feat, prev, *ku_args)
        pf = _cross_link(sd, cfg, pf, prev_upd, 'attention_previous_link.', 'attention_previous_norm_link.',
                         'link_ffn_link.', 'link_ffn_norm_link.', B, N)
    obj_feat = _update_attend_ffn(sd, cfg, x_feat, pf, B, N)                       # :372-386
    track = None
    if prev is not None:
        if cfg.get('previous_type') == 'ffn':                                      # :394-415
            track = _cross_link(sd, cfg, obj_feat, prev, 'attention_previous.', 'attention_previous_norm.',
                                'link_ffn.', 'link_ffn_norm.', B, N)
        elif cfg.get('previous_type') == 'update':                                 # :417-444
            prev_trk = kernel_updator(sd, 'attention_previous_update_track.', x_feat, prev, *ku_args)
            # the reference reshapes [B*N,1,C] -> [B,N,C,-1] -> permute(0,1,3,2) -> [B,N,C] (:422-429): a
            # no-op relabelling for K=1, reproduced literally here.
            prev_trk = prev_trk.reshape(B, N, C, -1).permute(0, 1, 3, 2)
            track = _cross_link(sd, cfg, obj_feat, prev_trk, 'attention_previous_track.',
                                'attention_previous_norm_track.', 'link_ffn_track.', 'link_ffn_norm_track.', B, N)
    cls_score, new_mask_preds = _heads_and_conv(sd, cfg, obj_feat, x, B, N)        # :477-532
    obj_out = obj_feat.permute(0, 1, 3, 2).reshape(B, N, C, K, K)
    if track is not None:
        track = track.permute(0, 1, 3, 2).reshape(B, N, C, K, K)
    return cls_score, new_mask_preds, obj_out, x_feat, track                       # :534-541


# ----------------------------------------------------------------------------------------
# KernelUpdateHeadVideo.forward  (knet_vis/tracker/kernel_update_head.py:209-374), query_merge_method='mean'
# ----------------------------------------------------------------------------------------
def kernel_update_head_video_forward(sd, cfg, x, proposal_feat, mask_preds):
    """x [B,F,C,H,W], mask_preds [B,F,N,H,W].  proposal_feat 5-D [B,N,C,1,1] -> gathered mode
    (with_cls=True): (cls [B,N,ncls], masks [B,F,N,H,W], obj [B,N,C,1,1]); 6-D [B,F,N,C,1,1] -> per-frame
    mode (with_cls=False): (None, masks, obj [B,F,N,C,1,1])."""
    B, Fr, C, H, W = x.shape
    gathered = proposal_feat.dim() != 6                                            # :217-225
    N = proposal_feat.shape[1] if gathered else proposal_feat.shape[2]
    xt = _feat_transform(sd, x.reshape(B * Fr, C, H, W)).reshape(B, Fr, C, H, W)   # :227-228
    m = hard_mask(mask_preds, cfg.get('hard_mask_thr', 0.5)).to(xt.dtype)          # :238-240
    x_feat = torch.einsum('bfnhw,bfchw->bfnc', m, xt)                              # :246 / :266
    if gathered:
        x_feat = x_feat.mean(1)                                                    # :246
        pf = proposal_feat.reshape(B, N, C, -1).permute(0, 1, 3, 2)                # :270
        sets = B
    else:
        x_feat = x_feat.reshape(B * Fr, N, C)                                      # :275
        pf = proposal_feat.reshape(B * Fr, N, C, -1).permute(0, 1, 3, 2)           # :274
        sets = B * Fr
    obj_feat = _update_attend_ffn(sd, cfg, x_feat, pf, sets, N)                    # :271-291
    mask_feat = obj_feat
    cls_score = None
    if gathered:                                                                   # :295-301
        cls_feat = obj_feat.sum(-2)
        for i in range(cfg.get('num_cls_fcs', 1)):
            cls_feat = torch.relu(layer_norm(linear(cls_feat, sd['cls_fcs.%d.weight' % (3 * i)]),
                                             sd['cls_fcs.%d.weight' % (3 * i + 1)], sd['cls_fcs.%d.bias' % (3 * i + 1)]))
        cls_score = linear(cls_feat, sd['fc_cls.weight'], sd['fc_cls.bias']).view(sets, N, -1)
    for i in range(cfg.get('num_mask_fcs', 1)):                                    # :303-304
        mask_feat = torch.relu(layer_norm(linear(mask_feat, sd['mask_fcs.%d.weight' % (3 * i)]),
                                          sd['mask_fcs.%d.weight' % (3 * i + 1)], sd['mask_fcs.%d.bias' % (3 * i + 1)]))
    mask_feat = linear(mask_feat, sd['fc_mask.weight'], sd['fc_mask.bias']).permute(0, 1, 3, 2).reshape(sets, N, C, 1, 1)
    if gathered:                                                                   # :329-339
        new_mask = torch.stack([F.conv2d(xt[i], mask_feat[i]) for i in range(B)], dim=0)
        return cls_score, new_mask, obj_feat.permute(0, 1, 3, 2).reshape(B, N, C, 1, 1)
    new_mask = torch.cat([F.conv2d(xt[i][j][None], mask_feat[i * Fr + j])          # :340-352
                          for i in range(B) for j in range(Fr)], dim=0).reshape(B, Fr, N, H, W)
    return None, new_mask, obj_feat.permute(0, 1, 3, 2).reshape(B, Fr, N, C, 1, 1)


# ----------------------------------------------------------------------------------------
# next row: tail of ConvKernelHead._decode_init_proposals  (knet/det/kernel_head.py:196-265)
# ----------------------------------------------------------------------------------------
def init_proposals(init_w, init_b, loc_feats, x_feats, use_binary=True):
    """init_w [N,C,1,1], init_b [N] | None, loc_feats / x_feats [B,C,H,W].
    -> (proposal_feats [B,N,C,1,1], mask_preds [B,N,H,W]) for proposal_feats_with_obj=True."""
    mask_preds = F.conv2d(loc_feats, init_w.to(loc_feats.dtype), None if init_b is None else init_b.to(loc_feats.dtype))  # :212
    sigmoid_masks = mask_preds.sigmoid()                                           # :241
    nonzero = sigmoid_masks > 0.5                                                  # :242
    m = nonzero.to(x_feats.dtype) if use_binary else nonzero.to(x_feats.dtype) * sigmoid_masks   # :243-246
    obj_feats = torch.einsum('bnhw,bchw->bnc', m, x_feats)                         # :247
    B, N = mask_preds.shape[:2]
    proposal_feats = init_w[None].expand(B, *init_w.shape) + obj_feats.view(B, N, -1, 1, 1)     # :236-238, 252-254
    return proposal_feats, mask_preds


# ----------------------------------------------------------------------------------------
# the S-stage loop  (knet/det/kernel_iter_head.py:118-137, 246-253; forward_dummy :317-330)
# ----------------------------------------------------------------------------------------
def iter_forward(sds, cfgs, x, proposal_feat, mask_preds, mask_round=None):
    """Chains the stages exactly as KernelIterHead.simple_test does.  `mask_round`, when given, is
    applied to each stage's new_mask_preds (e.g. a bf16 round-trip: what a bf16 module would hand
    to the next stage); returns the per-stage outputs."""
    outs = []
    obj = proposal_feat
    for sd, cfg in zip(sds, cfgs):
        cls_score, mask_preds, obj = kernel_update_head_forward(sd, cfg, x, obj, mask_preds)
        if mask_round is not None:
            mask_preds = mask_round(mask_preds)
        outs.append((cls_score, mask_preds, obj))
    return outs


def dummy_inputs(B, N, C, H, W, seed=1, dtype=torch.float32):
    """forward_dummy recipe (knet/det/kernel_iter_head.py:317-330): x, proposal_feats ~ N(0,1),
    mask_preds = proposal_feats @ x."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float32)
    pf = torch.randn(B, N, C, 1, 1, generator=g, dtype=torch.float32)
    mask = pf.view(B, N, C).bmm(x.view(B, C, -1)).view(B, N, H, W)
    return x.to(dtype), pf.to(dtype), mask.to(dtype)


# ----------------------------------------------------------------------------------------
# weights: same shapes / init as the reference modules (kernel_update_head.py:151-168)
# ----------------------------------------------------------------------------------------
def default_cfg(num_classes=19, in_channels=256, feedforward_channels=2048, num_heads=8, **over):
    """The mask_head block of configs/det/_base_/models/knet_kitti_step_s3_r50_fpn.py:88-136."""
    cfg = dict(num_classes=num_classes, num_ffn_fcs=2, num_heads=num_heads, num_cls_fcs=1, num_mask_fcs=1,
               feedforward_channels=feedforward_channels, in_channels=in_channels, out_channels=in_channels,
               dropout=0.0, mask_thr=0.5, conv_kernel_size=1, mask_upsample_stride=2,
               ffn_act_cfg=dict(type='ReLU', inplace=True), with_ffn=True,
               feat_transform_cfg=dict(conv_cfg=dict(type='Conv2d'), act_cfg=None),
               kernel_updator_cfg=dict(type='KernelUpdator', in_channels=in_channels, feat_channels=in_channels,
                                       out_channels=in_channels, input_feat_shape=3,
                                       act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN')),
               loss_rank=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.1),
               loss_mask=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
               loss_dice=dict(type='DiceLoss', loss_weight=4.0),
               loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0))
    cfg.update(over)
    return cfg


def _xavier(g, *shape):
    fan_out, fan_in = shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    a = math.sqrt(6.0 / ((fan_in + fan_out) * rf))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def _updator_params(g, sd, p, C, Fc, out):
    sd[p + 'dynamic_layer.weight'] = _xavier(g, 2 * Fc, C)
    sd[p + 'dynamic_layer.bias'] = torch.randn(2 * Fc, generator=g) * 0.05
    sd[p + 'input_layer.weight'] = _xavier(g, 2 * Fc, C)
    sd[p + 'input_layer.bias'] = torch.randn(2 * Fc, generator=g) * 0.05
    for nm in ('input_gate', 'update_gate'):
        sd[p + nm + '.weight'] = _xavier(g, Fc, C)
        sd[p + nm + '.bias'] = torch.randn(Fc, generator=g) * 0.05
    for nm in ('norm_in', 'norm_out', 'input_norm_in', 'input_norm_out'):
        sd[p + nm + '.weight'] = 1 + 0.1 * torch.randn(Fc, generator=g)
        sd[p + nm + '.bias'] = 0.1 * torch.randn(Fc, generator=g)
    sd[p + 'fc_layer.weight'] = _xavier(g, out, Fc)
    sd[p + 'fc_layer.bias'] = torch.randn(out, generator=g) * 0.05
    sd[p + 'fc_norm.weight'] = 1 + 0.1 * torch.randn(out, generator=g)
    sd[p + 'fc_norm.bias'] = 0.1 * torch.randn(out, generator=g)


def _mha_params(g, sd, p, E):
    sd[p + 'attn.in_proj_weight'] = _xavier(g, 3 * E, E)
    sd[p + 'attn.in_proj_bias'] = torch.randn(3 * E, generator=g) * 0.05
    sd[p + 'attn.out_proj.weight'] = _xavier(g, E, E)
    sd[p + 'attn.out_proj.bias'] = torch.randn(E, generator=g) * 0.05


def _ln_params(g, sd, p, C):
    sd[p + 'weight'] = 1 + 0.1 * torch.randn(C, generator=g)
    sd[p + 'bias'] = 0.1 * torch.randn(C, generator=g)


def _ffn_params(g, sd, p, C, Fh):
    sd[p + 'layers.0.0.weight'] = _xavier(g, Fh, C)
    sd[p + 'layers.0.0.bias'] = torch.randn(Fh, generator=g) * 0.05
    sd[p + 'layers.1.weight'] = _xavier(g, C, Fh)
    sd[p + 'layers.1.bias'] = torch.randn(C, generator=g) * 0.05


def random_state_dict(cfg, seed=0):
    """A state_dict with the reference's keys and shapes.  Matrices are
    xavier-uniform like init_weights (:151-168); biases / LN affine are perturbed away from the
    0 / 1 defaults so that parity tests exercise every term (a freshly initialised reference
    module has all-zero biases, which would hide a dropped bias)."""
    g = torch.Generator().manual_seed(seed)
    C = cfg['in_channels']
    Fh = cfg['feedforward_channels']
    ku = cfg['kernel_updator_cfg']
    sd = {}
    _mha_params(g, sd, 'attention.', C)
    _ln_params(g, sd, 'attention_norm.', C)
    _updator_params(g, sd, 'kernel_update_conv.', ku['in_channels'], ku['feat_channels'], ku['out_channels'])
    if cfg.get('feat_transform_cfg') is not None:
        sd['feat_transform.conv.weight'] = _xavier(g, C, C, 1, 1)
        sd['feat_transform.conv.bias'] = torch.randn(C, generator=g) * 0.05
    _ffn_params(g, sd, 'ffn.', C, Fh)
    _ln_params(g, sd, 'ffn_norm.', C)
    sd['cls_fcs.0.weight'] = _xavier(g, C, C)
    _ln_params(g, sd, 'cls_fcs.1.', C)
    sd['fc_cls.weight'] = _xavier(g, cfg['num_classes'], C)
    sd['fc_cls.bias'] = torch.full((cfg['num_classes'],), -math.log(99.0)) + 0.05 * torch.randn(cfg['num_classes'], generator=g)
    sd['mask_fcs.0.weight'] = _xavier(g, C, C)
    _ln_params(g, sd, 'mask_fcs.1.', C)
    sd['fc_mask.weight'] = _xavier(g, cfg['out_channels'], C)
    sd['fc_mask.bias'] = torch.randn(cfg['out_channels'], generator=g) * 0.05
    if cfg.get('previous') is not None:
        if cfg.get('previous_type') == 'ffn':
            _mha_params(g, sd, 'attention_previous.', C)
            _ln_params(g, sd, 'attention_previous_norm.', C)
            _ffn_params(g, sd, 'link_ffn.', C, Fh)
            _ln_params(g, sd, 'link_ffn_norm.', C)
        elif cfg.get('previous_type') == 'update':
            _updator_params(g, sd, 'attention_previous_update_track.', ku['in_channels'], ku['feat_channels'], ku['out_channels'])
            _mha_params(g, sd, 'attention_previous_track.', C)
            _ln_params(g, sd, 'attention_previous_norm_track.', C)
            _ffn_params(g, sd, 'link_ffn_track.', C, Fh)
            _ln_params(g, sd, 'link_ffn_norm_track.', C)
        if cfg.get('previous_link') == 'update_dynamic_cov':
            _updator_params(g, sd, 'attention_previous_update_link.', ku['in_channels'], ku['feat_channels'], ku['out_channels'])
            _mha_params(g, sd, 'attention_previous_link.', C)
            _ln_params(g, sd, 'attention_previous_norm_link.', C)
            _ffn_params(g, sd, 'link_ffn_link.', C, Fh)
            _ln_params(g, sd, 'link_ffn_norm_link.', C)
    return sd


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def round_state_dict_bf16(sd):
    return {k: round_bf16(v) for k, v in sd.items()}


# ---- post-loop mask path ------------------------------------------------------------------
def rescale_masks(masks, img_meta, mask_upsample_stride=1):
    """Last-stage upsample of _mask_forward (knet/det/kernel_iter_head.py:122-128) followed by
    KernelUpdateHead.rescale_masks (knet/det/kernel_update_head.py:443-458).  masks [K,H,W] logits -> [K,ori_h,ori_w]."""
    m = masks.float().unsqueeze(0)
    if mask_upsample_stride > 1:
        m = F.interpolate(m, scale_factor=mask_upsample_stride, align_corners=False, mode='bilinear')
    h, w = img_meta['img_shape'][:2]
    m = F.interpolate(m.sigmoid(), size=tuple(img_meta['batch_input_shape'][:2]), mode='bilinear', align_corners=False)
    m = m[:, :, :h, :w]
    return F.interpolate(m, size=tuple(img_meta['ori_shape'][:2]), mode='bilinear', align_corners=False).squeeze(0)


def panoptic_merge_joint(thing_masks, thing_labels, thing_scores, stuff_masks, stuff_labels, stuff_scores, num_thing_classes,
                         instance_score_thr, overlap_thr):
    """Score-weighted argmax merge of thing and stuff probability maps into one id map
    (VideoKernelIterHead.merge_stuff_thing_stuff_joint, knet/video/kernel_iter_head.py:818-882; the shipped video configs set
    merge_joint=True).  masks [K,H,W] / [M,H,W] probabilities, labels int, scores float.
    Returns (panoptic_seg int32 [H,W], segments_info list of dicts, kept thing indices into the concatenated list)."""
    masks = torch.cat([thing_masks, stuff_masks], 0).float()
    scores = torch.cat([thing_scores, stuff_scores], 0).float()
    labels = torch.cat([thing_labels, stuff_labels], 0)
    owner = (scores.view(-1, 1, 1) * masks).argmax(0)               # the highest score-weighted probability wins a pixel
    seg = torch.zeros(masks.shape[-2:], dtype=torch.int32)
    info, kept, next_id = [], [], 0
    for k in torch.argsort(-scores).tolist():
        cls = int(labels[k])
        thing = cls < num_thing_classes
        if thing and float(scores[k]) < instance_score_thr:
            continue
        won = owner == k
        area = int(won.sum())
        full = int((masks[k] >= 0.5).sum())
        if area == 0 or full == 0 or area / full < overlap_thr:
            continue
        next_id += 1
        seg[won] = next_id
        if thing:
            info.append(dict(id=next_id, isthing=True, score=float(scores[k]), category_id=cls, instance_id=k))
            kept.append(k)
        else:
            info.append(dict(id=next_id, isthing=False, category_id=cls - num_thing_classes + 1, area=area))
    return seg, info, kept


# ----------------------------------------------------------------------------------------
# tracking embeddings + association
# ----------------------------------------------------------------------------------------
def mlp(layers, x):
    """layers: list of (weight [out,in], bias | None, ln_weight | None, ln_bias | None, relu) -- the embedding stacks
    knet/video/knet_quansi_dense_embed_fc_joint_train.py:113-126, 572-580 and knet/video/track_heads.py:632-642."""
    for w, b, g, be, relu in layers:
        x = linear(x, w, b)
        if g is not None:
            x = layer_norm(x, g, be)
        if relu:
            x = torch.relu(x)
    return x


def bbox_iou(a, b, eps=1e-6):
    """mmdet.core.bbox_overlaps(mode='iou', is_aligned=False) (mmdet v2.18: no +1 on the extents, union clamped at eps)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    ov = wh[..., 0] * wh[..., 1]
    return ov / (area_a[:, None] + area_b[None, :] - ov).clamp(min=eps)


def tracker_match(bboxes, labels, track_feats, memo_labels, memo_embeds, memo_ids, num_tracklets, obj_score_thr=0.5,
                  match_score_thr=0.5, init_score_thr=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3, nms_class_iou_thr=0.7,
                  with_cats=True):
    """QuasiDenseEmbedTracker.match up to update_memo (knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py:137-204),
    match_metric='bisoftmax'.  Returns (selected indices in score order, ids, number of new tracks)."""
    n = bboxes.shape[0]
    order = sorted(range(n), key=lambda i: (-float(bboxes[i, 4]), i))                      # :139
    b = bboxes[order]
    ious = bbox_iou(b[:, :4], b[:, :4]) if n else b.new_zeros((0, 0))
    keep = [True] * n
    for i in range(1, n):                                                                   # :147-151
        thr = nms_backdrop_iou_thr if float(b[i, 4]) < obj_score_thr else nms_class_iou_thr
        if bool((ious[i, :i] > thr).any()):
            keep[i] = False
    sel = [order[i] for i in range(n) if keep[i]]
    b, lab, emb = bboxes[sel], labels[sel], track_feats[sel]
    ids = torch.full((len(sel),), -1, dtype=torch.long)
    m = 0 if memo_embeds is None else memo_embeds.shape[0]
    if len(sel) > 0 and m > 0:
        feats = emb @ memo_embeds.t()                                                        # :166-170
        scores = (feats.softmax(dim=1) + feats.softmax(dim=0)) / 2
        if with_cats:
            scores = scores * (lab.view(-1, 1) == memo_labels.view(1, -1)).float()           # :182-184
        for i in range(len(sel)):                                                            # :186-198
            conf, memo_ind = torch.max(scores[i, :], dim=0)
            tid = int(memo_ids[memo_ind])
            if float(conf) > match_score_thr and tid > -1:
                if float(b[i, 4]) > obj_score_thr:
                    ids[i] = tid
                    scores[:i, memo_ind] = 0
                    scores[i + 1:, memo_ind] = 0
                elif float(conf) > nms_conf_thr:
                    ids[i] = -2
    new = (ids == -1) & (b[:, 4] > init_score_thr) if len(sel) else torch.zeros(0, dtype=torch.bool)   # :199-204
    nnew = int(new.sum())
    ids[new] = torch.arange(num_tracklets, num_tracklets + nnew, dtype=torch.long)
    return torch.tensor(sel, dtype=torch.long), ids, nnew


# ---- row f4: cost matrix of MaskHungarianAssigner (training side) ------------------------------------------------------------
def focal_loss_cost(cls_pred, gt_labels, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
    """mmdet.core.bbox.match_costs.FocalLossCost.__call__ (mmdet v2.18, the version the reference README pins) -- third-party
    arithmetic restated: sigmoid, then pos_cost[:, labels] - neg_cost[:, labels]."""
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def dice_cost(mask_preds, gt_masks, weight=1.0, eps=1e-3):
    """DiceCost.__call__ with pred_act=True, act_mode='sigmoid' (knet/det/mask_hungarian_assigner.py:43-75)."""
    p = mask_preds.sigmoid().clamp(min=0.001, max=1.0)                         # :69
    inp = p.reshape(p.size(0), -1)                                            # :44-45
    tgt = gt_masks.reshape(gt_masks.size(0), -1).float()
    a = torch.einsum('nh,mh->nm', inp, tgt)                                   # :48
    b = torch.sum(inp * inp, 1) + eps                                         # :49
    c = torch.sum(tgt * tgt, 1) + eps                                         # :50
    return -((2 * a) / (b[:, None] + c[None, ...])) * weight                  # :51-53, :75


def mask_cost(mask_preds, gt_masks, weight=1.0):
    """MaskCost.__call__ with pred_act=True, act_mode='sigmoid' (knet/det/mask_hungarian_assigner.py:93-110)."""
    p = mask_preds.sigmoid().clamp(min=0.01, max=1.0)                          # :97
    _, H, W = gt_masks.shape                                                  # :101
    pos = torch.einsum('nhw,mhw->nm', p, gt_masks)                            # :104
    neg = torch.einsum('nhw,mhw->nm', 1 - p, 1 - gt_masks)                    # :105
    return (-(pos + neg) / (H * W)) * weight                                  # :109-110


def match_cost(mask_preds, cls_pred, gt_masks, gt_labels, w_cls=2.0, w_mask=1.0, w_dice=4.0):
    """The weighted cost MaskHungarianAssigner.assign hands to linear_sum_assignment (:228-247): cls + mask + dice."""
    cost = 0
    if w_cls != 0 and cls_pred is not None:
        cost = focal_loss_cost(cls_pred, gt_labels, w_cls)
    if w_mask != 0:
        cost = cost + mask_cost(mask_preds, gt_masks, w_mask)
    if w_dice != 0:
        cost = cost + dice_cost(mask_preds, gt_masks, w_dice)
    return cost
