"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libvknet.so), against
(a) the golden fixtures produced by the UNMODIFIED reference and (b) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): kernel tensors (obj_feat, cls_score, x_feat, link outputs)
within 1e-3 in fp32 / 1e-2 in bf16 storage; mask argmax over the N kernels identical to the
oracle's.  Logit maps are additionally checked to 1e-3 * max|logit| (they feed the next stage's
hard threshold).
"""
import numpy as np
import pytest
import torch

import knet_oracle as ko
from helpers import assert_masks_bf16, bf16_ulp, build_heads, golden_files, load_golden, maxabs, stagewise_vs_oracle_bf16, top2_gap
from vknet import _lib

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-3
TOL_BF16 = 1e-2


@pytest.fixture(scope='module')
def dev(built_lib):
    if not torch.cuda.is_available():
        pytest.skip('GPU tests need a CUDA device')
    return torch.device('cuda:0')


def assert_masks(got, ref, what, rel=1e-3):
    """logits close AND argmax over kernels identical (ties: pixels whose oracle top-2 gap is below the
    logit tolerance may legitimately resolve either way; they are counted and must stay rare)."""
    got = got.float().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, '%s: mask logits max-abs err %g (scale %g)' % (what, err, scale)
    if ref.shape[1] > 1:
        a, b = got.argmax(1), ref.argmax(1)
        bad = a != b
        if bad.any():
            gap = top2_gap(ref)[bad]
            assert gap.max().item() <= 2 * err + 1e-6 * scale, \
                '%s: %d argmax mismatches at non-tied pixels (max gap %g, logit err %g)' % (
                    what, int(bad.sum()), gap.max().item(), err)
            assert bad.float().mean().item() < 1e-3, '%s: too many tied-pixel argmax flips' % what


# ---- golden fixtures (outputs of the unmodified reference) ---------------------------------------
@pytest.mark.parametrize('path', golden_files('det_'), ids=lambda p: p.split('/')[-1][:-4])
def test_golden_det_chain(dev, path):
    g = load_golden(path)
    t = g['t']
    heads = build_heads('KernelUpdateHead', g['cfg'], g['sds'], dev)
    x, obj, m = t['x'].to(dev), t['proposal_feat'].to(dev), t['mask_preds'].to(dev)
    for s, h in enumerate(heads):
        cls, m, obj = h(x, obj, m)
        assert obj.shape == t['s%d.obj_feat' % s].shape and cls.shape == t['s%d.cls_score' % s].shape
        assert maxabs(cls, t['s%d.cls_score' % s]) < TOL_F32, 'stage %d cls_score' % s
        assert maxabs(obj, t['s%d.obj_feat' % s]) < TOL_F32, 'stage %d obj_feat' % s
        assert_masks(m, t['s%d.mask_preds' % s], 'stage %d' % s)


@pytest.mark.parametrize('path', golden_files('det_'), ids=lambda p: p.split('/')[-1][:-4])
def test_golden_det_stagewise(dev, path):
    """Each stage fed the REFERENCE's inputs for that stage (no error carry-over)."""
    g = load_golden(path)
    t = g['t']
    heads = build_heads('KernelUpdateHead', g['cfg'], g['sds'], dev)
    x = t['x'].to(dev)
    for s, h in enumerate(heads):
        obj_in = t['proposal_feat'] if s == 0 else t['s%d.obj_feat' % (s - 1)]
        m_in = t['mask_preds'] if s == 0 else t['s%d.mask_preds' % (s - 1)]
        cls, m, obj = h(x, obj_in.to(dev), m_in.to(dev))
        assert maxabs(cls, t['s%d.cls_score' % s]) < TOL_F32
        assert maxabs(obj, t['s%d.obj_feat' % s]) < TOL_F32
        assert_masks(m, t['s%d.mask_preds' % s], 'stage %d' % s)


@pytest.mark.parametrize('path', golden_files('video_'), ids=lambda p: p.split('/')[-1][:-4])
def test_golden_video_links(dev, path):
    g = load_golden(path)
    t = g['t']
    h = build_heads('VideoKernelUpdateHead', g['cfg'], g['sds'], dev)[0]
    x, pf, m, prev = (t[k].to(dev) for k in ('x', 'proposal_feat', 'mask_preds', 'previous_obj_feats'))
    cls, nm, obj, x_feat, track = h(x, pf, m, previous_obj_feats=prev)
    assert maxabs(x_feat, t['s0.x_feat']) < 1e-3 * max(1.0, t['s0.x_feat'].abs().max().item())
    assert maxabs(cls, t['s0.cls_score']) < TOL_F32
    assert maxabs(obj, t['s0.obj_feat']) < TOL_F32
    assert track is not None and track.shape == t['s0.obj_feat_track'].shape
    assert maxabs(track, t['s0.obj_feat_track']) < TOL_F32
    assert_masks(nm, t['s0.mask_preds'], 'linked stage')
    cls, nm, obj, x_feat, track = h(x, pf, m)          # no previous -> 5th output None (:540-541)
    assert track is None
    assert maxabs(cls, t['noprev.cls_score']) < TOL_F32
    assert maxabs(obj, t['noprev.obj_feat']) < TOL_F32
    assert_masks(nm, t['noprev.mask_preds'], 'unlinked stage')


# ---- clip head (knet_vis tracker): golden fixtures + larger clips on both engines --------------------------
@pytest.mark.parametrize('path', golden_files('clip_'), ids=lambda p: p.split('/')[-1][:-4])
def test_golden_clip_head(dev, path):
    import vknet
    g = load_golden(path)
    t = g['t']
    with_cls = 's0.cls_score' in t
    sd = g['sds'][0]
    h = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=with_cls, num_proposals=g['N'], **g['cfg']))
    h.load_state_dict(sd, strict=True)
    h = h.to(dev).eval()
    cls, nm, obj = h(t['x'].to(dev), t['proposal_feat'].to(dev), t['mask_preds'].to(dev))
    assert nm.shape == t['s0.mask_preds'].shape and obj.shape == t['s0.obj_feat'].shape
    assert maxabs(obj, t['s0.obj_feat']) < TOL_F32
    if with_cls:
        assert maxabs(cls, t['s0.cls_score']) < TOL_F32
    else:
        assert cls is None
    B, Fr = nm.shape[:2]
    assert_masks(nm.reshape(B * Fr, *nm.shape[2:]), t['s0.mask_preds'].reshape(B * Fr, *nm.shape[2:]), 'clip head')


@pytest.mark.parametrize('B,Fr,N,C,H,W,with_cls,dt', [(1, 4, 100, 256, 96, 160, True, 'bf16'), (2, 3, 100, 128, 16, 24, True, 'f32'),
                                                   (1, 4, 100, 256, 24, 40, False, 'bf16')])
def test_clip_head_vs_oracle(dev, B, Fr, N, C, H, W, with_cls, dt):
    """BASELINE cfg2-like clip (4 frames, 96x160) on the tcgen05 engine; fp32 storage on the SIMT engine."""
    import vknet
    cfg = ko.default_cfg(num_classes=40, in_channels=C, feedforward_channels=256)
    sd = ko.random_state_dict(cfg, seed=33)
    if dt == 'bf16':
        sd = ko.round_state_dict_bf16(sd)
    if not with_cls:
        sd = {k: v for k, v in sd.items() if not (k.startswith('cls_fcs') or k.startswith('fc_cls'))}
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, Fr, C, H, W, generator=gen)
    if with_cls:
        pf = torch.randn(B, N, C, 1, 1, generator=gen)
        mask = torch.einsum('bnc,bfchw->bfnhw', pf.view(B, N, C), x)
    else:
        pf = torch.randn(B, Fr, N, C, 1, 1, generator=gen)
        mask = torch.einsum('bfnc,bfchw->bfnhw', pf.view(B, Fr, N, C), x)
    if dt == 'bf16':
        x, mask = ko.round_bf16(x), ko.round_bf16(mask)
    want = ko.kernel_update_head_video_forward(sd, cfg, x, pf, mask)
    h = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=with_cls, num_proposals=N, **cfg))
    h.load_state_dict(sd, strict=True)
    tdt = torch.bfloat16 if dt == 'bf16' else torch.float32
    h = h.to(device=dev, dtype=tdt).eval()
    cls, nm, obj = h(x.to(dev).to(tdt), pf.to(dev), mask.to(dev).to(tdt))
    tol = TOL_BF16 if dt == 'bf16' else TOL_F32
    assert maxabs(obj, want[2]) < tol
    if with_cls:
        assert maxabs(cls, want[0]) < tol
    if dt == 'f32':
        assert_masks(nm.reshape(B * Fr, N, H, W), want[1].reshape(B * Fr, N, H, W), 'clip')
    else:
        assert_masks_bf16(nm.reshape(B * Fr, N, H, W), want[1].reshape(B * Fr, N, H, W), 'clip bf16')


# ---- individual operators against the oracle ------------------------------------------------------
@pytest.mark.parametrize('B,N,C,H,W,Fh', [(1, 10, 64, 8, 8, 64), (2, 37, 128, 7, 19, 96), (1, 100, 256, 24, 40, 2048)])
def test_operators(dev, B, N, C, H, W, Fh):
    from vknet import ops
    cfg = ko.default_cfg(num_classes=11, in_channels=C, feedforward_channels=Fh)
    sd = ko.random_state_dict(cfg, seed=5)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=9)
    g = torch.Generator().manual_seed(4)
    rows = torch.randn(B, N, C, generator=g)
    # a3+a4 pooling (with feat_transform)
    xt, x_feat, pf_r = ko._pool(sd, cfg, x, pf, mask)
    got = ops.mask_pool(h, x.to(dev), mask.to(dev))
    assert maxabs(got, x_feat) < 1e-4 * x_feat.abs().max().item(), 'mask_pool'
    # a5 KernelUpdator
    want = ko.kernel_updator(sd, 'kernel_update_conv.', x_feat, pf_r, C, C).reshape(B, N, C)
    assert maxabs(ops.kernel_update(h, x_feat.to(dev), pf.to(dev)), want) < 1e-4, 'kernel_update'
    # a6 MHSA + LN
    seq = rows.permute(1, 0, 2)
    want = ko.layer_norm(ko.mha_block(seq, seq, seq, seq, sd, 'attention.', 8), sd['attention_norm.weight'],
                         sd['attention_norm.bias']).permute(1, 0, 2)
    assert maxabs(ops.mhsa_ln(h, rows.to(dev)), want) < 1e-4, 'mhsa_ln'
    # a7 FFN + LN
    want = ko.layer_norm(ko.ffn_block(rows, sd, 'ffn.'), sd['ffn_norm.weight'], sd['ffn_norm.bias'])
    assert maxabs(ops.ffn_ln(h, rows.to(dev)), want) < 1e-4, 'ffn_ln'
    # a8 heads
    cls_f = torch.relu(ko.layer_norm(ko.linear(rows, sd['cls_fcs.0.weight']), sd['cls_fcs.1.weight'], sd['cls_fcs.1.bias']))
    mk_f = torch.relu(ko.layer_norm(ko.linear(rows, sd['mask_fcs.0.weight']), sd['mask_fcs.1.weight'], sd['mask_fcs.1.bias']))
    want_cls = ko.linear(cls_f, sd['fc_cls.weight'], sd['fc_cls.bias'])
    want_mk = ko.linear(mk_f, sd['fc_mask.weight'], sd['fc_mask.bias'])
    cls, mk = ops.heads(h, rows.to(dev))
    assert maxabs(cls, want_cls) < 1e-4 and maxabs(mk, want_mk) < 1e-4, 'heads'
    # a9 dynamic mask conv (with feat_transform folded in)
    want = torch.einsum('bnc,bchw->bnhw', want_mk, xt)
    assert_masks(ops.mask_gemm(h, x.to(dev), want_mk.to(dev)), want, 'mask_gemm', rel=1e-5)


# ---- shape sweep: real kernel counts, ragged H*W, frame batches ------------------------------------
@pytest.mark.parametrize('B,N,C,H,W', [(1, 10, 64, 64, 64), (3, 117, 64, 5, 13), (2, 166, 128, 11, 9),
                                       (1, 100, 256, 33, 17), (4, 100, 256, 12, 20)])
def test_stage_shapes(dev, B, N, C, H, W):
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256)
    sd = ko.random_state_dict(cfg, seed=B * 1000 + N)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=N)
    want = ko.kernel_update_head_forward(sd, cfg, x, pf, mask)
    cls, nm, obj = h(x.to(dev), pf.to(dev), mask.to(dev))
    assert cls.shape == want[0].shape and nm.shape == want[1].shape and obj.shape == want[2].shape
    assert maxabs(cls, want[0]) < TOL_F32 and maxabs(obj, want[2]) < TOL_F32
    assert_masks(nm, want[1], 'stage')


def test_variants_no_ffn_no_transform_more_fcs(dev):
    """with_ffn=False, feat_transform_cfg=None, num_mask_fcs=3 (the class default) and a non-default threshold."""
    cfg = ko.default_cfg(num_classes=5, in_channels=64, feedforward_channels=64, with_ffn=False,
                         feat_transform_cfg=None, num_mask_fcs=3, num_cls_fcs=2, hard_mask_thr=0.7)
    sd = ko.random_state_dict(cfg, seed=1)
    g = torch.Generator().manual_seed(8)
    for i in (1, 2):   # extra FC layers
        sd['mask_fcs.%d.weight' % (3 * i)] = ko._xavier(g, 64, 64)
        ko._ln_params(g, sd, 'mask_fcs.%d.' % (3 * i + 1), 64)
    sd['cls_fcs.3.weight'] = ko._xavier(g, 64, 64)
    ko._ln_params(g, sd, 'cls_fcs.4.', 64)
    for k in [k for k in sd if k.startswith('ffn')]:
        del sd[k]
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    x, pf, mask = ko.dummy_inputs(2, 14, 64, 9, 10, seed=2)
    want = ko.kernel_update_head_forward(sd, cfg, x, pf, mask)
    cls, nm, obj = h(x.to(dev), pf.to(dev), mask.to(dev))
    assert maxabs(cls, want[0]) < TOL_F32 and maxabs(obj, want[2]) < TOL_F32
    assert_masks(nm, want[1], 'variant stage')


# ---- bf16 storage: oracle = fp32 math on bf16-rounded x / masks / weights, masks rounded on output ----
@pytest.mark.parametrize('B,N,C,H,W,S', [(1, 20, 64, 16, 24, 2), (2, 100, 256, 40, 24, 2)])
def test_bf16_storage(dev, B, N, C, H, W, S):
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=512)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=70 + s)) for s in range(S)]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=6)
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    outs, _ = stagewise_vs_oracle_bf16(heads, sds, cfg, x.to(dev).bfloat16(), pf.to(dev), mask.to(dev).bfloat16(), 'bf16 storage')
    assert outs[-1][1].dtype == torch.bfloat16 and outs[-1][2].dtype == torch.float32


# ---- BASELINE.json cfg1 at full size: N=100, C=256, 200x88, S=3 -------------------------------------
def test_cfg1_full_size_loop(dev):
    import vknet
    B, N, C, H, W, S = 1, 100, 256, 200, 88, 3
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048)
    sds = [ko.random_state_dict(cfg, seed=s) for s in range(S)]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=1)
    want = ko.iter_forward(sds, [cfg] * S, x, pf, mask)
    heads = build_heads('KernelUpdateHead', cfg, sds, dev)
    # stage-wise with the oracle's inputs: isolates each stage from threshold flips upstream
    for s, h in enumerate(heads):
        obj_in = pf if s == 0 else want[s - 1][2]
        m_in = mask if s == 0 else want[s - 1][1]
        cls, m, obj = h(x.to(dev), obj_in.to(dev), m_in.to(dev))
        assert maxabs(cls, want[s][0]) < TOL_F32 and maxabs(obj, want[s][2]) < TOL_F32
        assert_masks(m, want[s][1], 'cfg1 stage %d' % s)
    # chained loop in one call == chained modules (same kernels, bit-identical)
    loop = vknet.KernelIterLoop(heads)
    cls_l, m_l, obj_l = loop(x.to(dev), pf.to(dev), mask.to(dev))
    obj_c, m_c = pf.to(dev), mask.to(dev)
    flips = []
    for s, h in enumerate(heads):
        cls_c, m_c, obj_c = h(x.to(dev), obj_c, m_c)
        flips.append(int(((m_c.float().cpu() > 0) != (want[s][1] > 0)).sum()))
    assert torch.equal(m_l, m_c) and torch.equal(obj_l, obj_c) and torch.equal(cls_l, cls_c)
    # end of the loop vs the oracle: threshold disagreements upstream are counted, not assumed zero ...
    print('cfg1 threshold disagreements per stage vs oracle:', flips)
    assert sum(flips[:-1]) <= 8, 'too many hard-mask disagreements feeding later stages: %s' % flips
    # ... and the end of the loop is checked UNCONDITIONALLY against the oracle chained on the CUDA path's own hard masks
    # (each oracle stage is fed the previous CUDA stage's outputs, so a flipped pixel upstream cannot cascade)
    obj_c, m_c = pf.to(dev), mask.to(dev)
    for s, h in enumerate(heads):
        ref = ko.kernel_update_head_forward(sds[s], cfg, x, obj_c.cpu().reshape(B, N, C, 1, 1), m_c.cpu())
        cls_c, m_c, obj_c = h(x.to(dev), obj_c, m_c)
        assert maxabs(obj_c, ref[2]) < TOL_F32 and maxabs(cls_c, ref[0]) < TOL_F32, 'cfg1 chained stage %d' % s
        assert_masks(m_c, ref[1], 'cfg1 chained stage %d' % s)
    assert torch.equal(m_l, m_c) and torch.equal(obj_l, obj_c)


def test_bench_operating_point_bf16_cfg1_batch64(dev):
    """The configuration bench.py quotes its headline on: bf16 storage, cfg1 (N=100, C=256, 200x88), F=2048, S=3, a batch of
    64 frames through the one-call loop (tcgen05 engines, chain kernel, 1-bit hard-mask hand-off between stages), also as
    the CUDA graph FramesInFlight replays.  Every stage is checked against the oracle on its actual inputs (kernel tensors
    1e-2, logits within one bf16 ulp, argmax identical up to bf16 near-ties -- counted and printed), and the one-call loop
    must reproduce the stage-wise module calls bit for bit."""
    import vknet
    B, N, C, H, W, S = 64, 100, 256, 200, 88, 3
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=50 + s)) for s in range(S)]
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(B, C, H, W, generator=g, device=dev)
    pfd = torch.randn(B, N, C, generator=g, device=dev)
    mb = pfd.bmm(x.view(B, C, -1)).view(B, N, H, W).bfloat16()
    xb = x.bfloat16()
    outs, ties = stagewise_vs_oracle_bf16(heads, sds, cfg, xb, pfd, mb, 'bench point (64 x cfg1)')
    loop = vknet.KernelIterLoop(heads)
    cls_l, m_l, obj_l = loop(xb, pfd, mb)
    assert torch.equal(m_l, outs[-1][1]) and torch.equal(obj_l.reshape(B, N, C), outs[-1][2].reshape(B, N, C)) \
        and torch.equal(cls_l, outs[-1][0]), 'one-call loop (bit-mask hand-off) != stage-wise modules'
    fif = vknet.FramesInFlight(heads, branches=1, batch=B).capture([(xb, pfd, mb)])
    cls_g, m_g, obj_g = fif.replay()[0]
    torch.cuda.synchronize()
    assert torch.equal(m_g, m_l) and torch.equal(obj_g.reshape(B, N, C), obj_l.reshape(B, N, C)) and torch.equal(cls_g, cls_l)
    print('bench operating point: %d bf16 near-tie pixels of %d resolved differently over %d stages' % (ties, S * B * H * W, S))


def test_graph_replay_matches_eager_and_is_deterministic(dev):
    import vknet
    cfg = ko.default_cfg(num_classes=19, in_channels=64, feedforward_channels=128)
    sds = [ko.random_state_dict(cfg, seed=s) for s in range(3)]
    heads = build_heads('KernelUpdateHead', cfg, sds, dev)
    x, pf, mask = (t.to(dev) for t in ko.dummy_inputs(2, 30, 64, 20, 28, seed=3))
    loop = vknet.KernelIterLoop(heads)
    a = [t.clone() for t in loop(x, pf, mask)]
    b = [t.clone() for t in loop(x, pf, mask)]
    assert all(torch.equal(u, v) for u, v in zip(a, b)), 'two eager runs differ bitwise'
    loop.capture(x, pf, mask)
    c = [t.clone() for t in loop.replay(x, pf, mask)]
    assert all(torch.equal(u, v) for u, v in zip(a, c)), 'graph replay differs from the eager run'
    x2, pf2, mask2 = (t.to(dev) for t in ko.dummy_inputs(2, 30, 64, 20, 28, seed=4))
    d = [t.clone() for t in loop.replay(x2, pf2, mask2)]
    e = loop(x2, pf2, mask2)
    assert all(torch.equal(u, v) for u, v in zip(d, e))


# ---- size-independent properties at full size ---------------------------------------------------------
def test_properties_full_size(dev):
    from vknet import ops
    B, N, C, H, W = 1, 100, 256, 200, 88
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048)
    sd = ko.random_state_dict(cfg, seed=2)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    x, pf, mask = (t.to(dev) for t in ko.dummy_inputs(B, N, C, H, W, seed=5))
    # pooling is additive over disjoint pixel sets: pool(mask) = pool(mask on left half) + pool(right half)
    neg = torch.full_like(mask, -1.0)
    left, right = neg.clone(), neg.clone()
    left[..., : W // 2] = mask[..., : W // 2]
    right[..., W // 2:] = mask[..., W // 2:]
    full = ops.mask_pool(h, x, mask)
    parts = ops.mask_pool(h, x, left) + ops.mask_pool(h, x, right)
    assert maxabs(full, parts) < 1e-4 * full.abs().max().item()
    # an all-negative mask pools to exactly zero; an all-positive mask row equals the feature sum
    assert ops.mask_pool(h, x, neg).abs().max().item() == 0.0
    # mask conv is linear in the kernels
    g = torch.Generator().manual_seed(0)
    k1 = torch.randn(B, N, C, generator=g).to(dev)
    k2 = torch.randn(B, N, C, generator=g).to(dev)
    m12 = ops.mask_gemm(h, x, k1 + k2)
    m1, m2 = ops.mask_gemm(h, x, k1), ops.mask_gemm(h, x, k2)
    assert maxabs(m12, m1 + m2) < 1e-4 * m12.abs().max().item()
    # permuting the kernels permutes the outputs (attention is permutation equivariant over N)
    perm = torch.randperm(N, generator=g).to(dev)
    cls, nm, obj = h(x, pf, mask)
    cls_p, nm_p, obj_p = h(x, pf[:, perm], mask[:, perm])
    assert maxabs(obj[:, perm], obj_p) < 1e-4 and maxabs(cls[:, perm], cls_p) < 1e-4
    assert maxabs(nm[:, perm], nm_p) < 1e-4 * nm.abs().max().item()


def test_error_behaviour(dev):
    import vknet
    cfg = ko.default_cfg(num_classes=3, in_channels=64, feedforward_channels=64)
    h = build_heads('KernelUpdateHead', cfg, [ko.random_state_dict(cfg)], dev)[0]
    x, pf, mask = (t.to(dev) for t in ko.dummy_inputs(1, 6, 64, 4, 4))
    with pytest.raises(vknet.VknError):
        h(x.cpu(), pf.cpu(), mask.cpu())                       # no CPU path
    with pytest.raises(vknet.VknError):
        h(x[:, :32], pf, mask)                                 # channel mismatch
    with pytest.raises(vknet.VknError):
        h(x.double(), pf, mask)                                # unsupported dtype
    bad = vknet.build_head(dict(type='KernelUpdateHead', **ko.default_cfg(in_channels=64, conv_kernel_size=3)))
    with pytest.raises(NotImplementedError):
        bad.to(dev)(x, torch.zeros(1, 6, 64, 3, 3, device=dev), mask)


# ---- tcgen05/TMA engine vs the CUDA-core engine and the oracle (bf16 storage) ---------------------------
@pytest.mark.parametrize('B,N,C,H,W', [(1, 100, 256, 200, 88), (2, 117, 256, 48, 156), (1, 166, 128, 16, 24),
                                       (3, 10, 64, 8, 8), (4, 100, 256, 96, 160),
                                       # cfg4 (VIP-Seg) shape with its real kernel count 100 + 66 stuff: two pooling M tiles, Npad 176
                                       (4, 166, 256, 120, 216)])
def test_tc_engine_matches_simt_engine_and_oracle(dev, B, N, C, H, W):
    from vknet import _lib, ops
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=21))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=12)
    x, mask = ko.round_bf16(x), ko.round_bf16(mask)
    xb, mb = x.to(dev).bfloat16(), mask.to(dev).bfloat16()
    out = {}
    for name, eng in (('simt', _lib.ENGINE_SIMT), ('tc', _lib.ENGINE_TC)):
        h.engine = eng
        g = torch.Generator().manual_seed(2)
        mk = torch.randn(B, N, C, generator=g)
        out[name] = (ops.mask_pool(h, xb, mb), ops.mask_gemm(h, xb, mk.to(dev)), h(xb, pf.to(dev), mb))
    xt, x_feat, _ = ko._pool(sd, cfg, x, pf, mask)
    # pooling: every product is exact (0/1 x bf16); only the fp32 summation order differs
    scale = x_feat.abs().max().item()
    assert maxabs(out['tc'][0], x_feat) < 2e-5 * scale, 'tc pooling vs oracle'
    assert maxabs(out['tc'][0], out['simt'][0]) < 2e-5 * scale, 'tc pooling vs simt'
    # mask conv: 3-plane bf16 split of the fp32 kernels keeps fp32-level accuracy -> after the bf16 output
    # rounding the two engines agree except for rare 1-ulp roundings
    a, b = out['tc'][1].float(), out['simt'][1].float()
    diff = (a - b).abs()
    assert diff.max().item() <= 2 ** -7 * b.abs().max().item(), 'tc mask conv differs from simt by more than 1 bf16 ulp'
    assert (diff > 0).float().mean().item() < 2e-3, 'tc mask conv: too many 1-ulp differences vs simt'
    want = ko.kernel_update_head_forward(sd, cfg, x, pf, mask)
    for name in ('simt', 'tc'):
        cls, nm, obj = out[name][2]
        assert maxabs(cls, want[0]) < TOL_BF16 and maxabs(obj, want[2]) < TOL_BF16, name
        assert_masks_bf16(nm, want[1], '%s engine stage' % name)


@pytest.mark.parametrize('wide', ['0', '1'])
@pytest.mark.parametrize('B,N,C,H,W', [(300, 20, 64, 8, 16), (160, 100, 128, 16, 16), (37, 100, 256, 24, 40)])
def test_persistent_mask_conv_balanced_ranges_cross_frames(dev, monkeypatch, B, N, C, H, W, wide):
    """The persistent mask conv cuts the launch's tiles into 148 equal ranges: with 1-2 tiles per frame a CTA's range
    spans several frames (plane + bias swap per segment).  Must equal the SIMT engine up to 1-ulp bf16 roundings."""
    from vknet import _lib, ops
    monkeypatch.setenv('VKN_MASK_WIDE', wide)        # '1': the pixels-as-N (256-pixel tile) form of the persistent kernel
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=64)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=4))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    g = torch.Generator().manual_seed(5)
    xb = torch.randn(B, C, H, W, generator=g).to(dev).bfloat16()
    mk = torch.randn(B, N, C, generator=g).to(dev)
    out = {}
    for name, eng in (('simt', _lib.ENGINE_SIMT), ('tc', _lib.ENGINE_TC)):
        h.engine = eng
        out[name] = ops.mask_gemm(h, xb, mk).float()
    diff = (out['tc'] - out['simt']).abs()
    assert diff.max().item() <= 2 ** -7 * out['simt'].abs().max().item()
    assert (diff > 0).float().mean().item() < 2e-3
    per_frame = diff.flatten(1).max(1).values           # a wrong plane / bias swap would wreck whole frames
    assert (per_frame <= 2 ** -7 * out['simt'].abs().max().item()).all()


def test_profile_hooks_and_launch_count(dev):
    from vknet import _lib
    cfg = ko.default_cfg(num_classes=3, in_channels=64, feedforward_channels=64)
    h = build_heads('KernelUpdateHead', cfg, [ko.random_state_dict(cfg)], dev)[0]
    x, pf, mask = (t.to(dev) for t in ko.dummy_inputs(1, 6, 64, 8, 8))
    n0 = _lib.launch_count()
    with _lib.profile() as p:
        h(x, pf, mask)
    assert _lib.launch_count() - n0 == len(p.records) >= 10
    assert all(name.startswith('vkn_') and ms >= 0 for name, ms in p.records)


def test_frames_in_flight_graph_matches_single_frame_loop(dev):
    """FramesInFlight (branches x batch in ONE graph, optional in-graph host I/O) == per-frame loop results."""
    import vknet
    cfg = ko.default_cfg(num_classes=19, in_channels=64, feedforward_channels=128)
    sds = [ko.random_state_dict(cfg, seed=s) for s in range(2)]
    heads = build_heads('KernelUpdateHead', cfg, sds, dev)
    loop = vknet.KernelIterLoop(heads)
    branches, batch = 3, 2
    ins = []
    for b in range(branches):
        parts = [ko.dummy_inputs(1, 20, 64, 16, 24, seed=10 * b + i) for i in range(batch)]
        ins.append(tuple(torch.cat(p) for p in zip(*parts)))
    want = [[t.clone() for t in loop(*(u.to(dev) for u in bt))] for bt in ins]
    fif = vknet.FramesInFlight(heads, branches=branches, batch=batch).capture([tuple(u.to(dev) for u in bt) for bt in ins])
    got = fif.replay()
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1]) and torch.equal(w[2].reshape(g[2].shape), g[2])
    pinned = [tuple(u.pin_memory() for u in bt) for bt in ins]
    hf = vknet.FramesInFlight(heads, branches=branches, batch=batch).capture(pinned, host_io=True)
    for p in pinned:                      # new host contents are picked up by the next replay
        p[0].mul_(1.0)
    hf.replay()
    torch.cuda.synchronize()
    for w, ho in zip(want, hf.host_out):
        assert torch.equal(w[0].cpu(), ho[0]) and torch.equal(w[1].cpu(), ho[1])


@pytest.mark.parametrize('N', [100, 117, 128, 166, 176])
def test_persistent_mask_conv_equals_tiled_mask_conv(dev, monkeypatch, N):
    """The persistent (resident-planes) tcgen05 mask conv issues the same MMAs per tile as the one-tile-per-CTA
    kernel, so their outputs must be bit-identical; both are exercised on a ragged last tile.  N > 112 (the real VPS kernel
    counts 117 / 166) runs as two kernel groups with their own resident planes."""
    from vknet import _lib, ops
    B, C, H, W = 3, 256, 40, 52                  # HW = 2080 = 16 tiles + a 32-pixel tail
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=4))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    h.engine = _lib.ENGINE_TC
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=3)
    xb = x.to(dev).bfloat16()
    mk = torch.randn(B, N, C, generator=torch.Generator().manual_seed(1)).to(dev)
    monkeypatch.setenv('VKN_MASK_PERSIST', '0')
    a = ops.mask_gemm(h, xb, mk).clone()
    monkeypatch.setenv('VKN_MASK_PERSIST', '1')
    b = ops.mask_gemm(h, xb, mk).clone()
    assert torch.equal(a, b)
    xt = ko._feat_transform(sd, ko.round_bf16(x))
    want = torch.einsum('bnc,bchw->bnhw', mk.cpu(), xt)
    assert (b.float().cpu() - want).abs().max().item() <= 2 ** -7 * want.abs().max().item()


# ---- next row: ConvKernelHead tail = the same two contractions with static kernels ------------
def test_init_proposals_golden_and_full_size(dev):
    import numpy as np
    from vknet import _lib, ops
    z = np.load(golden_files('init_')[0])
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    N, C = t['init_w'].shape[:2]
    conv = torch.nn.Conv2d(C, N, 1, bias=False)
    conv.weight.data.copy_(t['init_w'])
    prop, mask = ops.init_proposals(conv.to(dev), t['loc_feats'].to(dev), t['x_feats'].to(dev))
    assert maxabs(prop, t['proposal_feats']) < 1e-4 * t['proposal_feats'].abs().max().item()
    assert_masks(mask, t['mask_preds'], 'init masks', rel=1e-5)
    # BASELINE shapes, bf16 storage, both engines
    B, N, C, H, W = 2, 100, 256, 200, 88
    g = torch.Generator().manual_seed(0)
    w = ko.round_bf16(torch.randn(N, C, 1, 1, generator=g) * 0.2)
    loc, sem = ko.round_bf16(torch.randn(B, C, H, W, generator=g)), ko.round_bf16(torch.randn(B, C, H, W, generator=g))
    xf = ko.round_bf16(loc + sem)
    want_p, want_m = ko.init_proposals(w, None, loc, xf)
    conv = torch.nn.Conv2d(C, N, 1, bias=False)
    conv.weight.data.copy_(w)
    for eng in (_lib.ENGINE_SIMT, _lib.ENGINE_TC):
        p, m = ops.init_proposals(conv.to(dev), loc.to(dev).bfloat16(), xf.to(dev).bfloat16(), engine=eng)
        ref_m = ko.round_bf16(want_m)
        assert (m.float().cpu() - ref_m).abs().max().item() <= 2 ** -7 * ref_m.abs().max().item()
        flips = ((m.float().cpu() > 0) != (want_m > 0)).sum().item()
        assert flips <= 4, 'threshold disagreements vs the fp32 oracle: %d' % flips
        if flips == 0:
            assert maxabs(p, want_p) < 1e-4 * want_p.abs().max().item()


# ---- tcgen05 row engine (planes handed between row GEMMs) vs the warp-MMA chain and the oracle -------------------
@pytest.mark.parametrize('B,N,C,H,W,Fh,S', [(1, 20, 64, 16, 24, 512, 2), (2, 100, 256, 40, 24, 512, 2),
                                             (5, 100, 256, 24, 40, 2048, 1), (3, 117, 128, 16, 24, 192, 2),
                                             # P = 4000 rows: several tiles per persistent CTA (accumulator ring wraps)
                                             (40, 100, 256, 8, 16, 2048, 1)])
def test_row_engine_tc_vs_warp_mma_chain_and_oracle(dev, monkeypatch, B, N, C, H, W, Fh, S):
    import vknet
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=Fh)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=90 + s)) for s in range(S)]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=16)
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
    outs = {}
    for mode, rows_min in (('warp', '0'), ('tc', '1')):
        monkeypatch.setenv('VKN_ROWS_TC_MIN', rows_min)
        obj, m = pfd, mb
        res = []
        for h in heads:
            cls, m, obj = h(xb, obj, m)
            res.append((cls.clone(), m.clone(), obj.clone()))
        outs[mode] = res
        loop = vknet.KernelIterLoop(heads)          # one-call loop: obj planes ping-pong between stages
        cls_l, m_l, obj_l = loop(xb, pfd, mb)
        assert torch.equal(m_l, res[-1][1]) and torch.equal(obj_l, res[-1][2]) and torch.equal(cls_l, res[-1][0]), mode
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    tc_outs, _ = stagewise_vs_oracle_bf16(heads, sds, cfg, xb, pfd, mb, 'row engine')
    for s in range(S):
        for a, b in zip(tc_outs[s], outs['tc'][s]):
            assert torch.equal(a, b)
    # both engines multiply exact bf16 products with fp32 accumulation: first-stage kernels agree to fp32 round-off
    assert maxabs(outs['tc'][0][2], outs['warp'][0][2].cpu()) < 2e-4
    assert maxabs(outs['tc'][0][0], outs['warp'][0][0].cpu()) < 2e-4


@pytest.mark.parametrize('thr,wide', [(0.5, '0'), (0.7, '0'), (0.5, '1'), (0.7, '1')])
def test_loop_bitmask_handoff_equals_stagewise_modules(dev, monkeypatch, thr, wide):
    _loop_bitmask_case(dev, monkeypatch, thr, wide, 100)


@pytest.mark.parametrize('N', [117, 166])
def test_loop_bitmask_handoff_real_kernel_counts(dev, monkeypatch, N):
    """N = 117 (KITTI / Cityscapes-STEP: 100 things + 17 stuff) and 166 (VIP-Seg): the mask conv runs as two kernel groups,
    the pooling over two row tiles (N > 128); the loop still hands bit masks between its stages"""
    _loop_bitmask_case(dev, monkeypatch, 0.5, '0', N)


def _loop_bitmask_case(dev, monkeypatch, thr, wide, N):
    """Frame batches: the one-call loop hands the thresholded BIT per (kernel, pixel) from a stage's mask conv to the next
    stage's pooling instead of bf16 logits.  It thresholds the bf16-rounded logit, so the results must equal the
    stage-by-stage module calls (which store and re-read the logits) bit for bit -- also for a non-default threshold."""
    import vknet
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    monkeypatch.setenv('VKN_MASK_WIDE', wide)
    B, C, H, W, S = 6, 256, 96, 80, 3
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256, hard_mask_thr=thr)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=70 + s)) for s in range(S)]
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=17)
    xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
    obj, m = pfd, mb
    for h in heads:
        cls, m, obj = h(xb, obj, m)
    for bits in ('1', '0'):
        monkeypatch.setenv('VKN_LOOP_BITS', bits)
        cls_l, m_l, obj_l = vknet.KernelIterLoop(heads)(xb, pfd, mb)
        assert torch.equal(m_l, m) and torch.equal(obj_l, obj) and torch.equal(cls_l, cls), 'VKN_LOOP_BITS=%s' % bits
    # first stage against the oracle (later stages are covered by the threshold-aware full-size loop test)
    want = ko.kernel_update_head_forward(sds[0], cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
    cls0, _, obj0 = heads[0](xb, pfd, mb)
    assert maxabs(cls0, want[0]) < TOL_BF16 and maxabs(obj0, want[2]) < TOL_BF16


@pytest.mark.parametrize('bn', ['32', '64', '128', '256'])
def test_row_engine_tc_column_tiles(dev, monkeypatch, bn):
    """Every column-tile width of the persistent row GEMM (VKN_RG_BN) gives the same stage as the default choice."""
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    B, N, C, H, W = 12, 100, 256, 8, 16
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=1024)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=5))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=6)
    xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
    base = [t.clone() for t in h(xb, pfd, mb)]
    monkeypatch.setenv('VKN_RG_BN', bn)
    got = h(xb, pfd, mb)
    for a, b in zip(got, base):
        assert maxabs(a, b) <= 1e-5 * max(1.0, b.float().abs().max().item()), 'BN=%s' % bn
    want = ko.kernel_update_head_forward(sd, cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
    assert maxabs(got[0], want[0]) < TOL_BF16 and maxabs(got[2], want[2]) < TOL_BF16


def test_row_engine_tc_fused_layernorm_epilogue(dev, monkeypatch):
    """VKN_RG_FUSE_LN=1: fc_norm / attention_norm / the FC-head LayerNorms run in the row-GEMM epilogues (opt-in)."""
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    for (B, N, C, H, W, Fh) in ((3, 100, 256, 8, 16, 512), (2, 37, 128, 8, 16, 128)):
        cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=Fh)
        sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=21))
        h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
        x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=22)
        xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
        monkeypatch.setenv('VKN_RG_FUSE_LN', '0')
        base = [t.clone() for t in h(xb, pfd, mb)]
        monkeypatch.setenv('VKN_RG_FUSE_LN', '1')
        got = h(xb, pfd, mb)
        for i in (0, 2):         # cls_score, obj_feat (fp32): same arithmetic up to the LayerNorm summation order
            assert maxabs(got[i], base[i]) <= 5e-5 * max(1.0, base[i].float().abs().max().item())
        # mask logits are rounded to bf16 at the boundary: one-ulp flips at most
        assert maxabs(got[1], base[1]) <= 2 ** -7 * base[1].float().abs().max().item()
        want = ko.kernel_update_head_forward(sd, cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
        assert maxabs(got[0], want[0]) < TOL_BF16 and maxabs(got[2], want[2]) < TOL_BF16


@pytest.mark.parametrize('B,N,C,H,W,Fh', [(3, 100, 256, 8, 16, 512), (2, 37, 128, 8, 16, 128), (5, 117, 256, 8, 16, 2048),
                                         # 20 000 rows = 157 row tiles on 148 CTAs: some CTAs walk the program twice
                                         (200, 100, 64, 8, 8, 64)])
def test_row_engine_chain_kernel_equals_separate_launches(dev, monkeypatch, B, N, C, H, W, Fh):
    """Chain form (default): the row operators between pooling / attention / mask conv run as ONE kernel per segment, a CTA
    per 128-row tile walking the operator program (vkn_chain_tc_kernel).  Same arithmetic as one launch per operator
    (VKN_CHAIN=0) up to the summation order of the second FFN Linear (whole K in one pass instead of split-K)."""
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=Fh)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=41))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=42)
    xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
    monkeypatch.setenv('VKN_CHAIN', '0')
    monkeypatch.setenv('VKN_RG_FUSE_LN', '1')
    with _lib.profile() as prof0:
        base = [t.clone() for t in h(xb, pfd, mb)]
    monkeypatch.setenv('VKN_CHAIN', '1')
    with _lib.profile() as prof1:
        got = [t.clone() for t in h(xb, pfd, mb)]
    again = h(xb, pfd, mb)
    names0, names1 = [n for n, _ in prof0.records], [n for n, _ in prof1.records]
    assert names1.count('vkn_chain_tc_kernel') == 2 and 'vkn_rowgemm_tc_kernel' not in names1, names1
    assert 'vkn_chain_tc_kernel' not in names0 and len(names1) <= 7 < len(names0), (names0, names1)
    for a, b in zip(got, again):
        assert torch.equal(a, b), 'chain kernel must be deterministic'
    for i in (0, 2):
        assert maxabs(got[i], base[i]) <= 5e-5 * max(1.0, base[i].float().abs().max().item())
    assert maxabs(got[1], base[1]) <= 2 ** -7 * base[1].float().abs().max().item()
    if B <= 8:
        want = ko.kernel_update_head_forward(sd, cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
        assert maxabs(got[0], want[0]) < TOL_BF16 and maxabs(got[2], want[2]) < TOL_BF16


@pytest.mark.parametrize('ptype', ['ffn', 'update'])
def test_row_engine_tc_link_block(dev, monkeypatch, ptype):
    """VideoKernelUpdateHead link block (video/kernel_update_head.py:394-444) on the tcgen05 row engine == warp-MMA chain,
    both within tolerance of the oracle (bf16 storage)."""
    B, N, C, H, W = 6, 100, 256, 8, 16
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=512, previous='placeholder', previous_type=ptype)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=31))
    h = build_heads('VideoKernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=32)
    prev = torch.randn(B, N, C, 1, 1, generator=torch.Generator().manual_seed(33))
    x, mask = ko.round_bf16(x), ko.round_bf16(mask)
    want = ko.video_kernel_update_head_forward(sd, cfg, x, pf, mask, previous_obj_feats=prev)
    outs = {}
    for mode, rows_min in (('warp', '0'), ('tc', '1')):
        monkeypatch.setenv('VKN_ROWS_TC_MIN', rows_min)
        outs[mode] = h(x.to(dev).bfloat16(), pf.to(dev), mask.to(dev).bfloat16(), previous_obj_feats=prev.to(dev))
    for mode in outs:
        assert maxabs(outs[mode][4], want[4]) < TOL_BF16, mode
        assert maxabs(outs[mode][2], want[2]) < TOL_BF16, mode
    assert maxabs(outs['tc'][4], outs['warp'][4].cpu()) < 5e-4


def test_row_engine_tc_variants(dev, monkeypatch):
    """with_ffn=False, no feat_transform, deeper FC stacks, non-default threshold, video head (x_feat handed in)."""
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    C = 64
    cfg = ko.default_cfg(num_classes=5, in_channels=C, feedforward_channels=64, with_ffn=False,
                         feat_transform_cfg=None, num_mask_fcs=3, num_cls_fcs=2, hard_mask_thr=0.7)
    sd = ko.random_state_dict(cfg, seed=1)
    g = torch.Generator().manual_seed(8)
    for i in (1, 2):
        sd['mask_fcs.%d.weight' % (3 * i)] = ko._xavier(g, C, C)
        ko._ln_params(g, sd, 'mask_fcs.%d.' % (3 * i + 1), C)
    sd['cls_fcs.3.weight'] = ko._xavier(g, C, C)
    ko._ln_params(g, sd, 'cls_fcs.4.', C)
    for k in [k for k in sd if k.startswith('ffn')]:
        del sd[k]
    sd = ko.round_state_dict_bf16(sd)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(2, 14, C, 8, 16, seed=2)
    x, mask = ko.round_bf16(x), ko.round_bf16(mask)
    want = ko.kernel_update_head_forward(sd, cfg, x, pf, mask)
    cls, nm, obj = h(x.to(dev).bfloat16(), pf.to(dev), mask.to(dev).bfloat16())
    assert maxabs(cls, want[0]) < TOL_BF16 and maxabs(obj, want[2]) < TOL_BF16
    ref = ko.round_bf16(want[1])
    assert (nm.float().cpu() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize('B,N,C', [(1, 10, 64), (2, 37, 128), (3, 117, 64), (16, 100, 256), (2, 166, 128)])
def test_attention_four_queries_per_warp_kernel(dev, monkeypatch, B, N, C):
    monkeypatch.setenv('VKN_ATT_TC', '0')            # the SIMT kernels are the subject here
    """The batched attention kernel (4 queries per warp, picked when many (head, frame) pairs are in flight) against
    the oracle and against the one-query-per-warp kernel, on ragged N (partial last query group / key sweep)."""
    from vknet import ops
    cfg = ko.default_cfg(num_classes=11, in_channels=C, feedforward_channels=64)
    sd = ko.random_state_dict(cfg, seed=5)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    rows = torch.randn(B, N, C, generator=torch.Generator().manual_seed(4))
    seq = rows.permute(1, 0, 2)
    want = ko.layer_norm(ko.mha_block(seq, seq, seq, seq, sd, 'attention.', 8), sd['attention_norm.weight'],
                         sd['attention_norm.bias']).permute(1, 0, 2)
    monkeypatch.setenv('VKN_ATT4', '1')
    got4 = ops.mhsa_ln(h, rows.to(dev)).clone()
    monkeypatch.setenv('VKN_ATT4', '0')
    got1 = ops.mhsa_ln(h, rows.to(dev)).clone()
    assert maxabs(got4, want) < 1e-4 and maxabs(got1, want) < 1e-4
    assert maxabs(got4, got1.cpu()) < 2e-5


# ---- next row 8f-2: post-loop mask path (upsample + rescale_masks + threshold in one launch) -----------------------
@pytest.mark.parametrize('path', golden_files('rescale_'), ids=lambda p: p.split('/')[-1][:-4])
def test_rescale_masks_golden(dev, path):
    import numpy as np
    from vknet import ops
    z = np.load(path)
    K, H, W, up, Hb, Wb, h, w, Ho, Wo = (int(v) for v in z['meta'])
    meta = dict(img_shape=(h, w, 3), batch_input_shape=(Hb, Wb), ori_shape=(Ho, Wo, 3))
    want = torch.from_numpy(z['seg'])
    probs, bits = ops.rescale_masks(torch.from_numpy(z['masks']).to(dev), meta, up, 0.5)
    assert probs.shape == (K, Ho, Wo) and bits.dtype == torch.bool
    assert maxabs(probs, want) < 2e-6
    near = (want - 0.5).abs() < 2e-6
    assert torch.equal(bits.cpu() | near, (want > 0.5) | near)


@pytest.mark.parametrize('K,H,W,up,batch,img,ori,dt', [
    (100, 48, 156, 2, (384, 1248), (375, 1242), (375, 1242), 'bf16'),     # KITTI-STEP: stride-8 logits -> full resolution
    (10, 96, 160, 2, (768, 1280), (720, 1280), (360, 640), 'f32'),        # down-scaling to the original size
    (7, 25, 44, 1, (200, 352), (200, 350), (480, 840), 'f32'),            # no loop upsample, up-scaling, ragged tiles
    (3, 50, 88, 4, (200, 352), (190, 333), (97, 171), 'bf16')])
def test_rescale_masks_vs_oracle(dev, K, H, W, up, batch, img, ori, dt):
    from vknet import ops
    g = torch.Generator().manual_seed(K)
    masks = torch.randn(K, H, W, generator=g) * 4.0
    if dt == 'bf16':
        masks = ko.round_bf16(masks)
    meta = dict(img_shape=img + (3,), batch_input_shape=batch, ori_shape=ori + (3,))
    want = ko.rescale_masks(masks, meta, up)
    md = masks.to(dev).bfloat16() if dt == 'bf16' else masks.to(dev)
    probs, bits = ops.rescale_masks(md, meta, up, 0.5)
    assert maxabs(probs, want) < 5e-6
    near = (want - 0.5).abs() < 5e-6
    assert torch.equal(bits.cpu() | near, (want > 0.5) | near)
    assert near.float().mean().item() < 1e-3
    only_bits = ops.rescale_masks(md, meta, up, 0.5, probs=False)
    assert only_bits[0] is None and torch.equal(only_bits[1], bits)
    # the module method keeps the reference signature (scaled masks in, probabilities out)
    import vknet
    cfg = ko.default_cfg(num_classes=19, in_channels=64, feedforward_channels=64)
    head = vknet.build_head(dict(type='KernelUpdateHead', **cfg)).to(dev)
    if up == 1:
        assert maxabs(head.rescale_masks(md, meta), want) < 5e-6


def test_rescale_masks_errors(dev):
    from vknet import _lib, ops
    meta = dict(img_shape=(8, 8, 3), batch_input_shape=(8, 8), ori_shape=(8, 8, 3))
    with pytest.raises(_lib.VknError):
        ops.rescale_masks(torch.zeros(2, 4, 4), meta)                      # CPU tensor: no fallback
    with pytest.raises(_lib.VknError):                                     # extreme down-scaling: dependency cone too large
        ops.rescale_masks(torch.zeros(1, 2000, 2000, device=dev), dict(img_shape=(2000, 2000, 3), batch_input_shape=(2000, 2000),
                                                                        ori_shape=(20, 20, 3)))


# ---- post-loop result assembly on the device -------------------------------------------------------------------------
@pytest.mark.parametrize('path', golden_files('panoptic_'), ids=lambda p: p.split('/')[-1][:-4])
def test_panoptic_merge_golden(dev, path):
    """vkn_panoptic_merge == the reference's merge_stuff_thing_stuff_joint on the fixtures it wrote
    (knet/video/kernel_iter_head.py:832-895): id map, segment table and kept thing indices, bit for bit."""
    import numpy as np
    from vknet import ops
    z = np.load(path)
    K, M, H, W, nthing = (int(v) for v in z['meta'])
    t = {k: torch.from_numpy(z[k]).to(dev) for k in ('thing_masks', 'stuff_masks', 'thing_scores', 'stuff_scores', 'thing_labels',
                                                     'stuff_labels')}
    seg, info, kept = ops.panoptic_merge(t['thing_masks'], t['thing_labels'], t['thing_scores'], t['stuff_masks'], t['stuff_labels'],
                                         t['stuff_scores'], nthing, float(z['thr'][0]), float(z['thr'][1]))
    assert np.array_equal(seg.cpu().numpy(), z['seg'])
    rows = np.array([[d['id'], int(d['isthing']), d['category_id'], d.get('instance_id', -1), d.get('area', -1)] for d in info],
                    dtype=np.int64).reshape(-1, 5)
    assert np.array_equal(rows, z['info'])
    assert kept == z['kept'].tolist()


def test_panoptic_merge_full_size_vs_oracle(dev):
    """KITTI-STEP size: 100 things + 17 stuff kernels at 375 x 1242, probabilities with large overlaps and exact ties"""
    from vknet import ops
    g = torch.Generator().manual_seed(4)
    K, M, H, W, nthing = 100, 17, 375, 1242, 2
    base = torch.rand(K + M, H // 15 + 1, W // 18 + 1, generator=g) * 0.3
    masks = torch.nn.functional.interpolate(base[None], size=(H, W), mode='bilinear', align_corners=False)[0]
    for k in range(K + M):                                           # every kernel: a confident box, boxes overlap their neighbours
        y0, x0 = int(torch.randint(0, H - 60, (1,), generator=g)), int(torch.randint(0, W - 200, (1,), generator=g))
        hh, ww = int(torch.randint(20, 60, (1,), generator=g)), int(torch.randint(40, 200, (1,), generator=g))
        masks[k, y0:y0 + hh, x0:x0 + ww] = 0.55 + 0.45 * torch.rand(hh, ww, generator=g)
    masks[5] = masks[4]                                              # identical maps: the first index must win the ties
    scores = 0.2 + 0.8 * torch.rand(K + M, generator=g)
    scores[5] = scores[4]
    labels = torch.cat([torch.randint(0, nthing, (K,), generator=g), torch.arange(M) + nthing])
    want_seg, want_info, want_kept = ko.panoptic_merge_joint(masks[:K], labels[:K], scores[:K], masks[K:], labels[K:], scores[K:],
                                                             nthing, 0.3, 0.5)
    seg, info, kept = ops.panoptic_merge(masks[:K].to(dev), labels[:K].to(dev), scores[:K].to(dev), masks[K:].to(dev),
                                         labels[K:].to(dev), scores[K:].to(dev), nthing, 0.3, 0.5)
    assert torch.equal(seg.cpu(), want_seg)
    assert kept == want_kept and len(info) == len(want_info) >= 5 and len(info) < K + M
    for a, b in zip(info, want_info):
        assert {k: v for k, v in a.items() if k != 'score'} == {k: v for k, v in b.items() if k != 'score'}
        if 'score' in a:
            assert abs(a['score'] - b['score']) < 1e-7


def test_mask_boxes_kernel(dev):
    """mask -> box reduction of VideoKernelUpdateHead.segm2result (knet/video/kernel_update_head.py:734-744) == its torch
    formulation (which the CPU boundary tests pin to the reference's tensor_mask2box), bool and float masks, empty masks"""
    import vknet
    from vknet import ops
    g = torch.Generator().manual_seed(2)
    m = torch.rand(37, 90, 160, generator=g) > 0.97
    m[3] = False
    m[11, :, :] = False
    m[11, 89, 159] = True                                            # a single pixel in the far corner
    want = vknet.VideoKernelUpdateHead.mask_boxes_torch(m)
    assert torch.equal(ops.mask_boxes(m.to(dev)).cpu(), want)
    assert torch.equal(ops.mask_boxes(m.float().to(dev) * 0.25).cpu(), want)
    assert torch.equal(ops.mask_boxes(m.to(torch.uint8).to(dev)).cpu(), want)
    assert want[3].tolist() == [-1.0, -1.0, 10.0, 10.0] and want[11].tolist() == [159.0, 89.0, 159.0, 89.0]
    # odd sizes: every mask starts at a different alignment (H*W = 7 * 13), head / tail elements, groups spanning rows
    m2 = torch.rand(23, 7, 13, generator=g) > 0.8
    m2[5] = False
    m2[5, 0, 0] = True
    m2[6] = False
    m2[6, 6, 12] = True
    want2 = vknet.VideoKernelUpdateHead.mask_boxes_torch(m2)
    assert torch.equal(ops.mask_boxes(m2.to(dev)).cpu(), want2) and torch.equal(ops.mask_boxes(m2.float().to(dev)).cpu(), want2)
    m3 = torch.rand(5, 375, 1242, generator=g) > 0.999                 # KITTI size, row length not a multiple of 16
    assert torch.equal(ops.mask_boxes(m3.to(dev)).cpu(), vknet.VideoKernelUpdateHead.mask_boxes_torch(m3))
    head = vknet.build_head(dict(type='VideoKernelUpdateHead', **ko.default_cfg(num_classes=4, in_channels=64, feedforward_channels=64,
                                                                                 previous='p', previous_type='ffn')))
    labels, scores = torch.randint(0, 4, (37,), generator=g), torch.rand(37, generator=g)
    b_gpu = head.segm2result(m.to(dev), labels.to(dev), scores.to(dev))[0]
    b_cpu = head.segm2result(m, labels, scores)[0]
    assert (b_gpu == b_cpu).all() and b_gpu[3, :4].tolist() == [0, 0, 10, 10]


@pytest.mark.parametrize('B,N,C', [(8, 100, 256), (9, 117, 256), (16, 16, 64), (12, 37, 128), (150, 100, 256), (8, 128, 256)])
def test_attention_tcgen05_kernel(dev, monkeypatch, B, N, C):
    """tcgen05 attention (QK^T and PV on tensor cores, two fp16 planes per operand, softmax from TMEM) against the fp32 SIMT
    kernel and the oracle's MultiheadAttention + LayerNorm block (knet/det/kernel_update_head.py:204-208)."""
    import vknet
    from vknet import _lib, ops
    cfg = ko.default_cfg(num_classes=5, in_channels=C, feedforward_channels=64, num_heads=C // 32)
    sd = ko.random_state_dict(cfg, seed=12)
    h = build_heads('KernelUpdateHead', cfg, [sd], dev)[0]
    g = torch.Generator().manual_seed(13)
    qin = torch.randn(B, N, C, generator=g) * 1.5
    want = ko.layer_norm(ko.mha_block(qin.permute(1, 0, 2), qin.permute(1, 0, 2), qin.permute(1, 0, 2), qin.permute(1, 0, 2), sd,
                                      'attention.', C // 32), sd['attention_norm.weight'], sd['attention_norm.bias']).permute(1, 0, 2)
    monkeypatch.setenv('VKN_ATT_TC', '0')
    with _lib.profile() as p0:
        simt = ops.mhsa_ln(h, qin.to(dev)).clone()
    monkeypatch.setenv('VKN_ATT_TC', '1')
    monkeypatch.setenv('VKN_ATT_TC_MIN', '1')
    with _lib.profile() as p1:
        tc = ops.mhsa_ln(h, qin.to(dev)).clone()
    assert any('attention_tc' in n for n, _ in p1.records) and not any('attention_tc' in n for n, _ in p0.records)
    assert torch.equal(tc, ops.mhsa_ln(h, qin.to(dev))), 'deterministic'
    assert maxabs(simt, want) < 2e-5 and maxabs(tc, want) < 2e-5, (maxabs(simt, want), maxabs(tc, want))
    assert maxabs(tc, simt.cpu()) < 1e-5


@pytest.mark.parametrize('path', golden_files('track_match_'), ids=lambda p: p.split('/')[-1][:-4])
def test_track_match_golden(dev, path):
    """vkn_track_match == the reference's QuasiDenseEmbedTracker.match (fixtures from the unmodified class): kept detections
    in score order, track ids (matches, suppressed duplicates, births), frame by frame against the recorded memory."""
    import numpy as np
    from vknet import ops
    z = np.load(path)
    frames = int(z['meta'][0])
    thr = [float(v) for v in z['cfg']]
    for f in range(frames):
        t = {k: torch.from_numpy(z['f%d.%s' % (f, k)]) for k in ('bboxes', 'labels', 'feats', 'memo_labels', 'memo_embeds', 'memo_ids',
                                                               'out_bboxes', 'out_labels', 'out_ids')}
        n0, n1 = (int(v) for v in z['f%d.num_tracklets' % f])
        memo = (None, None, None) if t['memo_ids'].numel() == 0 else (t['memo_labels'].to(dev), t['memo_embeds'].to(dev), t['memo_ids'].to(dev))
        sel, ids, nnew = ops.track_match(t['bboxes'].to(dev), t['labels'].to(dev), t['feats'].to(dev), memo[0], memo[1], memo[2], n0, *thr)
        assert torch.equal(t['bboxes'][sel.cpu()], t['out_bboxes']) and torch.equal(t['labels'][sel.cpu()], t['out_labels']), f
        assert torch.equal(ids.cpu(), t['out_ids']) and n0 + nnew == n1, f
    sel, ids, nnew = ops.track_match(torch.zeros(0, 5, device=dev), torch.zeros(0, dtype=torch.long, device=dev),
                                     torch.zeros(0, 8, device=dev), None, None, None, 3, *thr)
    assert sel.numel() == 0 and ids.numel() == 0 and nnew == 0


@pytest.mark.parametrize('dt', ['f32', 'bf16'])
def test_tracking_embedding_mlp(dev, dt):
    """the tracking-embedding stack on the last-stage kernels: embed_fcs (Linear(no bias) -> LN -> ReLU) + fc_embed
    (knet/video/knet_quansi_dense_embed_fc_joint_train.py:113-126, 572-580) followed by the track head's fcs
    (Linear -> ReLU) x 2 + fc_embed (knet/video/track_heads.py:632-642), vs the same modules in torch fp32."""
    from vknet import ops
    torch.manual_seed(3)
    C = 256
    mods = [(torch.nn.Linear(C, C, bias=False), torch.nn.LayerNorm(C), True), (torch.nn.Linear(C, C), None, False),
            (torch.nn.Linear(C, C), None, True), (torch.nn.Linear(C, C), None, True), (torch.nn.Linear(C, C), None, False)]
    for lin, norm, _ in mods:
        torch.nn.init.xavier_uniform_(lin.weight)
        if norm is not None:
            torch.nn.init.normal_(norm.weight, 1.0, 0.2)
            torch.nn.init.normal_(norm.bias, 0.0, 0.2)
    if dt == 'bf16':
        for lin, _, _ in mods:
            lin.weight.data = lin.weight.data.bfloat16().float()
    x = torch.randn(100, C)
    want = x
    with torch.no_grad():
        for lin, norm, relu in mods:
            want = lin(want)
            want = norm(want) if norm is not None else want
            want = torch.relu(want) if relu else want
    if dt == 'bf16':
        for lin, _, _ in mods:
            lin.weight.data = lin.weight.data.bfloat16()
    got = ops.mlp([(lin.to(dev), None if norm is None else norm.to(dev), relu) for lin, norm, relu in mods], x.to(dev))
    assert got.shape == want.shape and maxabs(got, want) < 1e-4 * max(1.0, want.abs().max().item())
    want_k = ko.mlp([(lin.weight.float().cpu(), None if lin.bias is None else lin.bias.float().cpu(),
                      None if norm is None else norm.weight.cpu(), None if norm is None else norm.bias.cpu(), relu)
                     for lin, norm, relu in mods], x)
    assert maxabs(got, want_k) < 1e-4 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize('N', [100, 166])
def test_mask_conv_fp16_mode_vs_bf16_planes(dev, monkeypatch, N):
    """Persistent mask conv, fp16 mode (default for frame batches): the folded kernels as TWO fp16 planes (22 bits) and x
    converted bf16 -> fp16 in the ring (exact) -- a third fewer MMAs than three bf16 planes.  Same logits up to rare one-ulp
    roundings of the bf16 store, and both within one bf16 ulp of the oracle."""
    monkeypatch.setenv('VKN_ROWS_TC_MIN', '1')
    B, C, H, W = 6, 256, 96, 80
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=256)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=33))
    h = build_heads('KernelUpdateHead', cfg, [sd], dev, dtype=torch.bfloat16)[0]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=34)
    xb, mb, pfd = x.to(dev).bfloat16(), mask.to(dev).bfloat16(), pf.to(dev)
    outs = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('VKN_MASK_F16', mode)
        with _lib.profile() as p:
            outs[mode] = [t.clone() for t in h(xb, pfd, mb)]
        assert any('maskgemm_tc_persist' in n for n, _ in p.records)
    a, b = outs['0'][1].float(), outs['1'][1].float()
    assert torch.equal(outs['0'][2], outs['1'][2]) and torch.equal(outs['0'][0], outs['1'][0])     # kernels do not depend on the mode
    diff = (a - b).abs()
    assert diff.max().item() <= 2 ** -7 * a.abs().max().item()
    assert (diff > 0).float().mean().item() < 1e-3, 'fp16 vs bf16-plane mask conv: too many one-ulp differences'
    want = ko.kernel_update_head_forward(sd, cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
    for mode in ('0', '1'):
        assert_masks_bf16(outs[mode][1], want[1], 'mask conv VKN_MASK_F16=%s' % mode)


# ---- single-frame row engine: the cluster chain kernels (framechain.cu) -------------------------------------------------
@pytest.mark.parametrize('B,N,H,W,S,ncls', [(1, 100, 200, 88, 3, 19), (2, 117, 48, 156, 2, 19), (1, 166, 24, 40, 2, 124),
                                           (1, 10, 16, 24, 2, 133), (3, 100, 16, 24, 1, 40), (1, 176, 8, 16, 1, 1)])
def test_frame_chain_cluster_kernels(dev, monkeypatch, B, N, H, W, S, ncls):
    """One frame (or a few) per call -- the online VPS operating point: every row operator of a stage runs in the two cluster
    kernels of framechain.cu (5 launches per stage).  Each stage against the oracle on its actual inputs (bf16 rules), against
    the warp-MMA chain it replaces, and the one-call loop == the stage-wise modules bit for bit."""
    import vknet
    C = 256
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=2048)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=90 + s)) for s in range(S)]
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=8)
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    xb, pfd, mb = x.to(dev).bfloat16(), pf.to(dev).reshape(B, N, C), mask.to(dev).bfloat16()
    L = vknet._lib.lib()
    for h in heads:                       # first use packs the weights (one more launch per head)
        h.packed_weights(dev)
    n0 = L.vkn_launch_count()
    outs, ties = stagewise_vs_oracle_bf16(heads, sds, cfg, xb, pfd, mb, 'frame chain %dx%d' % (B, N))
    assert L.vkn_launch_count() - n0 == 5 * S, 'expected pool, reduce, chain A, chain B, mask conv per stage'
    # the engine it replaces: same math on the warp-MMA chain (different summation order -> tolerance, not bits)
    monkeypatch.setenv('VKN_FRAME_CHAIN', '0')
    n0 = L.vkn_launch_count()
    obj, m = pfd, mb
    for s, h in enumerate(heads):
        cls, m_new, obj_new = h(xb, obj, m)
        assert maxabs(cls, outs[s][0]) < 2e-3 and maxabs(obj_new, outs[s][2]) < 2e-3, 'stage %d: chain vs warp-MMA engine' % s
        # the two engines' logits agree to a bf16 ulp
        d = (m_new.float() - outs[s][1].float()).abs()
        assert bool((d <= bf16_ulp(m_new.float()) * 1.001 + 2.0 ** -16 * m_new.float().abs().max()).all())
        obj, m = outs[s][2].reshape(B, N, C), outs[s][1]
    assert L.vkn_launch_count() - n0 > 5 * S
    monkeypatch.delenv('VKN_FRAME_CHAIN')
    cls_l, m_l, obj_l = vknet.KernelIterLoop(heads)(xb, pfd, mb)
    assert torch.equal(m_l, outs[-1][1]) and torch.equal(obj_l.reshape(B, N, C), outs[-1][2].reshape(B, N, C)) \
        and torch.equal(cls_l, outs[-1][0]), 'one-call loop != stage-wise modules'
    # determinism (fixed-order reductions through distributed shared memory)
    cls_2, m_2, obj_2 = vknet.KernelIterLoop(heads)(xb, pfd, mb)
    assert torch.equal(m_l, m_2) and torch.equal(obj_l, obj_2) and torch.equal(cls_l, cls_2)
    # a C-ABI caller that does not provide the pre-laid weight image (VknHeadW.fc_pack = NULL): the ring is fed row by row,
    # same bits
    for h in heads:
        h.packed_weights(dev)[0].fc_pack = None
    cls_3, m_3, obj_3 = vknet.KernelIterLoop(heads)(xb, pfd, mb)
    assert torch.equal(m_l, m_3) and torch.equal(obj_l, obj_3) and torch.equal(cls_l, cls_3)
    for h in heads:
        h.invalidate_weight_cache()


def test_frame_chain_handoff_transports_agree(dev, monkeypatch):
    """The plane hand-off of the cluster chain has two transports into the same K-blocked buffers: one bulk copy shared ->
    distributed shared memory per destination (default) and 16-byte st.async stores (VKN_FC_BULK=0, the form compute-sanitizer can
    follow).  Same bits."""
    import vknet
    B, N, C, H, W, S = 2, 100, 256, 16, 24, 2
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048)
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=60 + s)) for s in range(S)]
    heads = build_heads('KernelUpdateHead', cfg, sds, dev, dtype=torch.bfloat16)
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=4)
    xb, pfd, mb = x.to(dev).bfloat16(), pf.to(dev).reshape(B, N, C), mask.to(dev).bfloat16()
    a = vknet.KernelIterLoop(heads)(xb, pfd, mb)
    monkeypatch.setenv('VKN_FC_BULK', '0')
    b = vknet.KernelIterLoop(heads)(xb, pfd, mb)
    assert all(torch.equal(u, v) for u, v in zip(a, b))


def test_frame_chain_video_head_and_clip_head_without_cls(dev):
    """The two other entries into the cluster chain: VideoKernelUpdateHead (pooled feature computed first, handed in as
    x_feat_in; returns x_feat) and KernelUpdateHeadVideo in per-frame mode (with_cls=False: no cls branch)."""
    import vknet
    B, N, C, H, W = 1, 100, 256, 24, 40
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=97))
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=9)
    xr, mr = ko.round_bf16(x), ko.round_bf16(mask)
    want = ko.kernel_update_head_forward(sd, cfg, xr, pf, mr)
    h = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
    h.load_state_dict(sd, strict=True)
    h = h.to(device=dev, dtype=torch.bfloat16).eval()
    L = vknet._lib.lib()
    h.packed_weights(dev)
    n0 = L.vkn_launch_count()
    out = h(xr.to(dev).bfloat16(), pf.to(dev), mr.to(dev).bfloat16())
    assert L.vkn_launch_count() - n0 <= 6, 'video head: pool, reduce, x_feat Linear, chain A, chain B, mask conv'
    cls, nm, obj, x_feat = out[0], out[1], out[2], out[3]
    assert maxabs(cls, want[0]) < TOL_BF16 and maxabs(obj, want[2]) < TOL_BF16
    assert_masks_bf16(nm, want[1], 'video head via frame chain')
    want_xf = ko._pool(sd, cfg, xr, pf, mr)[1]
    assert maxabs(x_feat, want_xf) < 1e-3 * max(1.0, want_xf.abs().max().item())
    # clip head, per-frame mode
    Fr = 2
    sd2 = {k: v for k, v in sd.items() if not (k.startswith('cls_fcs') or k.startswith('fc_cls'))}
    gen = torch.Generator().manual_seed(5)
    x2 = ko.round_bf16(torch.randn(B, Fr, C, H, W, generator=gen))
    pf2 = torch.randn(B, Fr, N, C, 1, 1, generator=gen)
    mask2 = ko.round_bf16(torch.einsum('bfnc,bfchw->bfnhw', pf2.view(B, Fr, N, C), x2))
    want2 = ko.kernel_update_head_video_forward(sd2, cfg, x2, pf2, mask2)
    h2 = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=False, num_proposals=N, **cfg))
    h2.load_state_dict(sd2, strict=True)
    h2 = h2.to(device=dev, dtype=torch.bfloat16).eval()
    h2.packed_weights(dev)
    n0 = L.vkn_launch_count()
    cls2, nm2, obj2 = h2(x2.to(dev).bfloat16(), pf2.to(dev), mask2.to(dev).bfloat16())
    assert L.vkn_launch_count() - n0 == 5
    assert maxabs(obj2, want2[2]) < TOL_BF16
    assert_masks_bf16(nm2.reshape(B * Fr, N, H, W), want2[1].reshape(B * Fr, N, H, W), 'clip head (per-frame) via frame chain')
    # clip head, gathered mode (BASELINE cfg2: one kernel set pools the mean over its 4 frames and convolves all of them)
    Fr = 4
    x3 = ko.round_bf16(torch.randn(B, Fr, C, H, W, generator=gen))
    pf3 = torch.randn(B, N, C, 1, 1, generator=gen)
    mask3 = ko.round_bf16(torch.einsum('bnc,bfchw->bfnhw', pf3.view(B, N, C), x3))
    want3 = ko.kernel_update_head_video_forward(sd, cfg, x3, pf3, mask3)
    h3 = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=True, num_proposals=N, **cfg))
    h3.load_state_dict(sd, strict=True)
    h3 = h3.to(device=dev, dtype=torch.bfloat16).eval()
    h3.packed_weights(dev)
    n0 = L.vkn_launch_count()
    cls3, nm3, obj3 = h3(x3.to(dev).bfloat16(), pf3.to(dev), mask3.to(dev).bfloat16())
    assert L.vkn_launch_count() - n0 == 5
    assert maxabs(obj3, want3[2]) < TOL_BF16 and maxabs(cls3, want3[0]) < TOL_BF16
    assert_masks_bf16(nm3.reshape(B * Fr, N, H, W), want3[1].reshape(B * Fr, N, H, W), 'clip head (gathered) via frame chain')


# ---- row f4: MaskHungarianAssigner cost matrix (vkn_match_cost) -----------------------------------------------------------
@pytest.mark.parametrize('path', golden_files('assign_'), ids=lambda p: p.split('/')[-1][:-4])
def test_match_cost_golden(dev, path):
    """The fused cost kernel against the reference's own cost objects (fixtures), each term and the sum, and the drop-in
    MaskHungarianAssigner.assign against the reference's assignment."""
    from vknet import assigner, ops
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    pred, cls, gt, lab = (t[k].to(dev) for k in ('mask_logits', 'cls_logits', 'gt_masks', 'gt_labels'))
    tol = 2e-5
    assert maxabs(ops.match_cost(pred, cls, gt, lab), t['cost']) < tol
    assert maxabs(ops.match_cost(pred, cls, gt, lab, w_mask=0.0, w_dice=0.0), t['cost_cls']) < tol
    assert maxabs(assigner.MaskCost(weight=1.0, pred_act=True)(pred, gt), t['cost_mask']) < tol
    assert maxabs(assigner.DiceCost(weight=4.0, pred_act=True)(pred, gt), t['cost_dice']) < tol
    asg = assigner.MaskHungarianAssigner(cls_cost=dict(type='FocalLossCost', weight=2.0),
                                         dice_cost=dict(type='DiceCost', weight=4.0, pred_act=True),
                                         mask_cost=dict(type='MaskCost', weight=1.0, pred_act=True))
    res = asg.assign(pred, cls, gt, lab)
    assert res.num_gts == gt.shape[0] and torch.equal(res.gt_inds.cpu(), t['gt_inds']) and torch.equal(res.labels.cpu(), t['labels'])
    # twice the same bits (fixed-order reduction of the pixel chunks)
    assert torch.equal(ops.match_cost(pred, cls, gt, lab), ops.match_cost(pred, cls, gt, lab))


@pytest.mark.parametrize('N,M,H,W,ncls', [(100, 30, 200, 304, 19), (166, 70, 97, 131, 124), (1, 1, 5, 3, 1), (130, 33, 64, 64, 8)])
def test_match_cost_vs_oracle(dev, N, M, H, W, ncls):
    """Training-size masks (1/4-resolution KITTI crop), ragged sizes (HW not a multiple of the 64-pixel block), more than one
    target / prediction block, single elements; plus the empty-set behaviour of assign (reference :218-224)."""
    from vknet import assigner, ops
    g = torch.Generator().manual_seed(N + M)
    pred, gt = 3 * torch.randn(N, H, W, generator=g), torch.rand(M, H, W, generator=g).round()
    cls, lab = torch.randn(N, ncls, generator=g), torch.randint(0, ncls, (M,), generator=g)
    want = ko.match_cost(pred, cls, gt, lab)
    got = ops.match_cost(pred.to(dev), cls.to(dev), gt.to(dev), lab.to(dev))
    assert maxabs(got, want) < 2e-5 * max(1.0, want.abs().max().item())
    asg = assigner.MaskHungarianAssigner(cls_cost=dict(type='FocalLossCost', weight=2.0),
                                         dice_cost=dict(type='DiceCost', weight=4.0, pred_act=True),
                                         mask_cost=dict(type='MaskCost', weight=1.0, pred_act=True))
    res = asg.assign(pred.to(dev), cls.to(dev), gt[:0].to(dev), lab[:0].to(dev))
    assert res.num_gts == 0 and bool((res.gt_inds == 0).all()) and bool((res.labels == -1).all())
    with pytest.raises(NotImplementedError):
        assigner.DiceCost(weight=1.0, pred_act=False)(pred.to(dev), gt.to(dev))
