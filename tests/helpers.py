"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import glob
import os

import numpy as np
import torch

import knet_oracle as ko

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load_golden(path):
    z = np.load(path)
    B, N, C, H, W, S, Fh, ncls = (int(v) for v in z['meta'])
    t = {k: torch.from_numpy(z[k]) for k in z.files if k not in ('meta', 'frames')}
    sds = []
    for s in range(S):
        pre = 's%d.w.' % s
        sds.append({k[len(pre):]: v for k, v in t.items() if k.startswith(pre)})
    video = 'previous_obj_feats' in t
    over = {}
    if video:
        sd = sds[0]
        over['previous'] = 'placeholder'
        over['previous_type'] = 'ffn' if 'link_ffn.layers.1.weight' in sd else (
            'update' if 'link_ffn_track.layers.1.weight' in sd else None)
        over['previous_link'] = 'update_dynamic_cov' if 'link_ffn_link.layers.1.weight' in sd else None
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=Fh, **over)
    return dict(B=B, N=N, C=C, H=H, W=W, S=S, cfg=cfg, sds=sds, t=t, video=video)


def build_heads(kind, cfg, sds, device, dtype=torch.float32):
    import vknet
    heads = []
    for sd in sds:
        h = vknet.build_head(dict(type=kind, **cfg))
        h.load_state_dict(sd, strict=True)
        heads.append(h.to(device=device, dtype=dtype).eval())
    return heads


def maxabs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def top2_gap(masks):
    """per-pixel gap between the best and second-best kernel logit ([B,N,H,W] -> [B,H,W])."""
    v = masks.float().topk(2, dim=1).values
    return v[:, 0] - v[:, 1]


def bf16_ulp(t):
    """one bf16 unit in the last place at the magnitude of each element of t (8 significand bits)."""
    a = t.float().abs().clamp_min(2.0 ** -126)
    return torch.exp2(torch.floor(torch.log2(a)) - 7)


def assert_masks_bf16(got, ref32, what, verbose=True):
    """bf16-storage mask logits of the CUDA path against the oracle (fp32 math, stored rounded to bf16).

    * every logit within ONE bf16 ulp of the oracle's stored value (plus 2^-16 of the largest logit: the fp32 round-off a whole
      stage accumulates -- pooling over ~10^4 pixels, LayerNorms, a dozen GEMMs -- which is all that is left of a logit that
      cancels to ~0; still 2^8 finer than the bf16 resolution the logits are stored at);
    * argmax over the kernels identical, except at pixels where the oracle's own top-2 stored logits are within that same
      distance of each other AND the kernel the CUDA path picked is one of those near-ties (the consumer's argmax,
      knet/video/kernel_iter_head.py:854, cannot distinguish them at bf16 resolution).  Returns the number of such pixels.
    """
    got = got.float().cpu()
    ref = ko.round_bf16(ref32.float())
    floor_ = 2.0 ** -16 * ref.abs().max().item()
    err = (got - ref).abs()
    tol = bf16_ulp(torch.maximum(ref.abs(), got.abs())) * (1 + 1e-6) + floor_
    worst = (err / tol).max().item()
    assert worst <= 1.0, '%s: a logit is %.2f x (one bf16 ulp + fp32 noise floor) away from the oracle' % (what, worst)
    nflip = int((err > floor_).sum())
    if ref.shape[1] == 1:
        return 0
    a, b = got.argmax(1), ref.argmax(1)
    bad = a != b
    nbad = int(bad.sum())
    if nbad:
        top2 = ref.topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1])[bad]
        u = bf16_ulp(top2[:, 0])[bad] + floor_
        assert bool((gap <= u).all()), '%s: %d argmax mismatches at pixels whose oracle top-2 gap exceeds one bf16 ulp ' \
            '(max gap %.3g ulp)' % (what, int((gap > u).sum()), (gap / u).max().item())
        picked = ref.gather(1, a.unsqueeze(1)).squeeze(1)[bad]
        assert bool((top2[:, 0][bad] - picked <= u).all()), '%s: the CUDA path picked a kernel that is not a near-tie' % what
    if verbose:
        print('%s: %d of %d logits differ by one bf16 ulp; %d of %d pixels resolve a bf16 near-tie differently' % (
            what, nflip, got.numel(), nbad, a.numel()))
    return nbad


def stagewise_vs_oracle_bf16(heads, sds, cfg, xb, pfd, mb, what, tol=1e-2):
    """Every stage of the CUDA path (bf16 storage) against the oracle evaluated on THAT stage's actual inputs -- the
    CUDA path's own previous outputs -- so no hard-threshold disagreement can cascade between the two and every stage is
    checked unconditionally: kernel tensors within `tol`, mask logits / argmax by assert_masks_bf16.
    Returns the per-stage CUDA outputs and the total number of near-tie pixels resolved differently."""
    obj, m = pfd, mb
    outs, ties = [], 0
    B, N = pfd.shape[:2]
    thr = cfg.get('hard_mask_thr', 0.5)
    for s, h in enumerate(heads):
        cls, m_new, obj_new = h(xb, obj, m)
        # The documented threshold sliver: the CUDA path thresholds the logit (m > logit(thr)), the
        # reference the fp32 sigmoid, which rounds 0 < m <~ 1e-7 to exactly 0.5.  With 10^8 logits per stage such a value does
        # occur; those pixels are resolved the CUDA path's way before the oracle runs, and counted.
        m_cpu = m.float().cpu()
        logit_thr = float(torch.logit(torch.tensor(thr, dtype=torch.float64)))
        sliver = (m_cpu > logit_thr) != (torch.sigmoid(m_cpu) > thr)
        if bool(sliver.any()):
            print('%s stage %d: %d input logit(s) in the sigmoid rounding sliver (%s)' % (
                what, s, int(sliver.sum()), m_cpu[sliver][:4].tolist()))
            m_cpu = torch.where(sliver, torch.where(m_cpu > logit_thr, torch.full_like(m_cpu, logit_thr + 1.0),
                                                    torch.full_like(m_cpu, logit_thr - 1.0)), m_cpu)
        want = ko.kernel_update_head_forward(sds[s], cfg, xb.float().cpu(), obj.float().cpu().reshape(B, N, -1, 1, 1), m_cpu)
        e_cls, e_obj = maxabs(cls, want[0]), maxabs(obj_new, want[2])
        assert e_cls < tol and e_obj < tol, '%s stage %d: cls err %g obj err %g' % (what, s, e_cls, e_obj)
        ties += assert_masks_bf16(m_new, want[1], '%s stage %d' % (what, s))
        outs.append((cls, m_new, obj_new))
        obj, m = obj_new, m_new
    return outs, ties
