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
