"""Phase timeline of the single-frame cluster chain (framechain.cu): for every chain launch of one 3-stage loop call, the time
each phase boundary was reached (microseconds from the first CTA's entry; min / median / max over the launch's CTAs).

    python tools/frame_chain_timeline.py [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]

import torch  # noqa: E402

import bench  # noqa: E402
import vknet  # noqa: E402
from vknet import _lib  # noqa: E402

NAMES_A = {0: 'entry', 1: 'ring+vec issued', 2: 'cluster sync 0', 3: 'dependency ok', 4: 'planes built', 5: 'x_feat GEMM, hand-off issued',
           6: 'input_layer GEMM, x_feat in', 7: 'gemm2 (dynamic_layer)', 8: 'gate_feats handed', 9: 'gemm3 (gates)', 10: 'LN stats', 11: 'features handed',
           12: 'gemm4 (fc_layer)', 13: 'LN stats', 14: 'obj0 handed', 30: 'end (in_proj stored)'}
NAMES_B = {0: 'entry', 1: 'ring+vec issued', 2: 'cluster sync 0', 3: 'dependency ok', 4: 'k/v loaded', 5: 'attention', 6: 'att handed',
           7: 'gemm out_proj', 8: 'LN stats', 9: 'o1 handed', 10: 'ffn1 a', 11: 'ffn1 b', 12: 'ffn2 a', 13: 'ffn2 b', 14: 'partials scattered',
           15: 'LN stats', 16: 'obj handed', 17: 'gemm cls/mask fc', 18: 'LN stats', 19: 'planes handed', 20: 'fc_mask GEMM, hand-off issued',
           21: 'fc_cls GEMM, mk in', 30: 'end (fold stored)'}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x, pf, m = torch.cat(xs).to(dev).bfloat16(), torch.cat(pfs).to(dev), torch.cat(ms).to(dev).bfloat16()
    loop = vknet.KernelIterLoop([h, h, h])
    for _ in range(5):
        loop(x, pf, m)
    torch.cuda.synchronize()
    nlaunch, stride = 6, 4096 * 8
    buf = torch.zeros(nlaunch * stride + 64, dtype=torch.int64, device=dev)
    _lib.lib().vkn_debug_timestamps(_lib.ptr(buf), buf.numel())
    loop(x, pf, m)
    torch.cuda.synchronize()
    _lib.lib().vkn_debug_timestamps(None, 0)
    ts = buf[: nlaunch * stride].reshape(nlaunch, 1024, 32).cpu()
    t00 = None
    for i in range(nlaunch):
        t = ts[i]
        t = t[t[:, 0] > 0]
        if t.numel() == 0:
            continue
        names = NAMES_A if i % 2 == 0 else NAMES_B
        t0 = int(t[:, 0].min())
        t00 = t0 if t00 is None else t00
        print('--- launch %d (%s), %d CTAs, entry at %.1f us' % (i, 'A' if i % 2 == 0 else 'B', t.shape[0], (t0 - t00) / 1e3))
        prev = 0.0
        for s_ in sorted(names):
            col = t[:, s_].double()
            col = col[col > 0]
            if col.numel() == 0:
                continue
            rel = (col - t0) / 1e3
            med = float(rel.median())
            print('  %-30s min %7.2f  med %7.2f  max %7.2f   (+%.2f)' % (names[s_], float(rel.min()), med, float(rel.max()), med - prev))
            prev = med


if __name__ == '__main__':
    main()
