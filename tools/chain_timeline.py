"""Step timeline of the chain kernel launches of ONE stage (vkn_debug_timestamps): for every vkn_chain_tc_kernel launch the
mean time (us, over CTAs) at which row tiles X / Y of the CTA's first pair finished each step, and the cycle counters of
the three roles (TMA warp, MMA warp, epilogue warps).

    python tools/chain_timeline.py [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]

import torch  # noqa: E402

import bench  # noqa: E402
import vknet  # noqa: E402
from vknet import _lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(min(B, 8))])
    rep = (B + len(xs) - 1) // len(xs)
    x = torch.cat(xs).repeat(rep, 1, 1, 1)[:B].to(dev).bfloat16()
    pf = torch.cat(pfs).repeat(rep, 1, 1)[:B].to(dev)
    m = torch.cat(ms).repeat(rep, 1, 1, 1)[:B].to(dev).bfloat16()
    for _ in range(3):
        h(x, pf, m)
    torch.cuda.synchronize()
    nlaunch, stride = 16, 4096 * 8
    buf = torch.zeros(nlaunch * stride + 64, dtype=torch.int64, device=dev)
    _lib.lib().vkn_debug_timestamps(_lib.ptr(buf), buf.numel())
    with _lib.profile() as prof:
        h(x, pf, m)
    _lib.lib().vkn_debug_timestamps(None, 0)
    torch.cuda.synchronize()
    for n, ms_ in prof.records:
        print('%-34s %8.1f us' % (n, ms_ * 1e3))
    ts = buf[: nlaunch * stride].reshape(nlaunch, stride // 64, 64).cpu()
    for i in range(nlaunch):
        t = ts[i]
        live = t[:, 0] > 0
        if not live.any():
            continue
        t = t[live].double()
        t0 = t[:, 0]
        print('chain launch %d: %d CTAs, CTA time %.1f us (mean), %.1f (max)' % (
            i, int(live.sum()), ((t[:, 31] - t0).mean()) / 1e3, ((t[:, 31] - t0).max()) / 1e3))
        prev = 0.0
        for si in range(15):
            ex, ey = t[:, 1 + 2 * si], t[:, 2 + 2 * si]
            if (ex > 0).any():
                mx, my = ((ex - t0)[ex > 0].mean()) / 1e3, ((ey - t0)[ey > 0].mean()) / 1e3
                print('   step %2d  X done %7.1f  Y done %7.1f   (+%.1f)' % (si, mx, my, my - prev))
                prev = my
        clk = 1.9e3    # cycles per us (approx.)
        names = {32: 'epi wait acc_full', 33: 'epi tile work', 34: 'epi edge', 35: 'epi row steps', 40: 'mma wait operands',
                 41: 'mma wait acc_empty', 48: 'tma wait slot', 49: 'tma wait steps'}
        print('   ' + ', '.join('%s %.1f us' % (v, t[:, k].mean() / clk) for k, v in names.items()))


if __name__ == '__main__':
    main()
