"""In-graph-free timeline of the row-operator launches of one stage loop: per launch, when its CTAs entered,
when the programmatic dependency resolved, when panel / tile were ready, when the main loop and the stores were
done (all relative to the first entry of the first launch, in microseconds; min..max over CTAs).

    python tools/linear_timeline.py [B]
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
    nlaunch = 33
    stride = 4096 * 8
    buf = torch.zeros(nlaunch * stride + 64, dtype=torch.int64, device=dev)
    _lib.lib().vkn_debug_timestamps(_lib.ptr(buf), buf.numel())
    loop(x, pf, m)
    torch.cuda.synchronize()
    _lib.lib().vkn_debug_timestamps(None, 0)
    ts = buf[: nlaunch * stride].reshape(nlaunch, 4096, 8).cpu()
    t0 = None
    names = ['entry', 'prefetch', 'dep ok', 'panel', 'visible', 'loop', 'stored']
    prev_end = None
    for i in range(nlaunch):
        t = ts[i]
        live = t[:, 0] > 0
        if not live.any():
            continue
        t = t[live]
        if t0 is None:
            t0 = int(t[:, 0].min())
        row = []
        for s_ in range(7):
            col = t[:, s_]
            col = col[col > 0]
            row.append('%s %6.1f..%6.1f' % (names[s_], (int(col.min()) - t0) / 1e3, (int(col.max()) - t0) / 1e3))
        end = int(t[:, 6].max())
        gap = '' if prev_end is None else ' | entry-prev_end %5.1f us' % ((int(t[:, 0].min()) - prev_end) / 1e3)
        print('launch %2d ctas %4d | %s%s' % (i, int(live.sum()), ' | '.join(row), gap))
        prev_end = end


if __name__ == '__main__':
    main()
