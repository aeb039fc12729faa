"""Phase timeline of the tcgen05 row-GEMM launches of ONE stage (eager call, PDL on): per launch, min..max over its
CTAs of entry / setup done / dependency resolved / first operands landed / MMAs issued / accumulators ready /
epilogue done / exit, in microseconds relative to the first entry of the stage.

    python tools/rowgemm_timeline.py [B]
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x, pf, m = torch.cat(xs).to(dev).bfloat16(), torch.cat(pfs).to(dev), torch.cat(ms).to(dev).bfloat16()
    loop = vknet.KernelIterLoop([h]).capture(x, pf, m)     # graph replay: launches back to back, PDL edges kept
    for _ in range(5):
        loop.replay()
    torch.cuda.synchronize()
    nlaunch = 48
    stride = 4096 * 8
    buf = torch.zeros(nlaunch * stride + 64, dtype=torch.int64, device=dev)
    _lib.lib().vkn_debug_timestamps(_lib.ptr(buf), buf.numel())
    loop2 = vknet.KernelIterLoop([h]).capture(x, pf, m)    # re-capture with the timestamp blocks baked in
    _lib.lib().vkn_debug_timestamps(None, 0)
    for _ in range(3):
        loop2.replay()
    torch.cuda.synchronize()
    ts = buf[: nlaunch * stride].reshape(nlaunch, 2048, 16).cpu()
    latest = max(int(ts[i][:, 0].max()) for i in range(nlaunch))
    names = ['entry', 'setup', 'dep', 'landed', 'issued', 'acc', 'epi', 'exit']
    t0 = None
    prev_end = None
    for i in range(nlaunch):
        t = ts[i]
        live = t[:, 0] > 0
        if not live.any() or int(t[:, 0].max()) < latest - 2_000_000:      # warm-up launches (stale blocks)
            continue
        t = t[live]
        if t0 is None:
            t0 = int(t[:, 0].min())
        row = []
        for s_ in range(8):
            col = t[:, s_]
            col = col[col > 0]
            if len(col):
                row.append('%s %6.1f..%6.1f' % (names[s_], (int(col.min()) - t0) / 1e3, (int(col.max()) - t0) / 1e3))
        # per-CTA durations
        dur = (t[:, 7] - t[:, 0]).float() / 1e3
        gap = '' if prev_end is None else ' | entry-prev_exit %5.1f' % ((int(t[:, 0].min()) - prev_end) / 1e3)
        extra = ' | epilogue kcycles/CTA: load %.1f store %.1f wait %.1f, tiles %.1f' % (t[:, 8].float().mean() / 1e3, t[:, 9].float().mean() / 1e3, t[:, 10].float().mean() / 1e3, t[:, 11].float().mean())
        print('launch %2d ctas %4d cta-time %4.1f..%4.1f | %s%s%s' % (i, int(live.sum()), dur.min(), dur.max(), ' | '.join(row), gap, extra))
        prev_end = int(t[:, 7].max())


if __name__ == '__main__':
    main()
