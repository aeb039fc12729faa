"""Where the warp roles of the persistent mask-conv kernel wait (vkn_debug_timestamps accounting), cfg1 shapes:

    python tools/maskgemm_waits.py [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]

import torch  # noqa: E402

import bench  # noqa: E402
import vknet  # noqa: E402
from vknet import _lib, ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    xs, pfs, _ = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x, mk = torch.cat(xs).to(dev).bfloat16(), torch.cat(pfs).to(dev)
    for _ in range(3):
        ops.mask_gemm(h, x, mk)
    torch.cuda.synchronize()
    stride = 4096 * 8
    buf = torch.zeros(8 * stride, dtype=torch.int64, device=dev)
    _lib.lib().vkn_debug_timestamps(_lib.ptr(buf), buf.numel())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.mask_gemm(h, x, mk)
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().vkn_debug_timestamps(None, 0)
    t = buf.reshape(8, 2048, 16).cpu()
    names = ['cta cycles', 'tma: ring slot free', 'mma: x landed', 'mma: acc free', 'epi: acc ready', 'epi: box free', 'tiles', 'marker', 'epi: tmem ld', 'epi: convert+sts', 'epi: fence+bar']
    for blk in t:
        live = blk[:, 7] == 0x6d61736b67656d6d
        if not live.any():
            continue
        r = blk[live].double()
        print('mask conv: %d CTAs, call %.1f us (incl. fold launches)' % (int(live.sum()), 1e3 * e0.elapsed_time(e1)))
        for i, n in enumerate(names):
            print('  %-22s mean %10.0f  min %10.0f  max %10.0f' % (n, r[:, i].mean(), r[:, i].min(), r[:, i].max()))


if __name__ == '__main__':
    main()
