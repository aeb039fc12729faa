"""Steady-state device time of each C-ABI operator at the BASELINE shapes: every op is enqueued `reps`
times back to back on one stream (L2-warm, launch overhead pipelined) and timed with CUDA events, plus
the per-kernel table from vkn_profile_begin/end.  Prints one JSON line per batch size.

    python tools/kernel_times.py [B ...]        (default: 1 4)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200')]

import torch  # noqa: E402

import vknet  # noqa: E402
from vknet import _lib, ops  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (head_cfg / dummy_inputs of the benchmark workload)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps      # us


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    N, C, H, W = (bench.CFG1[k] for k in 'NCHW')
    for B in [int(a) for a in sys.argv[1:]] or [1, 4]:
        xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
        x = torch.cat(xs).to(dev).bfloat16()
        pf = torch.cat(pfs).to(dev)
        m = torch.cat(ms).to(dev).bfloat16()
        rows = torch.randn(B, N, C, device=dev)
        out = {}
        out['mask_pool (pool+reduce+ft linear)'] = timeit(lambda: ops.mask_pool(h, x, m))
        xf = ops.mask_pool(h, x, m)
        out['kernel_update (3 linears + rowop)'] = timeit(lambda: ops.kernel_update(h, xf, pf))
        out['mhsa_ln (qkv, attn, out-proj, rowop)'] = timeit(lambda: ops.mhsa_ln(h, rows))
        out['ffn_ln (2 linears + rowop)'] = timeit(lambda: ops.ffn_ln(h, rows))
        out['heads (2 launches)'] = timeit(lambda: ops.heads(h, rows))
        out['mask_gemm (fold linear + conv)'] = timeit(lambda: ops.mask_gemm(h, x, rows))
        out['stage (eager module call)'] = timeit(lambda: h(x, pf, m))
        loop = vknet.KernelIterLoop([h, h, h]).capture(x, pf, m)
        out['3-stage loop (graph replay)'] = timeit(lambda: loop.replay())
        acc = {}
        for _ in range(10):
            with _lib.profile() as p:
                h(x, pf, m)
            for name, t in p.records:
                a = acc.setdefault(name, [0.0, 0])
                a[0] += t
                a[1] += 1
        out['profile_us'] = {k: round(1e3 * v[0] / v[1], 2) for k, v in acc.items()}
        print(json.dumps({'B': B, 'times_us': {k: (round(v, 2) if not isinstance(v, dict) else v) for k, v in out.items()}}))


if __name__ == '__main__':
    main()
