"""Single-frame latency probe: the S = 3 loop of one cfg1 frame as a CUDA-graph replay (median of 60) + the eager per-kernel
profile, for the environment it is started with (tuning knobs are read from the environment).

    VKN_POOL_CTAS=74 python tools/latency_probe.py [B]
"""
import json
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
    heads = []
    for s in range(3):
        h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
        h.init_weights()
        heads.append(h.to(dev).bfloat16().eval())
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x, pf, m = torch.cat(xs).to(dev).bfloat16(), torch.cat(pfs).to(dev), torch.cat(ms).to(dev).bfloat16()
    loop = vknet.KernelIterLoop(heads).capture(x, pf, m)
    for _ in range(10):
        loop.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    acc = {}
    eager = vknet.KernelIterLoop(heads)
    for _ in range(5):
        with _lib.profile() as p:
            eager(x, pf, m)
        for name, t in p.records:
            a = acc.setdefault(name, [0.0, 0])
            a[0] += t
            a[1] += 1
    env = {k: v for k, v in os.environ.items() if k.startswith('VKN_')}
    print(json.dumps(dict(B=B, env=env, loop_us_median=round(ts[30], 1), loop_us_min=round(ts[0], 1),
                          eager_profile_us={k: round(1e3 * v[0] / v[1], 2) for k, v in acc.items()})))


if __name__ == '__main__':
    main()
