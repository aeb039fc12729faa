"""Per-launch device time of ONE stage at the BASELINE cfg1 shapes, for reading under ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none \
        --csv --log-file gpurun_out/stage_b8.csv python tools/stage_launches.py 8

The profiled region (cudaProfilerStart/Stop) holds exactly one eager `KernelUpdateHead.forward` after three
warm-up calls, so the CSV is the ordered launch list of a stage with warm caches.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]

import torch  # noqa: E402

import vknet  # noqa: E402
import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
    h.init_weights()
    h = h.to(dev).bfloat16().eval()
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x = torch.cat(xs).to(dev).bfloat16()
    pf = torch.cat(pfs).to(dev)
    m = torch.cat(ms).to(dev).bfloat16()
    for _ in range(3):
        h(x, pf, m)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    h(x, pf, m)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
