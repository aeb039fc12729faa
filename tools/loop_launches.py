"""One call of the S=3 one-call loop (vkn_iter_forward) at the BASELINE cfg1 shapes inside a cudaProfilerStart/Stop
region, for ncu (launch lists and --set full captures of the loop's kernels, bit-mask hand-off included):

    ncu --profile-from-start off --set full -k regex:pool_tc|maskgemm -c 6 -o rep python tools/loop_launches.py 64
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]

import torch  # noqa: E402

import bench  # noqa: E402
import vknet  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    heads = []
    for _ in range(3):
        h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
        h.init_weights()
        heads.append(h.to(dev).bfloat16().eval())
    xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(B)])
    x, pf, m = torch.cat(xs).to(dev).bfloat16(), torch.cat(pfs).to(dev), torch.cat(ms).to(dev).bfloat16()
    loop = vknet.KernelIterLoop(heads)
    for _ in range(2):
        loop(x, pf, m)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loop(x, pf, m)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
