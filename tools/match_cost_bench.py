"""Row f4 timing: vkn_match_cost against the reference's arithmetic (oracle port: sigmoid + clamp x 2, three einsums) on the
GPU (torch) and on the host, at a training-size problem (100 predictions x 30 targets, 200 x 304 masks)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), os.path.join(ROOT, 'oracle')]

import torch  # noqa: E402

import knet_oracle as ko  # noqa: E402
from vknet import _lib, ops  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    N, M, H, W, ncls = 100, 30, 200, 304, 19
    g = torch.Generator().manual_seed(0)
    pred, gt = 3 * torch.randn(N, H, W, generator=g), torch.rand(M, H, W, generator=g).round()
    cls, lab = torch.randn(N, ncls, generator=g), torch.randint(0, ncls, (M,), generator=g)
    pd, gd, cd, ld = pred.to(dev), gt.to(dev), cls.to(dev), lab.to(dev)

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps

    ours = timeit(lambda: ops.match_cost(pd, cd, gd, ld), 50)
    torch_gpu = timeit(lambda: ko.match_cost(pd, cd, gd, ld), 20)
    t0 = time.perf_counter()
    for _ in range(3):
        ko.match_cost(pred, cls, gt, lab)
    cpu = (time.perf_counter() - t0) / 3 * 1e6
    byts = (N + M) * H * W * 4
    with _lib.profile() as p:
        ops.match_cost(pd, cd, gd, ld)
    kern = {name: round(1e3 * t, 1) for name, t in p.records}
    print(json.dumps(dict(shape=[N, M, H, W], ours_us=round(ours, 1), torch_gpu_us=round(torch_gpu, 1), cpu_us=round(cpu, 1),
                          algorithmic_mb=round(byts / 1e6, 2), ours_gbs=round(byts / ours / 1e3, 1),
                          cpu_threads=torch.get_num_threads(), kernels_us=kern)))


if __name__ == '__main__':
    main()
