"""Device time of the fused post-loop mask path (vkn_rescale_masks) at the KITTI-STEP shape, against its output-byte
roofline and the CPU oracle (the reference's torch calls) on the host cores.

    python tools/rescale_bench.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), os.path.join(ROOT, 'oracle'), ROOT]

import torch  # noqa: E402

import knet_oracle as ko  # noqa: E402  (CPU baseline leg only)
from vknet import ops  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    K, H, W, up = 100, 48, 156, 2
    meta = dict(img_shape=(375, 1242, 3), batch_input_shape=(384, 1248), ori_shape=(375, 1242, 3))
    g = torch.Generator().manual_seed(0)
    masks = (torch.randn(K, H, W, generator=g) * 4).bfloat16()
    md = masks.to(dev)
    peak = 6547.5
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = float(json.load(open(p))['hbm_gbs'])
    out = {}
    for name, kw in (('bits', dict(mask_thr=0.5, probs=False)), ('probs', dict(mask_thr=None, probs=True)),
                     ('probs+bits', dict(mask_thr=0.5, probs=True))):
        for _ in range(5):
            ops.rescale_masks(md, meta, up, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            ops.rescale_masks(md, meta, up, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        nbytes = K * 375 * 1242 * ((1 if kw['mask_thr'] is not None else 0) + (4 if kw['probs'] else 0)) + K * H * W * 2
        out[name] = dict(us=us, algorithmic_bytes=nbytes, gbs=nbytes / us / 1e3, frac_of_hbm_peak=nbytes / us / 1e3 / peak)
    t0 = time.perf_counter()
    ko.rescale_masks(masks.float(), meta, up)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    print(json.dumps(dict(shape='K=100 masks 48x156 bf16, x2, batch 384x1248, crop/ori 375x1242', hbm_peak_gbs=peak, kernel=out,
                          cpu_oracle_ms=cpu_ms, cpu_threads=torch.get_num_threads())))


if __name__ == '__main__':
    main()
