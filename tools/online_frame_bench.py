"""One online VPS frame through every row of SURVEY.md section 8 at the KITTI-STEP shapes (375 x 1242 image, 48 x 156 feature map,
100 things + 17 stuff kernels = 117, C = 256, S = 3, bf16 storage): device time of each row (CUDA events, steady state) next to
the reference's arithmetic on the host cores (the oracle port, CPU leg only).

    python tools/online_frame_bench.py
Rows: f1 init proposals | a1-a13 the S-stage loop (graph replay) | a11 link block (previous_type='ffn') | f2 rescale + threshold,
joint panoptic merge, mask -> box | f3 tracking embeddings, association.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), os.path.join(ROOT, 'oracle'), ROOT]

import torch  # noqa: E402

import knet_oracle as ko  # noqa: E402  (CPU baseline leg only)
import vknet  # noqa: E402
from vknet import ops  # noqa: E402


def gpu_us(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def cpu_us(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e6 * (time.perf_counter() - t0) / reps


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    N, C, H, W, S, ncls, nthing, nstuff = 117, 256, 48, 156, 3, 19, 100, 17
    meta = dict(img_shape=(375, 1242, 3), batch_input_shape=(384, 1248), ori_shape=(375, 1242, 3))
    cfg = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=2048)
    cfg_link = ko.default_cfg(num_classes=ncls, in_channels=C, feedforward_channels=2048, previous='placeholder', previous_type='ffn')
    cfgs = [cfg] * (S - 1) + [cfg_link]
    sds = [ko.round_state_dict_bf16(ko.random_state_dict(c, seed=s)) for s, c in enumerate(cfgs)]
    heads = []
    for s, (c, sd) in enumerate(zip(cfgs, sds)):
        h = vknet.build_head(dict(type='VideoKernelUpdateHead' if s == S - 1 else 'KernelUpdateHead', **c))
        h.load_state_dict(sd, strict=True)
        heads.append(h.to(dev).bfloat16().eval())
    x, pf, mask = ko.dummy_inputs(1, N, C, H, W, seed=1)
    xb, pfd, mb = x.to(dev).bfloat16(), pf.to(dev).reshape(1, N, C), mask.to(dev).bfloat16()
    out = {}

    # f1: init proposals (ConvKernelHead tail): static kernels conv + thresholded pooling
    init = torch.nn.Conv2d(C, nthing, 1).to(dev).bfloat16()
    out['f1 init_proposals'] = dict(gpu_us=gpu_us(lambda: ops.init_proposals(init, xb, xb)))

    # a: the loop.  Stages 0..S-2 through the one-call loop; the last (video) stage as its module call (pool, link, stage)
    loop = vknet.KernelIterLoop(heads[:-1]).capture(xb, pfd, mb)
    out['a loop stages 0..%d (graph)' % (S - 2)] = dict(gpu_us=gpu_us(lambda: loop.replay()))
    cls_s, m_s, obj_s = loop.replay()
    prev = torch.randn(1, N, C, 1, 1, device=dev)
    vh = heads[-1]
    out['a last stage, video head + link (eager)'] = dict(gpu_us=gpu_us(lambda: vh(xb, obj_s, m_s, previous_obj_feats=prev)))
    cls, nm, obj, x_feat, track = vh(xb, obj_s, m_s, previous_obj_feats=prev)
    cpu = [None]

    def cpu_loop():
        o, m = pf, ko.round_bf16(mask)
        for s in range(S - 1):
            _, m, o = ko.kernel_update_head_forward(sds[s], cfg, ko.round_bf16(x), o, m)
        cpu[0] = ko.video_kernel_update_head_forward(sds[-1], cfg_link, ko.round_bf16(x), o, m, previous_obj_feats=prev.cpu())
    out['a loop (all stages + link)'] = dict(cpu_us=cpu_us(cpu_loop, 2))

    # f2: last-stage x2 upsample + rescale + threshold; joint panoptic merge; mask -> box
    logits = nm[0]                                              # [N, 48, 156] bf16
    out['f2 rescale_masks (probs)'] = dict(gpu_us=gpu_us(lambda: ops.rescale_masks(logits, meta, 2, None, probs=True)),
                                           cpu_us=cpu_us(lambda: ko.rescale_masks(logits.float().cpu(), meta, 2), 2))
    probs = ops.rescale_masks(logits, meta, 2, None, probs=True)[0]
    scores = torch.rand(N, device=dev) * 0.8 + 0.2
    labels = torch.cat([torch.randint(0, 2, (nthing,)), torch.arange(nstuff) + 2]).to(dev)

    def merge():
        return ops.panoptic_merge(probs[:nthing], labels[:nthing], scores[:nthing], probs[nthing:], labels[nthing:], scores[nthing:], 2, 0.3, 0.5)
    pc, lc, sc = probs.cpu(), labels.cpu(), scores.cpu()
    out['f2 panoptic_merge (joint)'] = dict(gpu_us=gpu_us(merge, 20), cpu_us=cpu_us(
        lambda: ko.panoptic_merge_joint(pc[:nthing], lc[:nthing], sc[:nthing], pc[nthing:], lc[nthing:], sc[nthing:], 2, 0.3, 0.5), 2))
    bits = probs[:nthing] > 0.5
    out['f2 mask_boxes'] = dict(gpu_us=gpu_us(lambda: ops.mask_boxes(bits)))

    # f3: tracking embeddings (embed_fcs + fc_embed + track head) and the quasi-dense association against 30 memorised tracks
    mods = [(torch.nn.Linear(C, C, bias=False), torch.nn.LayerNorm(C), True), (torch.nn.Linear(C, C), None, False),
            (torch.nn.Linear(C, C), None, True), (torch.nn.Linear(C, C), None, True), (torch.nn.Linear(C, C), None, False)]
    mods = [(l.to(dev).bfloat16(), None if n is None else n.to(dev), r) for l, n, r in mods]
    rows = obj.reshape(N, C)[:nthing].float().contiguous()
    out['f3 tracking embeddings (5 Linear)'] = dict(gpu_us=gpu_us(lambda: ops.mlp(mods, rows)))
    emb = ops.mlp(mods, rows)
    g = torch.Generator().manual_seed(2)
    xy = torch.rand(nthing, 2, generator=g) * 300
    bboxes = torch.cat([xy, xy + 40 + 60 * torch.rand(nthing, 2, generator=g), torch.rand(nthing, 1, generator=g)], 1).to(dev)
    tl = torch.randint(0, 2, (nthing,), generator=g).to(dev)
    memo = (torch.randint(0, 2, (30,), generator=g).to(dev), torch.randn(30, C, generator=g).to(dev), torch.arange(30).to(dev))

    def match():
        return ops.track_match(bboxes, tl, emb, memo[0], memo[1], memo[2], 30, 0.3, 0.5, 0.35, 0.5, 0.3, 0.7, True)
    out['f3 track_match'] = dict(gpu_us=gpu_us(match, 20), cpu_us=cpu_us(
        lambda: ko.tracker_match(bboxes.cpu(), tl.cpu(), emb.cpu(), memo[0].cpu(), memo[1].cpu(), memo[2].cpu(), 30, 0.3, 0.5, 0.35, 0.5, 0.3, 0.7,
                                 True), 2))
    tot = sum(v.get('gpu_us', 0.0) for v in out.values())
    print(json.dumps(dict(shape='KITTI-STEP frame: 375x1242, features 48x156, N=117 (100 things + 17 stuff), C=256, S=3, bf16',
                          rows={k: {a: round(b, 1) for a, b in v.items()} for k, v in out.items()}, gpu_total_us=round(tot, 1),
                          cpu_threads=torch.get_num_threads(),
                          note='gpu_us include the Python / ctypes call overhead of eager calls (the loop stages are a graph replay); '
                               'panoptic_merge and track_match include their one device->host read')))


if __name__ == '__main__':
    main()
