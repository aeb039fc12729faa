#!/usr/bin/env python
"""bench.py -- frames/s/GPU of the S-stage KernelUpdateHead loop (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path over one batch of synthetic input: ONE frame per rank
(cfg1: N=100 kernels, C=256, 200x88 feature map, S=3 stages, bf16 storage).  Frames are independent, so the
throughput mode keeps VKN_STREAMS x VKN_BATCH (default 3 x 126) frames in flight per CUDA-graph launch and rank; the
timed region runs whole launches: at least 20 of them and at least K frames, per-launch CUDA events (median / p95 under
`timing`), and `ms_per_step` = timed device time / frames processed.
With more than one GPU every launch also performs the one exchange frame sharding needs (all-gather of the ranks'
shard-boundary kernels over NCCL + the `previous_type='ffn'` link block, cfg3), on a side stream.

  value     frames/s, whole job, inputs resident in HBM, CUDA-graph replay of the loop, CUDA-event timed,
            max over ranks.  Inputs rotate over R distinct sets whose footprint exceeds L2.
  e2e       same metric through the public API with HOST (pinned) inputs: H2D copies of x / kernels /
            masks and the D2H read of the result tuple are inside the timed region.
  roofline  the kernel family with the largest share of the step, timed live with CUDA events on its launch stream
            (vkn_profile_begin/end): row GEMMs -> algorithmic FLOPs vs the measured dense-bf16 peak; pooling / mask conv
            -> algorithmic bytes vs the measured HBM copy peak (all families under roofline.families).
  cpu_baseline  the CPU oracle port of the reference's PyTorch path (fp32, all host threads) on a bounded
            sample of the same workload.
--impl reference runs ONLY that CPU arm (rank 0) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200')]

CFG1 = dict(B=1, N=100, C=256, H=200, W=88, S=3, ncls=19, ffn=2048)
METRIC = 'frames/sec/GPU (100 kernels, C=256, 200x88, S=3)'
# the workload both arms are quoted on (BASELINE.json configs[1]); frames are independent, so throughput = frames/s/GPU
WORKLOAD = 'cfg1 KITTI-STEP R-50 shape per frame: N=100 kernels, C=256, 200x88 feature map, S=3 stages'
# --config: the other shapes BASELINE.json / SURVEY.md 8d name (per-frame KernelUpdateHead loop, same kernels); their lines
# are kept under profiles/.  cfg1 is the one the headline metric is quoted on.
CONFIGS = {
    'cfg1': dict(N=100, H=200, W=88, ncls=19, workload=WORKLOAD),
    'n117': dict(N=117, H=200, W=88, ncls=19, workload='cfg1 feature map with the real KITTI/Cityscapes-STEP kernel count: N=117 '
                                                       '(100 things + 17 stuff kernels), C=256, 200x88, S=3'),
    'n166': dict(N=166, H=200, W=88, ncls=124, workload='cfg1 feature map with the VIP-Seg kernel count: N=166 (100 + 66 stuff), '
                                                        'C=256, 200x88, S=3'),
    'cfg2': dict(N=100, H=96, W=160, ncls=40, workload='cfg2 YouTube-VIS frame shape: N=100, C=256, 96x160, S=3 (per-frame heads; '
                                                       'the clip head with frames_per_set=4 is covered by the parity tests)'),
    'cfg4': dict(N=100, H=120, W=216, ncls=124, workload='cfg4 VIP-Seg frame shape: N=100, C=256, 120x216, S=3'),
    'cfg4_n166': dict(N=166, H=120, W=216, ncls=124, workload='cfg4 VIP-Seg frame shape with its real kernel count: N=166, C=256, '
                                                              '120x216, S=3'),
}


def head_cfg(link=False):
    C = CFG1['C']
    cfg = dict(num_classes=CFG1['ncls'], num_thing_classes=2, num_stuff_classes=17, num_ffn_fcs=2, num_heads=8,
               num_cls_fcs=1, num_mask_fcs=1, feedforward_channels=CFG1['ffn'], in_channels=C, out_channels=C,
               dropout=0.0, mask_thr=0.5, conv_kernel_size=1, mask_upsample_stride=2,
               ffn_act_cfg=dict(type='ReLU', inplace=True), with_ffn=True,
               feat_transform_cfg=dict(conv_cfg=dict(type='Conv2d'), act_cfg=None),
               kernel_updator_cfg=dict(type='KernelUpdator', in_channels=C, feat_channels=C, out_channels=C,
                                       input_feat_shape=3, act_cfg=dict(type='ReLU', inplace=True),
                                       norm_cfg=dict(type='LN')),
               loss_rank=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.1),
               loss_mask=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
               loss_dice=dict(type='DiceLoss', loss_weight=4.0),
               loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0))
    if link:
        cfg.update(previous='placeholder', previous_type='ffn')
    return cfg


def dummy_inputs(torch, seed):
    """forward_dummy recipe (knet/det/kernel_iter_head.py:317-330)."""
    g = torch.Generator().manual_seed(seed)
    B, N, C, H, W = (CFG1[k] for k in 'BNCHW')
    x = torch.randn(B, C, H, W, generator=g)
    pf = torch.randn(B, N, C, generator=g)
    mask = pf.bmm(x.view(B, C, -1)).view(B, N, H, W)
    return x, pf, mask


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:   # noqa: BLE001 -- nvidia-smi missing: report empty clocks
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = []
        for i, name in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=reasons, samples=len(sm))


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, source)"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops', 1650.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's PyTorch path (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, budget_s=120.0):
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import knet_oracle as ko
    cores = os.cpu_count() or 1
    cfg = ko.default_cfg(num_classes=CFG1['ncls'], in_channels=CFG1['C'], feedforward_channels=CFG1['ffn'])
    sds = [ko.random_state_dict(cfg, seed=s) for s in range(CFG1['S'])]
    x, pf, mask = ko.dummy_inputs(CFG1['B'], CFG1['N'], CFG1['C'], CFG1['H'], CFG1['W'], seed=1)
    # "all the host threads it can use": torch's intra-op pool stops scaling (and then degrades) on these
    # small ops, so probe a few pool sizes and keep the fastest -- the reference arm gets its best case.
    best = (None, 1e30)
    with torch.no_grad():
        for nt in sorted({1, 4, 8, 16, 32, 64, cores} & set(range(1, cores + 1))):
            torch.set_num_threads(nt)
            ko.iter_forward(sds, [cfg] * CFG1['S'], x, pf, mask)
            t0 = time.perf_counter()
            ko.iter_forward(sds, [cfg] * CFG1['S'], x, pf, mask)
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (nt, dt)
    threads = best[0]
    torch.set_num_threads(threads)
    times = []
    with torch.no_grad():
        for _ in range(max(1, min(warmup, 3))):
            ko.iter_forward(sds, [cfg] * CFG1['S'], x, pf, mask)
        t_all = time.perf_counter()
        for _ in range(max(steps, 1)):
            t0 = time.perf_counter()
            ko.iter_forward(sds, [cfg] * CFG1['S'], x, pf, mask)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > budget_s:
                break
    times.sort()
    med = times[len(times) // 2]
    return dict(value=1.0 / med, unit='frames/s', cores=threads, kind='port', host_cores=cores,
                sample='%d frames of cfg1 (fp32, torch %s, best of 1..%d threads = %d), median %.2f ms/frame; the '
                       'reference is pure Python and cannot travel to this box: the port restates it op for op and is '
                       'pinned to it by tests/golden' % (len(times), torch.__version__, cores, threads, med * 1e3)), med


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb, med = cpu_arm(args.steps, args.warmup)
    line = dict(impl='reference', metric=METRIC, value=cb['value'], unit='frames/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=med * 1e3, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload=WORKLOAD,
                            note='reference CPU PyTorch path (oracle port, fp32), one frame per step, best host thread count'),
                cpu_baseline=cb,
                e2e=dict(value=cb['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def parity_check(torch, heads, fif, dev):
    """`parity` of the JSON line: the CUDA path against the CPU oracle on frames of the TIMED batch (rank 0, outside the
    timed region; the oracle is the checker here, never the thing measured).
    * stage-wise: every stage of the batch (module calls at the throughput batch size: tcgen05 engines + chain kernel)
      against the oracle evaluated on that stage's actual inputs, for `nf` frames: max-abs error of obj_feat / cls_score,
      logits more than one bf16 ulp (+2^-16 of the largest logit) away, argmax mismatches that are NOT bf16 near-ties of
      the oracle, near-tie pixels resolved differently, hard-threshold (logit > 0) disagreements handed to the next stage;
    * loop: the graph-replayed one-call loop (bit-mask hand-off) must equal those stage-wise calls bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import knet_oracle as ko
    st = fif.static[0]
    xb, pfd, mb = st['x'], st['pf'], st['mask']
    B, N, C = pfd.shape
    nf = min(2, B)
    cfg = ko.default_cfg(num_classes=CFG1['ncls'], in_channels=C, feedforward_channels=CFG1['ffn'])
    res = dict(frames_checked=nf, batch=B, obj_max_abs=0.0, cls_max_abs=0.0, logits_beyond_one_bf16_ulp=0,
               argmax_mismatch_not_near_tie=0, near_tie_pixels_resolved_differently=0, threshold_disagreements=0,
               pixels=0)

    def ulp(t):
        return torch.exp2(torch.floor(torch.log2(t.abs().clamp_min(2.0 ** -126))) - 7)
    obj, m = pfd, mb
    with torch.no_grad():
        for h in heads:
            sd = {k: v.detach().float().cpu() for k, v in h.state_dict().items()}
            cls, m_new, obj_new = h(xb, obj, m)[:3]
            m_cpu = m[:nf].float().cpu()
            sliver = (m_cpu > 0) != (torch.sigmoid(m_cpu) > 0.5)     # fp32 sigmoid rounds 0 < m <~ 1e-7 to 0.5
            res['threshold_sliver_logits'] = res.get('threshold_sliver_logits', 0) + int(sliver.sum())
            m_cpu = torch.where(sliver, torch.where(m_cpu > 0, torch.ones_like(m_cpu), -torch.ones_like(m_cpu)), m_cpu)
            want = ko.kernel_update_head_forward(sd, cfg, xb[:nf].float().cpu(), obj[:nf].float().cpu().reshape(nf, N, C, 1, 1), m_cpu)
            res['obj_max_abs'] = max(res['obj_max_abs'], float((obj_new[:nf].float().cpu().reshape(want[2].shape) - want[2]).abs().max()))
            res['cls_max_abs'] = max(res['cls_max_abs'], float((cls[:nf].float().cpu() - want[0]).abs().max()))
            got, ref = m_new[:nf].float().cpu(), ko.round_bf16(want[1])
            floor_ = 2.0 ** -16 * float(ref.abs().max())
            res['logits_beyond_one_bf16_ulp'] += int(((got - ref).abs() > ulp(torch.maximum(ref.abs(), got.abs())) * (1 + 1e-6) + floor_).sum())
            bad = got.argmax(1) != ref.argmax(1)
            top2 = ref.topk(2, dim=1).values
            tie = (top2[:, 0] - top2[:, 1]) <= ulp(top2[:, 0]) + floor_
            res['near_tie_pixels_resolved_differently'] += int((bad & tie).sum())
            res['argmax_mismatch_not_near_tie'] += int((bad & ~tie).sum())
            res['threshold_disagreements'] += int(((got > 0) != (want[1] > 0)).sum())
            res['pixels'] += bad.numel()
            obj, m = obj_new.reshape(B, N, C), m_new
        outs = fif.replay()[0]
        torch.cuda.synchronize()
        plain = all(type(h).__name__ == 'KernelUpdateHead' for h in heads)
        if plain:                   # same kernels on both routes: bit-identical
            res['loop_equals_stagewise_bitwise'] = bool(torch.equal(outs[1], m) and torch.equal(outs[2].reshape(B, N, C), obj) and
                                                        torch.equal(outs[0], cls))
        else:                       # the video head's module call pools through vkn_mask_pool first (a different, equally exact route)
            res['loop_vs_stagewise_obj_max_abs'] = float((outs[2].reshape(B, N, C) - obj).abs().max())
            res['loop_vs_stagewise_logits_differing'] = int((outs[1] != m).sum())
    res['tolerance'] = 'kernel tensors 1e-2 (bf16 storage); logits one bf16 ulp; argmax identical up to oracle near-ties'
    res['ok'] = bool(res['obj_max_abs'] < 1e-2 and res['cls_max_abs'] < 1e-2 and res['logits_beyond_one_bf16_ulp'] == 0 and
                     res['argmax_mismatch_not_near_tie'] == 0 and res.get('loop_equals_stagewise_bitwise', True) and
                     res.get('loop_vs_stagewise_obj_max_abs', 0.0) < 1e-3)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import vknet
    from vknet import _lib
    from vknet import dist as vdist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG >= VERSION
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG', None)
        os.environ.setdefault('NCCL_DEBUG_FILE', os.devnull)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    S = CFG1['S']
    B, N, C, H, W = (CFG1[k] for k in 'BNCHW')
    HW = H * W
    with_link = world > 1 or bool(os.environ.get('VKN_BENCH_LINK'))     # cfg3: + the previous_type='ffn' link block

    # random-init weights of the KITTI-STEP R-50 head (init_weights semantics), bf16 storage
    torch.manual_seed(0)
    heads = []
    for s in range(S):
        last_linked = with_link and s == S - 1
        h = vknet.build_head(dict(type='VideoKernelUpdateHead' if last_linked else 'KernelUpdateHead',
                                  **head_cfg(link=last_linked)))
        h.init_weights()
        eng = os.environ.get('VKN_ENGINE', 'auto')
        h.engine = dict(auto=_lib.ENGINE_AUTO, simt=_lib.ENGINE_SIMT, tc=_lib.ENGINE_TC)[eng]
        heads.append(h.to(dev).bfloat16().eval())
    loop = vknet.KernelIterLoop(heads)

    # Frames in flight: one CUDA graph = NS concurrent branches x BF frames each (vknet.FramesInFlight).  Frames are
    # independent; a branch of 126 frames = 12 600 kernel rows = 99 row tiles keeps the chain kernel, the
    # pooling and the mask conv on (nearly) every SM, three branches overlap the HBM-bound and the tensor-bound kernels of
    # different frames.  Two such groups alternate; one group's inputs + outputs exceed L2 (126 MB) many times over, so x and
    # the masks stream from HBM.  Measured sweep (profiles/r2_inflight_sweep.md): 3x64 42.5k, 2x148 47.0k, 3x126 46.6k,
    # 2x378 47.4k frames/s; 3x126 has the best end-to-end rate.
    NS = max(1, int(os.environ.get('VKN_STREAMS', '3')))
    BF = max(1, int(os.environ.get('VKN_BATCH', '126')))
    quick = bool(os.environ.get('VKN_BENCH_QUICK'))       # profiler runs: small group, no CPU arm
    if quick and 'VKN_STREAMS' not in os.environ and 'VKN_BATCH' not in os.environ:
        NS, BF = 1, 1
    FPG = NS * BF                                           # frames per graph launch (per rank)
    set_bytes = (C * HW + 2 * N * HW) * 2
    G = max(2, int(160e6 // (set_bytes * FPG)) + 1)
    seeds = iter(range(1 + rank * 100000, 10 ** 9))

    def frame_batch():
        """BF frames of the forward_dummy recipe (knet/det/kernel_iter_head.py:317-330), generated on the device (setup, not
        timed; the e2e groups below are copied to pinned host memory): x ~ N(0,1), kernels ~ N(0,1), masks = kernels . x"""
        g = torch.Generator(device=dev).manual_seed(next(seeds))
        x = torch.randn(BF, C, H, W, generator=g, device=dev)
        pf = torch.randn(BF, N, C, generator=g, device=dev)
        mask = pf.bmm(x.view(BF, C, -1)).view(BF, N, H, W)
        return x.bfloat16().cpu(), pf.cpu(), mask.bfloat16().cpu()

    groups, host_groups = [], []
    for g_ in range(G):
        batches = [frame_batch() for _ in range(NS)]
        fif = vknet.FramesInFlight(heads, branches=NS, batch=BF)
        fif.capture([tuple(t_.to(dev) for t_ in bt) for bt in batches])
        groups.append(fif)
        if g_ < 2:                                          # e2e groups: same frames, pinned host buffers in/out
            pinned = [tuple(t_.pin_memory() for t_ in bt) for bt in batches]
            hf = vknet.FramesInFlight(heads, branches=NS, batch=BF)
            hf.loops = fif.loops                            # share workspaces
            hf.capture(pinned, host_io=True)
            host_groups.append(hf)
    x1, pf1, m1 = dummy_inputs(torch, 7)
    st0 = groups[0].static[0]
    before = _lib.launch_count()
    loop(st0['x'], st0['pf'], st0['mask'])                      # one vkn_iter_forward at the throughput batch size
    launches_per_frame_call = _lib.launch_count() - before      # kernels of ONE call = one branch of a graph launch

    # ---- cfg3 exchange: the ranks' frames are consecutive blocks of one clip; frame t's tracking kernels attend to frame
    #      t-1's kernels.  A rank needs ONE frame it does not own (its left neighbour's last): one all-gather of world x
    #      [N, C] (102 KB each) over NCCL, then the link block on the rank's own FPG frames.  Both run on a side stream so
    #      that they overlap the next graph launch (the groups alternate, so its inputs / outputs are different buffers).
    if with_link:
        link_head = heads[-1]
    else:                                                   # one GPU: a link block of its own for the cfg3 side measurement below
        torch.manual_seed(1)
        link_head = vknet.build_head(dict(type='VideoKernelUpdateHead', **head_cfg(link=True)))
        link_head.init_weights()
        link_head = link_head.to(dev).bfloat16().eval()
    do_link = [with_link]                                  # switched on for the cfg3 pass at one GPU
    side = torch.cuda.Stream(device=dev)
    link_done = [None] * len(groups)
    track_host = torch.empty(FPG, N, C).pin_memory()
    link_state = {}

    def link_fn_for(nf):
        if nf not in link_state:
            w, links, wd = link_head.packed_weights(dev)
            shape = link_head._shape(nf, N, H, W, _lib.VKN_BF16, wd)
            ws = _lib.Workspace()
            link_state[nf] = (links, shape, ws)
        links, shape, ws = link_state[nf]
        wsp, wsb = ws.get(shape, dev)
        return lambda cur, prev: link_head._link(shape, links['track'], cur.contiguous(), prev.contiguous(), None, wsp, wsb)

    def exchange(outs):
        obj_local = torch.cat([o[2].reshape(-1, N, C) for o in outs], dim=0)
        return vdist.link_sharded_clip_boundary(link_fn_for(obj_local.shape[0]), obj_local, world * obj_local.shape[0], rank, world)

    def run(launches, grp, host=False, times=None, start=0):
        """`launches` graph launches (FPG frames each), rotating over the groups; per-launch CUDA events into `times`."""
        cur = torch.cuda.current_stream(dev)
        for i in range(start, start + launches):
            gi = i % len(grp)
            if do_link[0] and link_done[gi] is not None:
                cur.wait_event(link_done[gi])                   # the link of this group's previous outputs has read them
            if times is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(cur)
                times.append(ev)
            outs = grp[gi].replay()
            if do_link[0]:
                ready = torch.cuda.Event()
                ready.record(cur)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    track = exchange(outs)
                    if host:
                        track_host.copy_(track, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                link_done[gi] = done
        if do_link[0]:
            cur.wait_stream(side)
        return launches * FPG

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Timed region: at least 20 graph launches and at least `--steps` frames, every group visited; per-launch events give
    # the median / p95 of a launch.  `ms_per_step` = total device time / frames processed (a step = one frame per rank).
    n_launch = max(20, -(-args.steps // FPG), len(groups))
    run(max(3, -(-args.warmup // FPG)), groups)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    e1 = torch.cuda.Event(enable_timing=True)
    evs = []
    barrier()
    frames_done = run(n_launch, groups, times=evs, start=1)
    e1.record()
    barrier()
    total_ms = evs[0].elapsed_time(e1)
    per_launch = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1))
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    ms = total_ms / frames_done                             # device time per frame and rank (max over ranks)

    # ---- one GPU: the same timed loop WITH the cfg3 link block on every launch (what each rank of an N-GPU run does besides
    #      the 102 KB exchange): separates the cost of the link arithmetic from the cost of the collective in the 1 -> N curve
    cfg3_single = None
    if not with_link and not quick:
        do_link[0] = True
        run(3, groups)
        barrier()
        evs3 = []
        e3 = torch.cuda.Event(enable_timing=True)
        fr3 = run(n_launch, groups, times=evs3, start=1)
        e3.record()
        barrier()
        ms3 = evs3[0].elapsed_time(e3) / fr3
        cfg3_single = dict(value=1.0 / (ms3 * 1e-3), ms_per_step=ms3,
                           note='cfg1 + the previous_type=ffn link block on every frame (side stream), one GPU, no exchange: the '
                                'per-rank workload of the N-GPU lines; value(N) / (N x this) isolates the collective')
        do_link[0] = False
        for i_ in range(len(link_done)):
            link_done[i_] = None

    # ---- e2e: host buffers in, result tuple out, copies inside the timed region ---------------------
    h2d = (C * HW + N * HW) * 2 + N * C * 4
    d2h = (N * CFG1['ncls'] + N * C) * 4 + N * HW * 2 + (N * C * 4 if with_link else 0)
    n_e2e = max(6, -(-args.steps // FPG))
    for i_ in range(len(link_done)):
        link_done[i_] = None
    run(2, host_groups, host=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    frames_e2e = run(n_e2e, host_groups, host=True)
    f1.record()
    barrier()
    sampler.stop()
    e2e_ms = f0.elapsed_time(f1)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / frames_e2e

    # ---- sharded link == sequential link (N > 1: every rank checks its shard against the sequential run of the whole clip)
    shard_parity = None
    if with_link:
        outs = groups[0].replay()
        obj_local = torch.cat([o[2].reshape(-1, N, C) for o in outs], dim=0)
        track = exchange(outs)
        obj_all = vdist.all_gather_kernels(obj_local, world * FPG) if world > 1 else obj_local
        prev_all = torch.cat([obj_all[:1], obj_all[:-1]], dim=0)
        seq = torch.cat([link_fn_for(FPG)(obj_all[r * FPG:(r + 1) * FPG], prev_all[r * FPG:(r + 1) * FPG]) for r in range(world)])
        seq[0] = obj_all[0]
        d = (track - seq[rank * FPG:(rank + 1) * FPG]).abs().max().reshape(1)
        if world > 1:
            dist.all_reduce(d, op=dist.ReduceOp.MAX)
        shard_parity = dict(max_abs_sharded_vs_sequential=float(d.item()), frames_per_rank=FPG, ranks=world,
                            exchange_bytes_per_rank=N * C * 4)

    # ---- single-frame latency (one stream, one frame per graph, no overlap) ----------------------------
    single = vknet.KernelIterLoop(heads[:S - 1] + [heads[-1]])
    single.capture(x1.to(dev).bfloat16(), pf1.to(dev), m1.to(dev).bfloat16())
    for _ in range(5):
        single.replay()
    torch.cuda.synchronize()
    before = _lib.launch_count()
    single.forward(x1.to(dev).bfloat16(), pf1.to(dev), m1.to(dev).bfloat16())
    single_launches = _lib.launch_count() - before
    lat = []
    for i in range(60):
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        single.replay()
        l1.record()
        torch.cuda.synchronize()
        lat.append(l0.elapsed_time(l1))
    lat.sort()
    latency_ms = lat[len(lat) // 2]
    launches_per_step = launches_per_frame_call / BF

    # ---- roofline of the dominant kernel: live per-kernel device times --------------------------------
    acc = {}
    reps = 10
    # profiled at the batch size of the throughput mode (BF frames per call): the operating point `value` is quoted on
    xb_, pfb_, mb_ = frame_batch()
    xs, pfs, ms_ = xb_.to(dev), pfb_.to(dev), mb_.to(dev)
    for _ in range(3):
        loop(xs, pfs, ms_)
    torch.cuda.synchronize()
    for _ in range(reps):
        with _lib.profile() as p:
            loop(xs, pfs, ms_)
        for name, t_ms in p.records:
            a = acc.setdefault(name, [0.0, 0])
            a[0] += t_ms
            a[1] += 1
    per_kernel = {k: dict(total_ms_per_step=v[0] / reps, launches_per_step=v[1] // reps,
                          avg_us=1e3 * v[0] / v[1]) for k, v in acc.items()}
    P = BF * N
    Fh, ncls = CFG1['ffn'], CFG1['ncls']
    Npad = (N + 15) // 16 * 16
    # Algorithmic bytes per call (BF frames, S stages) of each kernel family -- what the math has to move, not what a
    # particular tiling moves.  Row operators: every weight once (bf16) + fp32 rows in/out.
    w_stage = (2041856 + 257 * ncls) * 2
    rows_stage = 4 * (P * C * (2 + 6 + 6 + 5 + 5 + 3 + 2 + 8 + 9 + 3 + 2) + P * Fh * 2 + P * ncls) + BF * 3 * Npad * C * 2
    # frame batches hand 1-bit hard masks between the stages of the loop (vkn_iter_forward): inner-stage mask traffic is
    # N x ceil(HW/128) x 16 bytes per frame instead of N x HW x 2
    ntile = (HW + 127) // 128
    loop_bits = P >= 400 and Npad <= 192 and ntile * BF >= 296 and os.environ.get('VKN_LOOP_BITS', '1') != '0'
    m_logits, m_bits = N * HW * 2, N * ntile * 16
    m_in = [m_logits] + [m_bits if loop_bits else m_logits] * (S - 1)          # mask bytes read by pooling, per stage
    m_out = [m_bits if loop_bits else m_logits] * (S - 1) + [m_logits]         # mask bytes written by the mask conv
    fam_bytes = {
        'pool': BF * sum(C * HW * 2 + mi + N * C * 4 for mi in m_in),
        'maskgemm': BF * sum(C * HW * 2 + mo + 3 * Npad * C * 2 for mo in m_out),
        'linear': S * (w_stage + rows_stage),
        'attention': S * (P * 3 * C * 4 + P * C * 4),
        'pool_reduce': S * (P * C * 4),
    }

    def family(name):
        for key, fam_name in (('pool_reduce', 'pool_reduce'), ('pool', 'pool'), ('maskgemm', 'maskgemm'), ('chain', 'linear'),
                              ('rowgemm', 'linear'), ('linear', 'linear'), ('rowop', 'rowop'), ('attention', 'attention')):
            if key in name:
                return fam_name
        return name
    fam = {}
    for k, v in per_kernel.items():
        f_ = fam.setdefault(family(k), dict(total_ms_per_step=0.0, launches_per_step=0))
        f_['total_ms_per_step'] += v['total_ms_per_step']
        f_['launches_per_step'] += v['launches_per_step']
    peak, peak_tf, peak_src = measured_peaks()
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')       # dram bytes per launch from ncu --set full
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for k, f_ in fam.items():
        if k in fam_bytes:
            f_['algorithmic_bytes_per_launch'] = fam_bytes[k] / max(1, f_['launches_per_step'])
            f_['achieved_gbs'] = fam_bytes[k] / (f_['total_ms_per_step'] * 1e-3) / 1e9
            f_['frac'] = f_['achieved_gbs'] / peak
            f_['ncu_dram_bytes_per_launch'] = traffic.get(k)
    # the row operators are tensor-pipe work: algorithmic FLOPs of every nn.Linear of a stage (SURVEY.md 8d row term without
    # the attention products); the kernels execute 3x that (three exact bf16 planes per fp32 operand)
    rows_flops = S * P * (28 * C * C + 4 * C * Fh + 2 * C * ncls)
    if 'linear' in fam:
        f_ = fam['linear']
        f_['algorithmic_flops_per_launch'] = rows_flops / max(1, f_['launches_per_step'])
        f_['achieved_tflops'] = rows_flops / (f_['total_ms_per_step'] * 1e-3) / 1e12
        f_['frac_tensor'] = f_['achieved_tflops'] / peak_tf
        f_['executed_mma_flops_factor'] = 3
    dom = max(fam, key=lambda k: fam[k]['total_ms_per_step'])
    d_ = fam[dom]
    note = ('times are CUDA-event brackets on the launch stream (vkn_profile_begin/end), one call of %d frame(s), eager launches; ' % BF +
            'the family with the largest share of the call is reported, all families under "families" (their *_per_step fields '
            'are per call of %d frames)' % BF)
    if dom == 'linear':
        kname = ('vkn_chain_tc_kernel' if any('chain' in k for k in per_kernel) else 'vkn_rowgemm_tc_kernel') + \
                ' (row operators, %d launches/call)' % d_['launches_per_step']
        roof = dict(bound='tensor', kernel=kname, achieved=d_['achieved_tflops'], peak=peak_tf, unit='TFLOP/s',
                    frac=d_['frac_tensor'], traffic=traffic.get('linear'), peak_source=peak_src + ', dense bf16 burst',
                    algorithmic_flops_per_launch=d_['algorithmic_flops_per_launch'],
                    avg_launch_us=1e3 * d_['total_ms_per_step'] / max(1, d_['launches_per_step']), profiled_batch=BF,
                    note=note + '; algorithmic FLOPs (fp32 Linear math) -- the kernel issues 3x as many bf16 MMA FLOPs (exact '
                                '3-plane split of the fp32 rows), so tensor-pipe occupancy is about 3x this fraction',
                    families=fam)
    else:
        roof = dict(bound='hbm', kernel={'pool': 'vkn_pool_tc_kernel', 'maskgemm': 'vkn_maskgemm_tc_persist_kernel'}.get(dom, dom),
                    achieved=d_.get('achieved_gbs'), peak=peak, unit='GB/s', frac=d_.get('frac'), traffic=traffic.get(dom),
                    peak_source=peak_src, algorithmic_bytes_per_launch=d_.get('algorithmic_bytes_per_launch'),
                    avg_launch_us=1e3 * d_['total_ms_per_step'] / max(1, d_['launches_per_step']),
                    profiled_batch=BF, note=note, families=fam)
    # whole-step figure: module-boundary algorithmic bytes per frame: 60.4 MB bf16
    # (module-boundary figure: a stage reads x and its masks and writes its masks; the bit-mask hand-off moves fewer bytes)
    step_bytes = S * ((C * HW + 2 * N * HW) * 2 + (2041856 + 257 * ncls) * 2)
    step_gbs = step_bytes / (ms * 1e-3) / 1e9
    step_roof = dict(bound='hbm', kernel='whole step: the S-stage loop of one frame (every kernel)',
                     algorithmic_bytes_per_frame=step_bytes, achieved=step_gbs, peak=peak, unit='GB/s', frac=step_gbs / peak,
                     peak_source=peak_src)
    roof['step'] = step_roof
    lat_gbs = step_bytes / (latency_ms * 1e-3) / 1e9

    if rank == 0:
        cb = dict(value=None, note='skipped (VKN_BENCH_QUICK)') if quick else cpu_arm(steps=8, warmup=2, budget_s=20.0)[0]
        parity = None if quick else parity_check(torch, heads, groups[0], dev)
        value = world / (ms * 1e-3)
        line = dict(metric=METRIC, value=value, unit='frames/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                    data='synthetic', value_per_gpu=value / world,
                    config=dict(workload=WORKLOAD + ('; + cfg3 link: per graph launch the ranks exchange their shard-boundary '
                                                     'kernels (one all-gather of world x [N,C]) and run the previous_type=ffn '
                                                     'link block on their %d frames, on a side stream' % FPG if with_link else ''),
                                storage='bf16 x / masks / weights, fp32 arithmetic (exact 3-plane bf16 products, fp32 accumulation)',
                                mode='CUDA-graph replay of vkn_iter_forward (S stages); %d frames per graph launch = %d concurrent '
                                     'branches x batch %d' % (FPG, NS, BF),
                                value_is='whole job: frames/s summed over the %d GPU(s); value_per_gpu = value / n_gpus' % world,
                                single_stream_ms_per_frame=latency_ms,
                                l2='inputs rotate over %d groups of %d frames inside the timed region (%.0f MB > 126 MB L2); '
                                   'weights stay hot' % (G, FPG, G * FPG * set_bytes / 1e6),
                                engine='tcgen05+TMA' if _lib.lib() and heads[0].engine != _lib.ENGINE_SIMT else 'simt',
                                parallelism='frame-shard x%d' % world),
                    timing=dict(timed_graph_launches=n_launch, timed_frames_per_rank=frames_done, timed_ms=total_ms,
                                launch_ms_median=per_launch[len(per_launch) // 2], launch_ms_p95=per_launch[int(0.95 * (len(per_launch) - 1))],
                                launch_ms_min=per_launch[0], launch_ms_max=per_launch[-1],
                                note='CUDA events on the launch stream around every graph launch; ms_per_step = timed_ms / '
                                     'timed_frames_per_rank (max over ranks)'),
                    e2e=dict(value=world / (e2e_ms * 1e-3), unit='frames/s', h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=d2h, ms_per_step=e2e_ms,
                             h2d_gbs_per_rank=h2d / (e2e_ms * 1e-3) / 1e9, d2h_gbs_per_rank=d2h / (e2e_ms * 1e-3) / 1e9,
                             timed_frames_per_rank=frames_e2e),
                    latency=dict(ms_per_frame=latency_ms, frames_per_s=1e3 / latency_ms, launches=int(single_launches),
                                 ms_p95=lat[int(0.95 * (len(lat) - 1))], roofline_frac=lat_gbs / peak,
                                 note='ONE frame per call (the online VPS operating point, knet/video/kernel_iter_head.py:435-468): '
                                      'CUDA-graph replay of the S-stage loop on one stream, median of 60'),
                    gpu_launches=int(round(launches_per_step * frames_done)),
                    clocks=sampler.summary(), roofline=roof, step_roofline=step_roof, cpu_baseline=cb, parity=parity,
                    kernels=per_kernel)
        if shard_parity is not None:
            line['shard_parity'] = shard_parity
        if cfg3_single is not None:
            line['cfg3_single_gpu'] = cfg3_single
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=7560)
    ap.add_argument('--warmup', type=int, default=1134)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg1', choices=sorted(CONFIGS),
                    help='workload shape (BASELINE.json configs); cfg1 is the one the metric is quoted on')
    args = ap.parse_args()
    CFG1.update(CONFIGS[args.config])
    global WORKLOAD
    WORKLOAD = CONFIGS[args.config]['workload']
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
