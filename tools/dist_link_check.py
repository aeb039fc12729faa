"""torchrun worker: frame-sharded clip with the REAL CUDA link block over NCCL vs the sequential frame-by-frame
run on one GPU (knet/video/kernel_update_head.py:394-415 + vknet/dist.py).  Every rank prints max|diff|.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_link_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), os.path.join(ROOT, 'oracle')]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import knet_oracle as ko  # noqa: E402  (weights + the sequential CPU reference only)
import vknet  # noqa: E402
from vknet import _lib  # noqa: E402
from vknet import dist as vd  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    N, C, frames = 100, 256, 4 * world + 1          # uneven split on purpose
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=2048, previous='p', previous_type='ffn')
    sd = ko.random_state_dict(cfg, seed=3)
    head = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
    head.load_state_dict(sd, strict=True)
    head = head.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    obj_all = torch.randn(frames, N, C, generator=g)
    w, links, wd = head.packed_weights(dev)

    def link_fn(cur, prev):
        shape = head._shape(cur.shape[0], N, 8, 8, _lib.VKN_F32, wd)
        ws, wsb = head._ws.get(shape, dev)
        return head._link(shape, links['track'], cur.contiguous(), prev.contiguous(), None, ws, wsb)

    start, end = vd.shard_frames(frames, rank, world)
    track = vd.link_sharded_clip(link_fn, obj_all[start:end].to(dev), frames)
    # the minimal exchange (one frame per rank: the shard-boundary kernels) must give the same tracking kernels
    track_b = vd.link_sharded_clip_boundary(link_fn, obj_all[start:end].to(dev), frames)
    same = bool(torch.equal(track, track_b)) if end > start else True
    # sequential oracle: frame t links to frame t-1, frame 0 keeps its own kernels
    want = [obj_all[0]]
    for t in range(1, frames):
        want.append(ko._cross_link(sd, cfg, obj_all[t:t + 1].reshape(1, N, 1, C), obj_all[t - 1:t].reshape(1, N, 1, C),
                                   'attention_previous.', 'attention_previous_norm.', 'link_ffn.', 'link_ffn_norm.',
                                   1, N).reshape(N, C))
    want = torch.stack(want)[start:end]
    err = (track.cpu() - want).abs().max().item() if end > start else 0.0
    print('rank %d frames [%d,%d) max|diff| %.3e  boundary-exchange == full all-gather: %s' % (rank, start, end, err, same), flush=True)
    ok = torch.tensor([1.0 if (err < 1e-3 and same) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
