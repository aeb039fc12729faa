"""CPU, world_size 2, gloo: the host logic of frame sharding + the one all-gather the link step needs
(vknet/dist.py).  The link block itself is the oracle's restatement here (the CUDA kernel needs a GPU);
the property under test is: sharded clip == sequential frame-by-frame run."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import knet_oracle as ko


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_frames, q, boundary=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, 'video-k-net_b200'), os.path.join(root, 'oracle')]
    from vknet import dist as vd
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    N, C = 6, 64
    cfg = ko.default_cfg(num_classes=3, in_channels=C, feedforward_channels=64, previous='p', previous_type='ffn')
    sd = ko.random_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(0)
    obj_all = torch.randn(num_frames, N, C, generator=g)      # every rank can regenerate the whole clip

    def link(cur, prev):   # previous_type='ffn' block (knet/video/kernel_update_head.py:394-415)
        F = cur.shape[0]
        return ko._cross_link(sd, cfg, cur.reshape(F, N, 1, C), prev.reshape(F, N, 1, C), 'attention_previous.',
                              'attention_previous_norm.', 'link_ffn.', 'link_ffn_norm.', F, N).reshape(F, N, C)

    start, end = vd.shard_frames(num_frames, rank, world)
    fn = vd.link_sharded_clip_boundary if boundary else vd.link_sharded_clip
    track = fn(link, obj_all[start:end].clone(), num_frames)
    # sequential reference: frame t links to frame t-1; frame 0 keeps its own kernels
    seq = [obj_all[0]] + [link(obj_all[t:t + 1], obj_all[t - 1:t])[0] for t in range(1, num_frames)]
    want = torch.stack(seq)[start:end]
    q.put((rank, start, end, float((track - want).abs().max()) if end > start else 0.0))
    dist.barrier()
    dist.destroy_process_group()


def _run(num_frames, world=2, boundary=False):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_frames, q, boundary)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_sharded_link_equals_sequential_even_split():
    res = _run(8)
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 8)]
    assert all(r[3] < 1e-5 for r in res), res


def test_sharded_link_equals_sequential_uneven_split():
    res = _run(5)
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[3] < 1e-5 for r in res), res


def test_boundary_exchange_equals_sequential():
    """the minimal exchange (one frame per rank: the shard-boundary kernels) gives the same tracking kernels as the
    sequential run -- even / uneven splits and a clip shorter than the world size (empty shards forward nothing)"""
    for frames, world in ((8, 2), (5, 2), (2, 3)):
        res = _run(frames, world, boundary=True)
        assert all(r[3] < 1e-5 for r in res), (frames, world, res)


def test_shard_partition_properties():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'video-k-net_b200'))
    from vknet.dist import shard_frames, shard_sizes
    for F in (0, 1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_frames(F, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert sum(shard_sizes(F, w)) == F and max(shard_sizes(F, w)) - min(shard_sizes(F, w)) <= 1
