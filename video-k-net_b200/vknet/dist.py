"""Frame sharding of a clip across the GPUs of one box, and the ONE exchange it needs.

Inside a stage every op is per frame (the MHA batch dimension is the frame dimension,
knet/det/kernel_update_head.py:204-208; the mask conv loops over frames :252-257), so the F frames
of a clip are block-partitioned over the ranks with no data-path collective.  The only cross-frame
dependency of the shipped `previous_type='ffn'` configs is the tracking-kernel link
(knet/video/kernel_update_head.py:394-415): frame t attends to frame t-1's last-stage kernels
(the un-linked `obj_feat`, knet/video/knet_quansi_dense_embed_fc_joint_train.py:525).  One
all-gather over NCCL (NVLink/NVSwitch) provides them: of the whole per-rank kernels [F_local, N, C]
(`link_sharded_clip`) or -- all a block partition needs -- of each rank's LAST frame only
(`link_sharded_clip_boundary`: world x 102 KB, latency-bound).  Frame 0 of the clip has no predecessor: `is_first` semantics
(knet/video/kernel_iter_head.py:478-479) -> its tracking kernels are its own obj_feat.

One process per GPU (torchrun); works with the gloo backend on CPU tensors for the host-logic tests.
"""
import torch
import torch.distributed as dist


def shard_frames(num_frames, rank, world_size):
    """Contiguous block partition; the first `num_frames % world_size` ranks take one extra frame."""
    base, rem = divmod(num_frames, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(num_frames, world_size):
    return [shard_frames(num_frames, r, world_size)[1] - shard_frames(num_frames, r, world_size)[0]
            for r in range(world_size)]


def all_gather_kernels(obj_local, num_frames, group=None):
    """obj_local [F_local, N, C] on every rank -> obj_all [F, N, C] (same on every rank).
    Uneven shards are padded to the largest shard so a single fixed-size all_gather suffices."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj_local
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_frames, world)
    fmax = max(sizes)
    pad = obj_local
    if obj_local.shape[0] < fmax:
        pad = torch.zeros((fmax,) + tuple(obj_local.shape[1:]), dtype=obj_local.dtype, device=obj_local.device)
        pad[: obj_local.shape[0]] = obj_local
    out = torch.empty((world * fmax,) + tuple(obj_local.shape[1:]), dtype=obj_local.dtype, device=obj_local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(s == fmax for s in sizes):
        return out
    return torch.cat([out[r * fmax: r * fmax + sizes[r]] for r in range(world)], dim=0)


def previous_for_shard(obj_all, start, end):
    """Kernels of frames start-1 .. end-2 (the predecessors of frames start .. end-1).  For the clip's
    first frame the slot holds the frame's own kernels; callers overwrite that output (is_first)."""
    idx = torch.arange(start, end, device=obj_all.device) - 1
    return obj_all[idx.clamp_(min=0)]


def gather_boundary_kernels(obj_local, group=None):
    """The ONE exchange of the sharded clip, minimal form: with a block partition a rank needs exactly one frame it does
    not own -- the LAST frame of its left neighbour (the predecessor of its first frame).  Every rank contributes that one
    [N, C] block (102 KB at N=100, C=256) to a single all-gather over NCCL: world x 102 KB on the wire instead of the whole
    clip's kernels.  Returns [world, N, C]: entry r = last-frame kernels of rank r (ranks with an empty shard contribute
    zeros and forward nothing: callers use `boundary_prev`)."""
    world = dist.get_world_size(group)
    last = obj_local[-1] if obj_local.shape[0] > 0 else torch.zeros(obj_local.shape[1:], dtype=obj_local.dtype,
                                                                     device=obj_local.device)
    out = torch.empty((world * last.shape[0],) + tuple(last.shape[1:]), dtype=obj_local.dtype, device=obj_local.device)
    dist.all_gather_into_tensor(out, last.contiguous(), group=group)
    return out.reshape((world,) + tuple(last.shape))


def boundary_prev(boundary, sizes, rank):
    """kernels of the frame preceding this rank's first frame: the last frame of the nearest non-empty shard to the left
    (None for the clip's first frame)."""
    for r in range(rank - 1, -1, -1):
        if sizes[r] > 0:
            return boundary[r]
    return None


def link_sharded_clip_boundary(link_fn, obj_local, num_frames, rank=None, world_size=None, group=None):
    """Same result as `link_sharded_clip`, exchanging only the shard-boundary frames (`gather_boundary_kernels`): the
    predecessors of a rank's frames are its own frames shifted by one, plus ONE frame from the left neighbour."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    start, end = shard_frames(num_frames, rank, world_size)
    assert obj_local.shape[0] == end - start, 'obj_local does not match this rank\'s shard'
    left = None
    if world_size > 1:
        boundary = gather_boundary_kernels(obj_local, group)
        left = boundary_prev(boundary, shard_sizes(num_frames, world_size), rank)
    if end == start:
        return obj_local
    prev = torch.empty_like(obj_local)
    prev[1:] = obj_local[:-1]
    prev[0] = left if left is not None else obj_local[0]
    track = link_fn(obj_local, prev)
    if left is None:                     # this shard starts the clip
        track = track.clone()
        track[0] = obj_local[0]          # is_first: object_feats_track = object_feats
    return track


def link_sharded_clip(link_fn, obj_local, num_frames, rank=None, world_size=None, group=None):
    """Tracking kernels for this rank's frames.

    link_fn(cur [F_local,N,C], prev [F_local,N,C]) -> [F_local,N,C] is the link block
    (VideoKernelUpdateHead's `previous_type='ffn'` branch -- on the GPU `vkn_link_attend`).
    Equals the sequential frame-by-frame run exactly, because the memory handed from frame to
    frame is the un-linked obj_feat.
    """
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    start, end = shard_frames(num_frames, rank, world_size)
    assert obj_local.shape[0] == end - start, 'obj_local does not match this rank\'s shard'
    obj_all = all_gather_kernels(obj_local, num_frames, group)
    if end == start:
        return obj_local
    prev = previous_for_shard(obj_all, start, end)
    track = link_fn(obj_local, prev)
    if start == 0:
        track = track.clone()
        track[0] = obj_local[0]          # is_first: object_feats_track = object_feats
    return track
