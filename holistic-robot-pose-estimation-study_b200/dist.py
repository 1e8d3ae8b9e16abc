"""Batch sharding across the GPUs of one box and the output gather (the only collective on this path).

Frames are independent (eval-mode BN, per-(frame,keypoint) softmax, per-pose FK), so rank r runs frames
[r*B/G, (r+1)*B/G) through its own handle and the packed output records are all-gathered once per forward
(the reference's equivalent is nn.DataParallel's gather, scripts/test.py:159). Works on any
torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

from . import capi


def shard_range(total, rank, world):
    """Contiguous split of `total` frames; the first total % world ranks take one extra frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def field_widths(dof, nkpt, depth_num=0):
    """Floats per frame of every field of the output record (HRP_F_*); depth_num > 0: the multi_kp variant's HRP_F_DEPTHS."""
    return (dof, 6, 3, 2, 1, nkpt * 3, nkpt * 3, nkpt * 3, nkpt * 2, nkpt * 2, depth_num)


def record_offsets(batch, dof, nkpt, depth_num=0):
    """Mirror of hrp_output_offsets: struct-of-arrays record, every field 16-byte aligned."""
    offs, o = [], 0
    for w in field_widths(dof, nkpt, depth_num):
        offs.append(o)
        o = (o + batch * w + 3) & ~3
    offs.append(o)
    return offs


def gather_records(rec, batch, dof, nkpt, group=None, depth_num=0):
    """All-gather equal-sized per-rank records and re-assemble global per-field tensors {name: [G*batch, ...]}
    (depth_num > 0: models built with multi_kp, whose record carries every regressed depth as `depths`)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    out = out.view(world, rec.numel())
    offs = record_offsets(batch, dof, nkpt, depth_num)
    res = {}
    names = tuple(capi.FIELD_NAMES) + (("depths",) if depth_num > 0 else ())
    for f, (name, w) in enumerate(zip(names, field_widths(dof, nkpt, depth_num))):
        t = out[:, offs[f]:offs[f] + batch * w].reshape(world * batch, w)
        if f in (5, 6, 7):
            t = t.view(world * batch, nkpt, 3)
        elif f in (8, 9):
            t = t.view(world * batch, nkpt, 2)
        res[name] = t
    return res
