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


class PeerGather:
    """All-gather of equal-sized fp32 records over peer memory (csrc/p2p_gather.cu): every rank stores its record straight into
    windows its peers exposed through CUDA IPC. One instance per (record size, process group); the IPC handles are exchanged
    once, here, with a torch.distributed all_gather. Every rank must call `all_gather` in the same order."""

    def __init__(self, numel, device, group=None):
        import ctypes as C
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel = int(numel)
        self.padded = (self.numel * 4 + 15) // 16 * 4                  # floats per rank, record rounded up to 16 bytes
        self.device = torch.device(device)
        h = C.c_void_p()
        capi.check(capi.lib().hrp_p2p_create(self.rank, self.world, self.padded * 4, self.device.index, C.byref(h)))
        self._h = h
        mine = (C.c_uint8 * 64)()
        capi.check(capi.lib().hrp_p2p_handle(self._h, mine))
        backend = dist.get_backend(group)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device if backend == "nccl" else "cpu")
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allh, t, group=group)
        buf = (C.c_uint8 * (self.world * 64))(*allh.cpu().tolist())
        capi.check(capi.lib().hrp_p2p_connect(self._h, buf))
        dist.barrier(group)                                            # every window is mapped before the first store

    def all_gather(self, rec):
        """rec: CUDA fp32 [numel] -> [world, numel] on the current stream."""
        import ctypes as C
        if rec.numel() != self.numel or rec.dtype != torch.float32 or not rec.is_cuda:
            raise ValueError("PeerGather.all_gather: expected a CUDA fp32 record of %d elements" % self.numel)
        src = rec.contiguous()
        if self.padded != self.numel or src.data_ptr() % 16:
            pad = torch.zeros(self.padded, dtype=torch.float32, device=rec.device)
            pad[:self.numel] = src
            src = pad
        out = torch.empty(self.world, self.padded, dtype=torch.float32, device=rec.device)
        st = torch.cuda.current_stream(rec.device).cuda_stream
        capi.check(capi.lib().hrp_p2p_all_gather(self._h, C.c_void_p(src.data_ptr()), self.padded * 4, C.c_void_p(out.data_ptr()), C.c_void_p(st)))
        src.record_stream(torch.cuda.current_stream(rec.device))
        return out[:, :self.numel]

    def check(self):
        """Synchronises; raises if a peer timed out in any gather so far."""
        capi.check(capi.lib().hrp_p2p_status(self._h))

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().hrp_p2p_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


def gather_records(rec, batch, dof, nkpt, group=None, depth_num=0, peer=None):
    """All-gather equal-sized per-rank records and re-assemble global per-field tensors {name: [G*batch, ...]}
    (depth_num > 0: models built with multi_kp, whose record carries every regressed depth as `depths`). peer: a PeerGather
    for this record size -- the gather then runs as P2P stores over NVLink instead of an NCCL all-gather."""
    world = dist.get_world_size(group)
    if peer is not None:                              # stores into the peers' windows over NVLink (PeerGather) instead of NCCL
        out = peer.all_gather(rec)
    else:
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
