"""Index-range sharding of a block array over the ranks of one job.

The decode is a pure map over independent blocks, so multi-GPU operation needs no data-path
collective: rank g owns a contiguous range of blocks (TILED) or of block rows (LINEAR, so that
its output is a contiguous span of image rows), decodes it locally, and only the 1-bit
"all blocks valid" result is reduced across ranks.

The one optional exchange is the final-image gather (`gather_image`): when some rank needs the whole
decoded image, the slabs -- contiguous byte ranges of it -- are all-gathered (NCCL over NVLink /
NVSwitch on GPUs).  It is not part of the decode and is timed separately (`bench.py --gather`).
"""


def tiled_shard(n_blocks, rank, world, align=1):
    """[lo, hi) of rank over a flat block array; boundaries are multiples of `align` blocks
    (2 for 8-byte formats keeps every shard 16-byte aligned)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    lo = (n_blocks * rank // world) // align * align
    hi = n_blocks if rank == world - 1 else (n_blocks * (rank + 1) // world) // align * align
    return lo, hi


def linear_shard(height_in_blocks, rank, world):
    """[row_lo, row_hi) block rows of rank for a row-major image."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return height_in_blocks * rank // world, height_in_blocks * (rank + 1) // world


def linear_shard_view(width, height, width_in_blocks, height_in_blocks, block_bytes, pixel_bytes, rank, world):
    """Byte offsets and sizes of rank's slab: dict(row_lo, rows, height, in_offset, in_bytes,
    out_offset, out_bytes) -- the same view arithmetic as the row-sharded CPU baseline driver
    and the host shim's slab pipeline (detex_b200/csrc/api.cu)."""
    r0, r1 = linear_shard(height_in_blocks, rank, world)
    h = max(0, min(height - r0 * 4, (r1 - r0) * 4))
    return dict(row_lo=r0, rows=r1 - r0, height=h,
                in_offset=r0 * width_in_blocks * block_bytes, in_bytes=(r1 - r0) * width_in_blocks * block_bytes,
                out_offset=r0 * 4 * width * pixel_bytes, out_bytes=h * width * pixel_bytes)


def all_valid(local_ok, group=None):
    """AND-reduce the per-rank validity flag (the only cross-rank datum of the decode)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(local_ok)
    t = torch.tensor([1 if local_ok else 0], dtype=torch.int32)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def gather_image(local, out=None, group=None):
    """Optional final-image gather: every rank contributes its decoded slab `local`
    (1-D uint8 tensor; slab r is the r-th contiguous byte range of the image, as linear_shard_view /
    tiled_shard lay it out) and every rank receives the whole image.  Returns the image (`out` if
    given: 1-D uint8, total size).  Equal slabs are gathered straight into the image buffer
    (all_gather_into_tensor, no staging); a ragged last slab goes through one padded staging buffer."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if out is None:
            return local
        out[: local.numel()].copy_(local)
        return out
    world = dist.get_world_size(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=local.device)
    mine = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    sizes = [int(x) for x in sizes.tolist()]
    total = sum(sizes)
    if out is None:
        out = torch.empty(total, dtype=local.dtype, device=local.device)
    if out.numel() < total:
        raise ValueError("gather_image: output holds %d bytes, the slabs %d" % (out.numel(), total))
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out[:total], local.contiguous(), group=group)
        return out
    pad = max(sizes)
    staged = torch.empty(world * pad, dtype=local.dtype, device=local.device)
    padded = torch.zeros(pad, dtype=local.dtype, device=local.device)
    padded[: local.numel()].copy_(local)
    dist.all_gather_into_tensor(staged, padded, group=group)
    off = 0
    for r, n in enumerate(sizes):
        out[off: off + n].copy_(staged[r * pad: r * pad + n])
        off += n
    return out


class PeerImage:
    """The gather fused into the decode ("peer-mapped destination pointers"): one image
    buffer per rank in symmetric memory (torch.distributed._symmetric_memory: every rank's buffer is
    mapped into every other rank's address space over NVLink / NVSwitch).  A rank that passes
    slab_pointer(dst_rank, offset) as the decode kernels' output pointer makes them store its slab
    directly into dst_rank's image -- no intermediate copy, no collective; a barrier before the reader
    touches the image is all the synchronisation there is.  GPU only, one process per GPU."""

    def __init__(self, total_bytes, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.total = int(total_bytes)
        self.local = symm_mem.empty(self.total, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.local, self.group)
        self._views = {}

    def view(self, rank):
        """rank's image as a tensor in this rank's address space"""
        import torch
        if rank not in self._views:
            self._views[rank] = self.handle.get_buffer(rank, (self.total,), torch.uint8)
        return self._views[rank]

    def slab_pointer(self, rank, offset):
        return self.view(rank).data_ptr() + int(offset)
