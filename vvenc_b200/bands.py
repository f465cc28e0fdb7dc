"""Sharding of the path across the GPUs of one box: CTU-row bands for the block-cost sweeps, neighbour pictures for the MCTF search.

Every block's candidates are independent of every other block once the pictures are finished, so a picture is cut
into contiguous bands of CTU rows, one band per rank (one process per GPU); pictures are replicated, never halo-exchanged.
The only exchange is ONE all-gather of the per-row result tables (best cost + vector per block, 16 bytes each) --
latency bound, a few MB at most -- so it goes through NCCL as is (torch.distributed; gloo in the CPU tests).
"""
import numpy as np


def split_ctu_rows(pic_height, ctu_size, world):
    """[(y0, y1)) pel-row ranges, CTU aligned, contiguous, as even as possible; ranks beyond the row count get empty bands"""
    rows = (pic_height + ctu_size - 1) // ctu_size
    per, extra = divmod(rows, world)
    out = []
    r0 = 0
    for r in range(world):
        n = per + (1 if r < extra else 0)
        y0 = min(pic_height, r0 * ctu_size); y1 = min(pic_height, (r0 + n) * ctu_size)
        out.append((y0, y1)); r0 += n
    return out


def band_blocks(block, pic_width, y0, y1):
    """top-left positions of the block x block grid cells whose rows lie inside [y0, y1)"""
    xs, ys = np.meshgrid(np.arange(0, pic_width - block + 1, block), np.arange(y0, y1 - block + 1, block))
    return xs.ravel().astype(np.int32), ys.ravel().astype(np.int32)


def all_gather_tables(local, counts=None):
    """local: 1-D torch tensor (uint8 view of the rank's result table).  Returns the concatenation over ranks in rank order.
    Uses one all_gather; unequal band sizes are padded to the largest band."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    if counts is None:
        dist.all_gather(sizes, n)
        counts = [int(s.item()) for s in sizes]
    mx = max(counts)
    pad = torch.zeros(mx, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def split_refs(num_refs, world):
    """MCTF: the neighbour pictures of one target picture are searched independently of each other (MCTF.cpp:788-796), so they -- not block rows, which
    depend on the row above through the upper-neighbour candidate (:1289-1306) -- are what is dealt out: rank r gets the reference indices r, r + world, ...
    (round robin keeps the near, cheap-to-match and the far pictures evenly mixed)."""
    return [list(range(r, num_refs, world)) for r in range(world)]


def all_gather_motion_fields(local_fields, num_refs, blocks, device=None):
    """device: torch device of the exchange buffers (a CUDA device under NCCL; None = CPU, the gloo tests).
    local_fields: {ref index: int32 array [blocks][4] (x, y, error, rmsme)} of this rank's share (split_refs).  One all-gather of fixed-size slots;
    returns the int32 array [num_refs][blocks][4] every rank needs for the apply stage (vvb_mctf_apply takes the fields of all neighbour pictures)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    per = (num_refs + world - 1) // world
    mine = split_refs(num_refs, world)[dist.get_rank()]
    slot = torch.zeros((per, blocks, 4), dtype=torch.int32)
    for k, ref in enumerate(mine):
        slot[k] = torch.from_numpy(np.ascontiguousarray(local_fields[ref], dtype=np.int32).reshape(blocks, 4))
    if device is not None:
        slot = slot.to(device)
    bufs = [torch.empty_like(slot) for _ in range(world)]
    dist.all_gather(bufs, slot)
    out = np.zeros((num_refs, blocks, 4), dtype=np.int32)
    for r, refs in enumerate(split_refs(num_refs, world)):
        for k, ref in enumerate(refs):
            out[ref] = bufs[r][k].cpu().numpy()
    return out


def band_pyramid_lists(base, levels, pic_width, y0, y1):
    """quad-tree block lists (candidates.pyramid_lists) of the band [y0, y1): the band is tiled on its own -- bands start on CTU rows, so the tiling is the
    picture's -- and the block rows are shifted to picture coordinates.  Returns [(xs, ys)] per level, level 0 = base size."""
    from .candidates import pyramid_lists
    out = []
    for xs, ys in pyramid_lists(base, levels, pic_width, y1 - y0):
        out.append((xs, (ys + y0).astype(np.int32)))
    return out


class BandGather:
    """The one exchange of the sharded sweep: every rank's per-block result table (vvb_best, 16 bytes per block; all levels of the band
    concatenated) into one buffer on every rank.  Band sizes are known on every rank (they follow from split_ctu_rows), so no size exchange is needed:
    slots are padded to the largest band and ONE all_gather_into_tensor moves them.  launch() snapshots the local table on `compute_stream` and issues
    the collective on a side stream so that it overlaps the next picture's kernels; wait() makes `compute_stream` wait for the last collective;
    table(rank) returns the gathered bytes of one band."""

    def __init__(self, bytes_per_rank, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.counts = [int(c) for c in bytes_per_rank]
        self.world = len(self.counts)
        self.slot = max(self.counts)
        self.src = torch.zeros(self.slot, dtype=torch.uint8, device=device)
        self.buf = torch.zeros(self.world * self.slot, dtype=torch.uint8, device=device)
        self.comm = torch.cuda.Stream(device=device) if device.type == 'cuda' else None
        if self.comm is not None:
            self.ev_snap = torch.cuda.Event(); self.ev_done = torch.cuda.Event()
            self.ev_done.record(self.comm)

    def launch(self, pieces, compute_stream=None):
        """pieces: the rank's result tensors (uint8), concatenated in order into its slot"""
        torch, dist = self.torch, self.dist
        if self.comm is None:                       # CPU / gloo (tests): synchronous
            off = 0
            for p in pieces:
                self.src[off:off + p.numel()] = p.reshape(-1); off += p.numel()
            dist.all_gather_into_tensor(self.buf, self.src)
            return
        with torch.cuda.stream(compute_stream):
            compute_stream.wait_event(self.ev_done)          # the previous collective has finished reading the snapshot
            off = 0
            for p in pieces:
                self.src[off:off + p.numel()].copy_(p.reshape(-1), non_blocking=True); off += p.numel()
            self.ev_snap.record(compute_stream)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_snap)
            dist.all_gather_into_tensor(self.buf, self.src)
            self.ev_done.record(self.comm)

    def wait(self, compute_stream=None):
        if self.comm is not None:
            compute_stream.wait_event(self.ev_done)

    def table(self, rank):
        return self.buf[rank * self.slot: rank * self.slot + self.counts[rank]]
