"""CTU-row band sharding of the candidate batch across the GPUs of one box.

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
