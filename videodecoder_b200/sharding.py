"""GOP sharding across the GPUs of one box.

Decoding shards only at IDR-delimited GOPs (or across independent streams): a P picture needs the
whole previous reconstructed picture, an IDR flushes every reference (reference: NAL::decode_PIC,
h264/NAL.cpp:189; Decoder::clear_DecodedPic, h264/Decoder.cpp:38-63).  GOP g goes to rank g mod N;
every rank owns its frame pool; nothing is exchanged on the data path.  The only collective of the
system gathers the per-frame 16-byte digests on rank 0 (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def assign_gops(n_gops, world_size):
    """GOP indices owned by each rank: g -> g mod N."""
    return [[g for g in range(n_gops) if g % world_size == r] for r in range(world_size)]


def gather_digests(local_digests, local_gops, n_gops, frames_per_gop, rank, world_size, device="cpu"):
    """local_digests: (len(local_gops) * frames_per_gop, 16) uint8 in the order of local_gops.
    Returns on rank 0 the (n_gops * frames_per_gop, 16) array in GOP order, None elsewhere."""
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return np.asarray(local_digests, dtype=np.uint8).reshape(n_gops * frames_per_gop, 16)
    per_rank = (n_gops + world_size - 1) // world_size
    buf = torch.zeros((per_rank * frames_per_gop, 16), dtype=torch.uint8, device=device)
    mine = torch.from_numpy(np.ascontiguousarray(local_digests, dtype=np.uint8).reshape(-1, 16))
    buf[: mine.shape[0]] = mine.to(device)
    gathered = [torch.zeros_like(buf) for _ in range(world_size)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    out = np.zeros((n_gops * frames_per_gop, 16), dtype=np.uint8)
    owners = assign_gops(n_gops, world_size)
    for r in range(world_size):
        g_r = gathered[r].cpu().numpy()
        for k, g in enumerate(owners[r]):
            out[g * frames_per_gop:(g + 1) * frames_per_gop] = g_r[k * frames_per_gop:(k + 1) * frames_per_gop]
    return out
