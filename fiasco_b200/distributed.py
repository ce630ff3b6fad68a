"""Multi-GPU plumbing for the tile-split mode: independent tiles / frames are
dealt round-robin to the ranks (tile i -> rank i mod world), every rank encodes its own tiles with
no data-path collective, and ONE gather at the end brings the per-tile FIASCO streams (a few
hundred bytes to a few KB each) to rank 0.  torch.distributed is plumbing only: backend "nccl"
on GPUs (uint8 device tensors over NVLink), "gloo" in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard(n_tiles, rank, world):
    """Tile indices owned by `rank` (round robin keeps the load balanced when tiles differ)."""
    return list(range(rank, n_tiles, world))


def gather_streams(local, n_tiles, rank, world, device="cpu"):
    """`local`: {tile index: bytes} of this rank.  Returns on rank 0 the list of all n_tiles byte
    strings in tile order (None elsewhere).  One all_gather of the sizes, one all_gather of the
    padded payloads."""
    if world == 1:
        return [local[i] for i in range(n_tiles)]
    mine = shard(n_tiles, rank, world)
    per_rank = (n_tiles + world - 1) // world
    sizes = torch.zeros(per_rank, dtype=torch.int64, device=device)
    for k, i in enumerate(mine):
        sizes[k] = len(local[i])
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    width = int(max(int(s.max()) for s in all_sizes)) or 1
    payload = torch.zeros((per_rank, width), dtype=torch.uint8, device=device)
    for k, i in enumerate(mine):
        b = np.frombuffer(local[i], dtype=np.uint8)
        payload[k, :len(b)] = torch.from_numpy(b.copy()).to(device)
    all_payload = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(all_payload, payload)
    if rank != 0:
        return None
    out = [None] * n_tiles
    for r in range(world):
        pl = all_payload[r].cpu().numpy()
        sz = all_sizes[r].cpu().numpy()
        for k, i in enumerate(shard(n_tiles, r, world)):
            out[i] = pl[k, :int(sz[k])].tobytes()
    return out
