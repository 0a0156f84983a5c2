"""Multi-GPU plumbing.  One process per GPU; frames and camera streams are independent, so
the hot path has NO collective: ranks just take disjoint frame / stream ranges.  The single optional
exchange is the box-wide union of voxel keys (BASELINE.json config 4): every rank all-gathers its sorted
canonical keys (variable length) and runs one local sort-unique over the concatenation.

torch.distributed is plumbing only; the gather works on whatever device the keys live on (NCCL for CUDA
tensors, gloo for the CPU tests of this host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [start, stop) of `n_items` independent frames / streams for `rank`."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def stream_owner(stream: int, world: int) -> int:
    """config 4: camera stream c is processed by GPU c mod G."""
    return stream % world


def all_gather_varlen(t: torch.Tensor, group=None) -> list:
    """All-gather 1-D tensors of different lengths: lengths first, then one padded all_gather."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros(cap, dtype=t.dtype, device=t.device)
    padded[:t.numel()] = t.reshape(-1)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return [p[:c] for p, c in zip(parts, counts)]


def all_gather_ragged(flat: torch.Tensor, lengths, group=None) -> list:
    """All-gather a batch of variable-length segments in TWO collectives, however many segments there are: `flat` is the
    concatenation of S segments of `lengths[s]` elements (one segment per rig frame of a step).  Returns, per rank, the list
    of its S segments (views into the gathered buffers)."""
    world = dist.get_world_size(group)
    len_t = torch.as_tensor(list(lengths), dtype=torch.int64, device=flat.device)
    all_len = [torch.empty_like(len_t) for _ in range(world)]
    dist.all_gather(all_len, len_t, group=group)
    all_len = torch.stack(all_len).tolist()                      # the one host read of the exchange
    cap = max(max(sum(l) for l in all_len), 1)
    padded = torch.zeros(cap, dtype=flat.dtype, device=flat.device)
    padded[:flat.numel()] = flat.reshape(-1)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = []
    for p, lens in zip(parts, all_len):
        segs, at = [], 0
        for n in lens:
            segs.append(p[at:at + n])
            at += n
        out.append(segs)
    return out


def box_wide_keys(local_keys: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenation of every rank's canonical voxel keys (not yet deduplicated)."""
    return torch.cat(all_gather_varlen(local_keys, group))


def box_wide_voxels(points: torch.Tensor, voxel_size: float, group=None, flavour=None):
    """Union over all ranks of the voxel sets of `points` (each rank's own cameras): returns the sorted
    lattice points of the union, identical on every rank."""
    from . import ops
    _, keys, layout = ops.voxel_snap_unique(points, voxel_size, flavour=flavour, return_keys=True)
    canon, canon_layout = ops.keys_canonical(keys, layout)
    merged = ops.keys_sort_unique(box_wide_keys(canon, group), 63)
    return ops.keys_decode(merged, canon_layout, voxel_size)
