"""Multi-GPU partitioning of the path (one process per GPU).

* tokenize / find / count / fragments: queries are sharded in contiguous blocks (whole files for tokenize), the
  index is replicated, no collective; concatenating rank outputs in rank order reproduces the single-GPU output.
* LOLA count matrices: the DATABASE is sharded by region set — rank r owns sets [r*C, min((r+1)*C, n)), C = ceil(n/W) —
  every rank sees all query sets, computes its column block, and the blocks are all-gathered (NCCL on device inside
  gtgpu_igd_count_sharded; the helpers here are the host-side arithmetic and are exercised on CPU with gloo).
"""
from __future__ import annotations

import numpy as np


def block_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of n items for `rank`: sizes differ by at most one, earlier ranks get the larger blocks."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def db_set_range(n_sets: int, world: int, rank: int) -> tuple[int, int]:
    """Region sets of the LOLA database owned by `rank` (fixed column width C so the all-gather is regular)."""
    cols = -(-n_sets // world) if n_sets else 0
    lo = min(rank * cols, n_sets)
    return lo, min(lo + cols, n_sets)


def shard_files(file_offsets: np.ndarray, world: int, rank: int):
    """(first file, last file, first query, last query) of this rank and its rebased file offsets."""
    n_files = len(file_offsets) - 1
    f0, f1 = block_range(n_files, world, rank)
    q0, q1 = int(file_offsets[f0]), int(file_offsets[f1])
    return f0, f1, q0, q1, (file_offsets[f0:f1 + 1] - file_offsets[f0]).astype(np.uint64)


def merge_tokenized(parts):
    """Concatenate per-rank (token_offsets, ids) in rank order into one (token_offsets, ids)."""
    offs, ids, base = [np.zeros(1, dtype=np.uint64)], [], 0
    for o, i in parts:
        offs.append(np.asarray(o[1:], dtype=np.uint64) + np.uint64(base))
        ids.append(np.asarray(i, dtype=np.uint32))
        base += int(o[-1])
    return np.concatenate(offs), (np.concatenate(ids) if ids else np.zeros(0, dtype=np.uint32))


def assemble_column_blocks(gathered: np.ndarray, n_sets_db: int) -> np.ndarray:
    """gathered[world, n_query_sets, C] (rank r's block, zero-padded to C columns) -> [n_query_sets, n_sets_db]."""
    world, nq, cols = gathered.shape
    return np.ascontiguousarray(gathered.transpose(1, 0, 2).reshape(nq, world * cols)[:, :n_sets_db])


def init_comm_from_torch(ctx) -> None:
    """Create the library's NCCL communicator for `ctx`, shipping rank 0's unique id through torch.distributed."""
    import torch
    import torch.distributed as dist

    from . import ffi
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    device = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.tensor(list(ffi.comm_unique_id()), dtype=torch.uint8, device=device)
    dist.broadcast(uid, src=0)
    ffi.comm_init(ctx, world, rank, bytes(uid.cpu().tolist()))
