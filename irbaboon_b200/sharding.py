"""Stream sharding across the GPUs of one box.

Streams never interact (fp/convolution.cpp:160-215 is a per-channel loop), so rank g of G owns the contiguous
stream range [g*S/G, (g+1)*S/G) with its FDL rings; shared IR spectra are replicated.  There is NO data-path
collective: torch.distributed is used for the launch barrier, the max-over-ranks time and the final host-side
gather of output blocks only.  Works with the gloo backend on CPU (tests) and nccl on GPUs (bench.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def stream_range(rank, world, n_streams):
    """[begin, end) of the streams rank owns; ranges are contiguous, disjoint, cover [0, n_streams)."""
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def aligned_stream_range(rank, world, n_streams, tile):
    """Same, with every boundary on a multiple of `tile` (the kernel tile of irb_engine_tile_channels) so that a
    tile never straddles two ranks' IR bindings."""
    tiles = (n_streams + tile - 1) // tile
    b, e = stream_range(rank, world, tiles)
    return min(b * tile, n_streams), min(e * tile, n_streams)


def shard_streams(x, rank, world, axis=-2):
    """Slice of a [..., streams, B] array owned by rank."""
    b, e = stream_range(rank, world, x.shape[axis])
    idx = [slice(None)] * x.ndim
    idx[axis] = slice(b, e)
    return x[tuple(idx)]


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (the time every multi-GPU number is reported with)."""
    if _world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_streams(local, n_streams, device="cpu"):
    """Final host-side gather: every rank passes its [local_streams, B] float32 block (numpy); rank 0 receives the
    [n_streams, B] array in global stream order, other ranks None."""
    world = _world()
    local = np.ascontiguousarray(local, np.float32)
    if world == 1:
        return local
    rank = dist.get_rank()
    sizes = [stream_range(r, world, n_streams) for r in range(world)]
    cap = max(e - b for b, e in sizes)
    buf = torch.zeros((cap,) + local.shape[1:], dtype=torch.float32, device=device)
    buf[: local.shape[0]] = torch.from_numpy(local).to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([out[r][: e - b].cpu().numpy() for r, (b, e) in enumerate(sizes)], axis=0)
