"""Frame-batch sharding across ranks: frames are independent
(lib/simple_modulator_cc_impl.cc:72-76 loops frame by frame with no carried state), so
rank r of G processes the contiguous slice [bounds(r), bounds(r+1)) with its own handles and
there is NO collective on the data path.  torch.distributed is used only to gather results
on rank 0 when a caller wants them in one place, and for timing barriers in bench.py."""
import numpy as np


def shard_bounds(n_frames, world_size, rank):
    """Contiguous split, remainder spread over the first ranks: [lo, hi) of `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError('rank MUST lie in [0, world_size)')
    base, rem = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def process_sharded(work, frames, out_size=None, gather=True):
    """Run `work(local_frames) -> local_out` on this rank's shard of `frames` ([n][size] array).
    With gather=True rank 0 returns the full [n][out_size] result in frame order (others None);
    uses the default process group when torch.distributed is initialised, else runs unsharded."""
    try:
        import torch.distributed as dist
        active = dist.is_available() and dist.is_initialized()
    except ImportError:
        active = False
    frames = np.asarray(frames)
    if not active:
        return work(frames)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(frames.shape[0], world, rank)
    local = np.ascontiguousarray(work(frames[lo:hi]))
    if not gather:
        return local
    out_size = local.shape[1] if out_size is None else out_size
    sizes = [shard_bounds(frames.shape[0], world, r) for r in range(world)]
    # gather needs equal-sized tensors: pad every shard to the largest one, trim on rank 0
    most = max(b - a for a, b in sizes)
    t_local = torch.zeros(most * out_size * 2, dtype=torch.float32)
    t_local[:local.size * 2] = torch.from_numpy(local.view(np.float32).reshape(-1))
    if rank == 0:
        parts = [torch.empty_like(t_local) for _ in sizes]
        dist.gather(t_local, parts, dst=0)
        return np.concatenate([p.numpy().view(np.complex64).reshape(most, out_size)[:b - a]
                               for p, (a, b) in zip(parts, sizes)], axis=0)
    dist.gather(t_local, None, dst=0)
    return None
