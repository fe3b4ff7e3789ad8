"""Frame-level data parallelism: one process per GPU, contiguous block sharding, result gather.

The path shards by frame with no data-path exchange; ``torch.distributed`` (NCCL on
GPUs, gloo in the CPU tests) is used only to gather the variable-length per-frame results:
``counts[B/R]`` then the padded payload ``[B/R, max_k, k]``.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of ``total`` frames owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_gather_cache = {}


def gather_results(counts, payload, total, group=None):
    """All-gather per-frame ``counts`` [b] and padded ``payload`` [b, max_k, ...] from every rank.

    Ranks may own different numbers of frames (block sharding of ``total``).  Equal shards (the usual
    case) go through one ``all_gather_into_tensor`` per array into cached output buffers; ragged shards
    are padded to the largest shard for the collective and trimmed afterwards.  Returns (counts [total],
    payload [total, max_k, ...]) in global frame order on every rank (the buffers are reused by the next
    call with the same shapes)."""
    world = dist.get_world_size(group)
    if world == 1:
        return counts, payload
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    bmax = max(sizes)
    if min(sizes) == bmax and counts.shape[0] == bmax:
        key = (counts.device, counts.dtype, payload.dtype, tuple(payload.shape), world)
        bufs = _gather_cache.get(key)
        if bufs is None:
            bufs = (counts.new_empty((world * bmax,)), payload.new_empty((world * bmax,) + tuple(payload.shape[1:])))
            _gather_cache[key] = bufs
        dist.all_gather_into_tensor(bufs[0], counts.contiguous(), group=group)
        dist.all_gather_into_tensor(bufs[1], payload.contiguous(), group=group)
        return bufs
    pad = bmax - counts.shape[0]
    if pad:
        counts = torch.cat([counts, counts.new_zeros((pad,))])
        payload = torch.cat([payload, payload.new_zeros((pad,) + tuple(payload.shape[1:]))])
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    all_payload = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(all_counts, counts.contiguous(), group=group)
    dist.all_gather(all_payload, payload.contiguous(), group=group)
    counts_out = torch.cat([c[:s] for c, s in zip(all_counts, sizes)])
    payload_out = torch.cat([p[:s] for p, s in zip(all_payload, sizes)])
    return counts_out, payload_out


def canny_hough_sharded(engine, frames_local, total, lo, hi, rho, theta, threshold, blur=True, max_lines=1024,
                        group=None):
    """Run the Canny + HoughLines chain on this rank's frames and gather (counts, lines) of the whole batch."""
    engine.front_canny(frames_local, lo, hi, blur=blur)
    lines, _, counts, _ = engine.hough_lines(None, rho, theta, threshold, max_lines=max_lines, want_votes=False)
    return gather_results(counts, lines, total, group)
