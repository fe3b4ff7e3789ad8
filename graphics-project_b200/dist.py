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


class PendingGather:
    """Handle of an asynchronous gather: ``wait()`` makes the current stream wait for the collectives and
    returns (counts [total], payload [total, max_k, ...])."""

    def __init__(self, works, result):
        self.works, self.result = works, result

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        return self.result


def gather_results(counts, payload, total, group=None, async_op=False, slot=0):
    """All-gather per-frame ``counts`` [b] and padded ``payload`` [b, max_k, ...] from every rank.

    With ``async_op`` the collectives are only enqueued (they start once the current stream reaches this
    point and run beside whatever is enqueued next) and a ``PendingGather`` is returned; the inputs must
    stay untouched until its ``wait()``.  ``slot`` selects one of several cached output buffer sets, so a
    caller that pipelines steps can keep the previous step's result alive.

    Ranks may own different numbers of frames (block sharding of ``total``).  Equal shards (the usual
    case) go through one ``all_gather_into_tensor`` per array into cached output buffers; ragged shards
    are padded to the largest shard for the collective and trimmed afterwards.  Returns (counts [total],
    payload [total, max_k, ...]) in global frame order on every rank (the buffers are reused by the next
    call with the same shapes)."""
    world = dist.get_world_size(group)
    if world == 1:
        return PendingGather([], (counts, payload)) if async_op else (counts, payload)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    bmax = max(sizes)
    if min(sizes) == bmax and counts.shape[0] == bmax:
        key = (counts.device, counts.dtype, payload.dtype, tuple(payload.shape), world, slot)
        bufs = _gather_cache.get(key)
        if bufs is None:
            bufs = (counts.new_empty((world * bmax,)), payload.new_empty((world * bmax,) + tuple(payload.shape[1:])))
            _gather_cache[key] = bufs
        w0 = dist.all_gather_into_tensor(bufs[0], counts.contiguous(), group=group, async_op=async_op)
        w1 = dist.all_gather_into_tensor(bufs[1], payload.contiguous(), group=group, async_op=async_op)
        return PendingGather([w0, w1], bufs) if async_op else bufs
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
    return PendingGather([], (counts_out, payload_out)) if async_op else (counts_out, payload_out)


def canny_hough_sharded(engine, frames_local, total, lo, hi, rho, theta, threshold, blur=True, max_lines=1024,
                        group=None):
    """Run the Canny + HoughLines chain on this rank's frames and gather (counts, lines) of the whole batch."""
    engine.front_canny(frames_local, lo, hi, blur=blur)
    lines, _, counts, _ = engine.hough_lines(None, rho, theta, threshold, max_lines=max_lines, want_votes=False)
    return gather_results(counts, lines, total, group)


def gather_vp(pcounts, fit, hist, total, group=None, scene_hist=False):
    """Gather the `lines` pipeline's per-frame results: ``pcounts`` [b] with ``fit``
    [b, 3] = (phi, theta, loss), and the Gaussian-sphere histograms ``hist`` [b, n_el, n_az].  With
    ``scene_hist`` the histograms are summed over every frame of the batch instead (one ``all_reduce`` of
    [n_el, n_az]: frames of one scene share their vanishing points).  Returns (pcounts [total], fit [total, 3],
    hist [total, n_el, n_az] or [n_el, n_az]) on every rank."""
    all_counts, all_fit = gather_results(pcounts, fit, total, group, slot="vp_fit")
    if scene_hist:
        scene = hist.sum(0)
        if dist.get_world_size(group) > 1:
            dist.all_reduce(scene, group=group)
        return all_counts, all_fit, scene
    _, all_hist = gather_results(pcounts, hist, total, group, slot="vp_hist")
    return all_counts, all_fit, all_hist


def lines_vp_sharded(engine, frames_local, total, iterations=1000, refinement_iterations=500, max_segs=2048,
                     group=None, scene_hist=False):
    """BASELINE config 5 on this rank's frames (gray -> LSD -> polar lines -> sphere histogram -> fit), then the
    gather of (pcounts, fit, histograms) of the whole batch."""
    from . import vp
    fit, hist, _, pcounts, _, _ = vp.get_camera_angles_batch(frames_local, iterations, refinement_iterations,
                                                             max_segs=max_segs, engine=engine)
    return gather_vp(pcounts, fit, hist, total, group, scene_hist)
