"""Frame sharding across the GPUs of one node.

Frames (and frame pairs) are independent units: the reference itself is batch-1 with no cross-frame state in
extraction.  Rank r owns frames r, r + world, r + 2*world, ... (a pair (i-1, i) is matched on the rank that
owns frame i, which keeps the previous frame's descriptors resident).  There is NO collective on the data
path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) only all-gathers four int64 counters per rank
and max-reduces the elapsed time, so that one consistent whole-job number is reported.
"""
import torch
import torch.distributed as dist


def frame_indices(rank, world, n_frames):
    """Global indices of the frames rank `rank` processes."""
    return list(range(rank, n_frames, world))


def owner_of(frame_idx, world):
    return frame_idx % world


def gather_counters(frames, keypoints, matches, elapsed_ms, device=None):
    """All ranks call this; returns (counters [world, 4] int64 on CPU, max elapsed ms over ranks).
    counters columns: frames, keypoints, matches, elapsed_ns."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = device if device is not None else torch.device("cpu")
    mine = torch.tensor([int(frames), int(keypoints), int(matches), int(elapsed_ms * 1e6)], dtype=torch.int64, device=dev)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    if world == 1:
        return mine.cpu()[None], float(t[0])
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return torch.stack(out).cpu(), float(t[0])


def whole_job_fps(counters, max_elapsed_ms):
    """Whole-job frames/s = frames of ALL ranks / slowest rank's time."""
    return float(counters[:, 0].sum()) / (max_elapsed_ms * 1e-3)
