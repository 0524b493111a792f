"""Data-parallel plumbing for the render-and-compare path: the batch of (image, object)
samples is independent, so ranks take contiguous slices and the only collectives are a SUM all-reduce of a small
loss / pose-gradient vector (mirrors core/utils/my_comm.py:43-62 of the reference) and a MAX over ranks for timing.
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """contiguous [lo, hi) slice of `total` samples owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pose_grad_checksum(grad_R, grad_t):
    """12-float summary of a rank's pose gradients: column sums of dL/dR (9) and dL/dt (3)."""
    return torch.cat([grad_R.reshape(-1, 9).sum(0), grad_t.reshape(-1, 3).sum(0)])


def allreduce_sum(vec):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def allreduce_min_scalar(x):
    """batch-global minimum across ranks (the `_ren_norms.min()` coupling of renderer_dibr.py:284)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x, op=dist.ReduceOp.MIN)
    return x


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
