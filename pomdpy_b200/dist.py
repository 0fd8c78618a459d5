"""Root-parallel POMCP across GPUs: independent trees, one all-reduce(sum) of the root edge
statistics per move.  The operands are integers (visit counts and 2^-24 fixed-point return sums), so the merged
result does not depend on the reduction order and every rank computes the same greedy action."""
import numpy as np


def world():
    """(rank, world_size) of the initialised default process group, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_budget(n_sims, rank, world_size):
    """Per-rank share of a move's simulation budget and the offset of its random-stream ids."""
    per_rank = -(-int(n_sims) // int(world_size))
    return per_rank, rank * per_rank


def merge_host(n, qfix):
    """Sum (n, qfix) over the ranks with host tensors (gloo, or any backend that takes CPU tensors)."""
    import torch
    import torch.distributed as dist
    A = len(n)
    buf = np.zeros(64, dtype=np.int64)
    buf[:A], buf[32:32 + A] = n, qfix
    t = torch.from_numpy(buf)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return buf[:A].copy(), buf[32:32 + A].copy()


def merge_device(planner, n_actions, stream=None):
    """Sum the root statistics over the ranks on the device: the planner writes its 64 int64 into a CUDA buffer
    on `stream`, NCCL all-reduces it over NVLink, one small copy brings the result to the host."""
    import torch
    import torch.distributed as dist
    buf = torch.empty(64, dtype=torch.int64, device="cuda")
    planner.root_stats_device(buf.data_ptr(), stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    host = buf.cpu().numpy()
    return host[:n_actions].copy(), host[32:32 + n_actions].copy()
