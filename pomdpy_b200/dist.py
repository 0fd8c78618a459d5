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


class FusedMerge(object):
    """Exchange buffers for the merge that runs inside the search kernel (include/pomcp_b200.h, pomcp_set_peer_merge):
    one symmetric-memory allocation per rank, every rank's buffer mapped into every process (NVLink peer access), so
    the kernel's tail can push its 64 root words to all peers and sum theirs without a separate collective."""

    def __init__(self, planner, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import native
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("the fused merge serves the <= 8 GPUs of one box")
        words = native.MERGE_BYTES // 8
        self.buf = symm_mem.empty((words,), dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize()
        self.planner, self.group = planner, group

    def activate(self):
        """Collective: call on every rank once all of them hold a FusedMerge (try_fused_merge checks that first)."""
        import torch.distributed as dist
        dist.barrier(self.group)                         # every buffer is zero before anyone's kernel writes a flag
        self.planner.set_peer_merge(self.world, self.rank, [int(p) for p in self.handle.buffer_ptrs])
        return self


def try_fused_merge(planner):
    """FusedMerge if the process group is NCCL on CUDA and symmetric memory can be set up, else None (the caller
    then merges with merge_device).  The decision is collective: either every rank gets one or none does."""
    import torch
    import torch.distributed as dist
    ok, fm = 1, None
    try:
        fm = FusedMerge(planner)
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None
    return fm.activate()
