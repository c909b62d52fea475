"""Replica data-parallelism.  Replicas never interact, so the
path shards with NO data-path collective: rank g owns the contiguous block of
global replica ids [g*R_local, (g+1)*R_local) and seeds are keyed by the GLOBAL
id, so 1/2/4/8-GPU runs give identical per-replica results.  The only exchange
is one all-gather of the per-replica episode-return record at the end of a
rollout (NCCL on GPUs; gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist

RETURN_FIELDS = 6       # OrderNum, RejectNum, TotallyWaitTime, SumOrderValue, DispatchNum, TotallyDispatchCost


class ReplicaShard:
    def __init__(self, replicas_per_rank, rank=None, world_size=None):
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
        self.local = int(replicas_per_rank)
        self.first_replica = self.rank * self.local
        self.total = self.world_size * self.local

    def global_ids(self):
        return range(self.first_replica, self.first_replica + self.local)

    def all_gather_returns(self, stats):
        """stats: int64 [R_local, >=6] tensor (device for NCCL, CPU for gloo).
        Returns int64 [R_total, 6] ordered by global replica id, on every rank."""
        rec = stats[:, :RETURN_FIELDS].contiguous()
        if self.world_size == 1 or not dist.is_initialized():
            return rec
        out = torch.empty((self.total, RETURN_FIELDS), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec)
        return out
