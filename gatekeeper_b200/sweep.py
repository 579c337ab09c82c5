"""Sharded audit sweep: every rank reviews its own contiguous range of the cluster's objects and the violation
bitmaps are all-gathered once; per-constraint totals are all-reduced.

Objects are independent units, so there is no data-path collective besides that final exchange.  On GPUs the
kernel writes straight into torch-owned device buffers on torch's current stream and NCCL gathers them; the
same function runs under `gloo` with host tensors (tests use the CPU test backend there).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous object range [lo, hi) of a rank (GPU g gets objects [g*N/G, (g+1)*N/G))."""
    return n_total * rank // world, n_total * (rank + 1) // world


class ShardedSweep:
    """Holds a resident batch (this rank's shard) and the torch buffers for the exchange."""

    def __init__(self, drv, resident_batch, n_local: int, n_constraints: int, device, world: int = 1):
        import torch
        self.torch = torch
        self.drv, self.rb, self.n, self.C = drv, resident_batch, n_local, n_constraints
        self.words = max(1, (n_constraints + 31) // 32)
        self.world = world
        self.device = device
        self.on_gpu = device.type == "cuda"
        self.viol = torch.zeros((n_local, self.words), dtype=torch.int32, device=device)
        self.err = torch.zeros((n_local, self.words), dtype=torch.int32, device=device)
        self.tot = torch.zeros((2, max(1, n_constraints)), dtype=torch.int64, device=device)
        self.gathered = torch.zeros((world * n_local, self.words), dtype=torch.int32, device=device) if world > 1 else None

    def evaluate(self, enforcement_point: str, stream=None):
        """One pass of the hot path over the shard: kernel -> (viol, err, totals) in this rank's buffers."""
        torch = self.torch
        if self.on_gpu:
            st = stream if stream is not None else torch.cuda.current_stream()
            self.rb.eval_device(enforcement_point, self.viol.data_ptr(), self.err.data_ptr(), self.tot[0].data_ptr(), self.tot[1].data_ptr(),
                                st.cuda_stream)
        else:
            r = self.rb.eval(enforcement_point)
            self.viol.copy_(torch.from_numpy(r.viol_bits.view("int32")))
            self.err.copy_(torch.from_numpy(r.err_bits.view("int32")))
            self.tot[0, :self.C] = torch.tensor(r.totals, dtype=torch.int64)
            self.tot[1, :self.C] = torch.tensor(r.err_totals, dtype=torch.int64)

    def exchange(self):
        """The one collective of the path: all-gather of the bitmap shards + all-reduce of the totals."""
        if self.world == 1:
            return self.viol, self.tot
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.gathered, self.viol)
        dist.all_reduce(self.tot)
        return self.gathered, self.tot

    def step(self, enforcement_point: str):
        self.evaluate(enforcement_point)
        return self.exchange()
