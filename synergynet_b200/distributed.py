"""Multi-GPU plumbing: faces are independent, so the batch is sharded on dim 0 with one process
per GPU and the only collective is ONE all-gather of the (B/N,3,68) landmarks
(the reference's analogue is nn.DataParallel scatter/gather,
main_train.py:176, which benchmark.py:127 bypasses).  Dense vertices are never gathered."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``total`` faces owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_process_group(backend: str) -> None:
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        dist.init_process_group(backend=backend)


def gather_landmarks(local: torch.Tensor, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """All ranks contribute (b,3,68) (equal b) and receive (world*b,3,68) in rank order."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class OverlappedGather:
    """The same single all-gather per step, issued on a side stream so that it runs under the NEXT step's backbone
    instead of at the tail of its own step (the gather moves 835 KB per rank after ~2 ms of compute; un-overlapped it is
    the whole 1 -> 8 GPU efficiency loss).  ``gather`` returns immediately; ``wait`` makes the current stream (and hence
    a later ``.cpu()`` / synchronize) see every gather issued so far."""

    def __init__(self, device: torch.device, group=None):
        self.device, self.group = device, group
        self.side = torch.cuda.Stream(device)

    def gather(self, local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)                       # landmarks of this step are complete
        local = local.contiguous()
        local.record_stream(self.side)                   # allocated on the compute stream, read on the side stream
        out.record_stream(self.side)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def wait(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.side)
