"""Multi-GPU sharding of the similarity + selection stage (one process per GPU).

Row i of R^T R needs only CSC column i and the CSR rows it touches, so item
rows are cut into contiguous blocks of equal *work* (sum of rater degrees, not
row count), every rank keeps the whole ratings layout (replicated; the
reference broadcasts its side tables the same way, assist.py:69,73) and the
SpGEMM itself needs no communication.  Two exchanges remain, both all-gathers
over NCCL (NVLink / NVSwitch):

  1. BB flags after pass 1  (replaces the SQL DISTINCT + collect + broadcast of
     assist.py:82-87)  -- n_items bytes;
  2. the neighbour tables after pass 2 (replaces the collectAsMap + broadcast of
     assist.py:121-132) -- n_items * 2k * 20 bytes.

Because the accumulators are order-free integers the gathered result is
bit-identical for any world size.  The exchange helpers are device-agnostic so
they can be exercised with the gloo backend on CPU tensors.
"""
import torch
import torch.distributed as dist


class RowShard(object):
    """Contiguous, work-balanced block of item rows owned by `rank`."""

    def __init__(self, row_work, rank=0, world=1):
        self.rank, self.world = rank, world
        n = int(row_work.numel())
        self.n_items = n
        if world == 1:
            self.bounds = [0, n]
        else:
            cw = torch.cumsum(row_work.double() + 1.0, 0)          # +1: empty rows still cost a visit
            total = float(cw[-1]) if n else 0.0
            targets = torch.tensor([total * r / world for r in range(1, world)], dtype=torch.float64,
                                   device=cw.device)
            cuts = torch.searchsorted(cw, targets).tolist() if n else [0] * (world - 1)
            self.bounds = [0] + [int(c) for c in cuts] + [n]
            for i in range(1, len(self.bounds)):
                self.bounds[i] = max(self.bounds[i], self.bounds[i - 1])
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.max_rows = max(self.bounds[r + 1] - self.bounds[r] for r in range(world))

    def rows(self, device):
        return torch.arange(self.lo, self.hi, dtype=torch.int32, device=device)


def allgather_rows(t, shard, group=None):
    """t: tensor whose dim 0 is the item axis, valid on [shard.lo, shard.hi).
    After the call every rank holds every row.  Blocks are padded to the largest
    block so a single equal-size all-gather moves them."""
    if shard.world == 1:
        return t
    tail = t.shape[1:]
    send = torch.zeros((shard.max_rows,) + tuple(tail), dtype=t.dtype, device=t.device)
    send[: shard.hi - shard.lo] = t[shard.lo:shard.hi]
    recv = torch.empty((shard.world * shard.max_rows,) + tuple(tail), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view((shard.world, shard.max_rows) + tuple(tail))
    for r in range(shard.world):
        lo, hi = shard.bounds[r], shard.bounds[r + 1]
        if r != shard.rank and hi > lo:
            t[lo:hi] = recv[r, : hi - lo]
    return t


def similarity_step(engine, shard, group=None):
    """Pass 1 on the owned rows, all-gather BB flags, pass 2, all-gather the tables."""
    rows = None if shard.world == 1 else shard.rows(engine.device)
    s1 = engine.pass1(rows)
    allgather_rows(engine.row_flags, shard, group)
    s2 = engine.pass2(engine.row_flags, rows)
    if shard.world > 1:
        for t in (engine.row_npairs, engine.row_nkept, engine.tab_len, engine.tab_idx, engine.tab_sim,
                  engine.tab_mutu, engine.tab_n):
            allgather_rows(t, shard, group)
    return engine.tables(dict(pass1=s1, pass2=s2, rows=(shard.lo, shard.hi)))
