"""Row sharding across the GPUs of one box: rows of U are independent
(src/U_NZentries.cpp:39-69 has no cross-iteration dependence; the reference uses
schedule(static)), so each rank owns a contiguous row range and no data-path collective is
needed for U.  Only the likelihood partial sums (3 doubles) are combined, by one all-reduce."""
import numpy as np


def row_cuts(n0, world):
    """Cut points (world+1) of contiguous row ranges balancing sum(n0^3) -- the factorisation
    cost -- so the trivial n0 = 1 rows of `zy`/prediction layouts do not skew the split."""
    n0 = np.asarray(n0, dtype=np.float64)
    w = np.cumsum(n0 ** 3)
    total = w[-1] if w.size else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(w, total * r / world, side="left") + 1) if total > 0 else 0)
    cuts.append(n0.size)
    cuts = np.minimum.accumulate(np.asarray(cuts[::-1]))[::-1]   # monotone
    return np.maximum.accumulate(cuts)


def uniform_cuts(nrows, world):
    return np.array([(nrows * r) // world for r in range(world + 1)], dtype=np.int64)


def allreduce_loglik(partial, group=None):
    """Sum (quadform.num, logdet.num, nfail) over ranks with torch.distributed (NCCL on GPUs,
    gloo in CPU tests).  partial: length-3 float64 torch tensor on the rank's device."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial
