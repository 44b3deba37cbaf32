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


# Cost of a row of the closed-form set kernel relative to a row whose neighbours are L2-resident, against
# the footprint of the arrays the neighbours are gathered from (row index x bytes per location) over the
# L2 size.  Measured on B200 (ranks of the N = 2 / 4 / 8 runs and one GPU at n = 8e6): late
# rows of a large problem gather from all earlier locations, which stop fitting the 126 MB L2.
_LOCALITY_X = (0.0, 0.10, 0.29, 0.67, 1.43)
_LOCALITY_W = (1.0, 1.0, 1.02, 1.137, 1.177)


def locality_weight(rows, bytes_per_loc, l2_bytes=126e6, penalty_scale=1.0):
    x = np.asarray(rows, dtype=np.float64) * float(bytes_per_loc) / float(l2_bytes)
    return 1.0 + penalty_scale * (np.interp(x, _LOCALITY_X, _LOCALITY_W) - 1.0)


def weighted_cuts(weights, world):
    """Cut points (world+1) of contiguous row ranges with equal sums of `weights`."""
    w = np.cumsum(np.asarray(weights, dtype=np.float64))
    n = w.size
    total = w[-1] if n else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(w, total * r / world, side="left") + 1) if total > 0 else 0)
    cuts.append(n)
    cuts = np.minimum.accumulate(np.asarray(cuts[::-1], dtype=np.int64))[::-1]
    return np.maximum.accumulate(cuts)


def locality_cuts(nrows, world, d, penalty_scale=1.0):
    """Contiguous row ranges of an ordered-conditioning problem with equal expected kernel time: every
    rank sees the average gather cost instead of the last rank the worst one (8 (d + 1) bytes are gathered
    per neighbour: coordinates and nugget)."""
    if world <= 1:
        return np.array([0, nrows], dtype=np.int64)
    return weighted_cuts(locality_weight(np.arange(nrows), 8 * (d + 1), penalty_scale=penalty_scale), world)


def allreduce_loglik(partial, group=None):
    """Sum (quadform.num, logdet.num, nfail) over ranks with torch.distributed (NCCL on GPUs,
    gloo in CPU tests).  partial: length-3 float64 torch tensor on the rank's device."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial
