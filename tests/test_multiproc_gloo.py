"""world_size-2 (and 3) `gloo` test of the N > 1 host logic on CPU: contiguous row shards
(gpvecchia_b200/shard.py), per-shard likelihood partial sums in the kernel's per-row form, one
all-reduce of three doubles, observation terms contributed by the shard that starts at row 0.
The per-shard numbers come from the CPU oracle here (no GPU in this container); on the GPU box the
same shard/all-reduce code runs under torchrun with NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from gpvecchia_b200 import harness as H
from gpvecchia_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(cond_yz):
    n, m = 500, 8
    locs = H.make_locs(n, 2, stream=70)
    z = H.make_data(n, stream=70)
    tau = H.make_nuggets(n, stream=70)
    if cond_yz == "zy":
        locs2, NN, Cond, obs = H.layout_zy(locs, m, n)
        va = H.make_vecchia_approx(locs2, NN, Cond, obs, "zy", U_sparsity=O.U_sparsity)
    else:
        NN = H.ordered_nn_kdtree(locs, m)
        va = H.make_vecchia_approx(locs, NN, H.layout_yz(NN, cond_yz), np.ones(n, dtype=bool), cond_yz,
                                   U_sparsity=O.U_sparsity)
    return va, z, tau, [1.1, 0.12, 1.5]


def _worker(rank, world, port, cond_yz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    va, z, tau, cp = _problem(cond_yz)
    prep = va["U_prep"]
    Uo = O.createU(va, cp, tau)
    n0 = (prep["revNNarray"] != 0).sum(axis=1)
    cuts = shard.row_cuts(n0, world)
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    L = Uo["U_entries"]["Lentries"]
    n = int(va["obs"].sum())
    zord = z[va["ord_z"] - 1]
    skip = n if cond_yz == "zy" else 0
    qd, ld = O.loglik_numerator_rows(L[a:b], prep["revNNarray"], prep["revCond"], va["obs"], zord,
                                     Uo["nuggets_ord"], row_begin=a, row_end=b, skip_rows=skip,
                                     include_obs_terms=(a == 0))
    part = torch.tensor([qd, ld, 0.0], dtype=torch.float64)
    shard.allreduce_loglik(part)
    if rank == 0:
        qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
        q.put((part.tolist(), qr, lr, cuts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cond_yz", [(2, "z"), (2, "zy"), (3, "y")])
def test_sharded_numerator_allreduce_gloo(world, cond_yz):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cond_yz, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, qr, lr, cuts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert cuts[0] == 0 and len(cuts) == world + 1
    assert abs(got[0] - qr) <= 1e-10 * abs(qr)
    assert abs(got[1] - lr) <= 1e-10 * abs(lr)


def test_per_row_form_equals_matrix_form():
    # the kernel's fused per-row numerator vs crossprod(U[!latent,], z) / diag(U)
    for cyz in ("y", "z", "zy"):
        va, z, tau, cp = _problem(cyz)
        Uo = O.createU(va, cp, tau)
        prep = va["U_prep"]
        n = int(va["obs"].sum())
        qd, ld = O.loglik_numerator_rows(Uo["U_entries"]["Lentries"], prep["revNNarray"], prep["revCond"],
                                         va["obs"], z[va["ord_z"] - 1], Uo["nuggets_ord"],
                                         skip_rows=n if cyz == "zy" else 0)
        qr, lr, _ = O.loglik_numerator_from_U(z, Uo)
        assert abs(qd - qr) <= 1e-11 * abs(qr) and abs(ld - lr) <= 1e-11 * abs(lr), cyz
