"""Multi-process likelihood over NCCL inside the library (gpv_loglik_z_dist) against the one-GPU value.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dist_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import gpvecchia_b200 as G
from gpvecchia_b200 import harness as H, shard

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, m = 300_000, 30
locs = H.make_locs(n, 2, stream=5)
tau, z = H.make_nuggets(n, stream=5), H.make_data(n, stream=5)
cuts = shard.uniform_cuts(n, world)
a, b = int(cuts[rank]), int(cuts[rank + 1])
revNN = H.ordered_nn_gpu(locs, m, a, b, device=lr)
revCond = np.zeros(revNN.shape, dtype=np.int32); revCond[revNN == 0] = np.iinfo(np.int32).min; revCond[:, -1] = 1
obs = np.ones(n, dtype=np.int32)
cp = [1.0, H.default_range(n, 2), 0.8]
h = G.UHandle(locs, revNN, revCond, obs=obs, row_begin=a, row_end=b, device=lr)
uid = torch.from_numpy(G.UHandle.dist_unique_id().copy()).cuda() if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
dist.broadcast(uid, src=0)
h.dist_init(uid.cpu().numpy(), rank, world)
got = h.loglik_z_dist("matern", cp, tau[a:b], tau[a:b], z[a:b], cuts, cuts)
again = h.loglik_z_dist("matern", cp, None, None, None, None, None)
ok = True
if rank == 0:
    full = H.ordered_nn_gpu(locs, m, device=lr)
    fc = np.zeros(full.shape, dtype=np.int32); fc[full == 0] = np.iinfo(np.int32).min; fc[:, -1] = 1
    with G.UHandle(locs, full, fc, obs=obs, device=lr) as h1:
        want = h1.loglik_z("matern", cp, tau, tau, z)
    for k in ("loglik", "quadform_num", "logdet_num", "quadform_denom", "logdet_denom"):
        e1, e2 = abs(got[k] - want[k]) / abs(want[k]), abs(again[k] - want[k]) / abs(want[k])
        ok &= e1 < 1e-11 and e2 < 1e-11
        print(f"{k:15s} dist {got[k]:.12e} one-GPU {want[k]:.12e} rel {e1:.1e} / resident {e2:.1e}")
    print("dist_check", "OK" if ok else "FAILED", "world", world)
h.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
