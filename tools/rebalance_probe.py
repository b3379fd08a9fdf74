"""Developer probe (torchrun): a heat-bath-grown GenCI space (tools/hci_grow.py's loop, row-sharded) built with the
uniform row partition (PYCI_B200_NO_REBALANCE=1) and with the nnz-balanced one; per-rank stored entries, SpMV time
(max over ranks) and time to E0 of both.

    torchrun --nproc-per-node N tools/rebalance_probe.py [K=24] [npair=6] [target_ndet=1500000] [eps0=0.4]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyci_b200 import cabi  # noqa: E402
from pyci_b200.distributed import exchange_unique_id  # noqa: E402
from pyci_b200.synthetic import spin_orbital_integrals, synthetic_integrals  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 6
TARGET = int(sys.argv[3]) if len(sys.argv) > 3 else 1500000
EPS0 = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", rank=rank, world_size=world)
ctx = cabi.Context(local)
ctx.init_comm(rank, world, exchange_unique_id(cabi.nccl_unique_id, rank, world))
_, one, two = synthetic_integrals(K, 1234)
h2, g2 = spin_orbital_integrals(one, two)
ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
hf = np.array([[((1 << NP) - 1) | (((1 << NP) - 1) << K)]], dtype=np.uint64)
wfn = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * NP, 0, hf)
c, eps = np.ones(1), EPS0
while wfn.ndet < TARGET and eps > 1e-8:
    new = wfn.add_hci(ham, c, eps)
    if len(new) == 0:
        eps *= 0.5
        continue
    op = cabi.Op(ctx, ham, wfn)
    es, cs, st = op.solve(n=1, tol=1e-6)
    c = cs[0]
    op.close()
    if rank == 0:
        sys.stderr.write("[probe] eps %.3g -> %d determinants, E0 %.10f\n" % (eps, wfn.ndet, es[0]))
    eps *= 0.5 if wfn.ndet * 30 < TARGET else 0.8


def gmax(v):
    t = torch.tensor([float(v)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


out = {"ndet": wfn.ndet, "world": world}
for tag, env in (("uniform", "1"), ("balanced", None)):
    if env:
        os.environ["PYCI_B200_NO_REBALANCE"] = env
    else:
        os.environ.pop("PYCI_B200_NO_REBALANCE", None)
    for rep in range(2):
        wfn.reindex()
        op = cabi.Op(ctx, ham, wfn)
        bt = op.build_times()
        if rep == 0:
            op.close()
    parts = [None] * world
    dist.all_gather_object(parts, (int(op.row_begin), int(op.row_count), int(op.stored_nnz)))
    dist.barrier()
    ms = float(np.mean(op.time_spmv(3, 10, 0)))
    es, cs, st = op.solve(n=1, tol=1e-9)
    tot = sum(p[2] for p in parts)
    out[tag] = {"rows": [p[1] for p in parts], "stored_nnz": [p[2] for p in parts],
                "worst_over_mean": max(p[2] for p in parts) * world / tot, "build_s": gmax(bt["total"]),
                "spmv_ms_max_over_ranks": gmax(ms), "E0": float(es[0]), "matvecs": st["matvecs"],
                "solve_s": gmax(st["seconds"])}
    op.close()
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
