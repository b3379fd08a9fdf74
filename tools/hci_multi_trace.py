"""Developer probe (torchrun): add_hci / compute_enpt2 of bench.py's selected-CI cases row-sharded, with
PYCI_B200_HCI_TRACE=1 phase timings on stderr.  Usage: torchrun --nproc-per-node N tools/hci_multi_trace.py [big]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyci_b200 import cabi  # noqa: E402
from pyci_b200.distributed import exchange_unique_id  # noqa: E402
from pyci_b200.synthetic import synthetic_integrals  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", rank=rank, world_size=world)
ctx = cabi.Context(local)
ctx.init_comm(rank, world, exchange_unique_id(cabi.nccl_unique_id, rank, world))
big = len(sys.argv) > 1 and sys.argv[1] == "big"
n, stride = (18, 3) if big else (16, 33)
_, one, two = synthetic_integrals(n, 1234)
full = cabi.Wfn(ctx, cabi.FULLCI, n, 4, 4)
dets = np.ascontiguousarray(full.download_dets()[::stride])
full.close()
c = np.random.default_rng(1).standard_normal(len(dets))
c /= np.linalg.norm(c)
ham = cabi.Ham(ctx, n, 0.0, one, two)
for rep in range(3):
    if rank == 0:
        sys.stderr.write("---- rep %d (%d determinants)\n" % (rep, len(dets)))
    wfn = cabi.Wfn(ctx, cabi.FULLCI, n, 4, 4, dets)
    pt, nt = wfn.compute_enpt2(ham, c, -10.0, 2.0e-4)
    t1 = wfn.ext_seconds()
    new = wfn.add_hci(ham, c, 2.0e-4)
    t2 = wfn.ext_seconds()
    if rank == 0:
        sys.stderr.write("enpt2 %.2f ms (%d external), add_hci %.2f ms (%d added)\n" % (1e3 * t1, nt, 1e3 * t2, len(new)))
    wfn.close()
dist.barrier()
dist.destroy_process_group()
