"""Host logic of the row-sharded (N > 1) path, run as world_size-2 `gloo` processes on CPU: the rank
environment, the row partition rule shared with the C library, the communicator-id exchange, and the
rank-order concatenation of per-shard CSR exports (checked against the oracle's whole operator)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from pyci_b200.distributed import concat_csr, row_partition, shard_of_row

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition_covers_rows_once():
    for nrow, ncol, R in [(10, 10, 1), (10, 10, 3), (7, 20, 4), (20, 7, 8), (1, 1, 2), (0, 5, 2), (3312400, 3312400, 8)]:
        parts = row_partition(nrow, ncol, R)
        assert len(parts) == R
        assert sum(c for _, c in parts) == nrow
        pos = 0
        for lo, cnt in parts:
            assert cnt >= 0 and (cnt == 0 or lo == pos)
            pos += cnt
        npad = max(1, -(-max(nrow, ncol) // R))
        assert all(cnt <= npad for _, cnt in parts)
        for row in (0, nrow // 2, nrow - 1):
            if row >= 0 and nrow:
                r = shard_of_row(row, nrow, ncol, R)
                assert parts[r][0] <= row < parts[r][0] + parts[r][1]


def test_concat_csr_of_oracle_row_blocks_equals_whole_operator(small):
    from oracle import oracle as O
    dets = small["h6_sto_3g.fullci42.dets"]
    from conftest import datafile
    _, one, two = O.read_fcidump(datafile("h6_sto_3g"))
    whole = O.sparse_op(O.FULLCI, 6, 4, 2, dets, (one, two), symmetric=False)
    ndet = dets.shape[0]
    shards = []
    for lo, cnt in row_partition(ndet, ndet, 3):
        ip, ix, dv = O.sparse_op(O.FULLCI, 6, 4, 2, dets, (one, two), nrow=lo + cnt, symmetric=False)
        a, b = ip[lo], ip[lo + cnt]
        shards.append((ip[lo:lo + cnt + 1] - a, ix[a:b], dv[a:b]))
    got = concat_csr(shards)
    for g, w in zip(got, whole):
        assert np.array_equal(g, w)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from pyci_b200.distributed import env_ranks, exchange_unique_id, row_partition
    rank, world, local = env_ranks()
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    made = []
    def make_id():
        made.append(1)
        return bytes(range(128))
    uid = exchange_unique_id(make_id, rank, world)
    assert uid == bytes(range(128)), uid
    assert len(made) == (1 if rank == 0 else 0)
    # every rank derives the same partition and owns a disjoint block
    parts = row_partition(1001, 1001, world)
    mine = [None] * world
    dist.all_gather_object(mine, parts[rank])
    assert mine == parts
    # a bad id is rejected on every rank
    try:
        exchange_unique_id(lambda: b"short", rank, world)
    except ValueError:
        pass
    else:
        raise AssertionError("short id accepted")
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_unique_id_exchange_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
