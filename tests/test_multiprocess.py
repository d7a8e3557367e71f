"""N > 1 path on CPU: two gloo ranks decode disjoint shares of an archive collection (emulated kernels), nothing is
exchanged on the data path; a checksum all-gather only verifies that every archive was decoded exactly once."""
import hashlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _collection():
    import _cases as K
    from conftest import read_golden
    arcs = [K.genome(200 + i, 20_000 + 3000 * i, level=3, records=1 + i % 2) for i in range(5)]
    arcs += [read_golden("phix.naf"), read_golden("masked.naf")]
    return arcs


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _harness import emul_library
        from nafcodec_b200.batch import decode_collection
        arcs = _collection()
        mine, res = decode_collection(arcs, rank, world, _library=emul_library())
        digest = torch.zeros(len(arcs), 4, dtype=torch.int64)
        for i, r in zip(mine, res):
            h = hashlib.sha256((r.sequence or b"") + (r.ids or b"")).digest()
            digest[i] = torch.tensor([int.from_bytes(h[k * 7:(k + 1) * 7], "little") for k in range(4)])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)          # verification only: the decode itself exchanged nothing
        if rank == 0:
            q.put((sum(gathered).tolist(), [sorted(mine)]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    from nafcodec_b200.batch import partition
    sizes = [9, 1, 7, 3, 3, 8, 2, 2]
    for w in (1, 2, 3, 8, 11):
        parts = partition(sizes, w)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(l for l in loads) <= max(sizes)


@pytest.mark.timeout(600)
def test_two_ranks_decode_disjoint_shares():
    import _oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, _ = q.get(timeout=500)
    for p in procs:
        p.join(timeout=100)
        assert p.exitcode == 0
    arcs = _collection()
    for i, a in enumerate(arcs):
        d = O.decode(a)
        ids_blob = b"".join(d.id(k) + b"\0" for k in range(d.n) if d.id_present[k])
        h = hashlib.sha256(d.sequence + ids_blob).digest()
        want = [int.from_bytes(h[k * 7:(k + 1) * 7], "little") for k in range(4)]
        assert total[i] == want, i
