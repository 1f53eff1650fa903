"""N > 1 path on CPU: world_size 2, gloo.  The data path has no collective except the final
all-gather of per-image records (mnc_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mnc_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    total = 6
    s, e = D.shard_range(total, r, w)
    g = torch.Generator().manual_seed(7)
    boxes = torch.rand(total, 600, 4, generator=g)
    masks = torch.rand(total, 600, 1, 21, 21, generator=g)
    scores = torch.rand(total, 600, 21, generator=g)
    valid = (torch.rand(total, 600, generator=g) > 0.5).to(torch.uint8)
    rec = D.pack_records(boxes[s:e], masks[s:e], scores[s:e], valid[s:e])
    allrec = D.all_gather_records(rec)
    c, b2, m2, s2 = D.unpack_records(allrec, e - s)
    ok = (tuple(allrec.shape) == (w, D.record_len(e - s)) and torch.equal(b2, boxes) and torch.equal(m2, masks)
          and torch.equal(s2, scores) and torch.equal(c, valid.sum(1).to(torch.int64)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
