"""CPU, world_size 2 over gloo: the N>1 path -- static image shards, no data-path collective, ragged gather of results."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from walkgpt_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_path(images, seg_hidden, offsets):
    """Stand-in for the per-rank hot path: any per-prompt function of (its image, its [SEG] row)."""
    n = seg_hidden.shape[0]
    img_of = torch.repeat_interleave(torch.arange(len(offsets) - 1), torch.tensor([offsets[i + 1] - offsets[i] for i in range(len(offsets) - 1)]))
    val = images[img_of].sum(1) * 10 + seg_hidden.sum(1)
    return {"scores": val, "masks": (val[:, None, None] + torch.zeros(n, 4, 4)).to(torch.uint8), "iou": val[:, None] * 0.5}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs = torch.arange(5, dtype=torch.float32)[:, None] + 1
    offs = [0, 2, 2, 5, 8, 9]
    seg = torch.arange(9, dtype=torch.float32)[:, None] * 0.25
    li, ls, lo = parallel.shard_batch(imgs, seg, offs, rank, world)
    out = parallel.gather_results(_fake_path(li, ls, lo), keys=("masks", "scores", "iou"))
    full = _fake_path(imgs, seg, offs)
    ok = all(torch.equal(out[k], full[k]) for k in ("masks", "scores", "iou"))
    q.put((rank, ok, out["scores"].shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and n == 9 for _, ok, n in res)
