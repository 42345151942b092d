"""N > 1 host logic on CPU (gloo, world_size 2, 127.0.0.1): clips are sharded across ranks with no data-path
collective; each rank's walk returns the mean over ITS clips and d(that mean)/d feats; averaging losses and
all-reducing (mean) the parameter gradients reproduces the single-process full-batch result.  The compute here is the
oracle (CPU) - the point is the sharding / reduction convention bench.py and a DDP training loop rely on."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import crw_oracle as O


def _full(B, N, T, Ce, seed):
    g = torch.Generator().manual_seed(seed)
    maps = torch.randn(B * N, Ce, T, 4, 4, generator=g)
    head = torch.randn(128, Ce, generator=g) / Ce ** 0.5
    u12 = torch.rand(T - 1, B, N, N, generator=g)
    u21p = torch.rand(T - 1, B, N, N, generator=g)
    return maps, head, u12, u21p


def _loss_and_grad(maps, head, u12, u21p, B):
    head = head.clone().requires_grad_(True)
    q = O.patch_nodes(maps, head, B)
    loss, *_ = O.walk_loss(q, 0.07, 0.1, u12, u21p)
    loss.sum().backward()
    return loss.detach(), head.grad


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, N, T, Ce = 4, 9, 4, 8
    maps, head, u12, u21p = _full(B, N, T, Ce, 3)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    loss, g = _loss_and_grad(maps.view(B, N, Ce, T, 4, 4)[sl].reshape(per * N, Ce, T, 4, 4), head,
                             u12[:, sl].contiguous(), u21p[:, sl].contiguous(), per)
    buf = torch.cat([g.flatten(), loss])
    dist.all_reduce(buf)
    buf /= world
    ms = torch.tensor([float(rank + 1)])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                     # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        torch.save({"grad": buf[:-1].view_as(g), "loss": buf[-1:], "max": ms}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_sharding_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    got = torch.load(out)
    maps, head, u12, u21p = _full(4, 9, 4, 8, 3)
    loss, g = _loss_and_grad(maps, head, u12, u21p, 4)
    torch.testing.assert_close(got["loss"], loss, rtol=1e-5, atol=0)
    torch.testing.assert_close(got["grad"], g, rtol=1e-4, atol=1e-7)
    assert float(got["max"]) == 2.0
