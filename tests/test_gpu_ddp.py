"""Data-parallel parity ON THE HARDWARE (SURVEY 4 "Distributed", VERDICT r1 missing #5): N ranks of the drop-in CRW under
DistributedDataParallel (NCCL) reproduce the single-GPU result on the concatenated batch - loss and the gradients of the head and
of the encoder.  Reference wrap: code/train.py:260-262 (nn.DataParallel), loss mean at :68.  Needs >= 2 GPUs: skipped otherwise
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`)."""
import argparse
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _args(dev):
    return argparse.Namespace(device=dev, dropout=0.1, featdrop=0.0, temp=0.07, head_depth=0, model_type="scratch", remove_layers=[],
                              dilate_superpixels=False, flip=False, sk_targets=False)


def _inputs(B, T, N):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, T, N * 3, 64, 64, generator=g)
    u12 = torch.rand(T - 1, B, N, N, generator=g)
    u21p = torch.rand(T - 1, B, N, N, generator=g)
    return x, u12, u21p


def _worker(rank, world, port, B, T, N, out):
    import torch.distributed as dist
    from sapienza_video_contrastive_b200 import CRW
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        torch.manual_seed(0)
        crw = CRW(_args(dev)).to(dev).eval()                 # eval(): BatchNorm uses its running statistics on every rank
        ddp = torch.nn.parallel.DistributedDataParallel(crw, device_ids=[rank])
        x, u12, u21p = _inputs(B, T, N)
        b = B // world
        sl = slice(rank * b, (rank + 1) * b)
        q, loss, diags = ddp(x[sl].to(dev), None, None, walk_uniforms=(u12[:, sl].to(dev), u21p[:, sl].to(dev)))
        loss.mean().backward()                                # DDP averages the gradients over the ranks (train.py:68: loss.mean())
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss.detach())
        if rank == 0:
            res = {"loss": torch.stack(losses).mean().cpu(), "ghead": crw.selfsim_fc[0].weight.grad.cpu(),
                   "gconv1": crw.encoder.model.conv1.weight.grad.cpu(), "glast": crw.encoder.model.layer3[-1].conv2.weight.grad.cpu()}
            # the single-GPU run on the concatenated batch, same weights
            torch.manual_seed(0)
            one = CRW(_args(dev)).to(dev).eval()
            one.load_state_dict(crw.state_dict())
            q1, loss1, _ = one(x.to(dev), None, None, walk_uniforms=(u12.to(dev), u21p.to(dev)))
            loss1.mean().backward()
            res.update(loss1=loss1.detach().mean().cpu(), ghead1=one.selfsim_fc[0].weight.grad.cpu(),
                       gconv11=one.encoder.model.conv1.weight.grad.cpu(), glast1=one.encoder.model.layer3[-1].conv2.weight.grad.cpu())
            # ... and the same single GPU fed the ranks' shards one after the other (identical kernel shapes to the ranks'):
            # gradient accumulation of loss_r / world is what DistributedDataParallel must reproduce to rounding
            for p_ in one.parameters():
                p_.grad = None
            for r in range(world):
                s2 = slice(r * b, (r + 1) * b)
                _, l2, _ = one(x[s2].to(dev), None, None, walk_uniforms=(u12[:, s2].to(dev), u21p[:, s2].to(dev)))
                (l2.mean() / world).backward()
            res.update(ghead2=one.selfsim_fc[0].weight.grad.cpu(), gconv12=one.encoder.model.conv1.weight.grad.cpu(),
                       glast2=one.encoder.model.layer3[-1].conv2.weight.grad.cpu())
            torch.save(res, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_ddp_gradients_equal_single_gpu_on_concatenated_batch(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    out = str(tmp_path / "ddp.pt")
    B, T, N = 4, 4, 9
    mp.spawn(_worker, args=(world, _free_port(), B, T, N, out), nprocs=world, join=True)
    r = torch.load(out)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert abs(float(r["loss"]) - float(r["loss1"])) <= 1e-5 * abs(float(r["loss1"]))
    # (1) the data-parallel plumbing: N ranks == the shards accumulated on one GPU, to rounding (same kernel shapes everywhere)
    assert rel(r["ghead"], r["ghead2"]) < 1e-5
    # the encoder's weight gradients come from cuDNN kernels that reduce with atomics (run-to-run noise): measured 3.7e-5 on the
    # last convolution, 1.3e-4 on the first; the head's gradient (this repository's kernels, deterministic) holds 1e-5 above
    assert rel(r["glast"], r["glast2"]) < 2e-4
    assert rel(r["gconv1"], r["gconv12"]) < 1e-3
    # (2) == the concatenated batch in one call.  The loss agrees to 1e-5 above; the gradients pass through cuDNN convolutions
    # whose algorithms differ between batch 2 and batch 4, and at random initialisation the walk's gradient is a small
    # difference of nearly equal terms (all node embeddings are almost parallel), which amplifies that 1e-6 feature noise
    assert rel(r["ghead"], r["ghead1"]) < 2e-2
    assert rel(r["glast"], r["glast1"]) < 2e-2
    assert rel(r["gconv1"], r["gconv11"]) < 5e-2
