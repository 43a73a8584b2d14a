"""CPU: host logic of the self-supervised training step (dsmnet_b200/selfsup.py) against the fixture produced by the
reference's own losses/loss.py (`depthmono-mask` pyramid loss), with the oracle's CPU warp injected in place of the CUDA
op; and the multi-GPU form of the step (DistributedDataParallel gradient averaging) on two `gloo` ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle.ops as O
from conftest import GOLDEN
from dsmnet_b200.selfsup import losses_pyramid1, selfsup_loss_for_batch, ssim_map, train_step
from dsmnet_b200.shard import shard_range


def cpu_warp(im_src, disp, fliplr=False, LeftTop=(0, 0), scale_factor=1):
    """imwrap_BCHW with the oracle's CPU grid_sample; draws `delt` like the reference (imwrap.py:70)"""
    delt = float(1e-4 * (torch.rand(1)[0] + 0.1))
    return O.imwrap(im_src, disp, fliplr, tuple(LeftTop), scale_factor, delt)


def _load():
    z = np.load(os.path.join(GOLDEN, "selfsup_loss.npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def test_pyramid_loss_matches_reference():
    g = _load()
    ne = g["nedge"]; batch = g["batch"]; b1 = torch.flip(batch, dims=[3])
    crop = lambda t: t[:, :, ne:-ne, ne:-ne]
    d = [g["disp%d" % l].clone().requires_grad_() for l in range(7)]
    d1 = [g["disp1_%d" % l].clone().requires_grad_() for l in range(7)]
    torch.manual_seed(g["seed"])
    loss = losses_pyramid1(batch[:, 3:6], crop(batch[:, :3]), d, list(range(7)), (ne, ne), b1[:, :3], crop(b1[:, 3:6]), d1, (ne, ne),
                           g["weight_levels"].tolist(), True, cpu_warp)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-6 * abs(float(g["loss"]))
    for l in range(7):
        assert torch.allclose(d[l].grad, g["g%d" % l], rtol=1e-5, atol=1e-9)
        assert torch.allclose(d1[l].grad, g["g1_%d" % l], rtol=1e-5, atol=1e-9)


def test_ssim_identity_and_range():
    torch.manual_seed(0)
    a = torch.rand(2, 3, 24, 40)
    assert torch.allclose(ssim_map(a, a), torch.ones(2, 1, 24, 40), atol=1e-5)
    assert float(ssim_map(a, torch.rand(2, 3, 24, 40)).max()) <= 1.0 + 1e-5


class _Toy(torch.nn.Module):
    """stand-in for DispNetC with the same call surface: a 7-level positive disparity pyramid from two small convs"""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(6, 8, 3, padding=1); self.c2 = torch.nn.Conv2d(8, 1, 3, padding=1)

    def forward(self, imL, imR):
        x = torch.nn.functional.softplus(self.c2(torch.relu(self.c1(torch.cat([imL, imR], 1)))))
        return list(range(7)), [torch.nn.functional.avg_pool2d(x, 2 ** l) / 2 ** l if l else x for l in range(7)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = _Toy()
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.0)            # lr 0: we only want the reduced gradients
        g = torch.Generator().manual_seed(5)
        batch = torch.rand(4, 6, 64 + 16, 128 + 16, generator=g)
        b, e = shard_range(4, rank, world)
        torch.manual_seed(100)                                      # same delt stream on both ranks
        train_step(ddp, opt, batch[b:e], 8, [1.0, 0.5, 0.3, 0.2, 0.1, 0.05, 0.01], cpu_warp)
        q.put((rank, [p.grad.numpy().copy() for p in model.parameters()]))      # numpy: no shared-memory handles to outlive the worker
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_gloo():
    """DDP over two ranks, each with half the batch: both ranks end up with the SAME gradients, equal to the mean of the two
    per-shard gradients computed in one process (the loss is a per-shard mean, so DDP averages shard gradients)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: [torch.from_numpy(a) for a in gs] for r, gs in (q.get(timeout=300) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    # single-process reference: average of the per-shard gradients
    torch.manual_seed(0)
    model = _Toy()
    g = torch.Generator().manual_seed(5)
    batch = torch.rand(4, 6, 64 + 16, 128 + 16, generator=g)
    grads = []
    for r in range(world):
        b, e = shard_range(4, r, world)
        model.zero_grad()
        torch.manual_seed(100)
        selfsup_loss_for_batch(model, batch[b:e], 8, [1.0, 0.5, 0.3, 0.2, 0.1, 0.05, 0.01], True, cpu_warp).backward()
        grads.append([p.grad.clone() for p in model.parameters()])
    for i, a in enumerate(res[0]):
        assert torch.allclose(a, (grads[0][i] + grads[1][i]) / 2, rtol=1e-5, atol=1e-8)
