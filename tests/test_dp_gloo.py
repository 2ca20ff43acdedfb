"""Host-side data-parallel logic on CPU: world_size 2, gloo backend (no GPU needed)."""
import os

import torch
import torch.multiprocessing as mp
import torch.nn as nn


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from fusiondepth_b200.training import FlatParams, reduce_gradients
    torch.manual_seed(0)                                   # same weights on every rank
    models = {"a": nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8)), "b": nn.Linear(5, 7)}
    models["a"][0].weight.data = models["a"][0].weight.data.contiguous(memory_format=torch.channels_last)
    flat = FlatParams(models)
    # views alias the flat buffers, conv weight keeps its channels-last storage
    w = models["a"][0].weight
    assert w.data_ptr() == flat.data.data_ptr() and w.is_contiguous(memory_format=torch.channels_last)
    assert all(p.grad.data_ptr() >= flat.grad.data_ptr() for p in flat.params)
    assert all(p.data_ptr() % 256 == flat.data.data_ptr() % 256 for p in flat.params)   # 256 B aligned
    # rank-dependent "micro-batch": autograd accumulates straight into the flat gradient buffer
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 3, 6, 6, generator=g)
    y = models["a"](x).mean() + models["b"](torch.randn(2, 5, generator=g)).sum()
    y.backward()
    local = flat.grad.clone()
    assert float(local.abs().sum()) > 0
    reduce_gradients(flat, world)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    torch.distributed.all_gather(gathered, local)
    assert torch.allclose(flat.grad, sum(gathered), rtol=1e-6, atol=1e-7)
    # every rank ends with identical gradients => identical Adam updates
    ref = flat.grad.clone()
    torch.distributed.broadcast(ref, 0)
    assert torch.equal(ref, flat.grad)
    out[rank] = 1
    torch.distributed.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29571, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]


def test_flat_params_zero_grad_and_padding():
    from fusiondepth_b200.training import FlatParams
    m = {"m": nn.Sequential(nn.Conv2d(4, 6, 3, bias=True), nn.Conv2d(6, 2, 1))}
    n_real = sum(p.numel() for p in m["m"].parameters())
    flat = FlatParams(m)
    assert flat.numel >= n_real and flat.numel % FlatParams.ALIGN == 0
    for p in flat.params:
        p.grad.fill_(1.0)
    assert float(flat.grad.sum()) == n_real          # padding slots stay zero
    flat.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in flat.params)
