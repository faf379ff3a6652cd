"""Host-side logic of the data-parallel fine-tuning path on CPU: the flat gradient buffer and its all-reduce
(world_size 2, gloo), and the synthetic teacher targets.  No kernels run here."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    train = importlib.import_module('vln_imagine_b200.train')
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.LayerNorm(3), torch.nn.Linear(3, 1, bias=False))
    net[1].bias.requires_grad = False                      # frozen parameters take no segment
    flat = train.FlatGradients(net)
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    local = flat.buffer.clone()
    for p, off in zip(flat.params, flat.offsets):          # autograd accumulated in place, into the views
        assert p.grad.data_ptr() == flat.buffer[off:].data_ptr()
    flat.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(flat.buffer, sum(gathered) / world, atol=1e-6)
    # second iteration: zero() keeps the views attached, gradients accumulate again from zero
    flat.zero()
    net(x).sum().backward()
    assert torch.allclose(flat.buffer, local, atol=1e-6)
    if rank == 0:
        out.put((flat.numel, len(flat.params)))
    dist.destroy_process_group()


def test_flat_gradients_all_reduce_gloo_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    numel, n = out.get(timeout=5)
    assert n == 4                                  # 2 Linear weights + 1 bias + LN weight (LN bias frozen)
    assert numel == 16 + 4 + 4 + 4                 # segments padded to 4 elements: 15->16, 3->4, 3->4, 3->4


def test_teacher_targets():
    train = importlib.import_module('vln_imagine_b200.train')
    masks = torch.tensor([[1, 1, 1, 1, 0], [1, 1, 0, 0, 0], [1, 1, 1, 1, 1]], dtype=torch.bool)
    visited = torch.tensor([[0, 1, 0, 0, 0], [0, 1, 0, 0, 0], [0, 1, 1, 1, 1]], dtype=torch.bool)
    assert train.teacher_targets(masks, visited).tolist() == [3, 0, 0]
