"""The data-parallel fine-tuning leg on real GPUs (BASELINE.json cfg-4): NCCL gradient all-reduce,
CUDA-graph replay of the iteration, and the drop-in module under torch's DistributedDataParallel as the reference wraps it
(VLN-DUET/map_nav_src/r2r/agent_base.py:115-117).  The two-rank tests skip on a box with one GPU."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model_and_episode(rank, train_mode=False, seed=21):
    synth = importlib.import_module('vln_imagine_b200.synth')
    duet = importlib.import_module('vln_imagine_b200.duet')
    config = importlib.import_module('vln_imagine_b200.config')
    model = duet.VLNBert(config.default_duet_args()).cuda()
    net = model.vln_bert
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth.synth_state_dict(shapes, seed=0))
    model.train(train_mode)
    ep = synth.to_torch(synth.duet_episode(synth.TINY, seed + rank))
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in ep.items()}
    G, P = ep['gmap_img_embeds'].shape[1], ep['vp_img_embeds'].shape[1]
    d['gmap_vpids'], d['vp_cand_vpids'] = net.intern_vpids(ep['gmap_vpids'], ep['vp_cand_vpids'], G, P, torch.device('cuda'))
    return model, net, d


def _allreduce_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    train = importlib.import_module('vln_imagine_b200.train')
    model, net, d = _model_and_episode(rank)
    flat = train.FlatGradients(net)
    loss, _, _, _ = train.duet_finetune_iteration(model, d, n_steps=2)
    local = flat.buffer.clone()
    flat.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = sum(gathered) / world
    err = float((flat.buffer - mean).abs().max() / mean.abs().max())
    differ = float((gathered[0] - gathered[1]).abs().max())          # the ranks really saw different episodes
    if rank == 0:
        out.put((err, differ, float(loss), flat.numel))
    dist.destroy_process_group()


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from torch.nn.parallel import DistributedDataParallel as DDP
    train = importlib.import_module('vln_imagine_b200.train')
    model, net, d = _model_and_episode(rank)
    # the reference: self.vln_bert = DDP(self.vln_bert, device_ids=[rank], find_unused_parameters=True)
    ddp = DDP(model, device_ids=[rank], find_unused_parameters=True)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-5)
    before = net.global_encoder.encoder.x_layers[0].visn_inter.dense.weight.detach().clone()
    # one imitation-learning iteration THROUGH the wrapper: every mode call goes through DDP.forward
    loss, _, _, _ = train.duet_finetune_iteration(lambda mode, batch: ddp(mode, batch), d, n_steps=1)
    opt.step()
    w = net.global_encoder.encoder.x_layers[0].visn_inter.dense.weight
    g = w.grad.detach().clone()
    gathered = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    same = float((gathered[0] - gathered[1]).abs().max())            # DDP averaged the gradients: identical on both ranks
    moved = float((w.detach() - before).abs().max())
    if rank == 0:
        out.put((same, moved, float(loss), float(g.abs().max())))
    dist.destroy_process_group()


def _overlap_worker(rank, world, port, out):
    """graph-replayed iteration with the all-reduce in two parts (everything but the text side under the language encoder's
    backward, train.GraphedIteration(tail_fn=...)) against the plain eager iteration + one all-reduce"""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    train = importlib.import_module('vln_imagine_b200.train')
    model, net, d = _model_and_episode(rank)
    flat = train.FlatGradients(net)
    flat.zero()
    train.duet_finetune_iteration(model, d, n_steps=2)
    flat.all_reduce()
    ref = flat.buffer.clone()
    end = flat.prefix_end(net)
    pending = {}

    def grad_fn(ep):
        flat.zero()
        o = train.duet_finetune_iteration(model, ep, n_steps=2, fused_accumulation=True, split_language_backward=True)
        pending['finish'] = o[4]
        return o[0].detach()

    def tail_fn():
        pending.pop('finish')()

    def before_tail():
        pending['work'] = flat.all_reduce_range(end, flat.numel, async_op=True)

    def reduce_rest():
        w2 = flat.all_reduce_range(0, end, async_op=True)
        for w in (pending.pop('work', None), w2):
            if w is not None:
                w.wait()
    holder = {}

    def update_fn():
        holder['grads'] = flat.buffer.clone()
    it = train.GraphedIteration(net, grad_fn, update_fn, d, between=reduce_rest, warmup=1, tail_fn=tail_fn, before_tail=before_tail)
    errs = []
    for _ in range(2):
        it.load(d)
        it.replay()
        torch.cuda.synchronize()
        errs.append(float((holder['grads'] - ref).abs().max() / ref.abs().max()))
    if rank == 0:
        out.put((max(errs), end, flat.numel))
    dist.destroy_process_group()


def _spawn(worker, world=2, timeout=600):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
        assert p.exitcode == 0, 'worker exited with %r' % p.exitcode
    return out.get(timeout=10)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_nccl_all_reduce_averages_the_rank_gradients(lib_built):
    err, differ, loss, numel = _spawn(_allreduce_worker)
    assert differ > 0 and loss == loss
    assert err < 1e-6, err
    assert numel > 100_000_000


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_module_trains_under_distributed_data_parallel(lib_built):
    same, moved, loss, gmax = _spawn(_ddp_worker)
    assert loss == loss and gmax > 0
    assert same == 0.0, 'DDP must leave identical (averaged) gradients on both ranks'
    assert moved > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_overlapped_all_reduce_equals_the_single_one(lib_built):
    err, end, numel = _spawn(_overlap_worker)
    assert 0 < end < numel
    assert err < 1e-4, err            # same gradients, the fp32 additions of the steps in another order


def test_graph_replayed_iteration_equals_the_eager_one(lib_built):
    """train.GraphedIteration.replay() against the same iteration launched eagerly: gradients of every parameter agree (the two
    GASA scalars are accumulated with floating-point atomics, everything else is deterministic), and after finish() an
    inference call sees the UPDATED weights (derived copies and inference graphs are rebuilt)."""
    train = importlib.import_module('vln_imagine_b200.train')
    model, net, d = _model_and_episode(0)                 # eval(): no dropout, deterministic
    flat = train.FlatGradients(net)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, fused=True, capturable=True)

    def grad_fn(ep):
        flat.zero()
        loss, _, _, _ = train.duet_finetune_iteration(model, ep, n_steps=2, fused_accumulation=True)
        return loss.detach()

    holder = {}

    def update_fn():
        holder['grads'] = flat.buffer.clone()               # inside graph 2: the gradients this replay produced
        opt.step()

    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    eager_loss = float(grad_fn(d))
    eager_grads = flat.buffer.clone()
    it = train.GraphedIteration(net, grad_fn, update_fn, d, warmup=1)
    net.load_state_dict(sd0)                                # the warm-up / capture iterations moved the weights: start over
    opt.state.clear() if hasattr(opt.state, 'clear') else None
    it.load(d)
    loss = float(it.replay())
    torch.cuda.synchronize()
    g = holder['grads']
    assert abs(loss - eager_loss) < 1e-5 * abs(eager_loss)
    denom = eager_grads.abs().max()
    assert float((g - eager_grads).abs().max() / denom) < 1e-5
    # validation after training: fresh derived weights
    it.finish()
    model.eval()
    with torch.no_grad():
        a = model('language', {'txt_ids': d['txt_ids'], 'txt_masks': d['txt_masks']})
        net.precision = 'fp32'
        b = model('language', {'txt_ids': d['txt_ids'], 'txt_masks': d['txt_masks']})    # fp32 masters: the ground truth
        net.precision = 'bf16'
    assert float((a - b).abs().max() / b.abs().max()) < 2e-2, 'inference after graph-replayed training used stale weights'
