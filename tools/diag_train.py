"""Loss trajectory of the fine-tuning iteration: eager vs graph replay (diagnostic).  python tools/diag_train.py [mode] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import config, duet, synth, train  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'eager'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device('cuda', 0)
model = duet.VLNBert(config.default_duet_args()).cuda()
net = model.vln_bert
shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
net.load_state_dict(synth.synth_state_dict(shapes, seed=0))
net.config.hidden_dropout_prob = net.config.attention_probs_dropout_prob = 0.0
model.drop_env.p = 0.0
model.train()
ep = synth.to_torch(synth.duet_episode(synth.CFG2, 1234))
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in ep.items()}
G, P = ep['gmap_img_embeds'].shape[1], ep['vp_img_embeds'].shape[1]
d['gmap_vpids'], d['vp_cand_vpids'] = net.intern_vpids(ep['gmap_vpids'], ep['vp_cand_vpids'], G, P, dev)
flat = train.FlatGradients(net)
opt = torch.optim.AdamW(net.parameters(), lr=1e-5, fused=True, capturable=True)
gn = torch.zeros((), device=dev)


def grad_fn(e):
    flat.zero()
    loss, ce, aux, _ = train.duet_finetune_iteration(model, e, n_steps=6)
    return loss.detach()


def update_fn():
    n = torch.nn.utils.clip_grad_norm_(net.parameters(), 40.)
    gn.copy_(n)
    opt.step()


out = []
if mode == 'eager':
    for i in range(iters):
        loss = grad_fn(d)
        update_fn()
        out.append((float(loss), float(gn)))
else:
    for i in range(4):
        loss = grad_fn(d)
        update_fn()
        out.append((float(loss), float(gn)))
    it = train.GraphedIteration(net, grad_fn, update_fn, d, warmup=0)
    out.append(('captured', 0))
    for i in range(iters - 5):
        it.g_grad.replay()
        torch.cuda.synchronize()
        names = [n for n, p in net.named_parameters() if p.requires_grad]
        bad = [n for n, p in zip(names, flat.params) if not torch.isfinite(p.grad).all()]
        if bad:
            print('after graph 1: non-finite grads in', len(bad), bad[:12])
            for n, p in zip(names, flat.params):
                if n in bad[:3]:
                    g = p.grad
                    print('   ', n, tuple(g.shape), 'nan', int(torch.isnan(g).sum()), 'inf', int(torch.isinf(g).sum()),
                          'first bad idx', torch.nonzero(~torch.isfinite(g.reshape(-1)))[:5].view(-1).tolist())
                    if g.dim() == 2:
                        badm = ~torch.isfinite(g)
                        rows_bad = torch.nonzero(badm.any(1)).view(-1).tolist()
                        cols_bad = torch.nonzero(badm.any(0)).view(-1)
                        print('    bad rows', rows_bad[:20], 'n bad cols', int(cols_bad.numel()), 'per-row bad counts',
                              badm.sum(1)[badm.any(1)].tolist()[:20], 'huge finite', int((g.abs() > 1e6).sum()))
        it.g_update.replay()
        loss = it.loss
        torch.cuda.synchronize()
        fin = bool(torch.isfinite(flat.buffer).all())
        out.append((float(loss), float(gn), fin))
        if not fin:
            names = [n for n, p in net.named_parameters() if p.requires_grad]
            bad = [n for n, p in zip(names, flat.params) if not torch.isfinite(p.grad).all()]
            print('non-finite grads in', len(bad), 'of', len(names), bad[:10], '...', bad[-5:])
            wbad = [n for n, p in zip(names, flat.params) if not torch.isfinite(p).all()]
            print('non-finite weights in', len(wbad), wbad[:5])
            break
print(mode, ' '.join(str(o) for o in out))
