"""graph 1 (fwd+bwd) and graph 2 (clip+AdamW) checked separately against eager on optimiser-updated weights (diagnostic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import config, duet, synth, train  # noqa: E402

dev = torch.device('cuda', 0)
model = duet.VLNBert(config.default_duet_args()).cuda()
net = model.vln_bert
shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
net.load_state_dict(synth.synth_state_dict(shapes, seed=0))
net.config.hidden_dropout_prob = net.config.attention_probs_dropout_prob = 0.0
model.drop_env.p = 0.0
model.train()
T = int(os.environ.get('T', '2'))
ep = synth.to_torch(synth.duet_episode(synth.CFG2, 1234))
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in ep.items()}
G, P = ep['gmap_img_embeds'].shape[1], ep['vp_img_embeds'].shape[1]
d['gmap_vpids'], d['vp_cand_vpids'] = net.intern_vpids(ep['gmap_vpids'], ep['vp_cand_vpids'], G, P, dev)
flat = train.FlatGradients(net)
opt = torch.optim.AdamW(net.parameters(), lr=1e-5, fused=True, capturable=True)
names = [n for n, p in net.named_parameters() if p.requires_grad]


def grad_fn(e=None):
    flat.zero()
    loss, ce, aux, _ = train.duet_finetune_iteration(model, d, n_steps=T)
    return loss.detach()


def update_fn():
    torch.nn.utils.clip_grad_norm_(net.parameters(), 40.)
    opt.step()


def pvec():
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


for _ in range(3):
    l = grad_fn(); update_fn()
torch.cuda.synchronize()
print('eager losses ok, last', float(l))
it = train.GraphedIteration(net, grad_fn, update_fn, d, warmup=0)
torch.cuda.synchronize()
w0 = pvec().clone()
# eager reference on the current weights (no update)
le = grad_fn(); torch.cuda.synchronize()
ge = flat.buffer.clone()
print('eager  loss', float(le), 'gnorm', float(ge.norm()), 'weights unchanged', bool(torch.equal(w0, pvec())))
net._packs = None
it.g_grad.replay(); torch.cuda.synchronize()
gg = flat.buffer.clone()
print('graph1 loss', float(it.loss), 'gnorm', float(gg.norm()), 'finite', bool(torch.isfinite(gg).all()),
      'max rel diff', float((gg - ge).abs().max() / ge.abs().max()), 'weights unchanged', bool(torch.equal(w0, pvec())))
# graph 2 vs eager update from identical state: copy state
flat.buffer.copy_(ge)
it.g_update.replay(); torch.cuda.synchronize()
w_graph = pvec().clone()
print('graph2 update: max |dw|', float((w_graph - w0).abs().max()), 'finite', bool(torch.isfinite(w_graph).all()))
it.g_grad.replay(); torch.cuda.synchronize()
print('graph1 after graph2: loss', float(it.loss), 'finite grads', bool(torch.isfinite(flat.buffer).all()))
net._packs = None
le2 = grad_fn(); torch.cuda.synchronize()
print('eager on the same weights: loss', float(le2), 'finite grads', bool(torch.isfinite(flat.buffer).all()))
