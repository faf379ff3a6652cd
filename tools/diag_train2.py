"""Eager vs graph-replayed fine-tuning forward/backward on FIXED weights (diagnostic)."""
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
T = int(os.environ.get('T', '1'))
ep = synth.to_torch(synth.duet_episode(synth.CFG2, 1234))
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in ep.items()}
G, P = ep['gmap_img_embeds'].shape[1], ep['vp_img_embeds'].shape[1]
d['gmap_vpids'], d['vp_cand_vpids'] = net.intern_vpids(ep['gmap_vpids'], ep['vp_cand_vpids'], G, P, dev)
flat = train.FlatGradients(net)
names = [n for n, p in net.named_parameters() if p.requires_grad]


def grad_fn(backward=True):
    flat.zero()
    loss, ce, aux, _ = train.duet_finetune_iteration(model, d, n_steps=T, backward=backward)
    return loss.detach()


for _ in range(3):
    le = grad_fn()
torch.cuda.synchronize()
ge = flat.buffer.clone()
print('eager loss', float(le), 'grad finite', bool(torch.isfinite(ge).all()), 'norm', float(ge.norm()))

# forward only
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    grad_fn(False)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
net._packs = None
g0 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g0):
    lf = grad_fn(False)
g0.replay(); torch.cuda.synchronize()
print('graph forward-only loss', float(lf))
g0.replay(); torch.cuda.synchronize()
print('graph forward-only loss (2nd replay)', float(lf))

net._packs = None
g1 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g1):
    lg = grad_fn(True)
for r in range(2):
    g1.replay(); torch.cuda.synchronize()
    gg = flat.buffer
    print('graph fwd+bwd replay', r, 'loss', float(lg), 'grad finite', bool(torch.isfinite(gg).all()), 'norm', float(gg.norm()))
    bad = []
    for n, p, off in zip(names, flat.params, flat.offsets):
        a, b = gg[off:off + p.numel()], ge[off:off + p.numel()]
        if not torch.isfinite(a).all():
            bad.append((n, 'nonfinite'))
        else:
            e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
            if e > 1e-2:
                bad.append((n, round(e, 4)))
    print('  params differing:', len(bad), bad[:12])
    print('  ... last:', bad[-6:])
