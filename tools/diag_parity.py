"""Diagnostic (not a test): error levels and argmax agreement of the bf16 product vs the fp32 CPU oracle
over many synthetic decisions.  Usage: python tools/diag_parity.py [n_seeds] [batch]"""
import dataclasses
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from parity_utils import manifest, max_rel, to_dev  # noqa: E402
from test_duet_parity_gpu import run_product  # noqa: E402
from oracle import duet_oracle as O  # noqa: E402
import vln_imagine_b200.synth as synth  # noqa: E402
from vln_imagine_b200 import duet, config  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.set_num_threads(os.cpu_count())
model = duet.VLNBert(config.default_duet_args()).cuda().eval()
sd = synth.synth_state_dict(manifest('duet'), seed=0)
model.vln_bert.load_state_dict(sd)
shape = dataclasses.replace(synth.CFG1, batch=batch)
rows = []
for precision in ('bf16', 'fp32'):
    model.vln_bert.precision = precision
    agree = total = 0
    errs = {}
    flips = []
    for seed in range(200, 200 + n_seeds):
        ep = synth.to_torch(synth.duet_episode(shape, seed))
        with torch.no_grad():
            o_txt, o_img, o_loss, o_img2 = O.episode_prelude(sd, ep)
            o_pano, _, o_nav = O.nav_step(sd, ep, o_txt, o_img2)
        out = run_product(model, to_dev(ep))
        ref = dict(txt_embeds=o_txt, aligned_imagine_embeds=o_img2, pano_embeds=o_pano, gmap_embeds=o_nav['gmap_embeds'],
                   vp_embeds=o_nav['vp_embeds'], global_logits=o_nav['global_logits'], local_logits=o_nav['local_logits'],
                   fused_logits=o_nav['fused_logits'])
        for k, v in ref.items():
            errs[k] = max(errs.get(k, 0.0), max_rel(out[k], v))
        errs['aux_loss'] = max(errs.get('aux_loss', 0.0), abs(float(out['aux_loss']) - float(o_loss)) / abs(float(o_loss)))
        f, r = out['fused_logits'].cpu(), o_nav['fused_logits']
        a, b = f.argmax(-1), r.argmax(-1)
        top2 = r.topk(2, -1).values
        gap = (top2[:, 0] - top2[:, 1])
        scale = r[torch.isfinite(r)].abs().max()
        for i in torch.nonzero(a != b).flatten().tolist():
            flips.append(dict(seed=seed, ep=i, gap=float(gap[i]), gap_rel=float(gap[i] / scale),
                              err=float((f[i] - r[i])[torch.isfinite(r[i])].abs().max())))
        agree += int((a == b).sum())
        total += a.numel()
    print(json.dumps(dict(precision=precision, decisions=total, agree=agree, rate=agree / total, max_rel_err=errs, flips=flips)))
