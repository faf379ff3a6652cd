"""Diagnostic (not a test): action-argmax agreement of the bf16 product with the fp32 CPU oracle at the BENCHMARK
shapes (cfg-2 DUET B=64, cfg-3 HAMT B=64) over >= 2000 decisions, with every flip listed (reference top-2 gap, error).
Optionally the same statistic for the oracle itself under torch.autocast(bfloat16) - the noise floor of "the reference
run in bf16".   Usage: python tools/diag_argmax.py [duet|hamt|both] [n_batches] [--autocast N] [--out path.json]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from parity_utils import manifest, max_rel, to_dev  # noqa: E402
import vln_imagine_b200.synth as synth  # noqa: E402
from vln_imagine_b200 import config, duet, hamt  # noqa: E402


def flips_of(p, r, seed):
    p, r = p.detach().float().cpu(), r.detach().float().cpu()
    a, b = p.argmax(-1), r.argmax(-1)
    top2 = r.topk(2, -1).values
    gap = top2[:, 0] - top2[:, 1]
    scale = float(r[torch.isfinite(r)].abs().max())
    out = []
    for i in torch.nonzero(a != b).flatten().tolist():
        fin = torch.isfinite(r[i])
        out.append(dict(seed=seed, ep=i, gap=float(gap[i]), gap_rel=float(gap[i]) / scale,
                        err=float((p[i] - r[i])[fin].abs().max())))
    return int((a == b).sum()), a.numel(), out


def run(model_name, n_batches, n_autocast):
    import test_duet_parity_gpu as TD
    import test_hamt_parity_gpu as TH
    if model_name == 'duet':
        from oracle import duet_oracle as O
        model = duet.VLNBert(config.default_duet_args()).cuda().eval()
        shape, make, runp, key = synth.CFG2, synth.duet_episode, TD.run_product, 'fused_logits'
    else:
        from oracle import hamt_oracle as O
        model = hamt.VLNBertCMT(config.default_hamt_args()).cuda().eval()
        shape, make, runp, key = synth.CFG3, synth.hamt_episode, TH.run_product, 'act_logits'
    sd = synth.synth_state_dict(manifest(model_name), seed=0)
    model.vln_bert.load_state_dict(sd)
    model.vln_bert.precision = 'bf16'
    agree = total = 0
    ac_agree = ac_total = 0
    flips, ac_flips = [], []
    worst = 0.0
    t_cpu = 0.0
    for j in range(n_batches):
        seed = 5000 + j
        ep = synth.to_torch(make(shape, seed))
        t0 = time.time()
        with torch.no_grad():
            o_txt, o_img, o_loss, o_img2 = O.episode_prelude(sd, ep)
            nav = O.nav_step(sd, ep, o_txt, o_img2)
        ref = nav[2]['fused_logits'] if model_name == 'duet' else nav[0]
        t_cpu += time.time() - t0
        out = runp(model, to_dev(ep))
        worst = max(worst, max_rel(out[key], ref))
        a, n, f = flips_of(out[key], ref, seed)
        agree, total = agree + a, total + n
        flips += f
        if j < n_autocast:
            with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
                a_txt, a_img, a_loss, a_img2 = O.episode_prelude(sd, ep)
                a_nav = O.nav_step(sd, ep, a_txt.float(), a_img2.float())
            a_ref = a_nav[2]['fused_logits'] if model_name == 'duet' else a_nav[0]
            a, n, f = flips_of(a_ref.float(), ref, seed)
            ac_agree, ac_total = ac_agree + a, ac_total + n
            ac_flips += f
    rec = dict(model=model_name, batch=shape.batch, decisions=total, agree=agree, rate=agree / total, max_rel_logit_err=worst,
               flips=flips, oracle_cpu_seconds=t_cpu)
    if ac_total:
        rec['autocast_reference'] = dict(decisions=ac_total, agree=ac_agree, rate=ac_agree / ac_total, flips=ac_flips)
    return rec


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('which', nargs='?', default='both')
    ap.add_argument('n_batches', nargs='?', type=int, default=32)
    ap.add_argument('--autocast', type=int, default=0)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    recs = [run(m, a.n_batches, a.autocast) for m in (('duet', 'hamt') if a.which == 'both' else (a.which,))]
    for r in recs:
        print(json.dumps(r))
    if a.out:
        with open(a.out, 'w') as f:
            json.dump(recs, f, indent=1)
