"""DUET-Imagine at the long-horizon sizes of BASELINE.json configs[4] (200-token instruction, 12 imaginations, 100-node
graph; 256 episodes over 8 GPUs = 32 per GPU).

The CPU oracle needs ~0.5 s per episode at these lengths, so parity is split as the tier prescribes:
  * 4 episodes at the full sequence lengths against the oracle (fp32 check mode 1e-4, bf16 mode 2e-2);
  * the full per-GPU batch of 32 through size-independent properties: episodes are independent (the first 4
    episodes of the 32-batch give the results of the 4-batch run), permuting the batch permutes the results, the
    -inf pattern of the logits is exactly the masks', and the two precisions agree within the bf16 tolerance.
Everything goes through the module API, i.e. the C ABI (long sequences take the one-shot attention path with
more than one 64-key chunk and up to 7 query tiles)."""
import dataclasses
import importlib

import pytest
import torch

from parity_utils import TOL, argmax_report, manifest, max_rel, to_dev
from test_duet_parity_gpu import run_product

pytestmark = pytest.mark.gpu

KEYS = ('txt_embeds', 'aligned_imagine_embeds', 'pano_embeds', 'gmap_embeds', 'vp_embeds', 'global_logits', 'local_logits',
        'fused_logits')


@pytest.fixture(scope='module')
def env(lib_built):
    synth = importlib.import_module('vln_imagine_b200.synth')
    duet = importlib.import_module('vln_imagine_b200.duet')
    config = importlib.import_module('vln_imagine_b200.config')
    model = duet.VLNBert(config.default_duet_args()).cuda().eval()
    model.use_cuda_graphs = False
    model.vln_bert.load_state_dict(synth.synth_state_dict(manifest('duet'), seed=0, gasa_stress=True))
    ep32 = synth.to_torch(synth.duet_episode(dataclasses.replace(synth.CFG5, batch=32), 555))
    return synth, model, ep32


def take(ep, idx):
    """episodes `idx` of a batch dict (tensors and per-episode lists alike)"""
    B = ep['txt_ids'].shape[0]
    out = {}
    for k, v in ep.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B:
            out[k] = v[torch.as_tensor(idx)]
        elif isinstance(v, list) and len(v) == B:
            out[k] = [v[i] for i in idx]
        else:
            out[k] = v
    return out


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_cfg5_lengths_against_the_oracle(env, precision):
    from oracle import duet_oracle as O
    synth, model, ep32 = env
    ep = take(ep32, [0, 1, 2, 3])
    assert ep['txt_ids'].shape[1] == 200 and ep['gmap_img_embeds'].shape[1] == 100 and ep['imagine_feats'].shape[1] == 12
    sd = synth.synth_state_dict(manifest('duet'), seed=0, gasa_stress=True)
    with torch.no_grad():
        txt, img, loss, img2 = O.episode_prelude(sd, ep)
        pano, pano_masks, nav = O.nav_step(sd, ep, txt, img2)
    ref = dict(txt_embeds=txt, aligned_imagine_embeds=img2, pano_embeds=pano, gmap_embeds=nav['gmap_embeds'],
               vp_embeds=nav['vp_embeds'], global_logits=nav['global_logits'], local_logits=nav['local_logits'],
               fused_logits=nav['fused_logits'])
    model.vln_bert.precision = precision
    out = run_product(model, to_dev(ep))
    tol = TOL[precision]
    for k in KEYS:
        assert max_rel(out[k], ref[k]) < tol, k
    assert abs(float(out['aux_loss']) - float(loss)) < tol * abs(float(loss))
    agree, total, decisive_bad = argmax_report(out['fused_logits'], ref['fused_logits'], tol)
    assert decisive_bad == 0 and (precision == 'bf16' or agree == total)


def test_cfg5_full_batch_properties(env):
    synth, model, ep32 = env
    model.vln_bert.precision = 'bf16'
    full = run_product(model, to_dev(ep32))
    # (1) episodes are independent: a sub-batch reproduces its rows of the full batch (same kernels, other row offsets)
    sub = run_product(model, to_dev(take(ep32, [0, 1, 2, 3])))
    for k in KEYS:
        assert max_rel(sub[k], full[k][:4]) < 1e-5, k
    # (2) permutation equivariance
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(3)).tolist()
    shuf = run_product(model, to_dev(take(ep32, perm)))
    for k in KEYS:
        assert max_rel(shuf[k], full[k][torch.as_tensor(perm)]) < 1e-5, k
    # (3) the -inf pattern is exactly the masks' (models/vilmodel.py:1188-1196)
    g = full['global_logits'].cpu()
    dead = ep32['gmap_visited_masks'] | ~ep32['gmap_masks']
    assert torch.equal(torch.isinf(g), dead)
    assert torch.equal(torch.isinf(full['local_logits'].cpu()), ~ep32['vp_nav_masks'])
    assert torch.isfinite(full['fused_logits'].cpu()[~dead]).all()
    # (4) the two precisions agree at full size
    model.vln_bert.precision = 'fp32'
    f32 = run_product(model, to_dev(ep32))
    for k in KEYS:
        assert max_rel(full[k], f32[k]) < TOL['bf16'], k
    agree, total, decisive_bad = argmax_report(full['fused_logits'], f32['fused_logits'], TOL['bf16'])
    assert decisive_bad == 0 and agree >= 0.97 * total          # 32 decisions: at most one near-tie may differ
