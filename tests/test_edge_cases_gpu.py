"""Edge cases of the navigation path through the module API (C ABI underneath) against the CPU oracle, fp32 check mode
(1e-4) and bf16 mode (2e-2): a single episode, a batch where no imagination qualifies for the alignment loss, an episode
whose imagination mask is all False, the smallest graph ([stop] + the current node, nothing left to visit), HAMT's first
step (history = the [cls] token only), and the attention kernel at its sequence-length limit."""
import dataclasses
import importlib

import pytest
import torch
import torch.nn.functional as F

from parity_utils import TOL, manifest, max_rel, to_dev
from test_duet_cfg5_gpu import take
from test_duet_parity_gpu import run_product as run_duet

pytestmark = pytest.mark.gpu

DUET_KEYS = ('txt_embeds', 'aligned_imagine_embeds', 'pano_embeds', 'gmap_embeds', 'vp_embeds', 'global_logits', 'local_logits',
             'fused_logits')


@pytest.fixture(scope='module')
def duet_env(lib_built):
    synth = importlib.import_module('vln_imagine_b200.synth')
    duet = importlib.import_module('vln_imagine_b200.duet')
    config = importlib.import_module('vln_imagine_b200.config')
    model = duet.VLNBert(config.default_duet_args()).cuda().eval()
    sd = synth.synth_state_dict(manifest('duet'), seed=0, gasa_stress=True)
    model.vln_bert.load_state_dict(sd)
    return synth, model, sd


def duet_oracle_outputs(sd, ep):
    from oracle import duet_oracle as O
    with torch.no_grad():
        txt, img, loss, img2 = O.episode_prelude(sd, ep)
        pano, pano_masks, nav = O.nav_step(sd, ep, txt, img2)
    return dict(txt_embeds=txt, aligned_imagine_embeds=img2, pano_embeds=pano, gmap_embeds=nav['gmap_embeds'],
                vp_embeds=nav['vp_embeds'], global_logits=nav['global_logits'], local_logits=nav['local_logits'],
                fused_logits=nav['fused_logits'], aux_loss=loss)


def check_duet(model, sd, ep, precisions=('fp32', 'bf16')):
    ref = duet_oracle_outputs(sd, ep)
    for precision in precisions:
        model.vln_bert.precision = precision
        out = run_duet(model, to_dev(ep))
        for k in DUET_KEYS:
            assert max_rel(out[k], ref[k]) < TOL[precision], (precision, k)
        assert abs(float(out['aux_loss']) - float(ref['aux_loss'])) <= TOL[precision] * max(abs(float(ref['aux_loss'])), 1e-6)
    return ref


def test_duet_single_episode(duet_env):
    synth, model, sd = duet_env
    ep = synth.to_torch(synth.duet_episode(synth.TINY, 21))
    check_duet(model, sd, take(ep, [1]))


def test_duet_no_imagination_qualifies_for_the_alignment_loss(duet_env):
    """every flag 'False': the reference returns the int 0 and leaves the embeddings alone (vilmodel.py:650-651); here a 0-d
    zero tensor.  One episode additionally has an all-False imagination mask: its imagination tokens are masked keys."""
    synth, model, sd = duet_env
    ep = synth.to_torch(synth.duet_episode(synth.TINY, 22))
    ep['sub_instr_imag_flag'] = [['False'] * len(f) for f in ep['sub_instr_imag_flag']]
    ep['imagine_masks'] = ep['imagine_masks'].clone()
    ep['imagine_masks'][0] = False
    ref = check_duet(model, sd, ep)
    assert float(ref['aux_loss']) == 0.0
    model.vln_bert.precision = 'fp32'
    out = run_duet(model, to_dev(ep))
    assert float(out['aux_loss']) == 0.0 and torch.equal(out['aligned_imagine_embeds'], out['imagine_embeds'])


def test_duet_smallest_graph_nothing_left_to_visit(duet_env):
    """[stop] + the current (visited) node only: every global logit but [stop] is -inf and the fused logits reduce to the stop
    scores (vilmodel.py:1188-1217)"""
    synth, model, sd = duet_env
    ep = synth.to_torch(synth.duet_episode(synth.TINY, 23))
    B = ep['txt_ids'].shape[0]
    for k in ('gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_visited_masks'):
        ep[k] = ep[k][:, :2].contiguous()
    ep['gmap_pair_dists'] = ep['gmap_pair_dists'][:, :2, :2].contiguous()
    ep['gmap_masks'][:] = True
    ep['gmap_visited_masks'][:, 0] = False
    ep['gmap_visited_masks'][:, 1] = True
    ep['gmap_vpids'] = [row[:2] for row in ep['gmap_vpids']]
    ref = check_duet(model, sd, ep)
    assert torch.isinf(ref['global_logits'][:, 1]).all() and torch.isfinite(ref['fused_logits'][:, 0]).all()
    assert ref['fused_logits'].shape == (B, 2)


def test_hamt_first_step_history_is_the_cls_token_only(lib_built):
    synth = importlib.import_module('vln_imagine_b200.synth')
    hamt = importlib.import_module('vln_imagine_b200.hamt')
    config = importlib.import_module('vln_imagine_b200.config')
    from oracle import hamt_oracle as O
    model = hamt.VLNBertCMT(config.default_hamt_args()).cuda().eval()
    sd = synth.synth_state_dict(manifest('hamt'), seed=0)
    model.vln_bert.load_state_dict(sd)
    ep = synth.to_torch(synth.hamt_episode(synth.TINY, 24))
    B = ep['txt_ids'].shape[0]
    with torch.no_grad():
        o_txt, o_img, o_loss, o_img2 = O.episode_prelude(sd, ep)
        cls = O.forward_history(sd, None, None, None, None, None, batch_size=B)          # (B, 768): the t = 0 history
        hm = torch.ones(B, 1, dtype=torch.bool)
        ref = O.forward_visual(sd, o_txt, ep['txt_masks'], cls[:, None], hm, ep['ob_img_feats'], ep['ob_ang_feats'],
                               ep['ob_nav_types'], ep['ob_masks'], o_img2, ep['imagine_masks'])[0]
    d = to_dev(ep)
    for precision in ('fp32', 'bf16'):
        model.vln_bert.precision = precision
        with torch.no_grad():
            txt = model('language', txt_ids=d['txt_ids'], txt_masks=d['txt_masks'])
            img = model('imagine', imagine_pano_img_feats=d['imagine_feats'], imagine_masks=None)
            loss, img2 = model('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=d['txt_masks'],
                               align_imagine_embeds=img.clone(), imagine_masks=d['imagine_masks'],
                               sub_instr_segs=d['sub_instr_segs'], sub_instr_imag_flag=d['sub_instr_imag_flag'],
                               noun_phrase_segs=d['noun_phrase_segs'], obs_instr_ids=d['obs_instr_ids'])
            h0 = model('history').expand(B, -1)                                        # agent_cmt.py:451-455
            (logits,) = model('visual', txt_embeds=txt, txt_masks=d['txt_masks'], hist_embeds=[h0], hist_lens=[1] * B,
                              ob_img_feats=d['ob_img_feats'], ob_ang_feats=d['ob_ang_feats'], ob_nav_types=d['ob_nav_types'],
                              ob_masks=d['ob_masks'], imagine_embeds=img2, imagine_masks=d['imagine_masks'])
        assert max_rel(logits, ref) < TOL[precision], precision


@pytest.mark.parametrize('Lq,Lk', [(100, 512), (1, 1), (17, 300)])
def test_attention_at_the_sequence_limits(lib_built, Lq, Lk):
    ops = importlib.import_module('vln_imagine_b200.ops')
    _lib = importlib.import_module('vln_imagine_b200._lib')
    B, H = 3, 12
    g = torch.Generator().manual_seed(Lq * 1000 + Lk)
    q = torch.randn(B * Lq, 768, generator=g).cuda().bfloat16()
    k = torch.randn(B * Lk, 768, generator=g).cuda().bfloat16()
    v = torch.randn(B * Lk, 768, generator=g).cuda().bfloat16()
    mask = (torch.rand(B, Lk, generator=g) > 0.3)
    mask[:, 0] = True
    out = ops.attention(q, k, v, B, Lq, Lk, key_mask=mask.to(torch.uint8).cuda())
    qf, kf, vf = (t.float().view(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) / 8.0 + (1.0 - mask.float().cuda())[:, None, None, :] * -10000.0
    ref = (F.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * Lq, 768)
    assert max_rel(out, ref) < 2e-2
    if Lk == 512:
        with pytest.raises(_lib.VlnImagineError, match='512'):
            ops.attention(q, torch.cat([k, k[:B]]), torch.cat([v, v[:B]]), B, Lq, Lk + 1)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_shape_buckets_are_invisible(lib_built, monkeypatch, precision):
    """graphs.ShapeBuckets: with the per-step shapes padded to multiples of 8 (G, P, O) / 4 (T) - what a rollout with changing
    shapes switches to so that a handful of CUDA graphs serves it - every output equals the unpadded call (the padding is masked
    exactly like batch padding) and still matches the fixtures of the real reference; a varying-shape sequence of calls ends up
    with fewer captured graphs than distinct shapes"""
    import importlib
    from parity_utils import TOL, golden, manifest, max_rel, to_dev
    from test_duet_parity_gpu import run_product as run_duet
    from test_hamt_parity_gpu import run_product as run_hamt
    synth = importlib.import_module('vln_imagine_b200.synth')
    config = importlib.import_module('vln_imagine_b200.config')
    tol = TOL[precision]
    for kind in ('duet', 'hamt'):
        outs = {}
        for mode in ('0', '1'):
            monkeypatch.setenv('VLN_IMAGINE_SHAPE_BUCKETS', mode)
            if kind == 'duet':
                model = importlib.import_module('vln_imagine_b200.duet').VLNBert(config.default_duet_args()).cuda().eval()
                ep = to_dev(synth.to_torch(synth.duet_episode(synth.CFG1, 1234)))
                run, keys = run_duet, ('gmap_embeds', 'vp_embeds', 'global_logits', 'local_logits', 'fused_logits')
            else:
                model = importlib.import_module('vln_imagine_b200.hamt').VLNBertCMT(config.default_hamt_args()).cuda().eval()
                ep = to_dev(synth.to_torch(synth.hamt_episode(synth.CFG1, 1234)))
                run, keys = run_hamt, ('act_logits', 'states')
            model.vln_bert.load_state_dict(synth.synth_state_dict(manifest(kind), seed=0))
            model.vln_bert.precision = precision
            run(model, ep)                                   # first call of a signature runs eagerly, the second one is captured
            outs[mode] = run(model, ep)
        gold = golden(kind + '_cfg1')
        for k in keys:
            assert outs['1'][k].shape == outs['0'][k].shape, (kind, k)
            assert max_rel(outs['1'][k], outs['0'][k]) < (1e-5 if precision == 'fp32' else tol), (kind, k)
        ref_key = 'fused_logits' if kind == 'duet' else 'act_logits'
        assert max_rel(outs['1'][ref_key], gold[ref_key]) < tol
    # a rollout-like sequence of shapes: bucketing switches itself on after the third distinct signature
    monkeypatch.setenv('VLN_IMAGINE_SHAPE_BUCKETS', 'auto')
    duet = importlib.import_module('vln_imagine_b200.duet')
    model = duet.VLNBert(config.default_duet_args()).cuda().eval()
    model.vln_bert.load_state_dict(synth.synth_state_dict(manifest('duet'), seed=0))
    model.vln_bert.precision = precision
    ep = to_dev(synth.to_torch(synth.duet_episode(synth.CFG1, 1234)))
    G0, P0 = ep['gmap_img_embeds'].shape[1], ep['vp_img_embeds'].shape[1]
    full = run_duet(model, ep)
    shapes = set()
    for cut in range(1, 7):
        G, P = G0 - cut, P0 - (cut % 3)
        e2 = dict(ep)
        for k in ('gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_visited_masks'):
            e2[k] = ep[k][:, :G].contiguous()
        e2['gmap_pair_dists'] = ep['gmap_pair_dists'][:, :G, :G].contiguous()
        e2['gmap_vpids'] = [row[:G] for row in ep['gmap_vpids']]
        for k in ('vp_img_embeds', 'vp_pos_fts', 'vp_masks', 'vp_nav_masks'):
            e2[k] = ep[k][:, :P].contiguous()
        e2['vp_cand_vpids'] = [row[:P] for row in ep['vp_cand_vpids']]
        shapes.add((G, P))
        for _ in range(2):
            out = run_duet(model, e2)
        assert out['fused_logits'].shape == (ep['txt_ids'].shape[0], G) and out['vp_embeds'].shape[1] == P
        assert torch.isfinite(out['gmap_embeds']).all()
    assert len(model._g_nav.entries) < len(shapes) + 1, (len(model._g_nav.entries), len(shapes))
