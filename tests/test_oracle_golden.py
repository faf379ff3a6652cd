"""CPU: the oracle (oracle/*_oracle.py) against the committed golden vectors of the REAL reference
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference in the build container).
This is what pins the oracle to the reference rather than to itself."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from parity_utils import golden, manifest, max_rel, sub16

synth = importlib.import_module('vln_imagine_b200.synth')


def test_generation_report_says_the_oracle_reproduced_the_reference():
    for m in ('duet', 'hamt'):
        rep = json.load(open(os.path.join(GOLDEN, '%s_oracle_vs_reference.json' % m)))
        assert rep and all(max(case.values()) < 2e-4 for case in rep.values())


@pytest.mark.parametrize('tag,shape,seed,stress', [('tiny', 'TINY', 7, False), ('tiny_gasa', 'TINY', 8, True),
                                                   ('cfg1', 'CFG1', 1234, False), ('cfg1_gasa', 'CFG1', 1234, True)])
def test_duet_oracle_matches_reference_golden(tag, shape, seed, stress):
    from oracle import duet_oracle as O
    sd = synth.synth_state_dict(manifest('duet'), seed=0, gasa_stress=stress)
    ep = synth.to_torch(synth.duet_episode(getattr(synth, shape), seed))
    gold = golden('duet_' + tag)
    with torch.no_grad():
        txt, img, loss, img2 = O.episode_prelude(sd, ep)
        pano, pano_masks, nav = O.nav_step(sd, ep, txt, img2)
        nce_loss, nce_img = O.forward_align_infonce(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'], 0.007)
    f = (lambda t: t) if tag.startswith('tiny') else sub16
    out = dict(txt_embeds=f(txt), imagine_embeds=f(img), aligned_imagine_embeds=f(img2), pano_embeds=f(pano),
               gmap_embeds=f(nav['gmap_embeds']), vp_embeds=f(nav['vp_embeds']), global_logits=nav['global_logits'],
               local_logits=nav['local_logits'], fused_logits=nav['fused_logits'], nce_imagine_embeds=f(nce_img))
    for k, v in out.items():
        assert max_rel(v, gold[k]) < 1e-5, k
    assert torch.equal(pano_masks, gold['pano_masks'])
    assert abs(float(loss) - float(gold['aux_loss'])) < 1e-6
    assert abs(float(nce_loss) - float(gold['nce_loss'])) < 1e-4 * abs(float(gold['nce_loss']))
    assert torch.equal(nav['fused_logits'].argmax(-1), gold['fused_logits'].argmax(-1))


@pytest.mark.parametrize('tag,shape,seed', [('tiny', 'TINY', 7), ('cfg1', 'CFG1', 1234)])
def test_hamt_oracle_matches_reference_golden(tag, shape, seed):
    from oracle import hamt_oracle as O
    sd = synth.synth_state_dict(manifest('hamt'), seed=0)
    ep = synth.to_torch(synth.hamt_episode(getattr(synth, shape), seed))
    gold = golden('hamt_' + tag)
    with torch.no_grad():
        txt, img, loss, img2 = O.episode_prelude(sd, ep)
        logits, txt_o, hist_o, ob_o, hist = O.nav_step(sd, ep, txt, img2)
        cls_hist = O.forward_history(sd, None, None, None, None, None)
    f = (lambda t: t) if tag == 'tiny' else sub16
    out = dict(txt_embeds=f(txt), aligned_imagine_embeds=f(img2), act_logits=logits, txt_out=f(txt_o), hist_out=f(hist_o),
               ob_out=f(ob_o), hist_embed=hist, cls_hist=cls_hist)
    for k, v in out.items():
        assert max_rel(v, gold[k]) < 1e-5, k
    assert abs(float(loss) - float(gold['aux_loss'])) < 1e-6


@pytest.mark.parametrize('tag,shape,seed', [('tiny', 'TINY', 7), ('cfg1', 'CFG1', 1234)])
def test_hamt_encoder_visual_variant_oracle_matches_reference_golden(tag, shape, seed):
    """the parser-default imagination flags of HAMT-Imagine: ImagineEmbeddings encoder + concat_imagine_with='visual'"""
    from oracle import hamt_oracle as O
    rep = json.load(open(os.path.join(GOLDEN, 'hamt_encvis_oracle_vs_reference.json')))
    assert all(max(case.values()) < 2e-4 for case in rep.values())
    sd = synth.synth_state_dict(manifest('hamt_encvis'), seed=0)
    ep = synth.to_torch(synth.hamt_episode(getattr(synth, shape), seed))
    gold = golden('hamt_encvis_' + tag)
    with torch.no_grad():
        txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        img = O.forward_imagination_encoder(sd, ep['imagine_feats'], ep['imagine_masks'])
        loss, img2 = O.forward_align_cosine(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'])
        hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
        logits, txt_o, hist_o, ob_o = O.forward_visual(sd, txt, ep['txt_masks'], ep['hist_embeds'], hm, ep['ob_img_feats'],
                                                       ep['ob_ang_feats'], ep['ob_nav_types'], ep['ob_masks'], img2,
                                                       ep['imagine_masks'], concat_imagine_with='visual')
    f = (lambda t: t) if tag == 'tiny' else sub16
    out = dict(imagine_embeds=f(img), aligned_imagine_embeds=f(img2), act_logits=logits, txt_out=f(txt_o), hist_out=f(hist_o),
               ob_out=f(ob_o))
    for k, v in out.items():
        assert max_rel(v, gold[k]) < 1e-5, k
    assert abs(float(loss) - float(gold['aux_loss'])) < 1e-6


@pytest.mark.parametrize('tag,shape,seed', [('tiny', 'TINY', 7), ('cfg1', 'CFG1', 1234)])
def test_hamt_margin_alignment_oracle_matches_reference_golden(tag, shape, seed):
    from oracle import duet_oracle as D
    from oracle import hamt_oracle as O
    sd = synth.synth_state_dict(manifest('hamt'), seed=0)
    ep = synth.to_torch(synth.hamt_episode(getattr(synth, shape), seed))
    gold = golden('hamt_margin_' + tag)
    with torch.no_grad():
        txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        img = O.forward_imagination(sd, ep['imagine_feats'])
        loss, img2 = D.forward_align_margin(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'], 0.5)
    assert abs(float(loss) - float(gold['margin_loss'])) < 1e-6
    assert max_rel(sub16(img2), gold['margin_imagine_embeds']) < 1e-5


@pytest.mark.parametrize('tag,shape,seed', [('tiny', 'TINY', 7), ('cfg1', 'CFG1', 1234)])
def test_duet_reverie_oracle_matches_reference_golden(tag, shape, seed):
    """REVERIE recipe of DUET-Imagine (scripts/run_reverie.sh): object boxes in the panorama, object-grounding head, one
    imagination per instruction with the REVERIE alignment modules (cosine and InfoNCE)"""
    from oracle import duet_oracle as O
    rep = json.load(open(os.path.join(GOLDEN, 'duet_reverie_oracle_vs_reference.json')))
    assert len(rep) == 4 and all(max(c.values()) < 2e-4 for c in rep.values())
    sd = synth.synth_state_dict(manifest('duet_reverie'), seed=0, gasa_stress=(tag == 'tiny'))
    ep = synth.to_torch(synth.duet_reverie_episode(getattr(synth, shape), seed))
    gold, gold_nce = golden('duet_reverie_%s_cos' % tag), golden('duet_reverie_%s_nce' % tag)
    with torch.no_grad():
        txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        img = O.forward_imagination(sd, ep['imagine_feats'])
        loss, img2 = O.forward_align_reverie(sd, txt, ep['txt_masks'], img, 'cosine')
        nce, nce_img2 = O.forward_align_reverie(sd, txt, ep['txt_masks'], img, 'contrastive-InfoNCE', 0.007)
        pano, pano_masks = O.forward_panorama(sd, ep['view_img_fts'], ep['loc_fts'], ep['nav_types'], ep['view_lens'],
                                              obj_img_fts=ep['obj_img_fts'], obj_lens=ep['obj_lens'])
        nav = O.forward_navigation(sd, txt, ep['txt_masks'], ep['gmap_img_embeds'], ep['gmap_step_ids'], ep['gmap_pos_fts'],
                                   ep['gmap_masks'], ep['gmap_pair_dists'], ep['gmap_visited_masks'], ep['gmap_vpids'],
                                   ep['vp_img_embeds'], ep['vp_pos_fts'], ep['vp_masks'], ep['vp_nav_masks'], ep['vp_cand_vpids'],
                                   img2, ep['imagine_masks'], vp_obj_masks=ep['vp_obj_masks'])
    f = (lambda t: t) if tag == 'tiny' else sub16
    assert torch.equal(pano_masks, gold['pano_masks'])
    for k, v in dict(pano_embeds=f(pano), vp_embeds=f(nav['vp_embeds']), gmap_embeds=f(nav['gmap_embeds']), fused_logits=nav['fused_logits'],
                     local_logits=nav['local_logits'], global_logits=nav['global_logits'], obj_logits=nav['obj_logits'],
                     aligned_imagine_embeds=img2).items():
        assert max_rel(v, gold[k]) < 1e-5, k
    assert abs(float(loss) - float(gold['aux_loss'])) < 1e-6 and abs(float(nce) - float(gold_nce['aux_loss'])) < 1e-5
    assert max_rel(nce_img2, gold_nce['aligned_imagine_embeds']) < 1e-5


def test_duet_soon_model_side_oracle_matches_reference_golden():
    """scripts/run_soon.sh on the model side: 2048-d object boxes through obj_linear / obj_layer_norm, no imagination"""
    from oracle import duet_oracle as O
    rep = json.load(open(os.path.join(GOLDEN, 'duet_soon_oracle_vs_reference.json')))
    assert max(rep['tiny'].values()) < 2e-4
    sd = synth.synth_state_dict(manifest('duet_soon'), seed=0, gasa_stress=True)
    ep = synth.to_torch(synth.duet_reverie_episode(synth.TINY, 9, obj_dim=2048))
    gold = golden('duet_soon_tiny')
    B = ep['txt_ids'].shape[0]
    with torch.no_grad():
        txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        pano, pano_masks = O.forward_panorama(sd, ep['view_img_fts'], ep['loc_fts'], ep['nav_types'], ep['view_lens'],
                                              obj_img_fts=ep['obj_img_fts'], obj_lens=ep['obj_lens'])
        nav = O.forward_navigation(sd, txt, ep['txt_masks'], ep['gmap_img_embeds'], ep['gmap_step_ids'], ep['gmap_pos_fts'],
                                   ep['gmap_masks'], ep['gmap_pair_dists'], ep['gmap_visited_masks'], ep['gmap_vpids'],
                                   ep['vp_img_embeds'], ep['vp_pos_fts'], ep['vp_masks'], ep['vp_nav_masks'], ep['vp_cand_vpids'],
                                   torch.zeros(B, 0, 768), torch.zeros(B, 0, dtype=torch.bool), vp_obj_masks=ep['vp_obj_masks'])
    for k, v in dict(pano_embeds=pano, vp_embeds=nav['vp_embeds'], fused_logits=nav['fused_logits'], obj_logits=nav['obj_logits']).items():
        assert max_rel(v, gold[k]) < 1e-5, k


def test_duet_pretraining_oracle_matches_reference_golden():
    """groundwork for SURVEY 8(f) N4: the DUET pre-training forward (mlm + mrc + sap heads over whole trajectories) restated in
    oracle/pretrain_oracle.py against outputs of the real GlocalTextPathCMTPreTraining"""
    from oracle import pretrain_oracle as P
    rep = json.load(open(os.path.join(GOLDEN, 'duet_pretrain_oracle_vs_reference.json')))
    assert len(rep) == 8 and max(rep.values()) < 2e-4
    sd = synth.synth_state_dict(manifest('duet_pretrain'), seed=0)
    sd['mlm_head.predictions.decoder.weight'] = sd['bert.embeddings.word_embeddings.weight']
    sd['bert.global_encoder.sprel_linear.weight'] = torch.full((1, 1), -0.3)
    ep = synth.to_torch(synth.duet_pretrain_batch())
    gold = golden('duet_pretrain')
    with torch.no_grad():
        gl, ll, fl = P.forward_sap(sd, ep)
        scores = P.forward_mlm(sd, ep)
        logits, _ = P.forward_mrc(sd, ep)
        mlm, mrc = P.losses(sd, ep)
    for k, v in dict(global_logits=gl, local_logits=ll, fused_logits=fl, mlm_scores=scores[:, ::64], mlm_loss=mlm, mrc_logits=logits,
                     mrc_loss=mrc).items():
        assert max_rel(v, gold[k]) < 1e-5, k


def test_hamt_action_token_variants_oracle_matches_reference_golden():
    from oracle import hamt_oracle as O
    rep = json.load(open(os.path.join(GOLDEN, 'hamt_actpred_oracle_vs_reference.json')))
    assert len(rep) == 16 and max(rep.values()) < 2e-4
    sd = synth.synth_state_dict(manifest('hamt'), seed=0)
    ep = synth.to_torch(synth.hamt_episode(synth.TINY, 7))
    gold = golden('hamt_actpred')
    with torch.no_grad():
        txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        img = O.forward_imagination(sd, ep['imagine_feats'])
        hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
        for concat in ('language', 'visual'):
            for tok in ('ob', 'ob_hist', 'ob_txt_hist', 'ob_imagine_text'):
                logits = O.forward_visual(sd, txt, ep['txt_masks'], ep['hist_embeds'], hm, ep['ob_img_feats'], ep['ob_ang_feats'],
                                          ep['ob_nav_types'], ep['ob_masks'], img, ep['imagine_masks'], concat_imagine_with=concat,
                                          act_pred_token=tok)[0]
                assert max_rel(logits, gold['tiny_%s_%s' % (concat, tok)]) < 1e-5, (concat, tok)


def test_oracle_edge_cases_empty_alignment_and_single_admissible_action():
    """no flagged imagination -> loss 0 and embeds untouched; one admissible action -> every other logit is -inf"""
    from oracle import duet_oracle as O
    sd = synth.synth_state_dict(manifest('duet'), seed=0)
    ep = synth.to_torch(synth.duet_episode(synth.TINY, 11))
    txt = torch.randn(3, 24, 768)
    img = torch.randn(3, 3, 768)
    flags = [['False'] * len(f) for f in ep['sub_instr_imag_flag']]
    loss, out = O.forward_align_cosine(sd, txt, img, flags, ep['noun_phrase_segs'])
    assert float(loss) == 0.0 and torch.equal(out, img)
    g = torch.full((1, 4), float('-inf')); g[0, 0] = 0.3
    l = torch.full((1, 3), float('-inf')); l[0, 0] = 0.2
    fused = O.fuse_logits(g, l, [[None, 'a', 'b', 'c']], torch.tensor([[False, True, False, False]]), [[None, 'b', 'c']])
    assert torch.isfinite(fused).sum() == 1 and abs(float(fused[0, 0]) - 0.5) < 1e-6


def test_gradient_fixtures_record_the_reference_autocast_noise():
    """every gradient fixture carries what the unmodified reference shows under torch.autocast(bfloat16) against its own fp32
    gradients (oracle/gen_golden.autocast_noise): the bf16-mode bounds of the GPU gradient tests are 1.5 x these figures"""
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, '*grads_*.npz')))
    assert len(files) >= 8
    for f in files:
        z = np.load(f)
        for k in ('autocast_elem_median', 'autocast_elem_max', 'autocast_norm_median', 'autocast_norm_max', 'autocast_cosine'):
            assert k in z.files, (f, k)
        # bf16 noise of a deep post-LN stack: per-parameter sampled-element errors of several percent, cosine ~0.99
        assert 0.02 < float(z['autocast_elem_median']) < 0.2 and 0.98 < float(z['autocast_cosine']) < 1.0, f
        assert abs(float(z['autocast_loss']) - float(z['loss'])) < 2e-2 * abs(float(z['loss'])), f
