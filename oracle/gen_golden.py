"""TEST INFRASTRUCTURE - generate golden vectors from the REAL reference (build container only).

Imports the unmodified reference modules from /root/reference (which does not exist on the GPU
box), loads the deterministic synthetic weights (vln-imagine_b200/synth.py), runs every mode
of the hot path on seeded synthetic episodes and writes small fixtures to tests/golden/.
In the same run it checks that the CPU oracle (oracle/*_oracle.py) reproduces the reference,
so the oracle is pinned to the reference and not to itself.

    python oracle/gen_golden.py --model duet
    python oracle/gen_golden.py --model hamt     # separate process: both trees call their package `models`

Harness-side shims only; no reference file is modified or copied.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get('VLN_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _shim_transformers():
    from transformers import BertPreTrainedModel, PreTrainedModel
    orig = PreTrainedModel.init_weights

    def init_weights(self):           # transformers>=5: legacy init_weights() needs post_init() first
        if 'all_tied_weights_keys' not in self.__dict__:
            return self.post_init()
        return orig(self)
    BertPreTrainedModel.init_weights = init_weights


DUET_CFG = dict(max_action_steps=100, image_feat_size=768, angle_feat_size=4, obj_feat_size=0, obj_loc_size=3,
                num_l_layers=9, num_pano_layers=2, num_x_layers=4, graph_sprels=True, glocal_fuse=True,
                fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False, update_lang_bert=True,
                output_attentions=True, pred_head_dropout_prob=0.1, use_lang2visn_attn=False,
                imagine_enc_pano=True, max_imagination_len=20, fix_imagine_embeds=False, bypass_imag_encoder=True,
                use_cosine_aux_loss=True, concat_imagine_with='language', fix_lang_inside_cosine_model=True,
                aux_loss_type='cosine', infonce_temperature=0.007, no_loss_test=False, dataset='r2r')
HAMT_CFG = dict(image_feat_size=768, angle_feat_size=4, num_l_layers=9, num_r_layers=0, num_h_layers=0,
                num_x_layers=4, hist_enc_pano=True, num_h_pano_layers=2, fix_lang_embedding=True,
                fix_hist_embedding=True, fix_obs_embedding=False, update_lang_bert=False, output_attentions=True,
                pred_head_dropout_prob=0.1, no_lang_ca=False, act_pred_token='ob_txt', max_action_steps=50,
                imagine_enc_pano=True, max_imagination_len=20, fix_imagine_embeds=False, bypass_imag_encoder=True,
                use_cosine_aux_loss=True, aux_loss_type='cosine', infonce_temperature=0.3,
                contrastive_margin_value=0.5, concat_imagine_with='language', no_loss_test=False)


def build_reference(model, overrides=None):
    from transformers import BertConfig
    _shim_transformers()
    cfg = BertConfig()                 # defaults == bert-base-uncased
    cfg.hidden_dropout_prob = 0.0
    cfg.attention_probs_dropout_prob = 0.0
    fields = dict(DUET_CFG if model == 'duet' else HAMT_CFG)
    fields.update(overrides or {})
    for k, v in fields.items():
        setattr(cfg, k, v)
    if model == 'duet':
        sys.path.insert(0, os.path.join(REF, 'VLN-DUET', 'map_nav_src'))
        from models.vilmodel import GlocalTextPathNavCMT as Net
    else:
        sys.path.insert(0, os.path.join(REF, 'VLN-HAMT', 'finetune_src'))
        from models.vilmodel_cmt import NavCMT as Net
        os.environ.pop('CUDA_LAUNCH_BLOCKING', None)     # import side effect, vilmodel_cmt.py:25
    torch.manual_seed(0)
    return Net(cfg).eval()


def _sub(t):
    """strided sample of a (B,N,768) tensor, enough to pin it without committing megabytes"""
    return t[..., ::16].contiguous()


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def maxdiff(a, b):
    fin = torch.isfinite(a)
    assert torch.equal(fin, torch.isfinite(b)), 'inf pattern differs'
    if fin.sum() == 0:
        return 0.0
    return float((a[fin] - b[fin]).abs().max())


def run_duet(args):
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import duet_oracle as O
    ref = build_reference('duet')
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLD, 'duet_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)
    report = {}
    for tag, shape, seed, stress in [('tiny', synth.TINY, 7, False), ('tiny_gasa', synth.TINY, 8, True),
                                     ('cfg1', synth.CFG1, 1234, False), ('cfg1_gasa', synth.CFG1, 1234, True)]:
        sd = synth.synth_state_dict(manifest, seed=0, gasa_stress=stress)
        ref.load_state_dict(sd)
        ep = synth.to_torch(synth.duet_episode(shape, seed))
        out = {}
        with torch.no_grad():
            txt = ref('language', {'txt_ids': ep['txt_ids'], 'txt_masks': ep['txt_masks']})
            img = ref('imagine', {'imagine_feats': ep['imagine_feats'], 'imagine_masks': None})
            loss, img2 = ref('align_with_contrastive_loss', {
                'align_txt_embeds': txt, 'txt_masks': ep['txt_masks'], 'align_imagine_embeds': img.clone(),
                'imagine_masks': ep['imagine_masks'], 'sub_instr_segs': ep['sub_instr_segs'],
                'sub_instr_imag_flag': ep['sub_instr_imag_flag'], 'noun_phrase_segs': ep['noun_phrase_segs'],
                'obs_instr_ids': ep['obs_instr_ids']})
            pano, pano_masks = ref('panorama', {'view_img_fts': ep['view_img_fts'], 'obj_img_fts': None,
                                                'loc_fts': ep['loc_fts'], 'nav_types': ep['nav_types'],
                                                'view_lens': ep['view_lens'], 'obj_lens': None})
            nav = ref('navigation', {k: ep[k] for k in (
                'txt_masks', 'gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks',
                'gmap_pair_dists', 'gmap_visited_masks', 'gmap_vpids', 'vp_img_embeds', 'vp_pos_fts',
                'vp_masks', 'vp_nav_masks', 'vp_cand_vpids', 'imagine_masks')} | {
                'txt_embeds': txt, 'imagine_embeds': img2, 'vp_obj_masks': None})
            # InfoNCE variant of the aux loss on the same inputs (models/vilmodel.py:689-779)
            ref.config.aux_loss_type = 'contrastive-InfoNCE'
            from models.vilmodel import AlignWithContrastiveLossWithNegativeSamples as NCE
            nce = NCE(ref.config)
            nce.image_proj.load_state_dict(ref.contrastive_alignment_model.image_proj.state_dict())
            nce_loss, nce_img = nce.eval()(txt, ep['txt_masks'], img.clone(), ep['imagine_masks'], ep['sub_instr_segs'],
                                           ep['sub_instr_imag_flag'], ep['noun_phrase_segs'], ep['obs_instr_ids'])
            ref.config.aux_loss_type = 'cosine'

            # the oracle on the same inputs
            o_txt, o_img, o_loss, o_img2 = O.episode_prelude(sd, ep)
            o_pano, o_pmask, o_nav = O.nav_step(sd, ep, o_txt, o_img2)
            o_nce_loss, o_nce_img = O.forward_align_infonce(sd, o_txt, o_img, ep['sub_instr_imag_flag'],
                                                            ep['noun_phrase_segs'], 0.007)
        diffs = {
            'txt': maxdiff(txt, o_txt), 'img': maxdiff(img, o_img), 'loss': abs(float(loss) - float(o_loss)),
            'img2': maxdiff(img2, o_img2), 'pano': maxdiff(pano, o_pano),
            'gmap': maxdiff(nav['gmap_embeds'], o_nav['gmap_embeds']), 'vp': maxdiff(nav['vp_embeds'], o_nav['vp_embeds']),
            'global': maxdiff(nav['global_logits'], o_nav['global_logits']),
            'local': maxdiff(nav['local_logits'], o_nav['local_logits']),
            'fused': maxdiff(nav['fused_logits'], o_nav['fused_logits']),
            'nce_loss': abs(float(nce_loss) - float(o_nce_loss)), 'nce_img': maxdiff(nce_img, o_nce_img),
        }
        assert torch.equal(pano_masks, o_pmask)
        report[tag] = diffs
        print(tag, json.dumps(diffs))
        assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
        full = tag.startswith('tiny')
        f = (lambda t: t) if full else _sub
        out.update(txt_embeds=f(txt), imagine_embeds=f(img), aux_loss=loss, aligned_imagine_embeds=f(img2),
                   pano_embeds=f(pano), pano_masks=pano_masks, gmap_embeds=f(nav['gmap_embeds']),
                   vp_embeds=f(nav['vp_embeds']), global_logits=nav['global_logits'],
                   local_logits=nav['local_logits'], fused_logits=nav['fused_logits'],
                   nce_loss=nce_loss, nce_imagine_embeds=f(nce_img))
        np.savez(os.path.join(GOLD, 'duet_%s.npz' % tag), **_np(out))
    with open(os.path.join(GOLD, 'duet_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


GRAD_SAMPLES = 32


def grad_sample_index(name, numel):
    """the seeded element positions of parameter ``name`` whose gradient values the fixture keeps"""
    import zlib
    g = np.random.Generator(np.random.PCG64([977, zlib.crc32(name.encode())]))
    return g.integers(0, numel, size=GRAD_SAMPLES)


def nav_targets_of(ep):
    """teacher action per episode for the fixture: the LAST admissible graph node (valid and unvisited), so the
    cross-entropy gradient reaches the local->global fusion; node 0 ([stop]) when there is no other"""
    ok = ep['gmap_masks'] & ~ep['gmap_visited_masks']
    G = ok.shape[1]
    idx = torch.arange(G, device=ok.device)[None, :].expand_as(ok)
    return torch.where(ok, idx, torch.zeros_like(idx)).max(1).values


def duet_train_step(model, ep, call):
    """One fine-tuning iteration of the reference agent on one navigation step (r2r/agent.py:407-449 prelude,
    :466-541 step, :617-622 loss):  loss = CE_sum(fused_logits, teacher) / B + cosine_weight * aux_loss  with
    cosine_weight 0.5 (scripts/run_r2r.sh:78); vp_img_embeds = [0 ; pano_embeds] as _nav_vp_variable builds it
    (r2r/agent.py:173-186).  ``call(mode, batch)`` is the model under test."""
    txt = call('language', {'txt_ids': ep['txt_ids'], 'txt_masks': ep['txt_masks']})
    img = call('imagine', {'imagine_feats': ep['imagine_feats'], 'imagine_masks': None})
    aux, img2 = call('align_with_contrastive_loss', {
        'align_txt_embeds': txt, 'txt_masks': ep['txt_masks'], 'align_imagine_embeds': img,
        'imagine_masks': ep['imagine_masks'], 'sub_instr_segs': ep['sub_instr_segs'],
        'sub_instr_imag_flag': ep['sub_instr_imag_flag'], 'noun_phrase_segs': ep['noun_phrase_segs'],
        'obs_instr_ids': ep['obs_instr_ids']})
    pano, pano_masks = call('panorama', {'view_img_fts': ep['view_img_fts'], 'obj_img_fts': None,
                                         'loc_fts': ep['loc_fts'], 'nav_types': ep['nav_types'],
                                         'view_lens': ep['view_lens'], 'obj_lens': None})
    vp_img = torch.cat([torch.zeros_like(pano[:, :1]), pano], 1)
    nav = call('navigation', {k: ep[k] for k in (
        'txt_masks', 'gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_pair_dists',
        'gmap_visited_masks', 'gmap_vpids', 'vp_pos_fts', 'vp_masks', 'vp_nav_masks', 'vp_cand_vpids',
        'imagine_masks')} | {'txt_embeds': txt, 'imagine_embeds': img2, 'vp_obj_masks': None, 'vp_img_embeds': vp_img})
    tgt = nav_targets_of(ep).to(nav['fused_logits'].device)
    ce = torch.nn.functional.cross_entropy(nav['fused_logits'], tgt, reduction='sum') / tgt.shape[0]
    loss = ce + 0.5 * aux
    return loss, ce, aux, nav


class _Clone(torch.nn.Module):
    """harness-side stand-in for Dropout(p) at p = 0 that keeps the reference's backward alive"""

    def forward(self, x):
        return x.clone()


def run_duet_grads(args):
    """Gradient fixtures for BASELINE.json cfg-4 from the REAL reference: every parameter's gradient L2 norm and
    GRAD_SAMPLES seeded element values, plus the loss terms."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    ref = build_reference('duet')
    ref.contrastive_alignment_model.image_proj.dropout = _Clone()
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    for tag, shape, seed, stress in [('tiny', synth.TINY, 7, True), ('cfg1', synth.CFG1, 1234, False)]:
        sd = synth.synth_state_dict(manifest, seed=0, gasa_stress=stress)
        ref.load_state_dict(sd)
        ref.zero_grad(set_to_none=True)
        ep = synth.to_torch(synth.duet_episode(shape, seed))
        loss, ce, aux, nav = duet_train_step(ref, ep, lambda mode, batch: ref(mode, batch))
        loss.backward()
        out = {'loss': loss.detach(), 'ce': ce.detach(), 'aux': aux.detach(), 'fused_logits': nav['fused_logits'].detach(),
               'nav_targets': nav_targets_of(ep)}
        names, norms, samples = [], [], []
        for name, p in ref.named_parameters():
            assert p.grad is not None, name
            names.append(name)
            g = p.grad.detach().double().reshape(-1)
            norms.append(float(g.norm()))
            samples.append(g[torch.from_numpy(grad_sample_index(name, g.numel()))].float().numpy())
        out['grad_norms'] = np.asarray(norms, np.float64)
        out['grad_samples'] = np.stack(samples)
        full = {name: dict(ref.named_parameters())[name].grad.detach().clone() for name in (
            'global_encoder.sprel_linear.weight', 'global_encoder.sprel_linear.bias', 'imagine_embeddings.type_embedding.weight',
            'embeddings.LayerNorm.weight', 'sap_fuse_linear.net.3.weight', 'local_encoder.vp_pos_embeddings.0.weight')}
        autocast_noise(ref, lambda: duet_train_step(ref, ep, lambda mode, batch: ref(mode, batch))[0], out)
        for name in ('global_encoder.sprel_linear.weight', 'global_encoder.sprel_linear.bias',
                     'imagine_embeddings.type_embedding.weight', 'embeddings.LayerNorm.weight',
                     'sap_fuse_linear.net.3.weight', 'local_encoder.vp_pos_embeddings.0.weight'):
            out['full::' + name] = full[name]
        np.savez(os.path.join(GOLD, 'duet_grads_%s.npz' % tag), **_np(out))
        with open(os.path.join(GOLD, 'duet_grads_names.json'), 'w') as f:
            json.dump(names, f, indent=0)
        print(tag, 'loss', float(loss), 'ce', float(ce), 'aux', float(aux), 'max grad norm', max(norms), 'min', min(norms))


def hamt_targets_of(ep):
    """teacher action for the HAMT fixture: the last admissible observation token (nav_type != 0), i.e. [stop]"""
    ok = ep['ob_nav_types'] != 0
    idx = torch.arange(ok.shape[1], device=ok.device)[None, :].expand_as(ok)
    return torch.where(ok, idx, torch.zeros_like(idx)).max(1).values


def hamt_train_step(call, ep, hist_masks):
    """One fine-tuning step of the HAMT agent on one navigation step (r2r/agent_cmt.py:371-605 + loss :610-640):
    loss = CE_sum(act_logits, teacher) / B + 0.5 * aux.  ``call(mode, **kw)`` is the NavCMT-level model under test; the
    history embeddings are inputs (fix_hist_embedding, the released configuration)."""
    txt = call('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
    img = call('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=None)
    aux, img2 = call('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=ep['txt_masks'],
                     align_imagine_embeds=img, imagine_masks=ep['imagine_masks'], sub_instr_segs=ep['sub_instr_segs'],
                     sub_instr_imag_flag=ep['sub_instr_imag_flag'], noun_phrase_segs=ep['noun_phrase_segs'],
                     obs_instr_ids=ep['obs_instr_ids'])
    logits, txt_o, hist_o, ob_o = call(
        'visual', txt_embeds=txt, txt_masks=ep['txt_masks'], hist_embeds=ep['hist_embeds'], hist_masks=hist_masks,
        ob_img_feats=ep['ob_img_feats'], ob_ang_feats=ep['ob_ang_feats'], ob_nav_types=ep['ob_nav_types'],
        ob_masks=ep['ob_masks'], imagine_embeds=img2, imagine_masks=ep['imagine_masks'])
    tgt = hamt_targets_of(ep)
    ce = torch.nn.functional.cross_entropy(logits, tgt, reduction='sum') / tgt.shape[0]
    return ce + 0.5 * aux, ce, aux, logits


def run_hamt_grads(args):
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import hamt_oracle as O
    ref = build_reference('hamt')
    ref.contrastive_alignment_model.image_proj.dropout = _Clone()
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    sd = synth.synth_state_dict(manifest, seed=0)
    ref.load_state_dict(sd)
    for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
        ref.zero_grad(set_to_none=True)
        ep = synth.to_torch(synth.hamt_episode(shape, seed))
        hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
        loss, ce, aux, logits = hamt_train_step(lambda mode, **kw: ref(mode, **kw), ep, hm)
        loss.backward()
        out = {'loss': loss.detach(), 'ce': ce.detach(), 'aux': aux.detach(), 'act_logits': logits.detach()}
        names, norms, samples = [], [], []
        for name, p in ref.named_parameters():
            if p.grad is None:
                continue
            names.append(name)
            g = p.grad.detach().double().reshape(-1)
            norms.append(float(g.norm()))
            samples.append(g[torch.from_numpy(grad_sample_index(name, g.numel()))].float().numpy())
        out['grad_norms'] = np.asarray(norms, np.float64)
        out['grad_samples'] = np.stack(samples)
        autocast_noise(ref, lambda: hamt_train_step(lambda mode, **kw: ref(mode, **kw), ep, hm)[0], out)
        np.savez(os.path.join(GOLD, 'hamt_grads_%s.npz' % tag), **_np(out))
        with open(os.path.join(GOLD, 'hamt_grads_names.json'), 'w') as f:
            json.dump(names, f, indent=0)
        print(tag, 'loss', float(loss.detach()), 'ce', float(ce.detach()), 'aux', float(aux.detach()), 'params with grad', len(names),
              'of', len(list(ref.named_parameters())))


# HAMT-Imagine fine-tuning variants beyond the released recipe: the flags each one overrides
HAMT_GRAD_VARIANTS = {
    # the PARSER defaults (r2r/parser.py:109,122): ImagineEmbeddings encoder trained, imagination tokens on the vision stream
    'encvis_imgtxt': dict(bypass_imag_encoder=False, concat_imagine_with='visual', act_pred_token='ob_imagine_text'),
    # margin form of the alignment loss (models/vilmodel_cmt.py:825-856) + the 'ob_txt_hist' action token (:1195-1196)
    'margin_txthist': dict(aux_loss_type='constrastive-margin', contrastive_margin_value=0.5, act_pred_token='ob_txt_hist'),
    # history embeddings trained (fix_hist_embedding off, the parser default, r2r/parser.py:57; models/vilmodel_cmt.py:576-618,1036)
    'hist_obhist': dict(fix_hist_embedding=False, act_pred_token='ob_hist'),
}


def hamt_variant_train_step(call, ep, hist_masks, over):
    """hamt_train_step for the variants: 'imagine' gets the masks when the encoder runs (r2r/agent_cmt.py:413-417); with trainable
    history embeddings the history tokens are [cls ; one step] from two 'history' calls (r2r/agent_cmt.py:424,596-603)."""
    txt = call('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
    enc = not over.get('bypass_imag_encoder', True)
    img = call('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=ep['imagine_masks'] if enc else None)
    aux, img2 = call('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=ep['txt_masks'],
                     align_imagine_embeds=img, imagine_masks=ep['imagine_masks'], sub_instr_segs=ep['sub_instr_segs'],
                     sub_instr_imag_flag=ep['sub_instr_imag_flag'], noun_phrase_segs=ep['noun_phrase_segs'],
                     obs_instr_ids=ep['obs_instr_ids'])
    hist_embeds = ep['hist_embeds']
    if not over.get('fix_hist_embedding', True):
        B = ep['txt_ids'].shape[0]
        cls = call('history').reshape(1, -1).expand(B, -1)
        step = call('history', hist_img_feats=ep['hist_img_feats'], hist_ang_feats=ep['hist_ang_feats'],
                    ob_step_ids=torch.full((1,), int(ep['ob_step']), dtype=torch.long, device=ep['txt_ids'].device),
                    hist_pano_img_feats=ep['hist_pano_img_feats'], hist_pano_ang_feats=ep['hist_pano_ang_feats'])
        hist_embeds = torch.stack([cls, step], 1)
        hist_masks = torch.ones((B, 2), dtype=torch.bool, device=hist_embeds.device)
    logits = call('visual', txt_embeds=txt, txt_masks=ep['txt_masks'], hist_embeds=hist_embeds, hist_masks=hist_masks,
                  ob_img_feats=ep['ob_img_feats'], ob_ang_feats=ep['ob_ang_feats'], ob_nav_types=ep['ob_nav_types'],
                  ob_masks=ep['ob_masks'], imagine_embeds=img2, imagine_masks=ep['imagine_masks'])[0]
    tgt = hamt_targets_of(ep)
    ce = torch.nn.functional.cross_entropy(logits, tgt, reduction='sum') / tgt.shape[0]
    return ce + 0.5 * aux, ce, aux, logits


def autocast_noise(ref, run_step, out):
    """What ANY bf16 evaluation of this step does to the gradients: the unmodified reference under torch.autocast(bfloat16) against
    its own fp32 gradients (already in .grad), measured exactly as the GPU tests measure the product (sampled elements scaled by
    max(|ref sample|, 3 rms), per-parameter norms, per-parameter cosine).  Stored in the fixture: the bf16-mode bounds of the tests
    are multiples of these numbers, not free constants."""
    fp32 = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    ref.zero_grad(set_to_none=True)
    with torch.autocast('cpu', dtype=torch.bfloat16):
        loss = run_step()
    loss.float().backward()
    top = max(float(g.double().norm()) for g in fp32.values())
    elem, nerr, dots = [], [], []
    for n, g32 in fp32.items():
        g = dict(ref.named_parameters())[n].grad
        ref_norm = float(g32.double().norm())
        if g is None or ref_norm < 1e-7 * top:
            continue
        idx = torch.from_numpy(grad_sample_index(n, g32.numel()))
        got, want = g.reshape(-1)[idx].float(), g32.reshape(-1)[idx].float()
        scale = max(float(want.abs().max()), 3.0 * ref_norm / np.sqrt(g32.numel()))
        elem.append(float((got - want).abs().max()) / scale)
        nerr.append(abs(float(g.double().norm()) - ref_norm) / ref_norm)
        if g32.numel() == 1:
            # scalar parameters (the GASA slope: a sum over all score gradients with heavy cancellation) are noise-dominated in
            # bf16 - their own relative error under autocast is recorded so that the test bound is not a free constant
            out['autocast_scalar::' + n] = np.float64(abs(float(g.reshape(-1)[0]) - float(g32.reshape(-1)[0])) / abs(float(g32.reshape(-1)[0])))
        if g32.numel() >= 32:
            dots.append(float((got * want).sum() / (got.norm() * want.norm()).clamp_min(1e-30)))
    out['autocast_elem_median'] = np.float64(np.median(elem))
    out['autocast_elem_max'] = np.float64(np.max(elem))
    out['autocast_norm_median'] = np.float64(np.median(nerr))
    out['autocast_norm_max'] = np.float64(np.max(nerr))
    out['autocast_cosine'] = np.float64(np.mean(dots))
    out['autocast_loss'] = loss.detach().float()
    for n, g32 in fp32.items():                                  # leave the fp32 gradients in place
        dict(ref.named_parameters())[n].grad = g32
    print('   reference under autocast(bf16): element err median %.3f max %.3f, norm err median %.4f max %.3f, cosine %.5f'
          % (out['autocast_elem_median'], out['autocast_elem_max'], out['autocast_norm_median'], out['autocast_norm_max'],
             out['autocast_cosine']))


def _grad_fixture(ref, out):
    names, norms, samples = [], [], []
    for name, p in ref.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        g = p.grad.detach().double().reshape(-1)
        norms.append(float(g.norm()))
        samples.append(g[torch.from_numpy(grad_sample_index(name, g.numel()))].float().numpy())
    out['grad_norms'] = np.asarray(norms, np.float64)
    out['grad_samples'] = np.stack(samples)
    return names


def run_hamt_variant_grads(args):
    """Gradient fixtures of the HAMT fine-tuning variants from the REAL reference (TINY episodes)."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import hamt_oracle as O
    all_names = {}
    for tag, over in HAMT_GRAD_VARIANTS.items():
        ref = build_reference('hamt', over)
        ref.contrastive_alignment_model.image_proj.dropout = _Clone()
        manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
        ref.load_state_dict(synth.synth_state_dict(manifest, seed=0))
        ref.zero_grad(set_to_none=True)
        ep = synth.to_torch(synth.hamt_episode(synth.TINY, 7))
        hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
        loss, ce, aux, logits = hamt_variant_train_step(lambda mode='history', **kw: ref(mode, **kw), ep, hm, over)
        loss.backward()
        out = {'loss': loss.detach(), 'ce': ce.detach(), 'aux': aux.detach(), 'act_logits': logits.detach()}
        all_names[tag] = _grad_fixture(ref, out)
        autocast_noise(ref, lambda: hamt_variant_train_step(lambda mode='history', **kw: ref(mode, **kw), ep, hm, over)[0], out)
        np.savez(os.path.join(GOLD, 'hamt_grads_%s.npz' % tag), **_np(out))
        print(tag, 'loss', float(loss.detach()), 'ce', float(ce.detach()), 'aux', float(aux.detach()), 'params with grad', len(all_names[tag]))
    with open(os.path.join(GOLD, 'hamt_grads_variant_names.json'), 'w') as f:
        json.dump(all_names, f, indent=0)


def duet_reverie_train_step(call, ep):
    """One fine-tuning step of the REVERIE agent (reverie/agent_obj.py:373-463,544-546): navigation cross-entropy + object-grounding
    cross-entropy over the episodes that see an object (the first box is the teacher's) + 0.5 * aux."""
    txt = call('language', {'txt_ids': ep['txt_ids'], 'txt_masks': ep['txt_masks']})
    img = call('imagine', {'imagine_feats': ep['imagine_feats'], 'imagine_masks': None})
    aux, img2 = call('align_with_contrastive_loss', {
        'align_txt_embeds': txt, 'txt_masks': ep['txt_masks'], 'align_imagine_embeds': img,
        'imagine_masks': ep['imagine_masks'], 'obs_instr_ids': ep['obs_instr_ids']})
    pano, pano_masks = call('panorama', {k: ep[k] for k in ('view_img_fts', 'obj_img_fts', 'loc_fts', 'nav_types', 'view_lens', 'obj_lens')})
    vp_img = torch.cat([torch.zeros_like(pano[:, :1]), pano], 1)
    nav = call('navigation', {**{k: ep[k] for k in (
        'txt_masks', 'gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_pair_dists', 'gmap_visited_masks',
        'gmap_vpids', 'vp_pos_fts', 'vp_masks', 'vp_nav_masks', 'vp_obj_masks', 'vp_cand_vpids', 'imagine_masks')},
        'txt_embeds': txt, 'imagine_embeds': img2, 'vp_img_embeds': vp_img})
    tgt = nav_targets_of(ep).to(nav['fused_logits'].device)
    B = tgt.shape[0]
    ce = torch.nn.functional.cross_entropy(nav['fused_logits'], tgt, reduction='sum') / B
    has = torch.nonzero(ep['obj_lens'].to(tgt.device) > 0).view(-1)
    og_t = (ep['view_lens'].to(tgt.device) + 1)[has]                      # first object token of the local stream
    og = torch.nn.functional.cross_entropy(nav['obj_logits'][has], og_t, reduction='sum') / B
    return ce + og + 0.5 * aux, ce, og, aux, nav


def run_duet_reverie_grads(args):
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    ref = build_reference('duet', dict(dataset='reverie', obj_feat_size=768, aux_loss_type='cosine'))
    ref.contrastive_alignment_model.image_proj.dropout = _Clone()
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    ref.load_state_dict(synth.synth_state_dict(manifest, seed=0, gasa_stress=True))
    ref.zero_grad(set_to_none=True)
    ep = synth.to_torch(synth.duet_reverie_episode(synth.TINY, 7))
    loss, ce, og, aux, nav = duet_reverie_train_step(lambda mode, batch: ref(mode, batch), ep)
    loss.backward()
    out = {'loss': loss.detach(), 'ce': ce.detach(), 'og': og.detach(), 'aux': aux.detach(), 'fused_logits': nav['fused_logits'].detach(),
           'obj_logits': nav['obj_logits'].detach()}
    names = _grad_fixture(ref, out)
    autocast_noise(ref, lambda: duet_reverie_train_step(lambda mode, batch: ref(mode, batch), ep)[0], out)
    np.savez(os.path.join(GOLD, 'duet_reverie_grads_tiny.npz'), **_np(out))
    with open(os.path.join(GOLD, 'duet_reverie_grads_names.json'), 'w') as f:
        json.dump(names, f, indent=0)
    print('reverie loss', float(loss.detach()), 'ce', float(ce.detach()), 'og', float(og.detach()), 'aux', float(aux.detach()),
          'params with grad', len(names), 'of', len(manifest))


# DUET-Imagine fine-tuning variants beyond the released recipe: the flags each one overrides
DUET_GRAD_VARIANTS = {
    # the text encoder trained through the alignment loss as well (fix_lang_inside_cosine_model off, the parser default,
    # r2r/parser.py:128; models/vilmodel.py:1256-1262): the noun-phrase means carry gradients
    'unfixlang': dict(fix_lang_inside_cosine_model=False),
}


def run_duet_variant_grads(args):
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    all_names = {}
    for tag, over in DUET_GRAD_VARIANTS.items():
        ref = build_reference('duet', over)
        ref.contrastive_alignment_model.image_proj.dropout = _Clone()
        manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
        ref.load_state_dict(synth.synth_state_dict(manifest, seed=0, gasa_stress=True))
        ref.zero_grad(set_to_none=True)
        ep = synth.to_torch(synth.duet_episode(synth.TINY, 7))
        loss, ce, aux, nav = duet_train_step(ref, ep, lambda mode, batch: ref(mode, batch))
        loss.backward()
        out = {'loss': loss.detach(), 'ce': ce.detach(), 'aux': aux.detach(), 'fused_logits': nav['fused_logits'].detach()}
        all_names[tag] = _grad_fixture(ref, out)
        autocast_noise(ref, lambda: duet_train_step(ref, ep, lambda mode, batch: ref(mode, batch))[0], out)
        np.savez(os.path.join(GOLD, 'duet_grads_%s.npz' % tag), **_np(out))
        print(tag, 'loss', float(loss.detach()), 'ce', float(ce.detach()), 'aux', float(aux.detach()), 'params with grad', len(all_names[tag]))
    with open(os.path.join(GOLD, 'duet_grads_variant_names.json'), 'w') as f:
        json.dump(all_names, f, indent=0)


def run_hamt(args):
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import hamt_oracle as O
    ref = build_reference('hamt')
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLD, 'hamt_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)
    sd = synth.synth_state_dict(manifest, seed=0)
    ref.load_state_dict(sd)
    report = {}
    for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
        ep = synth.to_torch(synth.hamt_episode(shape, seed))
        with torch.no_grad():
            txt = ref('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
            img = ref('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=None)
            loss, img2 = ref('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=ep['txt_masks'],
                             align_imagine_embeds=img.clone(), imagine_masks=ep['imagine_masks'],
                             sub_instr_segs=ep['sub_instr_segs'], sub_instr_imag_flag=ep['sub_instr_imag_flag'],
                             noun_phrase_segs=ep['noun_phrase_segs'], obs_instr_ids=ep['obs_instr_ids'])
            hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
            logits, txt_o, hist_o, ob_o = ref(
                'visual', txt_embeds=txt, txt_masks=ep['txt_masks'], hist_embeds=ep['hist_embeds'], hist_masks=hm,
                ob_img_feats=ep['ob_img_feats'], ob_ang_feats=ep['ob_ang_feats'], ob_nav_types=ep['ob_nav_types'],
                ob_masks=ep['ob_masks'], imagine_embeds=img2, imagine_masks=ep['imagine_masks'])
            hist = ref('history', hist_img_feats=ep['hist_img_feats'], hist_ang_feats=ep['hist_ang_feats'],
                       ob_step_ids=torch.LongTensor([ep['ob_step']]),
                       hist_pano_img_feats=ep['hist_pano_img_feats'], hist_pano_ang_feats=ep['hist_pano_ang_feats'])
            cls_hist = ref('history')
            o_txt, o_img, o_loss, o_img2 = O.episode_prelude(sd, ep)
            o_logits, o_txt_o, o_hist_o, o_ob_o, o_hist = O.nav_step(sd, ep, o_txt, o_img2)
            o_cls = O.forward_history(sd, None, None, None, None, None)
        diffs = {'txt': maxdiff(txt, o_txt), 'loss': abs(float(loss) - float(o_loss)), 'img2': maxdiff(img2, o_img2),
                 'logits': maxdiff(logits, o_logits), 'txt_o': maxdiff(txt_o, o_txt_o),
                 'hist_o': maxdiff(hist_o, o_hist_o), 'ob_o': maxdiff(ob_o, o_ob_o), 'hist': maxdiff(hist, o_hist),
                 'cls_hist': maxdiff(cls_hist, o_cls)}
        report[tag] = diffs
        print(tag, json.dumps(diffs))
        assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
        f = (lambda t: t) if tag == 'tiny' else _sub
        np.savez(os.path.join(GOLD, 'hamt_%s.npz' % tag), **_np(dict(
            txt_embeds=f(txt), aux_loss=loss, aligned_imagine_embeds=f(img2), act_logits=logits,
            txt_out=f(txt_o), hist_out=f(hist_o), ob_out=f(ob_o), hist_embed=hist, cls_hist=cls_hist)))
    with open(os.path.join(GOLD, 'hamt_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


def run_hamt_encvis(args):
    """HAMT-Imagine with the PARSER DEFAULTS of the imagination flags (r2r/parser.py:109,122): the ImagineEmbeddings encoder
    (bypass_imag_encoder=False) and the imagination tokens on the vision stream (concat_imagine_with='visual')."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import hamt_oracle as O
    ref = build_reference('hamt', dict(bypass_imag_encoder=False, concat_imagine_with='visual'))
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLD, 'hamt_encvis_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)
    sd = synth.synth_state_dict(manifest, seed=0)
    ref.load_state_dict(sd)
    report = {}
    for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
        ep = synth.to_torch(synth.hamt_episode(shape, seed))
        with torch.no_grad():
            txt = ref('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
            img = ref('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=ep['imagine_masks'])
            loss, img2 = ref('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=ep['txt_masks'],
                             align_imagine_embeds=img.clone(), imagine_masks=ep['imagine_masks'],
                             sub_instr_segs=ep['sub_instr_segs'], sub_instr_imag_flag=ep['sub_instr_imag_flag'],
                             noun_phrase_segs=ep['noun_phrase_segs'], obs_instr_ids=ep['obs_instr_ids'])
            hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
            logits, txt_o, hist_o, ob_o = ref(
                'visual', txt_embeds=txt, txt_masks=ep['txt_masks'], hist_embeds=ep['hist_embeds'], hist_masks=hm,
                ob_img_feats=ep['ob_img_feats'], ob_ang_feats=ep['ob_ang_feats'], ob_nav_types=ep['ob_nav_types'],
                ob_masks=ep['ob_masks'], imagine_embeds=img2, imagine_masks=ep['imagine_masks'])
            o_txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
            o_img = O.forward_imagination_encoder(sd, ep['imagine_feats'], ep['imagine_masks'])
            o_loss, o_img2 = O.forward_align_cosine(sd, o_txt, o_img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'])
            o_logits, o_txt_o, o_hist_o, o_ob_o = O.forward_visual(
                sd, o_txt, ep['txt_masks'], ep['hist_embeds'], hm, ep['ob_img_feats'], ep['ob_ang_feats'], ep['ob_nav_types'],
                ep['ob_masks'], o_img2, ep['imagine_masks'], concat_imagine_with='visual')
        diffs = {'img': maxdiff(img, o_img), 'loss': abs(float(loss) - float(o_loss)), 'img2': maxdiff(img2, o_img2),
                 'logits': maxdiff(logits, o_logits), 'txt_o': maxdiff(txt_o, o_txt_o), 'hist_o': maxdiff(hist_o, o_hist_o),
                 'ob_o': maxdiff(ob_o, o_ob_o)}
        report[tag] = diffs
        print(tag, json.dumps(diffs))
        assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
        f = (lambda t: t) if tag == 'tiny' else _sub
        np.savez(os.path.join(GOLD, 'hamt_encvis_%s.npz' % tag), **_np(dict(
            imagine_embeds=f(img), aux_loss=loss, aligned_imagine_embeds=f(img2), act_logits=logits,
            txt_out=f(txt_o), hist_out=f(hist_o), ob_out=f(ob_o))))
    with open(os.path.join(GOLD, 'hamt_encvis_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


def run_duet_reverie(args):
    """DUET-Imagine as released for REVERIE (scripts/run_reverie.sh: dataset reverie, obj_feat_size 768 -> objects through
    img_linear, og_head, one imagination per instruction, AlignWithContrastiveLossReverie); both aux-loss types."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import duet_oracle as O
    report = {}
    for aux in ('cosine', 'contrastive-InfoNCE'):
        ref = build_reference('duet', dict(dataset='reverie', obj_feat_size=768, aux_loss_type=aux))
        manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
        if aux == 'cosine':
            with open(os.path.join(GOLD, 'duet_reverie_manifest.json'), 'w') as f:
                json.dump(manifest, f, indent=0)
        for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
            sd = synth.synth_state_dict(manifest, seed=0, gasa_stress=(tag == 'tiny'))
            ref.load_state_dict(sd)
            ep = synth.to_torch(synth.duet_reverie_episode(shape, seed))
            with torch.no_grad():
                txt = ref('language', {'txt_ids': ep['txt_ids'], 'txt_masks': ep['txt_masks']})
                img = ref('imagine', {'imagine_feats': ep['imagine_feats'], 'imagine_masks': None})
                loss, img2 = ref('align_with_contrastive_loss', {
                    'align_txt_embeds': txt, 'txt_masks': ep['txt_masks'], 'align_imagine_embeds': img.clone(),
                    'imagine_masks': ep['imagine_masks'], 'obs_instr_ids': ep['obs_instr_ids']})
                o_txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
                o_img = O.forward_imagination(sd, ep['imagine_feats'])
                o_loss, o_img2 = O.forward_align_reverie(sd, o_txt, ep['txt_masks'], o_img, aux, 0.007)
                diffs = {'loss': abs(float(loss) - float(o_loss)), 'img2': maxdiff(img2, o_img2)}
                out = {'aux_loss': loss, 'aligned_imagine_embeds': img2}
                if aux == 'cosine':
                    pano, pano_masks = ref('panorama', {k: ep[k] for k in ('view_img_fts', 'obj_img_fts', 'loc_fts', 'nav_types',
                                                                               'view_lens', 'obj_lens')})
                    nav = ref('navigation', {**{k: ep[k] for k in (
                        'txt_masks', 'gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_pair_dists',
                        'gmap_visited_masks', 'gmap_vpids', 'vp_img_embeds', 'vp_pos_fts', 'vp_masks', 'vp_nav_masks',
                        'vp_obj_masks', 'vp_cand_vpids', 'imagine_masks')}, 'txt_embeds': txt, 'imagine_embeds': img2})
                    o_pano, o_pm = O.forward_panorama(sd, ep['view_img_fts'], ep['loc_fts'], ep['nav_types'], ep['view_lens'],
                                                      obj_img_fts=ep['obj_img_fts'], obj_lens=ep['obj_lens'])
                    o_nav = O.forward_navigation(
                        sd, o_txt, ep['txt_masks'], ep['gmap_img_embeds'], ep['gmap_step_ids'], ep['gmap_pos_fts'], ep['gmap_masks'],
                        ep['gmap_pair_dists'], ep['gmap_visited_masks'], ep['gmap_vpids'], ep['vp_img_embeds'], ep['vp_pos_fts'],
                        ep['vp_masks'], ep['vp_nav_masks'], ep['vp_cand_vpids'], o_img2, ep['imagine_masks'],
                        vp_obj_masks=ep['vp_obj_masks'])
                    assert torch.equal(pano_masks, o_pm)
                    diffs.update(pano=maxdiff(pano, o_pano), fused=maxdiff(nav['fused_logits'], o_nav['fused_logits']),
                                 obj=maxdiff(nav['obj_logits'], o_nav['obj_logits']), vp=maxdiff(nav['vp_embeds'], o_nav['vp_embeds']))
                    f = (lambda t: t) if tag == 'tiny' else _sub
                    out.update(pano_embeds=f(pano), pano_masks=pano_masks, vp_embeds=f(nav['vp_embeds']), gmap_embeds=f(nav['gmap_embeds']),
                               fused_logits=nav['fused_logits'], local_logits=nav['local_logits'], global_logits=nav['global_logits'],
                               obj_logits=nav['obj_logits'])
            key = '%s_%s' % (tag, 'cos' if aux == 'cosine' else 'nce')
            report[key] = diffs
            print(key, json.dumps(diffs))
            assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
            np.savez(os.path.join(GOLD, 'duet_reverie_%s.npz' % key), **_np(out))
    with open(os.path.join(GOLD, 'duet_reverie_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


def run_duet_soon(args):
    """The model side of the SOON recipe (scripts/run_soon.sh): 2048-d butd object boxes through their own obj_linear /
    obj_layer_norm (models/vilmodel.py:464-468), object-grounding head, NO imagination (imagine_enc_pano off)."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import duet_oracle as O
    ref = build_reference('duet', dict(dataset='soon', obj_feat_size=2048, imagine_enc_pano=False))
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLD, 'duet_soon_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)
    sd = synth.synth_state_dict(manifest, seed=0, gasa_stress=True)
    ref.load_state_dict(sd)
    ep = synth.to_torch(synth.duet_reverie_episode(synth.TINY, 9, obj_dim=2048))
    with torch.no_grad():
        txt = ref('language', {'txt_ids': ep['txt_ids'], 'txt_masks': ep['txt_masks']})
        pano, pano_masks = ref('panorama', {k: ep[k] for k in ('view_img_fts', 'obj_img_fts', 'loc_fts', 'nav_types', 'view_lens', 'obj_lens')})
        nav = ref('navigation', {**{k: ep[k] for k in (
            'txt_masks', 'gmap_img_embeds', 'gmap_step_ids', 'gmap_pos_fts', 'gmap_masks', 'gmap_pair_dists', 'gmap_visited_masks',
            'gmap_vpids', 'vp_img_embeds', 'vp_pos_fts', 'vp_masks', 'vp_nav_masks', 'vp_obj_masks', 'vp_cand_vpids')}, 'txt_embeds': txt})
        o_txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
        o_pano, o_pm = O.forward_panorama(sd, ep['view_img_fts'], ep['loc_fts'], ep['nav_types'], ep['view_lens'],
                                          obj_img_fts=ep['obj_img_fts'], obj_lens=ep['obj_lens'])
        B = ep['txt_ids'].shape[0]
        o_nav = O.forward_navigation(
            sd, o_txt, ep['txt_masks'], ep['gmap_img_embeds'], ep['gmap_step_ids'], ep['gmap_pos_fts'], ep['gmap_masks'],
            ep['gmap_pair_dists'], ep['gmap_visited_masks'], ep['gmap_vpids'], ep['vp_img_embeds'], ep['vp_pos_fts'], ep['vp_masks'],
            ep['vp_nav_masks'], ep['vp_cand_vpids'], torch.zeros(B, 0, 768), torch.zeros(B, 0, dtype=torch.bool),
            vp_obj_masks=ep['vp_obj_masks'])
    diffs = {'pano': maxdiff(pano, o_pano), 'fused': maxdiff(nav['fused_logits'], o_nav['fused_logits']),
             'obj': maxdiff(nav['obj_logits'], o_nav['obj_logits']), 'vp': maxdiff(nav['vp_embeds'], o_nav['vp_embeds'])}
    print('soon', json.dumps(diffs))
    assert torch.equal(pano_masks, o_pm) and max(diffs.values()) < 2e-4
    np.savez(os.path.join(GOLD, 'duet_soon_tiny.npz'), **_np(dict(
        pano_embeds=pano, pano_masks=pano_masks, vp_embeds=nav['vp_embeds'], fused_logits=nav['fused_logits'],
        local_logits=nav['local_logits'], global_logits=nav['global_logits'], obj_logits=nav['obj_logits'])))
    with open(os.path.join(GOLD, 'duet_soon_oracle_vs_reference.json'), 'w') as f:
        json.dump({'tiny': diffs}, f, indent=1)


def run_duet_pretrain(args):
    """DUET pre-training forward (R2R recipe: mlm + mrc + sap, VLN-DUET/pretrain_src/config/r2r_*.json) - groundwork for SURVEY 8(f) N4."""
    from importlib import import_module
    from transformers import BertConfig
    synth = import_module('vln_imagine_b200.synth')
    from oracle import pretrain_oracle as P
    _shim_transformers()
    src = os.path.join(REF, 'VLN-DUET', 'pretrain_src')
    sys.path.insert(0, src)
    from model.pretrain_cmt import GlocalTextPathCMTPreTraining as Net

    def tie(self, *a, **kw):          # transformers >= 5 calls tie_weights(recompute_mapping=...); _tie_or_clone_weights is gone
        if 'mlm' in self.config.pretrain_tasks:
            self.mlm_head.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight
    Net.tie_weights = tie
    cfg = BertConfig()
    for k, v in json.load(open(os.path.join(src, 'config', 'r2r_model_config.json'))).items():
        setattr(cfg, k, v)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    cfg.pretrain_tasks = ['mlm', 'mrc', 'sap']
    torch.manual_seed(0)
    ref = Net(cfg).eval()
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLD, 'duet_pretrain_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)
    sd = synth.synth_state_dict(manifest, seed=0, gasa_stress=False)
    sd['mlm_head.predictions.decoder.weight'] = sd['bert.embeddings.word_embeddings.weight']        # tied
    sd['bert.global_encoder.sprel_linear.weight'] = torch.full((1, 1), -0.3)
    ref.load_state_dict(sd)
    ep = synth.to_torch(synth.duet_pretrain_batch())
    batch = dict(ep)
    with torch.no_grad():
        gl, ll, fl, _, _ = ref(batch, 'sap', compute_loss=False)
        scores = ref(batch, 'mlm', compute_loss=False)
        mlm_loss = ref(batch, 'mlm', compute_loss=True)
        logits, targets, _, _ = ref(batch, 'mrc', compute_loss=False)
        mrc_loss = ref(batch, 'mrc', compute_loss=True)
        o_gl, o_ll, o_fl = P.forward_sap(sd, ep)
        o_scores = P.forward_mlm(sd, ep)
        o_logits, o_targets = P.forward_mrc(sd, ep)
        o_mlm, o_mrc = P.losses(sd, ep)
    diffs = {'global': maxdiff(gl, o_gl), 'local': maxdiff(ll, o_ll), 'fused': maxdiff(fl, o_fl), 'mlm_scores': maxdiff(scores, o_scores),
             'mlm_loss': maxdiff(mlm_loss, o_mlm), 'mrc_logits': maxdiff(logits, o_logits), 'mrc_targets': maxdiff(targets, o_targets),
             'mrc_loss': maxdiff(mrc_loss, o_mrc)}
    print('duet_pretrain', json.dumps(diffs))
    assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
    np.savez_compressed(os.path.join(GOLD, 'duet_pretrain.npz'), **_np(dict(
        global_logits=gl, local_logits=ll, fused_logits=fl, mlm_scores=scores[:, ::64].contiguous(), mlm_loss=mlm_loss,
        mrc_logits=logits, mrc_loss=mrc_loss)))
    with open(os.path.join(GOLD, 'duet_pretrain_oracle_vs_reference.json'), 'w') as f:
        json.dump(diffs, f, indent=1)


def run_hamt_actpred(args):
    """act_pred_token variants of HAMT's action head (r2r/parser.py:67, models/vilmodel_cmt.py:1189-1199) with the imagination
    tokens on either stream: only the logits change."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import hamt_oracle as O
    out, report = {}, {}
    for concat in ('language', 'visual'):
        for tok in ('ob', 'ob_hist', 'ob_txt_hist', 'ob_imagine_text'):
            ref = build_reference('hamt', dict(act_pred_token=tok, concat_imagine_with=concat))
            manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
            sd = synth.synth_state_dict(manifest, seed=0)
            ref.load_state_dict(sd)
            for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
                ep = synth.to_torch(synth.hamt_episode(shape, seed))
                with torch.no_grad():
                    txt = ref('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
                    img = ref('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=None)
                    hm = O.hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
                    logits = ref('visual', txt_embeds=txt, txt_masks=ep['txt_masks'], hist_embeds=ep['hist_embeds'], hist_masks=hm,
                                 ob_img_feats=ep['ob_img_feats'], ob_ang_feats=ep['ob_ang_feats'], ob_nav_types=ep['ob_nav_types'],
                                 ob_masks=ep['ob_masks'], imagine_embeds=img, imagine_masks=ep['imagine_masks'])[0]
                    o_txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
                    o_img = O.forward_imagination(sd, ep['imagine_feats'])
                    o_logits = O.forward_visual(sd, o_txt, ep['txt_masks'], ep['hist_embeds'], hm, ep['ob_img_feats'], ep['ob_ang_feats'],
                                                ep['ob_nav_types'], ep['ob_masks'], o_img, ep['imagine_masks'],
                                                concat_imagine_with=concat, act_pred_token=tok)[0]
                key = '%s_%s_%s' % (tag, concat, tok)
                report[key] = maxdiff(logits, o_logits)
                assert report[key] < 2e-4, key
                out[key] = logits
            print(concat, tok, 'ok')
    np.savez(os.path.join(GOLD, 'hamt_actpred.npz'), **_np(out))
    with open(os.path.join(GOLD, 'hamt_actpred_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


def run_hamt_margin(args):
    """HAMT-Imagine alignment loss with aux_loss_type 'constrastive-margin' (sic, r2r/parser.py:117)."""
    from importlib import import_module
    synth = import_module('vln_imagine_b200.synth')
    from oracle import duet_oracle as D
    from oracle import hamt_oracle as O
    ref = build_reference('hamt', dict(aux_loss_type='constrastive-margin', contrastive_margin_value=0.5))
    manifest = {k: list(v.shape) for k, v in ref.state_dict().items()}
    assert manifest == json.load(open(os.path.join(GOLD, 'hamt_manifest.json'))), 'the margin variant shares the parameter tree'
    sd = synth.synth_state_dict(manifest, seed=0)
    ref.load_state_dict(sd)
    report = {}
    for tag, shape, seed in [('tiny', synth.TINY, 7), ('cfg1', synth.CFG1, 1234)]:
        ep = synth.to_torch(synth.hamt_episode(shape, seed))
        with torch.no_grad():
            txt = ref('language', txt_ids=ep['txt_ids'], txt_masks=ep['txt_masks'])
            img = ref('imagine', imagine_pano_img_feats=ep['imagine_feats'], imagine_masks=None)
            loss, img2 = ref('align_with_contrastive_loss', align_txt_embeds=txt, txt_masks=ep['txt_masks'],
                             align_imagine_embeds=img.clone(), imagine_masks=ep['imagine_masks'],
                             sub_instr_segs=ep['sub_instr_segs'], sub_instr_imag_flag=ep['sub_instr_imag_flag'],
                             noun_phrase_segs=ep['noun_phrase_segs'], obs_instr_ids=ep['obs_instr_ids'])
            o_txt = O.forward_text(sd, ep['txt_ids'], ep['txt_masks'])
            o_img = O.forward_imagination(sd, ep['imagine_feats'])
            o_loss, o_img2 = D.forward_align_margin(sd, o_txt, o_img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'], 0.5)
        diffs = {'loss': abs(float(loss) - float(o_loss)), 'img2': maxdiff(img2, o_img2)}
        report[tag] = diffs
        print(tag, json.dumps(diffs), float(loss))
        assert max(diffs.values()) < 2e-4, 'oracle does not reproduce the reference'
        np.savez(os.path.join(GOLD, 'hamt_margin_%s.npz' % tag), **_np(dict(margin_loss=loss, margin_imagine_embeds=_sub(img2))))
    with open(os.path.join(GOLD, 'hamt_margin_oracle_vs_reference.json'), 'w') as f:
        json.dump(report, f, indent=1)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', choices=['duet', 'duet_variants', 'hamt', 'hamt_variants', 'hamt_encvis', 'hamt_margin', 'hamt_actpred', 'duet_reverie', 'duet_soon', 'duet_pretrain'], required=True)
    ap.add_argument('--grads', action='store_true', help='write the gradient fixtures (cfg-4) instead')
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if a.grads:
        {'duet': run_duet_grads, 'duet_variants': run_duet_variant_grads, 'hamt': run_hamt_grads, 'hamt_variants': run_hamt_variant_grads,
         'duet_reverie': run_duet_reverie_grads}[a.model](a)
    else:
        {'duet': run_duet, 'hamt': run_hamt, 'hamt_encvis': run_hamt_encvis, 'hamt_margin': run_hamt_margin, 'hamt_actpred': run_hamt_actpred, 'duet_reverie': run_duet_reverie, 'duet_soon': run_duet_soon, 'duet_pretrain': run_duet_pretrain}[a.model](a)
