"""TEST INFRASTRUCTURE - CPU oracle for the DUET-Imagine navigation hot path.

A plain fp32 restatement (functional PyTorch on CPU, no nn.Module, no custom kernels) of what
the reference's GlocalTextPathNavCMT computes, written from the reference's algorithm:
every function cites the reference file:line it follows (paths relative to
VLN-DUET/map_nav_src/).  It consumes a reference-layout ``state_dict`` so the product, the
oracle and the real reference all share weights by name.

Pinning: the reference ships no golden vectors, so this oracle is pinned
against outputs of the reference itself, generated in the build container by
``oracle/gen_golden.py`` (which imports /root/reference) and committed under tests/golden/.
``tests/test_oracle_golden.py`` replays them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (vln-imagine_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
NUM_HEADS = 12


# ---------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------

def lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def lnorm(sd: SD, p: str, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], eps)


def gelu_erf(x):
    """models/vilmodel.py:32-38."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def neg_mask(masks):
    """(N,L) bool -> (N,1,1,L) additive -10000 mask.  models/ops.py:25-34."""
    return (1.0 - masks[:, None, None, :].to(torch.float32)) * -10000.0


def split_heads(x, nh=NUM_HEADS):
    B, L, D = x.shape
    return x.view(B, L, nh, D // nh).permute(0, 2, 1, 3)


def merge_heads(x):
    B, H, L, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, L, H * d)


def attend(q, k, v, add_mask):
    """softmax(q k^T / sqrt(dh) + mask) v.  models/vilmodel.py:118-134 and :336-349."""
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if add_mask is not None:
        s = s + add_mask
    return torch.matmul(torch.softmax(s, -1), v)


# ---------------------------------------------------------------------------------------------
# BERT blocks
# ---------------------------------------------------------------------------------------------

def bert_attention(sd, p, x, add_mask, eps=1e-12):
    """BertAttention = BertSelfAttention + BertSelfOutput.  models/vilmodel.py:80-167."""
    q = split_heads(lin(sd, p + '.self.query', x))
    k = split_heads(lin(sd, p + '.self.key', x))
    v = split_heads(lin(sd, p + '.self.value', x))
    ctx = merge_heads(attend(q, k, v, add_mask))
    return lnorm(sd, p + '.output.LayerNorm', lin(sd, p + '.output.dense', ctx) + x, eps)


def bert_ffn(sd, p_inter, p_out, x, eps=1e-12):
    """BertIntermediate + BertOutput.  models/vilmodel.py:169-194."""
    h = gelu_erf(lin(sd, p_inter + '.dense', x))
    return lnorm(sd, p_out + '.LayerNorm', lin(sd, p_out + '.dense', h) + x, eps)


def bert_layer(sd, p, x, add_mask):
    """models/vilmodel.py:196-209."""
    a = bert_attention(sd, p + '.attention', x, add_mask)
    return bert_ffn(sd, p + '.intermediate', p + '.output', a)


def cross_attention(sd, p, x, ctx, ctx_add_mask, eps=1e-12):
    """BertXAttention: queries from x, keys/values from ctx.  models/vilmodel.py:302-364."""
    q = split_heads(lin(sd, p + '.att.query', x))
    k = split_heads(lin(sd, p + '.att.key', ctx))
    v = split_heads(lin(sd, p + '.att.value', ctx))
    o = merge_heads(attend(q, k, v, ctx_add_mask))
    return lnorm(sd, p + '.output.LayerNorm', lin(sd, p + '.output.dense', o) + x, eps)


def graph_x_layer(sd, p, ctx, ctx_add_mask, visn, visn_add_mask, sprels):
    """GraphLXRTXLayer.forward.  models/vilmodel.py:384-399."""
    v = cross_attention(sd, p + '.visual_attention', visn, ctx, ctx_add_mask)
    m = visn_add_mask if sprels is None else visn_add_mask + sprels
    v = bert_attention(sd, p + '.visn_self_att', v, m)
    return bert_ffn(sd, p + '.visn_inter', p + '.visn_output', v)


def crossmodal_encoder(sd, p, n_layers, ctx, ctx_masks, visn, visn_masks, sprels=None):
    """CrossmodalEncoder.forward.  models/vilmodel.py:444-453."""
    cm, vm = neg_mask(ctx_masks), neg_mask(visn_masks)
    for i in range(n_layers):
        visn = graph_x_layer(sd, '%s.x_layers.%d' % (p, i), ctx, cm, visn, vm, sprels)
    return visn


def cls_prediction(sd, p, x):
    """ClsPrediction: Linear -> ReLU -> LN(1e-12) -> Linear(.,1).  models/vilmodel.py:1009-1020."""
    h = torch.relu(lin(sd, p + '.net.0', x))
    return lin(sd, p + '.net.3', lnorm(sd, p + '.net.2', h, 1e-12))


# ---------------------------------------------------------------------------------------------
# modes
# ---------------------------------------------------------------------------------------------

def forward_text(sd, txt_ids, txt_masks, num_l_layers=9):
    """mode 'language'.  BertEmbeddings (models/vilmodel.py:49-78) + LanguageEncoder (:414-434);
    entry at :1075-1079."""
    B, L = txt_ids.shape
    pos = torch.arange(L, device=txt_ids.device)[None, :].expand(B, L)
    e = (F.embedding(txt_ids, sd['embeddings.word_embeddings.weight'])
         + F.embedding(pos, sd['embeddings.position_embeddings.weight'])
         + sd['embeddings.token_type_embeddings.weight'][0])
    x = lnorm(sd, 'embeddings.LayerNorm', e, 1e-12)
    m = neg_mask(txt_masks)
    for i in range(num_l_layers):
        x = bert_layer(sd, 'lang_encoder.layer.%d' % i, x, m)
    return x


def forward_imagination(sd, imagine_feats):
    """mode 'imagine' with bypass_imag_encoder: features + type embedding 0.
    models/vilmodel.py:562-573, :1081-1085."""
    return imagine_feats + sd['imagine_embeddings.type_embedding.weight'][0]


def pano_encoder_layer(sd, p, x, key_pad):
    """Pre-norm TransformerEncoderLayer with nn.MultiheadAttention (packed in_proj, bool
    key_padding_mask -> -inf) and F.gelu; LN eps 1e-5.  models/transformer.py:170-182."""
    h = lnorm(sd, p + '.norm1', x, 1e-5)
    qkv = F.linear(h, sd[p + '.self_attn.in_proj_weight'], sd[p + '.self_attn.in_proj_bias'])
    q, k, v = [split_heads(t) for t in qkv.chunk(3, -1)]
    add = torch.zeros(key_pad.shape, dtype=torch.float32, device=key_pad.device).masked_fill(key_pad, float('-inf'))[:, None, None, :]
    a = merge_heads(attend(q, k, v, add))
    x = x + lin(sd, p + '.self_attn.out_proj', a)
    h = lnorm(sd, p + '.norm2', x, 1e-5)
    return x + lin(sd, p + '.linear2', F.gelu(lin(sd, p + '.linear1', h)))


def forward_panorama(sd, view_img_fts, loc_fts, nav_types, view_lens, num_pano_layers=2, obj_img_fts=None, obj_lens=None):
    """mode 'panorama'.  models/vilmodel.py:1087-1131.  R2R has no object branch (obj_feat_size=0); REVERIE appends the
    object boxes of a panorama behind its views, per episode, zero padded to the longest (:1096-1114)."""
    p = 'img_embeddings'
    img = lnorm(sd, p + '.img_layer_norm', lin(sd, p + '.img_linear', view_img_fts), 1e-12)
    pano_lens = view_lens
    if obj_img_fts is not None:
        if (p + '.obj_linear.weight') in sd:                    # obj_feat_size != image_feat_size (:464-468)
            obj = lnorm(sd, p + '.obj_layer_norm', lin(sd, p + '.obj_linear', obj_img_fts), 1e-12)
        else:
            obj = lnorm(sd, p + '.img_layer_norm', lin(sd, p + '.img_linear', obj_img_fts), 1e-12)
        pano_lens = view_lens + obj_lens
        rows = torch.zeros(img.shape[0], int(pano_lens.max()), img.shape[2])
        for b in range(img.shape[0]):
            vl, ol = int(view_lens[b]), int(obj_lens[b])
            rows[b, :vl] = img[b, :vl]
            rows[b, vl:vl + ol] = obj[b, :ol]
        img = rows
    e = (img
         + lnorm(sd, p + '.loc_layer_norm', lin(sd, p + '.loc_linear', loc_fts), 1e-12)
         + F.embedding(nav_types, sd[p + '.nav_type_embedding.weight'])
         + sd['embeddings.token_type_embeddings.weight'][1])
    x = lnorm(sd, p + '.layer_norm', e, 1e-12)
    V = img.shape[1]
    pano_masks = torch.arange(V, device=pano_lens.device)[None, :] < pano_lens[:, None]      # models/ops.py:36-44
    pad = ~pano_masks
    for i in range(num_pano_layers):
        x = pano_encoder_layer(sd, '%s.pano_encoder.layers.%d' % (p, i), x, pad)
    x = lnorm(sd, p + '.pano_encoder.norm', x, 1e-12)               # models/ops.py:19-23
    return x, pano_masks


def fuse_logits(global_logits, local_logits, gmap_vpids, gmap_visited_masks, vp_cand_vpids):
    """Global/local action fusion.  models/vilmodel.py:1198-1217."""
    fused = global_logits.clone()
    fused[:, 0] += local_logits[:, 0]
    for i in range(global_logits.shape[0]):
        visited = set(vp for vp, m in zip(gmap_vpids[i], gmap_visited_masks[i].tolist()) if m)
        by_id, backtrack = {}, 0
        for j, vp in enumerate(vp_cand_vpids[i]):
            if j > 0:
                if vp in visited:
                    backtrack = backtrack + local_logits[i, j]
                else:
                    by_id[vp] = local_logits[i, j]
        for j, vp in enumerate(gmap_vpids[i]):
            if j > 0 and vp not in visited:
                fused[i, j] += by_id[vp] if vp in by_id else backtrack
    return fused


def forward_navigation(sd, txt_embeds, txt_masks, gmap_img_embeds, gmap_step_ids, gmap_pos_fts,
                       gmap_masks, gmap_pair_dists, gmap_visited_masks, gmap_vpids,
                       vp_img_embeds, vp_pos_fts, vp_masks, vp_nav_masks, vp_cand_vpids,
                       imagine_embeds, imagine_masks, num_x_layers=4, vp_obj_masks=None):
    """mode 'navigation' with imaginations concatenated to the text stream.
    models/vilmodel.py:1133-1235; vp_obj_masks (REVERIE) adds the object-grounding logits (:1220-1225)."""
    g = 'global_encoder'
    gmap = (gmap_img_embeds
            + F.embedding(gmap_step_ids, sd[g + '.gmap_step_embeddings.weight'])
            + lnorm(sd, g + '.gmap_pos_embeddings.1', lin(sd, g + '.gmap_pos_embeddings.0', gmap_pos_fts), 1e-12))
    w, b = sd[g + '.sprel_linear.weight'].reshape(()), sd[g + '.sprel_linear.bias'].reshape(())
    sprels = (gmap_pair_dists * w + b)[:, None]                       # (B,1,G,G)  :1145-1149
    l = 'local_encoder'
    vp = vp_img_embeds + lnorm(sd, l + '.vp_pos_embeddings.1', lin(sd, l + '.vp_pos_embeddings.0', vp_pos_fts), 1e-12)

    ctx = torch.cat([txt_embeds, imagine_embeds], 1)                  # :1157-1158
    ctx_masks = torch.cat([txt_masks, imagine_masks], 1)
    gmap = crossmodal_encoder(sd, g + '.encoder', num_x_layers, ctx, ctx_masks, gmap, gmap_masks, sprels)
    vp = crossmodal_encoder(sd, l + '.encoder', num_x_layers, ctx, ctx_masks, vp, vp_masks)

    fuse_w = torch.sigmoid(cls_prediction(sd, 'sap_fuse_linear', torch.cat([gmap[:, 0], vp[:, 0]], 1)))
    global_logits = cls_prediction(sd, 'global_sap_head', gmap).squeeze(2) * fuse_w
    global_logits = global_logits.masked_fill(gmap_visited_masks, float('-inf'))
    global_logits = global_logits.masked_fill(~gmap_masks, float('-inf'))
    local_logits = cls_prediction(sd, 'local_sap_head', vp).squeeze(2) * (1 - fuse_w)
    local_logits = local_logits.masked_fill(~vp_nav_masks, float('-inf'))
    fused = fuse_logits(global_logits, local_logits, gmap_vpids, gmap_visited_masks, vp_cand_vpids)
    obj_logits = None
    if vp_obj_masks is not None:
        obj_logits = cls_prediction(sd, 'og_head', vp).squeeze(2).masked_fill(~vp_obj_masks, float('-inf'))
    return {'gmap_embeds': gmap, 'vp_embeds': vp, 'global_logits': global_logits,
            'local_logits': local_logits, 'fused_logits': fused, 'obj_logits': obj_logits}


def mlp_projection(sd, p, x):
    """MLPProjectionHead without dropout: 768->512->512->768, no bias, ReLU.
    models/vilmodel.py:575-589."""
    h = torch.relu(F.linear(x, sd[p + '.fc1.weight']))
    h = torch.relu(F.linear(h, sd[p + '.fc2.weight']))
    return F.linear(h, sd[p + '.fc3.weight'])


def _noun_phrase_rows(sub_instr_imag_flag, noun_phrase_segs):
    """Rows (b, i, token-index list) that take part in the alignment loss."""
    rows = []
    for b, flags in enumerate(sub_instr_imag_flag):
        for i, f in enumerate(flags):
            if f != 'True':
                continue
            toks: List[int] = []
            for s, e in noun_phrase_segs[b][i]:
                toks.extend(range(s, e + 1))
            rows.append((b, i, toks, len(noun_phrase_segs[b][i])))
    return rows


def forward_align_cosine(sd, align_txt_embeds, align_imagine_embeds, sub_instr_imag_flag,
                         noun_phrase_segs):
    """mode 'align_with_contrastive_loss', aux_loss_type 'cosine'.
    models/vilmodel.py:598-655: every flagged imagination is projected; those whose
    sub-instruction has >=1 noun phrase are overwritten by their projection and contribute
    1-cos(proj, mean noun-phrase token embedding) (eps 1e-8); the loss is the mean."""
    out = align_imagine_embeds.clone()
    losses = []
    p = 'contrastive_alignment_model.image_proj'
    for b, i, toks, n_np in _noun_phrase_rows(sub_instr_imag_flag, noun_phrase_segs):
        proj = mlp_projection(sd, p, align_imagine_embeds[b, i])
        if n_np > 0:
            t = align_txt_embeds[b, toks].mean(0)
            out[b, i] = proj
            losses.append(1 - F.cosine_similarity(proj, t, dim=-1))
    loss = torch.stack(losses).mean() if losses else torch.zeros(())
    return loss, out


def forward_align_infonce(sd, align_txt_embeds, align_imagine_embeds, sub_instr_imag_flag,
                          noun_phrase_segs, temperature):
    """aux_loss_type 'contrastive-InfoNCE'.  models/vilmodel.py:657-779: negatives are the
    per-noun-phrase mean embeddings of every *other* episode in the batch (only noun phrases
    of flagged sub-instructions); loss_i = CE(cos(proj,[pos;negs])/T, 0)."""
    B = align_imagine_embeds.shape[0]
    per_ep: List[List[torch.Tensor]] = [[] for _ in range(B)]
    for b, flags in enumerate(sub_instr_imag_flag):
        for i, f in enumerate(flags):
            if f != 'True':
                continue
            for s, e in noun_phrase_segs[b][i]:
                if e >= s:
                    per_ep[b].append(align_txt_embeds[b, s:e + 1].mean(0))
    out = align_imagine_embeds.clone()
    losses = []
    p = 'contrastive_alignment_model.image_proj'
    for b, i, toks, n_np in _noun_phrase_rows(sub_instr_imag_flag, noun_phrase_segs):
        proj = mlp_projection(sd, p, align_imagine_embeds[b, i])
        if n_np > 0:
            pos = align_txt_embeds[b, toks].mean(0)
            negs = [t for bb in range(B) if bb != b for t in per_ep[bb]]
            allt = torch.stack([pos] + negs, 0)
            sim = F.cosine_similarity(proj[None], allt) / temperature
            out[b, i] = proj
            losses.append(F.cross_entropy(sim[None], torch.zeros(1, dtype=torch.long)))
    loss = torch.stack(losses).mean() if losses else torch.zeros(())
    return loss, out


def forward_align_reverie(sd, align_txt_embeds, txt_masks, align_imagine_embeds, aux_loss_type='cosine', temperature=0.007):
    """REVERIE alignment (one imagination per instruction, no sub-instruction annotation): the projected imagination
    against the mean of ALL valid instruction tokens (AlignWithContrastiveLossReverie, models/vilmodel.py:781-828); the
    InfoNCE form takes the instruction means of the OTHER episodes as negatives (:830-888, :657-687)."""
    B = align_imagine_embeds.shape[0]
    p = 'contrastive_alignment_model.image_proj'
    means = [align_txt_embeds[b, txt_masks[b]].mean(0) for b in range(B)]
    out = align_imagine_embeds.clone()
    losses = []
    for b in range(B):
        proj = mlp_projection(sd, p, align_imagine_embeds[b, 0])
        out[b, 0] = proj
        if aux_loss_type == 'cosine':
            losses.append(1 - F.cosine_similarity(proj, means[b], dim=-1))
        else:
            allt = torch.stack([means[b]] + [means[o] for o in range(B) if o != b], 0)
            sim = F.cosine_similarity(proj[None], allt) / temperature
            losses.append(F.cross_entropy(sim[None], torch.zeros(1, dtype=torch.long)))
    return torch.stack(losses).mean(), out


def forward_align_margin(sd, align_txt_embeds, align_imagine_embeds, sub_instr_imag_flag, noun_phrase_segs, margin):
    """aux_loss_type 'constrastive-margin' (HAMT only: VLN-HAMT/finetune_src/models/vilmodel_cmt.py:825-856, 858-950): the
    negatives of forward_align_infonce; loss_i = (1 - cos(proj, pos)) + mean_neg relu(margin + cos(proj, neg) - cos(proj, pos))."""
    B = align_imagine_embeds.shape[0]
    per_ep: List[List[torch.Tensor]] = [[] for _ in range(B)]
    for b, flags in enumerate(sub_instr_imag_flag):
        for i, f in enumerate(flags):
            if f != 'True':
                continue
            for s, e in noun_phrase_segs[b][i]:
                if e >= s:
                    per_ep[b].append(align_txt_embeds[b, s:e + 1].mean(0))
    out = align_imagine_embeds.clone()
    losses = []
    p = 'contrastive_alignment_model.image_proj'
    for b, i, toks, n_np in _noun_phrase_rows(sub_instr_imag_flag, noun_phrase_segs):
        proj = mlp_projection(sd, p, align_imagine_embeds[b, i])
        if n_np > 0:
            pos = align_txt_embeds[b, toks].mean(0)
            negs = torch.stack([t for bb in range(B) if bb != b for t in per_ep[bb]], 0)
            pos_sim = F.cosine_similarity(proj[None], pos[None]).squeeze()
            neg_sims = F.cosine_similarity(proj[None], negs)
            out[b, i] = proj
            losses.append((1 - pos_sim) + torch.relu(margin + neg_sims - pos_sim).mean())
    loss = torch.stack(losses).mean() if losses else torch.zeros(())
    return loss, out


# ---------------------------------------------------------------------------------------------
# whole-episode helpers used by tests / bench
# ---------------------------------------------------------------------------------------------

def nav_step(sd, ep, txt_embeds, imagine_embeds):
    """One decision: 'panorama' followed by 'navigation' on a synthetic episode dict
    (vln-imagine_b200/synth.py: duet_episode, already converted with to_torch)."""
    pano, pano_masks = forward_panorama(sd, ep['view_img_fts'], ep['loc_fts'], ep['nav_types'], ep['view_lens'])
    nav = forward_navigation(
        sd, txt_embeds, ep['txt_masks'], ep['gmap_img_embeds'], ep['gmap_step_ids'], ep['gmap_pos_fts'],
        ep['gmap_masks'], ep['gmap_pair_dists'], ep['gmap_visited_masks'], ep['gmap_vpids'],
        ep['vp_img_embeds'], ep['vp_pos_fts'], ep['vp_masks'], ep['vp_nav_masks'], ep['vp_cand_vpids'],
        imagine_embeds, ep['imagine_masks'])
    return pano, pano_masks, nav


def episode_prelude(sd, ep, aux='cosine', temperature=0.007):
    """Once per episode: 'language' -> 'imagine' -> 'align_with_contrastive_loss'
    (r2r/agent.py:408-449)."""
    txt = forward_text(sd, ep['txt_ids'], ep['txt_masks'])
    img = forward_imagination(sd, ep['imagine_feats'])
    if aux == 'cosine':
        loss, img2 = forward_align_cosine(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'])
    else:
        loss, img2 = forward_align_infonce(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'], temperature)
    return txt, img, loss, img2
