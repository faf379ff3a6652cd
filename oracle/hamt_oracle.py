"""TEST INFRASTRUCTURE - CPU oracle for the HAMT-Imagine (NavCMT) navigation hot path.

Plain fp32 functional PyTorch restatement of VLN-HAMT/finetune_src/models/vilmodel_cmt.py
(paths below are relative to VLN-HAMT/finetune_src/).  Same rules as duet_oracle.py: pinned
against outputs of the real reference via oracle/gen_golden.py -> tests/golden/; imported only
by tests/, smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .duet_oracle import (lin, lnorm, neg_mask, bert_layer, bert_attention, bert_ffn,
                          cross_attention, forward_align_cosine, forward_imagination)  # same block maths


def forward_text(sd, txt_ids, txt_masks, num_l_layers=9):
    """mode 'language'.  models/vilmodel_cmt.py:1008-1031 (BertEmbeddings :44-73 + encoder.layer)."""
    B, L = txt_ids.shape
    pos = torch.arange(L)[None, :].expand(B, L)
    e = (F.embedding(txt_ids, sd['embeddings.word_embeddings.weight'])
         + F.embedding(pos, sd['embeddings.position_embeddings.weight'])
         + sd['embeddings.token_type_embeddings.weight'][0])
    x = lnorm(sd, 'embeddings.LayerNorm', e, 1e-12)
    m = neg_mask(txt_masks)
    for i in range(num_l_layers):
        x = bert_layer(sd, 'encoder.layer.%d' % i, x, m)
    return x


def forward_history(sd, hist_img_feats, hist_ang_feats, ob_step, hist_pano_img_feats,
                    hist_pano_ang_feats, num_h_pano_layers=2, batch_size=1):
    """mode 'history'.  HistoryEmbeddings.forward, models/vilmodel_cmt.py:576-618.
    With no features returns the [cls] history token (batch_size, 768)."""
    p = 'hist_embeddings'
    type_emb = sd[p + '.type_embedding.weight'][0]
    if hist_img_feats is None:
        cls = sd[p + '.cls_token'][0, 0] + type_emb
        return lnorm(sd, p + '.layer_norm', cls[None].expand(batch_size, -1), 1e-12)
    e = (lnorm(sd, p + '.img_layer_norm', lin(sd, p + '.img_linear', hist_img_feats), 1e-12)
         + lnorm(sd, p + '.ang_layer_norm', lin(sd, p + '.ang_linear', hist_ang_feats), 1e-12)
         + sd[p + '.position_embeddings.weight'][ob_step]
         + type_emb)
    pe = (lnorm(sd, p + '.pano_img_layer_norm', lin(sd, p + '.pano_img_linear', hist_pano_img_feats), 1e-12)
          + lnorm(sd, p + '.pano_ang_layer_norm', lin(sd, p + '.pano_ang_linear', hist_pano_ang_feats), 1e-12))
    for i in range(num_h_pano_layers):                 # post-LN BertEncoder, all-ones mask
        pe = bert_layer(sd, '%s.pano_encoder.layer.%d' % (p, i), pe, None)
    e = e + pe.mean(1)
    return lnorm(sd, p + '.layer_norm', e, 1e-12)


def observation_embeddings(sd, ob_img_feats, ob_ang_feats, ob_nav_types):
    """ImageEmbeddings.forward with token type 1.  models/vilmodel_cmt.py:521-544, :1073-1077."""
    p = 'img_embeddings'
    e = (lnorm(sd, p + '.img_layer_norm', lin(sd, p + '.img_linear', ob_img_feats), 1e-12)
         + lnorm(sd, p + '.ang_layer_norm', lin(sd, p + '.ang_linear', ob_ang_feats), 1e-12)
         + sd['embeddings.token_type_embeddings.weight'][1]
         + F.embedding(ob_nav_types, sd[p + '.nav_type_embedding.weight']))
    return lnorm(sd, p + '.layer_norm', e, 1e-12)


def lxrt_x_layer(sd, p, lang, lang_add, visn, visn_add):
    """LXRTXLayer.forward, models/vilmodel_cmt.py:423-445: both cross-attention directions use
    the SAME visual_attention weights and read the layer inputs (:385-397); then per-stream
    self-attention (:399-407) and FFN (:409-421)."""
    lang_x = cross_attention(sd, p + '.visual_attention', lang, visn, visn_add)
    visn_x = cross_attention(sd, p + '.visual_attention', visn, lang, lang_add)
    lang_s = bert_attention(sd, p + '.lang_self_att', lang_x, lang_add)
    visn_s = bert_attention(sd, p + '.visn_self_att', visn_x, visn_add)
    lang_o = bert_ffn(sd, p + '.lang_inter', p + '.lang_output', lang_s)
    visn_o = bert_ffn(sd, p + '.visn_inter', p + '.visn_output', visn_s)
    return lang_o, visn_o


def next_action(sd, x):
    """NextActionPrediction: Linear -> ReLU -> LN -> (dropout) -> Linear(.,1).  :953-963."""
    h = torch.relu(lin(sd, 'next_action.net.0', x))
    return lin(sd, 'next_action.net.4', lnorm(sd, 'next_action.net.2', h, 1e-12))


def forward_imagination_encoder(sd, imagine_feats, imagine_masks, num_layers=2):
    """mode 'imagine' with bypass_imag_encoder=False: ImagineEmbeddings.forward, models/vilmodel_cmt.py:634-703
    (position + type embedding on the raw features, Linear + LN, a post-LN BertEncoder over the imaginations of an
    episode with the additive -10000 mask, final LN)."""
    p = 'imagine_embeddings'
    n = imagine_feats.shape[1]
    x = imagine_feats + sd[p + '.position_embeddings.weight'][:n][None] + sd[p + '.type_embedding.weight'][0]
    x = lnorm(sd, p + '.pano_img_layer_norm', lin(sd, p + '.pano_img_linear', x), 1e-12)
    m = neg_mask(imagine_masks)
    for i in range(num_layers):
        x = bert_layer(sd, '%s.pano_encoder.layer.%d' % (p, i), x, m)
    return lnorm(sd, p + '.layer_norm', x, 1e-12)


def forward_visual(sd, txt_embeds, txt_masks, hist_embeds, hist_masks, ob_img_feats, ob_ang_feats,
                   ob_nav_types, ob_masks, imagine_embeds, imagine_masks, num_x_layers=4, concat_imagine_with='language',
                   act_pred_token='ob_txt'):
    """mode 'visual', act_pred_token='ob_txt', no_lang_ca=False; the imagination tokens ride on the language stream
    (concat_imagine_with='language', the released recipe) or on the vision stream ('visual', the parser default).
    models/vilmodel_cmt.py:1056-1205."""
    ob = observation_embeddings(sd, ob_img_feats, ob_ang_feats, ob_nav_types)
    n_hist, n_ob = hist_embeds.shape[1], ob.shape[1]
    visn = torch.cat([hist_embeds, ob], 1)
    visn_add = torch.cat([neg_mask(hist_masks), neg_mask(ob_masks)], -1)
    L = txt_embeds.shape[1]
    lang, lang_add = txt_embeds, neg_mask(txt_masks)
    if concat_imagine_with == 'language':                       # :1109-1112
        lang = torch.cat([txt_embeds, imagine_embeds], 1)
        lang_add = torch.cat([lang_add, neg_mask(imagine_masks)], -1)
    else:                                                       # :1106-1108
        visn = torch.cat([visn, imagine_embeds], 1)
        visn_add = torch.cat([visn_add, neg_mask(imagine_masks)], -1)
    for i in range(num_x_layers):
        lang, visn = lxrt_x_layer(sd, 'encoder.x_layers.%d' % i, lang, lang_add, visn, visn_add)
    hist_out, ob_out = visn[:, :n_hist], visn[:, n_hist:n_hist + n_ob]      # :1173-1182
    txt_out = lang[:, :L]
    imagine_out = lang[:, L:] if concat_imagine_with == 'language' else visn[:, n_hist + n_ob:]
    gate = {'ob_txt': lambda: txt_out[:, :1], 'ob_hist': lambda: hist_out[:, :1],                  # :1189-1199
            'ob_txt_hist': lambda: txt_out[:, :1] + hist_out[:, :1],
            'ob_imagine_text': lambda: txt_out[:, :1] + imagine_out.mean(1).unsqueeze(1)}
    pred_in = ob_out if act_pred_token == 'ob' else ob_out * gate[act_pred_token]()
    logits = next_action(sd, pred_in).squeeze(-1)
    logits = logits.masked_fill(ob_nav_types == 0, float('-inf'))
    return logits, txt_out, hist_out, ob_out


def hist_masks_from_lens(hist_lens, size):
    """length2mask(...).logical_not().  utils/misc.py:12-17, models/model_HAMT.py:62-63."""
    return torch.arange(size)[None, :] < torch.as_tensor(hist_lens)[:, None]


def nav_step(sd, ep, txt_embeds, imagine_embeds):
    """One decision: 'visual' then 'history' (r2r/agent_cmt.py:538,604)."""
    hm = hist_masks_from_lens(ep['hist_lens'], ep['hist_embeds'].shape[1])
    logits, txt_o, hist_o, ob_o = forward_visual(
        sd, txt_embeds, ep['txt_masks'], ep['hist_embeds'], hm, ep['ob_img_feats'], ep['ob_ang_feats'],
        ep['ob_nav_types'], ep['ob_masks'], imagine_embeds, ep['imagine_masks'])
    h = forward_history(sd, ep['hist_img_feats'], ep['hist_ang_feats'], ep['ob_step'],
                        ep['hist_pano_img_feats'], ep['hist_pano_ang_feats'])
    return logits, txt_o, hist_o, ob_o, h


def episode_prelude(sd, ep):
    txt = forward_text(sd, ep['txt_ids'], ep['txt_masks'])
    img = forward_imagination(sd, ep['imagine_feats'])
    loss, img2 = forward_align_cosine(sd, txt, img, ep['sub_instr_imag_flag'], ep['noun_phrase_segs'])
    return txt, img, loss, img2
