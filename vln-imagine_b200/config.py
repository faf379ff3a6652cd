"""Model configuration, built offline from the agents' ``args`` namespace.

Mirrors ``get_vlnbert_models`` (VLN-DUET/map_nav_src/models/vlnbert_init.py:13-77 and
VLN-HAMT/finetune_src/models/vlnbert_init.py:13-83) without the Hugging Face hub round trip: the
reference downloads the ``bert-base-uncased`` config (no network here); its values are constants
and are inlined below.  ``default_duet_args`` / ``default_hamt_args`` carry the flag values of
the released run scripts (map_nav_src/scripts/run_r2r.sh:28-81, finetune_src/scripts/run_r2r.sh:18-76)
for tests and benchmarks.
"""
from __future__ import annotations

import types

BERT_BASE = dict(hidden_size=768, num_attention_heads=12, intermediate_size=3072, hidden_act='gelu',
                 layer_norm_eps=1e-12, vocab_size=30522, max_position_embeddings=512, type_vocab_size=2,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02)


class ModelConfig(types.SimpleNamespace):
    """Attribute bag like transformers.PretrainedConfig (callers toggle e.g. ``config.imagine_enc_pano``)."""

    def to_dict(self):
        return dict(self.__dict__)


def _get(args, name, default):
    return getattr(args, name, default)


def _check_bert_base(cfg):
    if cfg.hidden_size != 768 or cfg.num_attention_heads != 12 or cfg.intermediate_size != 3072:
        raise ValueError('libvlnimagine kernels are specialised for hidden 768 / 12 heads / FFN 3072 (bert-base)')


def duet_config(args) -> ModelConfig:
    c = ModelConfig(**BERT_BASE)
    c.max_action_steps = 100
    c.image_feat_size = _get(args, 'image_feat_size', 768)
    c.angle_feat_size = _get(args, 'angle_feat_size', 4)
    c.obj_feat_size = _get(args, 'obj_feat_size', 0)
    c.obj_loc_size = 3
    c.num_l_layers = _get(args, 'num_l_layers', 9)
    c.num_pano_layers = _get(args, 'num_pano_layers', 2)
    c.num_x_layers = _get(args, 'num_x_layers', 4)
    c.graph_sprels = _get(args, 'graph_sprels', True)
    c.glocal_fuse = _get(args, 'fusion', 'dynamic') == 'dynamic'
    c.fix_lang_embedding = _get(args, 'fix_lang_embedding', False)
    c.fix_pano_embedding = _get(args, 'fix_pano_embedding', False)
    c.fix_local_branch = _get(args, 'fix_local_branch', False)
    c.update_lang_bert = not c.fix_lang_embedding
    c.output_attentions = True
    c.pred_head_dropout_prob = 0.1
    c.use_lang2visn_attn = False
    c.imagine_enc_pano = _get(args, 'imagine_enc_pano', True)
    if c.imagine_enc_pano:
        c.max_imagination_len = _get(args, 'max_imagination_len', 20)
        c.fix_imagine_embeds = _get(args, 'fix_imagine_embeds', False)
        c.bypass_imag_encoder = _get(args, 'bypass_imag_encoder', True)
        c.use_cosine_aux_loss = _get(args, 'use_cosine_aux_loss', True)
        c.concat_imagine_with = _get(args, 'concat_imagine_with', 'language')
        c.fix_lang_inside_cosine_model = _get(args, 'fix_lang_inside_cosine_model', True)
        c.aux_loss_type = _get(args, 'aux_loss_type', 'cosine')
        c.infonce_temperature = _get(args, 'infonce_temperature', 0.007)
        c.no_loss_test = _get(args, 'no_loss_test', False)
        c.dataset = _get(args, 'dataset', 'r2r')
    _check_bert_base(c)
    return c


def hamt_config(args) -> ModelConfig:
    c = ModelConfig(**BERT_BASE)
    c.image_feat_size = _get(args, 'image_feat_size', 768)
    c.angle_feat_size = _get(args, 'angle_feat_size', 4)
    c.num_l_layers = _get(args, 'num_l_layers', 9)
    c.num_r_layers = 0
    c.num_h_layers = _get(args, 'num_h_layers', 0)
    c.num_x_layers = _get(args, 'num_x_layers', 4)
    c.hist_enc_pano = _get(args, 'hist_enc_pano', True)
    c.num_h_pano_layers = _get(args, 'hist_pano_num_layers', 2)
    c.fix_lang_embedding = _get(args, 'fix_lang_embedding', True)
    c.fix_hist_embedding = _get(args, 'fix_hist_embedding', True)
    c.fix_obs_embedding = _get(args, 'fix_obs_embedding', False)
    c.update_lang_bert = not c.fix_lang_embedding
    c.output_attentions = True
    c.pred_head_dropout_prob = 0.1
    c.no_lang_ca = _get(args, 'no_lang_ca', False)
    c.act_pred_token = _get(args, 'act_pred_token', 'ob_txt')
    c.max_action_steps = 50
    c.imagine_enc_pano = _get(args, 'imagine_enc_pano', True)
    if c.imagine_enc_pano:
        c.max_imagination_len = _get(args, 'max_imagination_len', 20)
        c.fix_imagine_embeds = _get(args, 'fix_imagine_embeds', False)
        c.bypass_imag_encoder = _get(args, 'bypass_imag_encoder', True)
        c.use_cosine_aux_loss = _get(args, 'use_cosine_aux_loss', True)
        c.aux_loss_type = _get(args, 'aux_loss_type', 'cosine')
        c.infonce_temperature = _get(args, 'infonce_temperature', 0.3)
        c.contrastive_margin_value = _get(args, 'contrastive_margin_value', 0.5)
        c.concat_imagine_with = _get(args, 'concat_imagine_with', 'language')
        c.no_loss_test = _get(args, 'no_loss_test', False)
    _check_bert_base(c)
    return c


def default_duet_args(**over):
    a = types.SimpleNamespace(
        bert_ckpt_file=None, tokenizer='bert', image_feat_size=768, angle_feat_size=4, obj_feat_size=0,
        num_l_layers=9, num_pano_layers=2, num_x_layers=4, graph_sprels=True, fusion='dynamic',
        fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False, feat_dropout=0.4, dropout=0.5,
        imagine_enc_pano=True, max_imagination_len=20, fix_imagine_embeds=False, bypass_imag_encoder=True,
        use_cosine_aux_loss=True, concat_imagine_with='language', fix_lang_inside_cosine_model=True,
        aux_loss_type='cosine', infonce_temperature=0.007, no_loss_test=False, dataset='r2r')
    a.__dict__.update(over)
    return a


def default_hamt_args(**over):
    a = types.SimpleNamespace(
        bert_ckpt_file=None, tokenizer='bert', dataset='r2r', image_feat_size=768, angle_feat_size=4,
        num_l_layers=9, num_h_layers=0, num_x_layers=4, hist_enc_pano=True, hist_pano_num_layers=2,
        fix_lang_embedding=True, fix_hist_embedding=True, fix_obs_embedding=False, no_lang_ca=False,
        act_pred_token='ob_txt', feat_dropout=0.4, dropout=0.5, imagine_enc_pano=True, max_imagination_len=20,
        fix_imagine_embeds=False, bypass_imag_encoder=True, use_cosine_aux_loss=True, aux_loss_type='cosine',
        infonce_temperature=0.3, contrastive_margin_value=0.5, concat_imagine_with='language', no_loss_test=False)
    a.__dict__.update(over)
    return a
