"""Parameter trees of the two models, with the reference's module / parameter names.

These modules are *containers only*: they own the fp32 master weights under the names a reference
checkpoint uses (so ``load_state_dict`` of a DUET-Imagine / HAMT-Imagine ``state_dict`` works
verbatim, and ``agent.load`` in the reference keeps working - VLN-DUET/map_nav_src/r2r/agent_base.py:250-282)
but none of their ``forward`` methods is ever called: all arithmetic runs in libvlnimagine kernels
driven by duet.py / hamt.py.  Names follow VLN-DUET/map_nav_src/models/vilmodel.py and
VLN-HAMT/finetune_src/models/vilmodel_cmt.py (see tests/golden/*_manifest.json for the full lists).
"""
from __future__ import annotations

import torch
import torch.nn as nn

H = 768
FF = 3072


class _Container(nn.Module):
    def forward(self, *a, **k):          # pragma: no cover
        raise RuntimeError('%s is a parameter container; the compute path is in libvlnimagine' % type(self).__name__)


class QKV(_Container):
    """.query / .key / .value  (BertSelfAttention :80-142, BertOutAttention :302-353)"""

    def __init__(self):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class DenseLN(_Container):
    """.dense / .LayerNorm  (BertSelfOutput :144-155, BertOutput :183-194)"""

    def __init__(self, d_in=H, eps=1e-12):
        super().__init__()
        self.dense = nn.Linear(d_in, H)
        self.LayerNorm = nn.LayerNorm(H, eps=eps)


class Dense(_Container):
    """.dense  (BertIntermediate :169-181)"""

    def __init__(self, d_in=H, d_out=FF):
        super().__init__()
        self.dense = nn.Linear(d_in, d_out)


class SelfAtt(_Container):
    """BertAttention :157-167  (attribute named ``self`` like the reference)"""

    def __init__(self):
        super().__init__()
        setattr(self, 'self', QKV())
        self.output = DenseLN()


class XAtt(_Container):
    """BertXAttention :355-364"""

    def __init__(self):
        super().__init__()
        self.att = QKV()
        self.output = DenseLN()


class BertLayerP(_Container):
    """BertLayer :196-209"""

    def __init__(self):
        super().__init__()
        self.attention = SelfAtt()
        self.intermediate = Dense()
        self.output = DenseLN(FF)


class LayerStack(_Container):
    def __init__(self, attr, layers):
        super().__init__()
        setattr(self, attr, nn.ModuleList(layers))


class GraphXLayerP(_Container):
    """GraphLXRTXLayer :366-412 (use_lang2visn_attn=False: no lang_* sub-modules)"""

    def __init__(self):
        super().__init__()
        self.visn_self_att = SelfAtt()
        self.visn_inter = Dense()
        self.visn_output = DenseLN(FF)
        self.visual_attention = XAtt()


class LXRTXLayerP(_Container):
    """HAMT LXRTXLayer, H/models/vilmodel_cmt.py:366-445"""

    def __init__(self):
        super().__init__()
        self.lang_self_att = SelfAtt()
        self.lang_inter = Dense()
        self.lang_output = DenseLN(FF)
        self.visn_self_att = SelfAtt()
        self.visn_inter = Dense()
        self.visn_output = DenseLN(FF)
        self.visual_attention = XAtt()


class PackedMHA(_Container):
    """nn.MultiheadAttention's parameters (in_proj_weight/bias, out_proj), D/models/transformer.py:138"""

    def __init__(self):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * H, H))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * H))
        self.out_proj = nn.Linear(H, H)


class PanoLayerP(_Container):
    """TransformerEncoderLayer (pre-norm use), D/models/transformer.py:133-190; LN eps 1e-5"""

    def __init__(self):
        super().__init__()
        self.self_attn = PackedMHA()
        self.linear1 = nn.Linear(H, FF)
        self.linear2 = nn.Linear(FF, H)
        self.norm1 = nn.LayerNorm(H, eps=1e-5)
        self.norm2 = nn.LayerNorm(H, eps=1e-5)


class PanoEncoderP(_Container):
    """TransformerEncoder with final norm, D/models/ops.py:11-23"""

    def __init__(self, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([PanoLayerP() for _ in range(n_layers)])
        self.norm = nn.LayerNorm(H, eps=1e-12)


class BertEmbeddingsP(_Container):
    """BertEmbeddings :49-78"""

    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, H, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, H)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, H)
        self.LayerNorm = nn.LayerNorm(H, eps=1e-12)


class DuetImageEmbeddingsP(_Container):
    """ImageEmbeddings :455-526 (obj branch absent when obj_feat_size == 0)"""

    def __init__(self, cfg):
        super().__init__()
        self.img_linear = nn.Linear(cfg.image_feat_size, H)
        self.img_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.loc_linear = nn.Linear(cfg.angle_feat_size + 3, H)
        self.loc_layer_norm = nn.LayerNorm(H, eps=1e-12)
        if cfg.obj_feat_size > 0 and cfg.obj_feat_size != cfg.image_feat_size:      # :464-468 (REVERIE / SOON object boxes)
            self.obj_linear = nn.Linear(cfg.obj_feat_size, H)
            self.obj_layer_norm = nn.LayerNorm(H, eps=1e-12)
        else:
            self.obj_linear = self.obj_layer_norm = None
        self.nav_type_embedding = nn.Embedding(3, H)
        self.layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.pano_encoder = PanoEncoderP(cfg.num_pano_layers)


def _pos_embed(d_in):
    return nn.Sequential(nn.Linear(d_in, H), nn.LayerNorm(H, eps=1e-12))


class XEncoderP(_Container):
    """CrossmodalEncoder :436-453"""

    def __init__(self, n, lang2visn=False):
        super().__init__()
        # use_lang2visn_attn (pre-training, VLN-DUET/pretrain_src/model/vilmodel.py:366-411): the layers also carry the
        # lang_self_att / lang_inter / lang_output blocks - the same parameter set as HAMT's LXRTXLayer
        self.x_layers = nn.ModuleList([(LXRTXLayerP() if lang2visn else GraphXLayerP()) for _ in range(n)])


class LocalVPEncoderP(_Container):
    """LocalVPEncoder :528-560"""

    def __init__(self, cfg):
        super().__init__()
        self.vp_pos_embeddings = _pos_embed(cfg.angle_feat_size * 2 + 6)
        self.encoder = XEncoderP(cfg.num_x_layers, bool(getattr(cfg, 'use_lang2visn_attn', False)))


class GlobalMapEncoderP(_Container):
    """GlobalMapEncoder :923-1006"""

    def __init__(self, cfg):
        super().__init__()
        self.gmap_pos_embeddings = _pos_embed(cfg.angle_feat_size + 3)
        self.gmap_step_embeddings = nn.Embedding(cfg.max_action_steps, H)
        self.encoder = XEncoderP(cfg.num_x_layers, bool(getattr(cfg, 'use_lang2visn_attn', False)))
        if cfg.graph_sprels:
            self.sprel_linear = nn.Linear(1, 1)
        else:
            self.sprel_linear = None


class ClsPredictionP(_Container):
    """ClsPrediction :1009-1020: Linear -> ReLU -> LN(1e-12) -> Linear(.,1)  (net.0 / net.2 / net.3)"""

    def __init__(self, input_size=H):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_size, H), nn.ReLU(), nn.LayerNorm(H, eps=1e-12), nn.Linear(H, 1))


class NextActionP(_Container):
    """HAMT NextActionPrediction, H/models/vilmodel_cmt.py:953-963  (net.0 / net.2 / net.4)"""

    def __init__(self, p_drop):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(H, H), nn.ReLU(), nn.LayerNorm(H, eps=1e-12), nn.Dropout(p_drop),
                                 nn.Linear(H, 1))


class BypassImagineEmbeddingsP(_Container):
    """BypassImagineEmbeddings :562-573"""

    def __init__(self):
        super().__init__()
        self.type_embedding = nn.Embedding(1, H)


class ImagineEmbeddingsP(_Container):
    """ImagineEmbeddings, H/models/vilmodel_cmt.py:634-703 (bypass_imag_encoder=False, imagine_enc_pano=True)"""

    def __init__(self, cfg):
        super().__init__()
        self.position_embeddings = nn.Embedding(cfg.max_imagination_len, H)
        self.type_embedding = nn.Embedding(1, H)
        self.layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.pano_img_linear = nn.Linear(cfg.image_feat_size, H)
        self.pano_img_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.pano_encoder = LayerStack('layer', [BertLayerP() for _ in range(cfg.num_h_pano_layers)])


class MLPProjectionHeadP(_Container):
    """MLPProjectionHead :575-589: 768 -> 512 -> 512 -> 768, no bias, dropout 0.15 on the input"""

    def __init__(self, d_in=H, d_hid=512, d_out=H, p_drop=0.15):
        super().__init__()
        self.fc1 = nn.Linear(d_in, d_hid, bias=False)
        self.fc2 = nn.Linear(d_hid, d_hid, bias=False)
        self.fc3 = nn.Linear(d_hid, d_out, bias=False)
        self.p_drop = p_drop


class AlignModelP(_Container):
    """AlignWithContrastiveLoss(/WithNegativeSamples) :591-779: owns image_proj only"""

    def __init__(self):
        super().__init__()
        self.image_proj = MLPProjectionHeadP()


class HamtImageEmbeddingsP(_Container):
    """HAMT ImageEmbeddings, H/models/vilmodel_cmt.py:521-544"""

    def __init__(self, cfg):
        super().__init__()
        self.img_linear = nn.Linear(cfg.image_feat_size, H)
        self.img_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.ang_linear = nn.Linear(cfg.angle_feat_size, H)
        self.ang_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.nav_type_embedding = nn.Embedding(3, H)
        self.layer_norm = nn.LayerNorm(H, eps=1e-12)


class HistoryEmbeddingsP(_Container):
    """HistoryEmbeddings, H/models/vilmodel_cmt.py:546-618 (hist_enc_pano branch)"""

    def __init__(self, cfg):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, H))
        self.img_linear = nn.Linear(cfg.image_feat_size, H)
        self.img_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.ang_linear = nn.Linear(cfg.angle_feat_size, H)
        self.ang_layer_norm = nn.LayerNorm(H, eps=1e-12)
        self.position_embeddings = nn.Embedding(cfg.max_action_steps, H)
        self.type_embedding = nn.Embedding(1, H)
        self.layer_norm = nn.LayerNorm(H, eps=1e-12)
        if cfg.hist_enc_pano:
            self.pano_img_linear = nn.Linear(cfg.image_feat_size, H)
            self.pano_img_layer_norm = nn.LayerNorm(H, eps=1e-12)
            self.pano_ang_linear = nn.Linear(cfg.angle_feat_size, H)
            self.pano_ang_layer_norm = nn.LayerNorm(H, eps=1e-12)
            self.pano_encoder = LayerStack('layer', [BertLayerP() for _ in range(cfg.num_h_pano_layers)])
        else:
            self.pano_encoder = None


class HamtEncoderP(_Container):
    """LxmertEncoder, H/models/vilmodel_cmt.py:448-519 (num_h_layers = num_r_layers = 0)"""

    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([BertLayerP() for _ in range(cfg.num_l_layers)])
        self.x_layers = nn.ModuleList([LXRTXLayerP() for _ in range(cfg.num_x_layers)])


def bert_init_(module: nn.Module, std: float = 0.02):
    """transformers' BertPreTrainedModel._init_weights: N(0, std) matrices / embeddings (padding row
    zeroed), zero biases, LayerNorm (1, 0).  Bare nn.Parameters (cls_token, in_proj_*) keep their
    constructor values like in the reference."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0.0, std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, 0.0, std)
            if m.padding_idx is not None:
                with torch.no_grad():
                    m.weight[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, PackedMHA):
            nn.init.xavier_uniform_(m.in_proj_weight)
            nn.init.zeros_(m.in_proj_bias)
