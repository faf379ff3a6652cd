"""Transformer blocks of the navigation hot path, expressed as sequences of libvlnimagine launches.

Shared by the DUET (duet.py) and HAMT (hamt.py) host modules.  Activations travel as a pair
(fp32 residual stream, bf16 GEMM operand copy); in the fp32 check mode only the fp32 tensor exists.
Several token streams that run the same layer shape with different weights (DUET global/local
encoders, HAMT language/vision streams) are stacked along the row axis, each stream starting on
a 128-row boundary, and go through ONE grouped launch per GEMM / LayerNorm.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import torch

from . import autograd_ops as ag
from . import ops
from .ops import BF16, F32, HIDDEN, EPI_GELU, EPI_NONE, EPI_RELU, MASK_ADD_NEG10000, MASK_NEG_INF  # noqa: F401


# ----------------------------------------------------------------------------------------------
# training switch: while it is on, every block below records autograd nodes whose forward AND backward are
# libvlnimagine kernels (autograd_ops.py); while it is off the blocks are plain launch sequences
# ----------------------------------------------------------------------------------------------
class _Mode:
    train = False
    p_hidden = 0.0          # nn.Dropout(hidden_dropout_prob) sites (train() mode only)
    p_attn = 0.0            # attention_probs_dropout_prob
    p_head = 0.0            # pred_head_dropout_prob (HAMT NextActionPrediction)


class grad_mode:
    """with blocks.grad_mode(flag, drop=(p_hidden, p_attn)): ...  (set by the host modules once per mode call)"""

    def __init__(self, flag: bool, drop=(0.0, 0.0)):
        self.flag = bool(flag)
        self.drop = drop if self.flag else (0.0, 0.0)

    def __enter__(self):
        self.prev = (_Mode.train, _Mode.p_hidden, _Mode.p_attn, _Mode.p_head)
        _Mode.train, _Mode.p_hidden, _Mode.p_attn = self.flag, float(self.drop[0]), float(self.drop[1])
        _Mode.p_head = float(self.drop[2]) if len(self.drop) > 2 else 0.0

    def __exit__(self, *a):
        _Mode.train, _Mode.p_hidden, _Mode.p_attn, _Mode.p_head = self.prev


def _attn_drop(device, p=None):
    """(p, site, seed tensor) for an attention problem, or None"""
    p = _Mode.p_attn if p is None else p
    return (p, ag.next_site(), ag.dropout_seed(device)) if p > 0 else None


def _dense_res_ln(x, lin: 'LinearPack', res32, ln: 'LNPack', eps, lowp, ends):
    """training: LN(dropout(x W^T + b) + res)  (BertSelfOutput / BertOutput, D/models/vilmodel.py:151-155,190-194)"""
    if _Mode.p_hidden > 0:
        d = ag.dropout(ag.linear(x, lin, lowp, out_dtype=F32, ends=ends), _Mode.p_hidden)
        return layer_norm(d, res32, ln, eps, lowp, ends)
    return layer_norm(ag.linear(x, lin, lowp, residual=res32, out_dtype=F32, ends=ends), None, ln, eps, lowp, ends)


def training() -> bool:
    return _Mode.train


# ----------------------------------------------------------------------------------------------
# derived weights (bf16 shadow copies, stacked groups), refreshed when a source parameter changes
# ----------------------------------------------------------------------------------------------
class _WeightsEpoch:
    """Counts optimiser steps process-wide.  Fused optimisers (torch.optim.AdamW(fused=True), apex) update parameters
    in place WITHOUT bumping their autograd version counters, so the version alone cannot tell that a derived copy is
    stale; a global optimizer-step hook can."""
    value = 0
    hooked = False

    @classmethod
    def install(cls):
        if cls.hooked:
            return
        cls.hooked = True
        try:
            from torch.optim.optimizer import register_optimizer_step_post_hook

            def _bump(optimizer, args, kwargs):
                cls.value += 1
                ag.advance_dropout_seeds()
            register_optimizer_step_post_hook(_bump)
        except ImportError:                                    # very old torch: version counters only
            pass


def weights_epoch() -> int:
    return _WeightsEpoch.value


def touch_weights():
    """call after updating parameters by any means that neither bumps tensor versions nor runs an optimizer step"""
    _WeightsEpoch.value += 1


class Pack:
    """Device tensors derived from parameters; rebuilt when any source's version / storage changes or an optimiser
    has stepped (in-place update, load_state_dict, .cuda()).  fp32 masters stay the nn.Parameters themselves."""

    def __init__(self, sources: Sequence[torch.Tensor], build):
        _WeightsEpoch.install()
        self.sources = list(sources)
        self._build = build
        self._key = None
        self._val = None

    def get(self):
        key = (_WeightsEpoch.value,) + tuple((p._version, p.data_ptr()) for p in self.sources)
        if key != self._key:
            with torch.no_grad():
                self._val = self._build()
            self._key = key
        return self._val


class LinearPack:
    """Rows-stacked weight [sum N_i, K] (+ bias) of one or several nn.Linear-shaped parameter sets."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Optional[Sequence[Optional[torch.Tensor]]] = None):
        self.weights = list(weights)
        self.biases = None if biases is None or all(b is None for b in biases) else list(biases)
        srcs = self.weights + ([b for b in self.biases if b is not None] if self.biases else [])
        self._pack = Pack(srcs, self._make)

    def _make(self):
        ws = [w.detach() for w in self.weights]
        w32 = (torch.cat(ws, 0) if len(ws) > 1 else ws[0]).contiguous()
        b = None
        if self.biases is not None:
            bs = [(b.detach() if b is not None else torch.zeros(w.shape[0], device=w.device)) for b, w in zip(self.biases, ws)]
            b = (torch.cat(bs, 0) if len(bs) > 1 else bs[0]).contiguous().float()
        return {'w32': w32, 'w16': ops.cast_bf16(w32), 'b': b}

    def get(self, lowp: bool):
        v = self._pack.get()
        return (v['w16'] if lowp else v['w32']), v['b']

    def get_token(self):
        """changes whenever a source parameter does (in-place update or re-allocation)"""
        self._pack.get()
        return self._pack._key

    def get_t(self, lowp: bool, n_groups: int = 1):
        """per-group transposed weight [n_groups * K, N] (the B operand of the input-gradient GEMM), cached with
        the pack"""
        v = self._pack.get()
        key = ('t', lowp, n_groups)
        if key not in v:
            w = v['w16'] if lowp else v['w32']
            n = w.shape[0] // n_groups
            with torch.no_grad():
                v[key] = torch.cat([ag.transpose(w[g * n:(g + 1) * n]) for g in range(n_groups)], 0).contiguous()
        return v[key]

    def grad_sources(self):
        """the parameters in the order LinearFn.backward returns their gradients"""
        return self.weights + ([b for b in self.biases if b is not None] if self.biases else [])


class StackPack:
    """[n, ...] stack of same-shaped vectors (LayerNorm gains/biases, head vectors) for grouped row kernels."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.tensors = list(tensors)
        self._pack = Pack(self.tensors, self._make)

    def _make(self):
        ts = [t.detach().reshape(-1) for t in self.tensors]
        return (torch.stack(ts, 0) if len(ts) > 1 else ts[0]).contiguous().float()

    def get(self):
        return self._pack.get()


class LNPack:
    def __init__(self, lns):
        self.g = StackPack([ln.weight for ln in lns])
        self.b = StackPack([ln.bias for ln in lns])

    def get(self):
        return self.g.get(), self.b.get()

    def grad_sources(self):
        return self.g.tensors + self.b.tensors


# ----------------------------------------------------------------------------------------------
# activations
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Act:
    f32: torch.Tensor                     # [rows, 768] fp32 residual stream
    b16: Optional[torch.Tensor] = None    # bf16 copy fed to the tensor-core GEMMs (None in fp32 mode)

    def operand(self, lowp: bool):
        return self.b16 if lowp else self.f32


@dataclasses.dataclass
class Stream:
    """One token stream inside a row-stacked activation."""
    row0: int                       # first row (multiple of 128)
    B: int
    L: int                          # tokens per episode (padded length)
    mask: Optional[torch.Tensor]    # uint8 [B, L] key-padding mask (1 = valid) or None
    group: int = 0                  # weight block this stream uses
    pair_dist: Optional[torch.Tensor] = None   # fp32 [B, L, L] graph distances (GASA) or None
    bias_affine: Optional[torch.Tensor] = None  # device {w, b}
    affine_params: Optional[tuple] = None       # (sprel_linear.weight, .bias): gradient routing in training

    @property
    def rows(self):
        return self.B * self.L

    def view(self, t: torch.Tensor):
        return t[self.row0:self.row0 + self.rows]


def stack_layout(token_counts: Sequence[int]):
    """Row offsets for streams stacked on ops.ROW_ALIGN-row boundaries -> (row0 list, group_row_end list, total rows)."""
    row0, ends, cur = [], [], 0
    for i, n in enumerate(token_counts):
        row0.append(cur)
        last = i == len(token_counts) - 1
        cur = cur + n if last else ops.pad_rows(cur + n)
        ends.append(cur)
    return row0, ends, cur


_CTX_BUFFERS = {}


def _ctx_buffer(rows: int, like: torch.Tensor, streams) -> torch.Tensor:
    """Attention output buffer for a row-stacked activation.  The rows between streams (padding up to the next
    128-row boundary) are never written by the attention kernel but do flow through the following GEMM, whose
    rows are independent: they are zeroed so that the padding stays finite.
    Inference consumes the buffer at once (the output projection that follows reads it on the same stream), so one
    buffer per layout is allocated and zeroed ONCE and reused by every layer and every step - it used to cost one fill
    launch per attention call (8 of the ~70 launches of a DUET step).  Autograd keeps the attention output for the
    backward pass and therefore gets a fresh buffer each time."""
    pads = [(a.row0 + a.rows, b.row0) for a, b in zip(streams[:-1], streams[1:]) if b.row0 > a.row0 + a.rows]
    key = None
    if not _Mode.train and not torch.is_grad_enabled():
        key = (rows, like.dtype, like.device, tuple((s.row0, s.rows) for s in streams))
        ctx = _CTX_BUFFERS.get(key)
        if ctx is not None:
            return ctx
    ctx = torch.empty((rows, HIDDEN), dtype=like.dtype, device=like.device)
    for r0, r1 in pads:
        ctx[r0:r1].zero_()
    # never keep memory of a graph's private pool, and never evict: captured graphs hold raw pointers to these buffers
    if key is not None and not torch.cuda.is_current_stream_capturing() and len(_CTX_BUFFERS) < 256:
        _CTX_BUFFERS[key] = ctx
    return ctx


def linear_residual_ln(x, lin: 'LinearPack', res32, ln: LNPack, eps, lowp, ends=None) -> Act:
    """LN(x W^T + b + res) (inference): one cluster GEMM with the LayerNorm in its epilogue in bf16 mode, GEMM + row
    kernel in the fp32 check mode (or with VLN_IMAGINE_FUSED_LN=0)."""
    w, b = lin.get(lowp)
    g, be = ln.get()
    if lowp and ops.fused_ln_enabled() and w.shape[0] == (1 if ends is None else len(ends)) * HIDDEN:
        _, y32, y16 = ops.gemm_ln(x, w, b, res32, g, be, eps, group_row_end=ends)
        return Act(y32, y16)
    ao = ops.gemm(x, w, b, residual=res32, out_dtype=F32, group_row_end=ends)
    y32, y16 = ops.add_ln(ao, None, g, be, eps, want16=lowp, group_row_end=ends)
    return Act(y32, y16)


def layer_norm(x32, res32, ln: LNPack, eps, lowp, ends=None) -> Act:
    if _Mode.train:
        y32, y16 = ag.layer_norm(x32, res32, ln, eps, lowp, ends)
        return Act(y32, y16)
    g, b = ln.get()
    y32, y16 = ops.add_ln(x32, res32, g, b, eps, want16=lowp, group_row_end=ends)
    return Act(y32, y16)


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
class SelfFFNPack:
    """Weights of BertAttention + BertIntermediate + BertOutput for one or several stacked streams."""

    def __init__(self, attns, inters, outs):
        # attns: modules with .self.{query,key,value} and .output.{dense,LayerNorm}
        self.qkv = LinearPack([w for a in attns for w in (a.self.query.weight, a.self.key.weight, a.self.value.weight)],
                              [b for a in attns for b in (a.self.query.bias, a.self.key.bias, a.self.value.bias)])
        self.o = LinearPack([a.output.dense.weight for a in attns], [a.output.dense.bias for a in attns])
        self.ln1 = LNPack([a.output.LayerNorm for a in attns])
        self.w1 = LinearPack([m.dense.weight for m in inters], [m.dense.bias for m in inters])
        self.w2 = LinearPack([m.dense.weight for m in outs], [m.dense.bias for m in outs])
        self.ln2 = LNPack([m.LayerNorm for m in outs])


def self_attn_ffn(x: Act, pk: SelfFFNPack, streams: List[Stream], ends, lowp: bool, eps=1e-12) -> Act:
    """Post-LN BERT layer: LN(x + O(attn(QKV(x)))) then LN(y + W2 gelu(W1 y)).
    Reference: BertLayer, VLN-DUET/map_nav_src/models/vilmodel.py:196-209 (and :80-194)."""
    xin = x.operand(lowp)
    rows = xin.shape[0]
    if _Mode.train:
        qkv = ag.linear(xin, pk.qkv, lowp, ends=ends)
        spec, extras = [], ()
        for s in streams:
            spec.append(dict(q=(0, s.row0, 0), k=(0, s.row0, HIDDEN), v=(0, s.row0, 2 * HIDDEN), B=s.B, Lq=s.L, Lk=s.L,
                             key_mask=s.mask, pair_dist=s.pair_dist, bias_affine=s.bias_affine, out_row0=s.row0,
                             drop=_attn_drop(xin.device)))
            if s.pair_dist is not None and s.affine_params is not None:
                extras = tuple(s.affine_params)
        ctx = ag.AttentionFn.apply(spec, rows, MASK_ADD_NEG10000, 1, qkv, *extras)
        y = _dense_res_ln(ctx, pk.o, x.f32, pk.ln1, eps, lowp, ends)
        return ffn(y, pk.w1, pk.w2, pk.ln2, ends, lowp, eps)
    w, b = pk.qkv.get(lowp)
    qkv = ops.gemm(xin, w, b, group_row_end=ends)                      # [rows, 2304]
    ctx = _ctx_buffer(rows, xin, streams)
    probs = []
    for s in streams:
        q = s.view(qkv)
        probs.append(dict(q=q[:, 0:HIDDEN], k=q[:, HIDDEN:2 * HIDDEN], v=q[:, 2 * HIDDEN:3 * HIDDEN], out=s.view(ctx),
                          B=s.B, Lq=s.L, Lk=s.L, key_mask=s.mask, pair_dist=s.pair_dist, bias_affine=s.bias_affine))
    ops.attention_multi(probs)
    y = linear_residual_ln(ctx, pk.o, x.f32, pk.ln1, eps, lowp, ends)
    return ffn(y, pk.w1, pk.w2, pk.ln2, ends, lowp, eps)


def ffn(y: Act, w1: LinearPack, w2: LinearPack, ln2: LNPack, ends, lowp: bool, eps=1e-12) -> Act:
    if _Mode.train:
        h = ag.ActFn.apply(ag.linear(y.operand(lowp), w1, lowp, ends=ends), EPI_GELU)
        return _dense_res_ln(h, w2, y.f32, ln2, eps, lowp, ends)
    w, b = w1.get(lowp)
    h = ops.gemm(y.operand(lowp), w, b, epilogue=EPI_GELU, group_row_end=ends)   # [rows, 3072]
    return linear_residual_ln(h, w2, y.f32, ln2, eps, lowp, ends)


class CrossPack:
    """BertXAttention weights: query / output per stream group; key+value stacked per context user."""

    def __init__(self, xatts):
        self.q = LinearPack([x.att.query.weight for x in xatts], [x.att.query.bias for x in xatts])
        self.kv = LinearPack([w for x in xatts for w in (x.att.key.weight, x.att.value.weight)],
                             [b for x in xatts for b in (x.att.key.bias, x.att.value.bias)])
        self.o = LinearPack([x.output.dense.weight for x in xatts], [x.output.dense.bias for x in xatts])
        self.ln = LNPack([x.output.LayerNorm for x in xatts])


def cross_attn(x: Act, kv: torch.Tensor, kv_col0: Sequence[int], ctx_len: int, ctx_mask, pk: CrossPack,
               streams: List[Stream], ends, lowp: bool, eps=1e-12) -> Act:
    """LN(x + O(attn(Q(x), K(ctx), V(ctx)))).  kv holds the projected context [B*ctx_len, *] with this
    stream's K at columns kv_col0[i] and V at kv_col0[i]+768.
    Reference: BertXAttention, VLN-DUET/map_nav_src/models/vilmodel.py:302-364."""
    xin = x.operand(lowp)
    rows = xin.shape[0]
    if _Mode.train:
        q = ag.linear(xin, pk.q, lowp, ends=ends)
        spec = [dict(q=(0, s.row0, 0), k=(1, 0, c0), v=(1, 0, c0 + HIDDEN), B=s.B, Lq=s.L, Lk=ctx_len, key_mask=ctx_mask,
                     out_row0=s.row0, drop=_attn_drop(xin.device)) for s, c0 in zip(streams, kv_col0)]
        ctx = ag.AttentionFn.apply(spec, rows, MASK_ADD_NEG10000, 2, q, kv)
        return _dense_res_ln(ctx, pk.o, x.f32, pk.ln, eps, lowp, ends)
    w, b = pk.q.get(lowp)
    q = ops.gemm(xin, w, b, group_row_end=ends)
    ctx = _ctx_buffer(rows, xin, streams)
    ops.attention_multi([dict(q=s.view(q), k=kv[:, c0:c0 + HIDDEN], v=kv[:, c0 + HIDDEN:c0 + 2 * HIDDEN], out=s.view(ctx),
                              B=s.B, Lq=s.L, Lk=ctx_len, key_mask=ctx_mask) for s, c0 in zip(streams, kv_col0)])
    return linear_residual_ln(ctx, pk.o, x.f32, pk.ln, eps, lowp, ends)


def linear(x: torch.Tensor, pk: LinearPack, lowp: bool, out_dtype=None, ends=None, residual=None) -> torch.Tensor:
    """y = x W^T + b [+ residual] for a LinearPack (no activation); differentiable in training."""
    if _Mode.train:
        return ag.linear(x, pk, lowp, residual=residual, out_dtype=out_dtype, ends=ends)
    w, b = pk.get(lowp)
    return ops.gemm(x, w, b, residual=residual, out_dtype=out_dtype or (BF16 if lowp else F32), group_row_end=ends)


def as_act(x32: torch.Tensor, lowp: bool) -> Act:
    """fp32 rows -> activation pair (adds the bf16 operand copy in bf16 mode)."""
    x32 = x32.contiguous()
    if _Mode.train:
        return Act(x32, ag.CastBf16Fn.apply(x32) if lowp else None)
    return Act(x32, ops.cast_bf16(x32) if lowp else None)


def operand(x32: torch.Tensor, lowp: bool) -> torch.Tensor:
    """fp32 rows -> GEMM operand of the current precision (differentiable in training)"""
    x32 = x32.contiguous()
    if not lowp:
        return x32
    return ag.CastBf16Fn.apply(x32) if _Mode.train else ops.cast_bf16(x32)


def embed(rows: int, device, *, a=None, a_ln=None, feat=None, feat_lin=None, feat_ln=None, idx=None, table=None,
          pos_table=None, pos_period=0, const_rows=(), out_ln=None, eps=1e-12, lowp=False, y32=None, y16=None,
          dropout=False) -> Act:
    """Input-embedding composer  LN_out([LN_a](a) + LN_f(W feat + b) + table[idx] + pos[row % period] + consts).
    *_ln are nn.LayerNorm-like parameter holders, feat_lin an nn.Linear-like holder.  Inference: ONE fused kernel
    (vi_embed_compose).  Training: the same sum built from differentiable primitives (autograd_ops)."""
    ln_pair = lambda m: (m.weight, m.bias) if m is not None else None      # noqa: E731
    if not _Mode.train:
        consts = list(const_rows) + [None, None]
        o32, o16 = ops.embed_compose(rows, device, a=a, a_ln=ln_pair(a_ln), feat=feat,
                                     feat_w=feat_lin.weight if feat_lin is not None else None,
                                     feat_b=feat_lin.bias if feat_lin is not None else None, feat_ln=ln_pair(feat_ln),
                                     idx=idx, table=table, pos_table=pos_table, pos_period=pos_period,
                                     const_row=consts[0], const_row2=consts[1], out_ln=ln_pair(out_ln), eps=eps,
                                     y32=y32, y16=y16, want16=lowp and y16 is None and (y32 is None))
        return Act(o32, o16)
    terms = []
    if a is not None:
        if a_ln is not None:
            t, _ = ag.LayerNormFn.apply(a, None, a_ln.weight, a_ln.bias, eps, False, None, 2, a_ln.weight, a_ln.bias)
            terms.append(t)
        else:
            terms.append(a)
    if feat is not None:
        t = ag.SmallLinearFn.apply(feat, feat_lin.weight, feat_lin.bias)
        if feat_ln is not None:
            t, _ = ag.LayerNormFn.apply(t, None, feat_ln.weight, feat_ln.bias, eps, False, None, 2, feat_ln.weight, feat_ln.bias)
        terms.append(t)
    if idx is not None:
        terms.append(ag.GatherRowsFn.apply(table, idx, 0, rows))
    if pos_table is not None:
        terms.append(ag.GatherRowsFn.apply(pos_table, None, pos_period, rows))
    consts = list(const_rows)
    while len(terms) > 3:                                     # the fused sum kernel takes three row tensors
        terms = [ag.SumRowsFn.apply(3, *terms[:3])] + terms[3:]
    s = ag.SumRowsFn.apply(len(terms), *terms, *consts) if (len(terms) > 1 or consts) else terms[0]
    drop = dropout and _Mode.p_hidden > 0                     # nn.Dropout after the embedding LayerNorm (vilmodel.py:77,1124)
    if out_ln is not None:
        o32, o16 = ag.LayerNormFn.apply(s, None, out_ln.weight, out_ln.bias, eps, lowp and not drop, None, 2, out_ln.weight, out_ln.bias)
        if drop:
            o32 = ag.dropout(o32, _Mode.p_hidden)
            return Act(o32, ag.CastBf16Fn.apply(o32) if lowp else None)
        return Act(o32, o16 if lowp else None)
    if drop:
        s = ag.dropout(s, _Mode.p_hidden)
    return Act(s, ag.CastBf16Fn.apply(s) if lowp else None)


def mask_u8(m: Optional[torch.Tensor]):
    """bool mask -> the uint8 view the kernels read (zero-copy for contiguous bool tensors)."""
    if m is None:
        return None
    if m.dtype == torch.bool and m.is_contiguous():
        return m.view(torch.uint8)
    return m.to(torch.uint8).contiguous()


class PanoLayerPack:
    """Weights of one pre-norm TransformerEncoderLayer of the panorama encoder."""

    def __init__(self, layer):
        self.norm1 = LNPack([layer.norm1])
        self.qkv = LinearPack([layer.self_attn.in_proj_weight], [layer.self_attn.in_proj_bias])
        self.o = LinearPack([layer.self_attn.out_proj.weight], [layer.self_attn.out_proj.bias])
        self.norm2 = LNPack([layer.norm2])
        self.w1 = LinearPack([layer.linear1.weight], [layer.linear1.bias])
        self.w2 = LinearPack([layer.linear2.weight], [layer.linear2.bias])


def pano_layer(x32: torch.Tensor, pk: PanoLayerPack, B: int, L: int, key_mask, lowp: bool, eps=1e-5) -> torch.Tensor:
    """Pre-norm layer: x += O(attn(QKV(LN1 x))) ; x += W2 gelu(W1 LN2 x).  The key-padding mask is the
    -inf kind of nn.MultiheadAttention.  Reference: TransformerEncoderLayer.forward_pre,
    VLN-DUET/map_nav_src/models/transformer.py:170-182."""
    h = layer_norm(x32, None, pk.norm1, eps, lowp)
    if _Mode.train:
        qkv = ag.linear(h.operand(lowp), pk.qkv, lowp)
        ph = _Mode.p_hidden
        spec = [dict(q=(0, 0, 0), k=(0, 0, HIDDEN), v=(0, 0, 2 * HIDDEN), B=B, Lq=L, Lk=L, key_mask=key_mask, out_row0=0,
                     drop=_attn_drop(x32.device, ph))]
        ctx = ag.AttentionFn.apply(spec, B * L, MASK_NEG_INF, 1, qkv)
        if ph > 0:                                            # src + dropout1(src2); linear2(dropout(act)); src + dropout2
            x32 = ag.SumRowsFn.apply(2, x32, ag.dropout(ag.linear(ctx, pk.o, lowp, out_dtype=F32), ph))
            h = layer_norm(x32, None, pk.norm2, eps, lowp)
            f = ag.dropout(ag.ActFn.apply(ag.linear(h.operand(lowp), pk.w1, lowp), EPI_GELU), ph)
            return ag.SumRowsFn.apply(2, x32, ag.dropout(ag.linear(f, pk.w2, lowp, out_dtype=F32), ph))
        x32 = ag.linear(ctx, pk.o, lowp, residual=x32, out_dtype=F32)
        h = layer_norm(x32, None, pk.norm2, eps, lowp)
        f = ag.ActFn.apply(ag.linear(h.operand(lowp), pk.w1, lowp), EPI_GELU)
        return ag.linear(f, pk.w2, lowp, residual=x32, out_dtype=F32)
    w, b = pk.qkv.get(lowp)
    qkv = ops.gemm(h.operand(lowp), w, b)
    ctx = ops.attention(qkv[:, 0:HIDDEN], qkv[:, HIDDEN:2 * HIDDEN], qkv[:, 2 * HIDDEN:3 * HIDDEN], B, L, L,
                        key_mask=key_mask, mask_mode=MASK_NEG_INF)
    w, b = pk.o.get(lowp)
    if lowp and ops.fused_ln_enabled():
        g2, b2 = pk.norm2.get()
        x32, _, h16 = ops.gemm_ln(ctx, w, b, x32, g2, b2, eps, want32=False, want16=True, want_pre=True)
        h = Act(None, h16)
    else:
        x32 = ops.gemm(ctx, w, b, residual=x32, out_dtype=F32)
        h = layer_norm(x32, None, pk.norm2, eps, lowp)
    w, b = pk.w1.get(lowp)
    f = ops.gemm(h.operand(lowp), w, b, epilogue=EPI_GELU)
    w, b = pk.w2.get(lowp)
    return ops.gemm(f, w, b, residual=x32, out_dtype=F32)


class ClsHeadPack:
    """ClsPrediction / NextActionPrediction weights for one or several row groups:
    Linear -> ReLU -> LayerNorm -> Linear(768, 1)."""

    def __init__(self, heads, last_index=3):
        self.w0 = LinearPack([h.net[0].weight for h in heads], [h.net[0].bias for h in heads])
        self.ln = LNPack([h.net[2] for h in heads])
        self.w1 = StackPack([h.net[last_index].weight for h in heads])
        self.b1 = StackPack([h.net[last_index].bias for h in heads])
        self.has_dropout = last_index == 4


def cls_head(x: torch.Tensor, pk: ClsHeadPack, lowp: bool, ends=None, eps=1e-12) -> torch.Tensor:
    """x [rows, K] operand (bf16 or fp32) -> raw logit per row (fp32)."""
    if _Mode.train:
        h = ag.ActFn.apply(ag.linear(x, pk.w0, lowp, out_dtype=F32, ends=ends), EPI_RELU)
        y32, _ = ag.layer_norm(h, None, pk.ln, eps, False, ends)
        if pk.has_dropout and _Mode.p_head > 0:               # NextActionPrediction net.3 (H/models/vilmodel_cmt.py:953-963)
            y32 = ag.dropout(y32, _Mode.p_head)
        n = len(pk.w1.tensors)
        return ag.RowDotFn.apply(y32, pk.w1.get(), pk.b1.get(), ends, n, *pk.w1.tensors, *pk.b1.tensors)
    w, b = pk.w0.get(lowp)
    h = ops.gemm(x, w, b, epilogue=EPI_RELU, out_dtype=F32, group_row_end=ends)
    g, be = pk.ln.get()
    return ops.ln_dot(h, g, be, eps, pk.w1.get(), pk.b1.get(), group_row_end=ends)
