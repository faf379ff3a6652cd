"""Thin tensor-level wrappers over the C ABI (include/vlnimagine.h).

PyTorch is used for device memory and streams only: every function here validates its tensors,
allocates the output with torch.empty and launches one libvlnimagine kernel on the current
stream.  Nothing falls back to ATen arithmetic.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, EPI_NONE, EPI_GELU, EPI_RELU, MASK_ADD_NEG10000, MASK_NEG_INF  # noqa: F401

HIDDEN = 768
HEADS = 12
BF16, F32 = torch.bfloat16, torch.float32


class _Counters:
    """Kernel launches issued through this module (bench.py reports them as gpu_launches) and optional
    per-launch CUDA-event traces: of the GEMM kernel with its shape (bench.py's roofline leg) and of every
    library entry point by name (bench.py's per-kernel breakdown)."""
    launches = 0
    gemm_trace = None          # list of (M, N, K, start_event, end_event) when enabled
    trace = None               # list of (entry point, start_event, end_event) when enabled


class _Lib:
    """The ctypes library, with an event pair around each call while _Counters.trace is a list."""

    def __getattr__(self, name):
        fn = getattr(_lib.lib, name)

        def call(*a):
            tr = _Counters.trace
            if tr is None:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            tr.append((name, e0, e1))
            return rc
        setattr(self, name, call)
        return call


lib = _Lib()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _launched(n: int = 1):
    _Counters.launches += n


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.VlnImagineError('%s must be a CUDA tensor: vln-imagine_b200 has no CPU path' % name)


def ensure_init(t: torch.Tensor):
    _need_cuda(t, 'tensor')
    _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.VlnImagineError('%s must be a 2-D row-major view (got shape %s strides %s)' % (name, tuple(t.shape), t.stride()))
    return t.shape[0], t.shape[1], t.stride(0)


def pad128(n: int) -> int:
    return (n + 127) // 128 * 128


ROW_ALIGN = 256     # streams of a row-stacked activation start on 256-row boundaries: a CTA-pair GEMM tile
                    # (256 rows) then never straddles two weight groups


def pad_rows(n: int) -> int:
    return (n + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


TILE_PAIR = 0x1000
_TILE_CANDIDATES = [256 | TILE_PAIR, 192 | TILE_PAIR, 128 | TILE_PAIR, 256, 192, 128, 96, 64]
_TILE_CACHE = {}           # GEMM signature -> tile code, filled by timing the candidates on first use
_TILE_TABLE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tile_table.json')


def _key_str(key):
    return repr(key[:-1])      # the device index is not part of the persisted signature


def load_tile_table(path: str = _TILE_TABLE_PATH):
    """Tile choices measured on a B200 by tools/tune_tiles.py; shapes not listed are tuned on first use."""
    import json
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


_TILE_TABLE = load_tile_table()


def autotune_enabled() -> bool:
    return os.environ.get('VLN_IMAGINE_AUTOTUNE', '1') != '0' and not os.environ.get('VI_GEMM_TILE')


def _tune_tile(key, launch, N: int, pair_ok: bool) -> int:
    """Time every admissible tile shape for this GEMM signature (CUDA events, back-to-back launches) and remember
    the fastest.  Skipped (library cost model) while a CUDA graph is being captured."""
    if key in _TILE_CACHE:
        return _TILE_CACHE[key]
    ks = _key_str(key)
    if ks in _TILE_TABLE and os.environ.get('VLN_IMAGINE_RETUNE', '0') == '0':
        _TILE_CACHE[key] = int(_TILE_TABLE[ks])
        return _TILE_CACHE[key]
    if torch.cuda.is_current_stream_capturing():
        return 0
    best, best_t = 0, float('inf')
    for tile in _TILE_CANDIDATES:
        bn, pair = tile & 0xFFF, bool(tile & TILE_PAIR)
        if N % bn or (pair and not pair_ok):
            continue
        for _ in range(2):
            launch(tile)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(400_000)                     # let the host queue the launches back to back
        e0.record()
        for _ in range(6):
            launch(tile)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        if t < best_t:
            best, best_t = tile, t
    _TILE_CACHE[key] = best
    return best


def gemm(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, epilogue: int = EPI_NONE,
         out_dtype: torch.dtype = BF16, group_row_end: Optional[Sequence[int]] = None,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = epi(X W^T + bias) + residual.  x bf16 -> tcgen05 kernel; x fp32 -> fp32 check-mode kernel.
    w is [n_groups*N, K]; group_row_end (python ints) splits the rows of x between the weight blocks."""
    M, K, ldx = _rows2d(x, 'x')
    n_groups = 1 if group_row_end is None else len(group_row_end)
    if w.dim() != 2 or not w.is_contiguous() or w.shape[1] != K or w.shape[0] % n_groups:
        raise _lib.VlnImagineError('weight shape %s does not match x %s / %d groups' % (tuple(w.shape), tuple(x.shape), n_groups))
    N = w.shape[0] // n_groups
    ends = _lib.int_array(list(group_row_end)) if group_row_end is not None else None
    ldr = residual.stride(0) if residual is not None else 0
    if x.dtype == BF16:
        if w.dtype != BF16:
            raise _lib.VlnImagineError('bf16 GEMM needs a bf16 weight')
        if out is None:
            out = torch.empty((M, N), dtype=out_dtype, device=x.device)
        ydt = _lib.DT_F32 if out.dtype == F32 else _lib.DT_BF16

        def launch(tile):
            check(lib.vi_gemm_bf16_tiled(x.data_ptr(), ldx, w.data_ptr(), _ptr(bias), _ptr(residual), ldr, out.data_ptr(),
                                         out.stride(0), ydt, M, N, K, epilogue, n_groups, ends, tile, _stream()),
                  'vi_gemm_bf16_tiled')
        tile = 0
        if autotune_enabled():
            ge = tuple(group_row_end) if group_row_end is not None else None
            pair_ok = ge is None or all(e % 256 == 0 for e in ge[:-1])
            key = (M, N, K, ge, epilogue, bias is not None, residual is not None, ydt, x.device.index)
            tile = _tune_tile(key, launch, N, pair_ok)
        tr = _Counters.gemm_trace
        if tr is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        launch(tile)
        _launched(1)
        if tr is not None:
            e1.record()
            tr.append((M, N, K, e0, e1))
    elif x.dtype == F32:
        if w.dtype != F32:
            raise _lib.VlnImagineError('fp32 GEMM needs an fp32 weight')
        if out is None:
            out = torch.empty((M, N), dtype=F32, device=x.device)
        check(lib.vi_gemm_f32(x.data_ptr(), ldx, w.data_ptr(), _ptr(bias), _ptr(residual), ldr, out.data_ptr(),
                              out.stride(0), M, N, K, epilogue, n_groups, ends, _stream()), 'vi_gemm_f32')
        _launched(1)
    else:
        raise _lib.VlnImagineError('unsupported GEMM dtype %s' % x.dtype)
    return out


def gemm_mc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            epilogue: int = EPI_NONE, out_dtype: torch.dtype = BF16, group_row_end: Optional[Sequence[int]] = None,
            tile: int = 256) -> torch.Tensor:
    """EXPERIMENTAL: ``gemm`` through vi_gemm_bf16_mc (W tile multicast across a 2-CTA cluster).  Not used by the model code;
    tools/gemm_mc_check.py compares it with ``gemm`` and times both."""
    M, K, ldx = _rows2d(x, 'x')
    n_groups = 1 if group_row_end is None else len(group_row_end)
    N = w.shape[0] // n_groups
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    ends = _lib.int_array(list(group_row_end)) if group_row_end is not None else None
    check(lib.vi_gemm_bf16_mc(x.data_ptr(), ldx, w.data_ptr(), _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
                              out.data_ptr(), out.stride(0), _lib.DT_F32 if out_dtype == F32 else _lib.DT_BF16, M, N, K, epilogue,
                              n_groups, ends, tile, _stream()), 'vi_gemm_bf16_mc')
    _launched(1)
    return out


def fused_ln_enabled() -> bool:
    """vi_gemm_ln_bf16 (cluster GEMM with residual + LayerNorm in the epilogue) is correct (tests/test_kernels_gpu.py::
    test_gemm_ln_rowblock) but not yet faster than the tuned GEMM + row kernel it replaces (B200, M = 4416: 44.9 vs 17.1 us
    at K = 768, 66.8 vs 29.2 us at K = 3072; 14.9 / 35.6 us with its epilogue I/O switched off - DESIGN.md section 5), so
    it is opt-in: VLN_IMAGINE_FUSED_LN=1."""
    return os.environ.get('VLN_IMAGINE_FUSED_LN', '0') == '1'


def gemm_ln(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
            gamma: torch.Tensor, beta: torch.Tensor, eps: float, want32: bool = True, want16: bool = True,
            want_pre: bool = False, group_row_end: Optional[Sequence[int]] = None):
    """(pre32 or None, y32 or None, y16 or None) with y = LayerNorm(x w^T + bias + residual) * gamma + beta over 768
    columns (vi_gemm_ln_bf16: cluster GEMM with the LayerNorm in its epilogue).  x bf16 [M, K], w bf16 [n_groups*768, K]."""
    M, K, ldx = _rows2d(x, 'x')
    n_groups = 1 if group_row_end is None else len(group_row_end)
    if x.dtype != BF16 or w.dtype != BF16 or not w.is_contiguous() or w.shape != (n_groups * HIDDEN, K):
        raise _lib.VlnImagineError('gemm_ln: x / w must be bf16 with w of shape [%d, %d] (got %s)' % (n_groups * HIDDEN, K, tuple(w.shape)))
    dev = x.device
    pre = torch.empty((M, HIDDEN), dtype=F32, device=dev) if want_pre else None
    y32 = torch.empty((M, HIDDEN), dtype=F32, device=dev) if want32 else None
    y16 = torch.empty((M, HIDDEN), dtype=BF16, device=dev) if want16 else None
    ends = _lib.int_array(list(group_row_end)) if group_row_end is not None else None
    ldr = residual.stride(0) if residual is not None else 0
    tr = _Counters.gemm_trace
    if tr is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.vi_gemm_ln_bf16(x.data_ptr(), ldx, w.data_ptr(), _ptr(bias), _ptr(residual), ldr, gamma.data_ptr(), beta.data_ptr(),
                              eps, _ptr(pre), _ptr(y32), _ptr(y16), M, K, n_groups, ends, _stream()), 'vi_gemm_ln_bf16')
    _launched(1)
    if tr is not None:
        e1.record()
        tr.append((M, HIDDEN, K, e0, e1))
    return pre, y32, y16


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, Lq: int, Lk: int,
              key_mask: Optional[torch.Tensor] = None, pair_dist: Optional[torch.Tensor] = None,
              bias_affine: Optional[torch.Tensor] = None, mask_mode: int = MASK_ADD_NEG10000,
              out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B*Lq, >=768 view], k/v [B*Lk, view] -> o [B*Lq, 768]; 12 heads of 64."""
    _, _, ldq = _rows2d(q, 'q')
    _, _, ldk = _rows2d(k, 'k')
    _, _, ldv = _rows2d(v, 'v')
    if out is None:
        out = torch.empty((B * Lq, HIDDEN), dtype=q.dtype, device=q.device)
    if key_mask is not None and (key_mask.dtype != torch.uint8 or not key_mask.is_contiguous()):
        raise _lib.VlnImagineError('key_mask must be a contiguous uint8 tensor')
    check(lib.vi_attn_fwd(q.data_ptr(), ldq, k.data_ptr(), ldk, v.data_ptr(), ldv, out.data_ptr(), out.stride(0),
                          _lib.DT_BF16 if q.dtype == BF16 else _lib.DT_F32, _ptr(key_mask), _ptr(pair_dist),
                          _ptr(bias_affine), _ptr(lse), B, HEADS, Lq, Lk, mask_mode, _stream()), 'vi_attn_fwd')
    _launched(1)
    return out


def attention_multi(problems, mask_mode: int = MASK_ADD_NEG10000):
    """Several attention problems in one launch.  Each problem is a dict with q, k, v (2-D row views), out,
    B, Lq, Lk and optional key_mask (uint8 [B, Lk]), pair_dist, bias_affine, lse."""
    n = len(problems)
    arr = (_lib.AttnProblem * n)()
    dtype = None
    for a, pr in zip(arr, problems):
        q, k, v, o = pr['q'], pr['k'], pr['v'], pr['out']
        _, _, a.ldq = _rows2d(q, 'q')
        _, _, a.ldk = _rows2d(k, 'k')
        _, _, a.ldv = _rows2d(v, 'v')
        a.q, a.k, a.v, a.o, a.ldo = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), o.stride(0)
        km = pr.get('key_mask')
        if km is not None and (km.dtype != torch.uint8 or not km.is_contiguous()):
            raise _lib.VlnImagineError('key_mask must be a contiguous uint8 tensor')
        a.key_mask, a.pair_dist, a.bias_affine, a.lse = _ptr(km), _ptr(pr.get('pair_dist')), _ptr(pr.get('bias_affine')), _ptr(pr.get('lse'))
        a.B, a.Lq, a.Lk = pr['B'], pr['Lq'], pr['Lk']
        drop = pr.get('drop')                      # (p, site, seed tensor) or None: attention-probability dropout
        if drop is not None and drop[0] > 0:
            a.drop_p, a.drop_site, a.drop_seed = float(drop[0]), int(drop[1]) & 0xFFFFFFFF, drop[2].data_ptr()
        if dtype is None:
            dtype = q.dtype
        elif dtype != q.dtype:
            raise _lib.VlnImagineError('attention problems of one launch must share a dtype')
    check(lib.vi_attn_fwd_multi(arr, n, HEADS, _lib.DT_BF16 if dtype == BF16 else _lib.DT_F32, mask_mode, _stream()),
          'vi_attn_fwd_multi')
    _launched(1 if dtype == BF16 else n)


def add_ln(a: torch.Tensor, b: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, eps: float,
           want16: bool, want32: bool = True, group_row_end: Optional[Sequence[int]] = None):
    """LayerNorm(a [+ b]); returns (y32 or None, y16 or None)."""
    rows = a.shape[0]
    y32 = torch.empty((rows, HIDDEN), dtype=F32, device=a.device) if want32 else None
    y16 = torch.empty((rows, HIDDEN), dtype=BF16, device=a.device) if want16 else None
    n_groups = 1 if group_row_end is None else len(group_row_end)
    ends = _lib.int_array(list(group_row_end)) if group_row_end is not None else None
    check(lib.vi_add_ln(a.data_ptr(), _ptr(b), gamma.data_ptr(), beta.data_ptr(), eps, _ptr(y32), _ptr(y16), rows,
                        n_groups, ends, _stream()), 'vi_add_ln')
    _launched(1)
    return y32, y16


def embed_compose(rows: int, device, *, a=None, a2=None, a3=None, a_ln=None, feat=None, feat_w=None, feat_b=None, feat_ln=None,
                  idx=None, table=None, pos_table=None, pos_period=0, const_row=None, const_row2=None,
                  out_ln=None, eps=1e-12, y32=None, y16=None, want16=False, want32=True):
    """See vi_embed_compose.  *_ln are (gamma, beta) pairs.  y32 / y16 may be preallocated row views."""
    if y32 is None and want32:
        y32 = torch.empty((rows, HIDDEN), dtype=F32, device=device)
    if y16 is None and want16:
        y16 = torch.empty((rows, HIDDEN), dtype=BF16, device=device)
    args = _lib.EmbedArgs()
    args.a, args.a2, args.a3 = _ptr(a), _ptr(a2), _ptr(a3)
    if a_ln is not None:
        args.a_gamma, args.a_beta = a_ln[0].data_ptr(), a_ln[1].data_ptr()
    if feat is not None:
        args.feat, args.feat_dim = feat.data_ptr(), feat.shape[-1]
        args.feat_w, args.feat_b = feat_w.data_ptr(), _ptr(feat_b)
        if feat_ln is not None:
            args.feat_gamma, args.feat_beta = feat_ln[0].data_ptr(), feat_ln[1].data_ptr()
    if idx is not None:
        args.idx, args.table = idx.data_ptr(), table.data_ptr()
    if pos_table is not None:
        args.pos_table, args.pos_period = pos_table.data_ptr(), pos_period
    args.const_row, args.const_row2 = _ptr(const_row), _ptr(const_row2)
    if out_ln is not None:
        args.out_gamma, args.out_beta = out_ln[0].data_ptr(), out_ln[1].data_ptr()
    args.eps = eps
    args.y32, args.y16, args.rows = _ptr(y32), _ptr(y16), rows
    check(lib.vi_embed_compose(args, _stream()), 'vi_embed_compose')
    _launched(1)
    return y32, y16


def ln_dot(h: torch.Tensor, gamma, beta, eps: float, w, b, group_row_end: Optional[Sequence[int]] = None):
    rows = h.shape[0]
    out = torch.empty((rows,), dtype=F32, device=h.device)
    n_groups = 1 if group_row_end is None else len(group_row_end)
    ends = _lib.int_array(list(group_row_end)) if group_row_end is not None else None
    check(lib.vi_ln_dot(h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, w.data_ptr(), _ptr(b), out.data_ptr(),
                        rows, n_groups, ends, _stream()), 'vi_ln_dot')
    _launched(1)
    return out


def mul_bcast(x: torch.Tensor, x_batch_stride: int, s: torch.Tensor, s_batch_stride: int, n_batches: int,
              rows_per_batch: int, want16: bool, want32: bool = True):
    """y[b, r] = x[b, r] * s[b] over 768-wide rows; x / s are base views with element batch strides."""
    rows = n_batches * rows_per_batch
    y32 = torch.empty((rows, HIDDEN), dtype=F32, device=x.device) if want32 else None
    y16 = torch.empty((rows, HIDDEN), dtype=BF16, device=x.device) if want16 else None
    check(lib.vi_mul_bcast(x.data_ptr(), x_batch_stride, s.data_ptr(), s_batch_stride, _ptr(y32), _ptr(y16), rows,
                           rows_per_batch, _stream()), 'vi_mul_bcast')
    _launched(1)
    return y32, y16


def duet_fuse_logits(g_raw, l_raw, fuse_raw, gmap_masks_u8, gmap_visited_u8, vp_nav_u8, gmap_ids, cand_ids,
                     B: int, G: int, P: int):
    dev = g_raw.device
    gl = torch.empty((B, G), dtype=F32, device=dev)
    ll = torch.empty((B, P), dtype=F32, device=dev)
    fl = torch.empty((B, G), dtype=F32, device=dev)
    check(lib.vi_duet_fuse_logits(g_raw.data_ptr(), l_raw.data_ptr(), _ptr(fuse_raw), gmap_masks_u8.data_ptr(),
                                  gmap_visited_u8.data_ptr(), vp_nav_u8.data_ptr(), gmap_ids.data_ptr(),
                                  cand_ids.data_ptr(), gl.data_ptr(), ll.data_ptr(), fl.data_ptr(), B, G, P,
                                  _stream()), 'vi_duet_fuse_logits')
    _launched(1)
    return gl, ll, fl


def mask_logits_navtype(raw: torch.Tensor, nav_types: torch.Tensor):
    out = torch.empty_like(raw)
    check(lib.vi_mask_logits_navtype(raw.data_ptr(), nav_types.data_ptr(), out.data_ptr(), raw.numel(), _stream()),
          'vi_mask_logits_navtype')
    _launched(1)
    return out


def gather_mean(src: torch.Tensor, offsets: torch.Tensor, row_idx: torch.Tensor, R: int, want16: bool, want32: bool = True):
    y32 = torch.empty((R, HIDDEN), dtype=F32, device=src.device) if want32 else None
    y16 = torch.empty((R, HIDDEN), dtype=BF16, device=src.device) if want16 else None
    check(lib.vi_gather_mean(src.data_ptr(), offsets.data_ptr(), row_idx.data_ptr(), _ptr(y32), _ptr(y16), R, _stream()),
          'vi_gather_mean')
    _launched(1)
    return y32, y16


def scatter_rows(src: torch.Tensor, dst_rows: torch.Tensor, dst: torch.Tensor):
    check(lib.vi_scatter_rows(src.data_ptr(), dst_rows.data_ptr(), dst.data_ptr(), src.shape[0], _stream()), 'vi_scatter_rows')
    _launched(1)
    return dst


def cosine_loss(proj: Optional[torch.Tensor], tgt: Optional[torch.Tensor], R: int, device):
    loss = torch.empty((), dtype=F32, device=device)
    rows = torch.empty((max(R, 1),), dtype=F32, device=device)
    check(lib.vi_cosine_loss(_ptr(proj), _ptr(tgt), rows.data_ptr(), loss.data_ptr(), R, _stream()), 'vi_cosine_loss')
    _launched(2)
    return loss, rows[:R]


def infonce_loss(proj, tgt, negs, row_episode, neg_episode, temperature: float, R: int, n_negs: int, device):
    loss = torch.empty((), dtype=F32, device=device)
    scratch = torch.empty((max(R, 1) * (n_negs + 2),), dtype=F32, device=device)
    check(lib.vi_infonce_loss(_ptr(proj), _ptr(tgt), _ptr(negs), _ptr(row_episode), _ptr(neg_episode), temperature,
                              scratch.data_ptr(), loss.data_ptr(), R, n_negs, _stream()), 'vi_infonce_loss')
    _launched(3)
    return loss


def margin_loss(proj, tgt, negs, row_episode, neg_episode, margin: float, R: int, n_negs: int, device):
    loss = torch.empty((), dtype=F32, device=device)
    scratch = torch.empty((max(R, 1) * (n_negs + 2),), dtype=F32, device=device)
    check(lib.vi_margin_loss(_ptr(proj), _ptr(tgt), _ptr(negs), _ptr(row_episode), _ptr(neg_episode), margin,
                             scratch.data_ptr(), loss.data_ptr(), R, n_negs, _stream()), 'vi_margin_loss')
    _launched(3)
    return loss


def infonce_loss_with_sims(proj, tgt, negs, row_episode, neg_episode, temperature: float, R: int, n_negs: int, device):
    """infonce_loss that also hands back the forward scratch (its first R * (n_negs + 1) floats are the scaled
    similarities the backward pass needs)."""
    loss = torch.empty((), dtype=F32, device=device)
    scratch = torch.empty((max(R, 1) * (n_negs + 2),), dtype=F32, device=device)
    check(lib.vi_infonce_loss(_ptr(proj), _ptr(tgt), _ptr(negs), _ptr(row_episode), _ptr(neg_episode), temperature,
                              scratch.data_ptr(), loss.data_ptr(), R, n_negs, _stream()), 'vi_infonce_loss')
    _launched(3)
    return loss, scratch


def copy_rows(src: torch.Tensor, src_bs: int, src_rs: int, n_batches: int, rows_per_batch: int,
              dst32: Optional[torch.Tensor], dst16: Optional[torch.Tensor], dst_bs: int, dst_rs: int):
    """dst[b, r] = src[b, r] over 768-wide rows with element strides (see vi_copy_rows)."""
    check(lib.vi_copy_rows(src.data_ptr(), src_bs, src_rs, _ptr(dst32), _ptr(dst16), dst_bs, dst_rs, n_batches,
                           rows_per_batch, _stream()), 'vi_copy_rows')
    _launched(1)


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None):
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    check(lib.vi_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), 'vi_cast_bf16')
    _launched(1)
    return dst
