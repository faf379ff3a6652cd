"""Validate and time the experimental multicast-cluster GEMM (vi_gemm_bf16_mc) against the default tcgen05 GEMM.
    python tools/gemm_mc_check.py          # prints one line per shape: max |diff| vs ops.gemm, error vs torch, both timings"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import ops  # noqa: E402

ops.ensure_init(torch.zeros(1, device='cuda'))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(4_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if '--gelu' in sys.argv:
    # what does the erf-GELU epilogue cost?  Same kernel, same shape, epilogue none / erf form / one-MUFU tanh form
    M, N, K = 4416, 3072, 768
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(N, K, device='cuda') * 0.05).bfloat16()
    b = torch.randn(N, device='cuda')
    ys = {e: ops.gemm_mc(x, w, b, epilogue=e, tile=256) for e in (0, 1, 3)}
    ts = {e: timeit(lambda e=e: ops.gemm_mc(x, w, b, epilogue=e, tile=256)) for e in (0, 1, 3)}
    d = (ys[1].float() - ys[3].float()).abs()
    print('FFN1 shape, multicast kernel: no activation %.1f us, erf GELU %.1f us, tanh-form GELU %.1f us; erf vs tanh outputs: '
          'max |diff| %.3g, %.2f %% of the bf16 outputs differ' % (ts[0], ts[1], ts[3], float(d.max()), 100 * float((d > 0).float().mean())))
    sys.exit(0)

CASES = [  # name, M, N, K, epilogue, residual, f32 out, row groups, tile
    ('odd sizes', 300, 256, 128, 0, False, False, None, 128),
    ('one row', 1, 128, 64, 2, False, True, None, 128),
    ('nav.ffn1', 4416, 3072, 768, 1, False, False, [2048, 4416], 256),
    ('nav.qkv', 4416, 2304, 768, 0, False, False, [2048, 4416], 192),
    ('nav.ffn2', 4416, 768, 3072, 0, True, True, [2048, 4416], 192),
    ('nav.kv', 5440, 3072, 768, 0, False, False, None, 256),
    ('hamt.ffn1', 9024, 3072, 768, 1, False, False, None, 256),
]
ok = True
for name, M, N, K, epi, res, f32, ends, tile in CASES:
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    ng = 1 if ends is None else len(ends)
    x = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    w = (torch.randn(ng * N, K, device='cuda', generator=g) * 0.05).bfloat16()
    b = torch.randn(ng * N, device='cuda', generator=g)
    r = torch.randn(M, N, device='cuda', generator=g) if res else None
    dt = torch.float32 if f32 else torch.bfloat16
    y_mc = ops.gemm_mc(x, w, b, residual=r, epilogue=epi, out_dtype=dt, group_row_end=ends, tile=tile)
    y_tc = ops.gemm(x, w, b, residual=r, epilogue=epi, out_dtype=dt, group_row_end=ends)
    torch.cuda.synchronize()
    bounds = [0] + (ends or [M])
    ref = torch.cat([F.linear(x[bounds[i]:bounds[i + 1]].float(), w[i * N:(i + 1) * N].float(), b[i * N:(i + 1) * N]) for i in range(ng)])
    ref = F.gelu(ref) if epi == 1 else (F.relu(ref) if epi == 2 else ref)
    if res:
        ref = ref + r
    d_tc = float((y_mc.float() - y_tc.float()).abs().max())
    e_ref = float((y_mc.float() - ref).abs().max() / ref.abs().max())
    t_mc = timeit(lambda: ops.gemm_mc(x, w, b, residual=r, epilogue=epi, out_dtype=dt, group_row_end=ends, tile=tile))
    t_tc = timeit(lambda: ops.gemm(x, w, b, residual=r, epilogue=epi, out_dtype=dt, group_row_end=ends))
    good = e_ref < 2e-2 and d_tc <= 1e-2 * float(ref.abs().max())
    ok &= good
    print('%-10s M=%5d N=%4d K=%4d tile %3d: max|mc - tc| %.3g, rel err vs torch %.2e, mc %.1f us, default %.1f us (%.0f TF/s vs %.0f) %s'
          % (name, M, N, K, tile, d_tc, e_ref, t_mc, t_tc, 2.0 * M * N * K / t_mc / 1e6, 2.0 * M * N * K / t_tc / 1e6, 'ok' if good else 'MISMATCH'))
print('ALL OK' if ok else 'FAILED')
