"""Time every GEMM shape of the DUET cfg-2 / HAMT cfg-3 step in isolation, for each tile width the kernel offers.
Usage: python tools/gemm_sweep.py [bn ...]   (default: auto 64 128 256)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import ops  # noqa: E402

ops.ensure_init(torch.zeros(1, device='cuda'))
# (name, M, N, K, groups(row ends) or None, epilogue, residual, out fp32)
SHAPES = [
    ('pano.img_linear', 2304, 768, 768, None, 0, False, True),
    ('pano.qkv', 2304, 2304, 768, None, 0, False, False),
    ('pano.o', 2304, 768, 768, None, 0, True, True),
    ('pano.ffn1', 2304, 3072, 768, None, 1, False, False),
    ('pano.ffn2', 2304, 768, 3072, None, 0, True, True),
    ('nav.kv', 5440, 3072, 768, None, 0, False, False),
    ('nav.q', 4288, 768, 768, [1920, 4288], 0, False, False),
    ('nav.o', 4288, 768, 768, [1920, 4288], 0, True, True),
    ('nav.qkv', 4288, 2304, 768, [1920, 4288], 0, False, False),
    ('nav.ffn1', 4288, 3072, 768, [1920, 4288], 1, False, False),
    ('nav.ffn2', 4288, 768, 3072, [1920, 4288], 0, True, True),
    ('nav.head', 4288, 768, 768, [1920, 4288], 2, False, True),
    ('hamt.x_qkv', 8832, 2304, 768, None, 0, False, False),
    ('hamt.ffn1', 8832, 3072, 768, [5440, 8832], 1, False, False),
    ('hamt.ffn2', 8832, 768, 3072, [5440, 8832], 0, True, True),
    ('lang.ffn1', 5120, 3072, 768, None, 1, False, False),
]
bns = sys.argv[1:] or ['auto', '64', '128', '256']
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
rows = []
for name, M, N, K, ends, epi, res, f32 in SHAPES:
    ng = 1 if ends is None else len(ends)
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(ng * N, K, device='cuda') * 0.05).bfloat16()
    b = torch.randn(ng * N, device='cuda')
    r = torch.randn(M, N, device='cuda') if res else None
    out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device='cuda')
    rec = {'name': name, 'M': M, 'N': N, 'K': K, 'gflop': 2.0 * M * N * K / 1e9}
    for bn in bns:
        if bn == 'auto':
            os.environ.pop('VI_GEMM_BN', None)
        else:
            os.environ['VI_GEMM_BN'] = bn
        for _ in range(3):
            ops.gemm(x, w, b, residual=r, epilogue=epi, out=out, group_row_end=ends)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(x, w, b, residual=r, epilogue=epi, out=out, group_row_end=ends)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        rec['us_' + bn] = round(t * 1e3, 1)
        rec['tf_' + bn] = round(rec['gflop'] / t, 1)
    rows.append(rec)
    print(json.dumps(rec), flush=True)
