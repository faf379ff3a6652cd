"""Time vi_gemm_ln_bf16 against the GEMM + LayerNorm pair it replaces (CUDA events, back-to-back launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import ops  # noqa: E402

ops.ensure_init(torch.zeros(1, device='cuda'))


def timeit(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(20_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M, K, ends in [(4416, 768, [2048, 4416]), (4416, 3072, [2048, 4416]), (2304, 768, None), (5120, 768, None)]:
    ng = 1 if ends is None else 2
    x = torch.randn(M, K, device='cuda').bfloat16()
    w = (torch.randn(ng * 768, K, device='cuda') * 0.05).bfloat16()
    b = torch.randn(ng * 768, device='cuda')
    res = torch.randn(M, 768, device='cuda')
    g = torch.ones(ng * 768, device='cuda')
    be = torch.zeros(ng * 768, device='cuda')
    t_f = timeit(lambda: ops.gemm_ln(x, w, b, res, g, be, 1e-12, group_row_end=ends))

    def pair():
        ao = ops.gemm(x, w, b, residual=res, out_dtype=torch.float32, group_row_end=ends)
        ops.add_ln(ao, None, g.view(ng, 768), be.view(ng, 768), 1e-12, want16=True, group_row_end=ends)
    t_p = timeit(pair)
    print('M=%d K=%d groups=%d: fused %.1f us, gemm+ln %.1f us (VI_RB_DEBUG=%s)' % (M, K, ng, t_f, t_p, os.environ.get('VI_RB_DEBUG', '0')))
