"""Launch the DUET cfg-2 cross-attention (2 streams) a few times, timed; for ncu.  python tools/attn_one.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vln_imagine_b200 import ops  # noqa: E402

ops.ensure_init(torch.zeros(1, device='cuda'))
B, G, P, C = 64, 30, 37, 85
q = torch.randn(4416, 768, device='cuda').bfloat16()
kv = torch.randn(B * C, 3072, device='cuda').bfloat16()
ctx = torch.empty(4416, 768, device='cuda', dtype=torch.bfloat16)
mask = torch.ones(B, C, dtype=torch.uint8, device='cuda')
mask[:, 70:] = 0
probs = [dict(q=q[:B * G], k=kv[:, 0:768], v=kv[:, 768:1536], out=ctx[:B * G], B=B, Lq=G, Lk=C, key_mask=mask),
         dict(q=q[2048:2048 + B * P], k=kv[:, 1536:2304], v=kv[:, 2304:3072], out=ctx[2048:2048 + B * P], B=B, Lq=P, Lk=C, key_mask=mask)]
for _ in range(3):
    ops.attention_multi(probs)
torch.cuda.synchronize()
torch.cuda._sleep(2_000_000)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention_multi(probs)
e1.record()
torch.cuda.synchronize()
print('cross-attention, 2 streams: %.2f us per launch' % (e0.elapsed_time(e1) / 20 * 1e3))
# self-attention of the same layer: global (GASA distances) | local
qkv = torch.randn(4416, 2304, device='cuda').bfloat16()
dist = torch.rand(B, G, G, device='cuda') * 20
aff = torch.tensor([-0.5, 0.0], device='cuda')
gm = torch.ones(B, G, dtype=torch.uint8, device='cuda'); gm[:, 25:] = 0
pm = torch.ones(B, P, dtype=torch.uint8, device='cuda')
sp = [dict(q=qkv[:B * G, 0:768], k=qkv[:B * G, 768:1536], v=qkv[:B * G, 1536:2304], out=ctx[:B * G], B=B, Lq=G, Lk=G, key_mask=gm,
           pair_dist=dist, bias_affine=aff),
      dict(q=qkv[2048:2048 + B * P, 0:768], k=qkv[2048:2048 + B * P, 768:1536], v=qkv[2048:2048 + B * P, 1536:2304],
           out=ctx[2048:2048 + B * P], B=B, Lq=P, Lk=P, key_mask=pm)]
for _ in range(3):
    ops.attention_multi(sp)
torch.cuda.synchronize()
torch.cuda._sleep(2_000_000)
e0.record()
for _ in range(20):
    ops.attention_multi(sp)
e1.record()
torch.cuda.synchronize()
print('self-attention, 2 streams: %.2f us per launch' % (e0.elapsed_time(e1) / 20 * 1e3))
