"""Shared helpers of the parity tests: tolerances from BASELINE.json's north_star and the comparison rules.

bf16 mode : logits / aux loss within 2e-2 (max-norm relative: max|d| / max|ref| over finite entries - the
            element-wise form is meaningless for logits that cross zero), identical -inf
            pattern, action argmax identical on >= 99.5 % of decisions.
fp32 mode : the same quantities within 1e-4.
"""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN

TOL = {'bf16': 2e-2, 'fp32': 1e-4}


def manifest(model):
    with open(os.path.join(GOLDEN, '%s_manifest.json' % model)) as f:
        return json.load(f)


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def max_rel(a, b):
    """max|a-b| / max|b| over finite entries; asserts the inf patterns agree."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    fa, fb = torch.isfinite(a), torch.isfinite(b)
    assert torch.equal(fa, fb), 'the -inf / nan pattern differs'
    assert torch.equal(a[~fa], b[~fb]), 'non-finite entries differ'
    if fb.sum() == 0:
        return 0.0
    return float((a[fb] - b[fb]).abs().max() / b[fb].abs().max().clamp_min(1e-30))


def argmax_agreement(a, b):
    return float((a.detach().cpu().argmax(-1) == b.detach().cpu().argmax(-1)).float().mean())


def sub16(t):
    """the committed cfg-1 fixtures keep every 16th hidden unit of (B, N, 768) tensors (oracle/gen_golden.py)"""
    return t[..., ::16]


def to_dev(ep, device='cuda'):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in ep.items()}


def argmax_report(product_logits, ref_logits, tol):
    """Action-argmax agreement of one batch.  Returns (agree, total, decisive_mismatches).  A decision is
    *decisive* when the reference's top-1 / top-2 gap exceeds the logit tolerance (tol x max|finite logit|): only
    there is "identical argmax" implied by "logits within tolerance".  Random-init logits are nearly flat
    (p10 of the top-2 gap is 2 % of the logit range), so a few non-decisive flips per thousand
    decisions are inherent to ANY bf16 evaluation, including torch.autocast of the reference itself."""
    p, r = product_logits.detach().float().cpu(), ref_logits.detach().float().cpu()
    a, b = p.argmax(-1), r.argmax(-1)
    top2 = r.topk(2, -1).values
    scale = r[torch.isfinite(r)].abs().max()
    gap = top2[:, 0] - top2[:, 1]                      # +inf when only one action is admissible
    decisive = gap >= tol * scale
    return int((a == b).sum()), a.numel(), int(((a != b) & decisive).sum())
