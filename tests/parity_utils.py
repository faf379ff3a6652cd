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


# ---- gradient parity (fine-tuning) ---------------------------------------------------------------------------------------------
# fp32 check mode: every parameter within 1e-3 (sampled elements and L2 norm).  bf16 mode: loss terms / logits within 2e-2; gradient
# element errors (median and max over parameters), the mean per-parameter cosine and the norm errors no worse than NOISE_MARGIN x
# what the UNMODIFIED REFERENCE shows under torch.autocast(bfloat16) on the same step against its own fp32 gradients - measured by
# oracle/gen_golden.autocast_noise with the same metrics and stored in every gradient fixture (autocast_*).  The scalar GASA slope
# (sprel_linear.weight, a sum with heavy cancellation) is held to 0.35 on its norm.
GRAD_TOL = {'fp32': 1e-3, 'bf16': 5e-2}
LOSS_TOL = {'fp32': 1e-4, 'bf16': 2e-2}
NOISE_MARGIN = 1.5
BF16_SCALAR_NORM = 0.35


def check_gradients(net, gold, names, precision, label):
    from oracle.gen_golden import grad_sample_index
    import numpy as np
    params = dict(net.named_parameters())
    with_grad = [n for n, p in params.items() if p.grad is not None]
    # parameters the reference leaves without a gradient (unused outputs of the last layer): the product may hand back exact zeros
    extra = [n for n in with_grad if n not in set(names)]
    assert [n for n in with_grad if n in set(names)] == names, (label, set(names) - set(with_grad))
    for n in extra:
        assert float(params[n].grad.abs().max()) == 0.0, (label, 'gradient for a parameter the reference does not train', n)
    tol = GRAD_TOL[precision]
    top = float(gold['grad_norms'].max())
    worst, dots = [], []
    for i, name in enumerate(names):
        g = params[name].grad
        assert torch.isfinite(g).all(), name
        ref_norm = float(gold['grad_norms'][i])
        got_norm = float(g.double().norm())
        if ref_norm < 1e-7 * top:
            assert got_norm < 1e-4 * top, (name, got_norm)
            continue
        idx = torch.from_numpy(grad_sample_index(name, g.numel())).cuda()
        got = g.reshape(-1)[idx].float().cpu()
        want = gold['grad_samples'][i]
        rms = ref_norm / np.sqrt(g.numel())
        scale = max(float(want.abs().max()), 3.0 * rms)
        err = float((got - want).abs().max()) / scale
        nerr = abs(got_norm - ref_norm) / ref_norm
        worst.append((max(err, nerr), name, err, nerr))
        if g.numel() >= 32:
            dots.append(float((got * want).sum() / (got.norm() * want.norm()).clamp_min(1e-30)))
    worst.sort(reverse=True)
    print('worst gradient errors (%s, %s):' % (label, precision))
    for w in worst[:5]:
        print('   %.3e  %s  (samples %.3e, norm %.3e)' % w)
    if precision == 'fp32':
        bad = [w for w in worst if w[0] >= tol]
        assert not bad, '%d of %d parameters outside %.0e: %s' % (len(bad), len(worst), tol, bad[:8])
    else:
        # bf16 mode: no worse than NOISE_MARGIN x what the unmodified reference shows under torch.autocast(bfloat16) on the same step
        # against its own fp32 gradients (recorded in the fixture by oracle/gen_golden.autocast_noise, same metrics)
        elem = np.array([w[2] for w in worst if not w[1].endswith('sprel_linear.weight')])
        cos = float(np.mean(dots))
        ac = {k: float(gold['autocast_' + k]) for k in ('elem_median', 'elem_max', 'norm_max', 'cosine')}
        print('   bf16: element error median %.3e max %.3e, mean per-parameter cosine %.5f  (reference under autocast: %.3e %.3e %.5f)'
              % (np.median(elem), elem.max(), cos, ac['elem_median'], ac['elem_max'], ac['cosine']))
        assert np.median(elem) < NOISE_MARGIN * ac['elem_median'] and elem.max() < NOISE_MARGIN * ac['elem_max']
        assert 1.0 - cos < NOISE_MARGIN * (1.0 - ac['cosine'])
        norm_tol = max(tol, NOISE_MARGIN * ac['norm_max'])
        # the scalar GASA slope: 0.35, or NOISE_MARGIN x the relative error the reference's own autocast run shows on it when that is
        # larger (DUET cfg-1: the autocast run gets the SIGN wrong, relative error 2.1)
        scalar_tol = {k[len('autocast_scalar::'):]: max(BF16_SCALAR_NORM, NOISE_MARGIN * float(v)) for k, v in gold.items()
                      if k.startswith('autocast_scalar::')}
        bad = [w for w in worst if w[3] >= (scalar_tol.get(w[1], BF16_SCALAR_NORM) if w[1].endswith('sprel_linear.weight') else norm_tol)]
        assert not bad, '%d of %d parameter norms outside tolerance: %s' % (len(bad), len(worst), bad[:8])


