"""HAMT-Imagine fine-tuning step on the GPU: forward + backward through the libvlnimagine kernels against gradient
fixtures of the REAL reference (tests/golden/hamt_grads_*.npz, ``oracle/gen_golden.py --model hamt --grads``): the 190
parameters that train in the released configuration (observation embeddings, the 4 cross-modal layers, the action head,
the imagination type embedding and the alignment head; language / history encoders frozen) are pinned by gradient norm
and 32 sampled elements.  Tolerances as in tests/test_duet_grads_gpu.py."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from parity_utils import golden, manifest, max_rel, to_dev
from test_duet_grads_gpu import BF16_COSINE, BF16_ELEM_MAX, BF16_ELEM_MEDIAN, GRAD_TOL, LOSS_TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env(lib_built):
    synth = importlib.import_module('vln_imagine_b200.synth')
    hamt = importlib.import_module('vln_imagine_b200.hamt')
    config = importlib.import_module('vln_imagine_b200.config')
    model = hamt.VLNBertCMT(config.default_hamt_args()).cuda().eval()
    return synth, model


@pytest.mark.parametrize('tag,shape,seed', [('tiny', 'TINY', 7), ('cfg1', 'CFG1', 1234)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_hamt_train_step_gradients(env, tag, shape, seed, precision):
    from oracle.gen_golden import grad_sample_index, hamt_train_step
    from oracle import hamt_oracle as O
    synth, model = env
    net = model.vln_bert
    net.load_state_dict(synth.synth_state_dict(manifest('hamt'), seed=0))
    net.precision = precision
    net.zero_grad(set_to_none=True)
    ep = to_dev(synth.to_torch(synth.hamt_episode(getattr(synth, shape), seed)))
    hm = O.hist_masks_from_lens(ep['hist_lens'].cpu(), ep['hist_embeds'].shape[1]).cuda()
    loss, ce, aux, logits = hamt_train_step(lambda mode, **kw: net(mode, **kw), ep, hm)
    loss.backward()
    torch.cuda.synchronize()
    gold = golden('hamt_grads_' + tag)
    with open(os.path.join(GOLDEN, 'hamt_grads_names.json')) as f:
        names = json.load(f)
    lt = LOSS_TOL[precision]
    assert max_rel(logits, gold['act_logits']) < lt
    for k, v in (('loss', loss), ('ce', ce), ('aux', aux)):
        assert abs(float(v) - float(gold[k])) < lt * abs(float(gold[k])), k
    params = dict(net.named_parameters())
    with_grad = [n for n, p in params.items() if p.grad is not None]
    assert with_grad == names, (set(with_grad) ^ set(names))
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
    print('worst gradient errors (hamt %s, %s):' % (tag, precision))
    for w in worst[:5]:
        print('   %.3e  %s  (samples %.3e, norm %.3e)' % w)
    if precision == 'fp32':
        bad = [w for w in worst if w[0] >= tol]
        assert not bad, '%d of %d parameters outside %.0e: %s' % (len(bad), len(worst), tol, bad[:8])
    else:
        elem = np.array([w[2] for w in worst])
        cos = float(np.mean(dots))
        print('   bf16: element error median %.3e max %.3e, mean per-parameter cosine %.5f' % (np.median(elem), elem.max(), cos))
        assert np.median(elem) < BF16_ELEM_MEDIAN and elem.max() < BF16_ELEM_MAX
        assert cos > BF16_COSINE
        bad = [w for w in worst if w[3] >= tol]
        assert not bad, '%d of %d parameter norms outside tolerance: %s' % (len(bad), len(worst), bad[:8])
