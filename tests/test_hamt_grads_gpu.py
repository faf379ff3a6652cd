"""HAMT-Imagine fine-tuning step on the GPU: forward + backward through the libvlnimagine kernels against gradient
fixtures of the REAL reference (tests/golden/hamt_grads_*.npz, ``oracle/gen_golden.py --model hamt --grads``): the 190
parameters that train in the released configuration (observation embeddings, the 4 cross-modal layers, the action head,
the imagination type embedding and the alignment head; language / history encoders frozen) are pinned by gradient norm
and 32 sampled elements.  Tolerances as in tests/test_duet_grads_gpu.py (bf16 mode: 1.5 x the reference's own autocast noise recorded in the fixture)."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from parity_utils import LOSS_TOL, check_gradients, golden, manifest, max_rel, to_dev

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
    check_gradients(net, gold, names, precision, 'hamt ' + tag)
