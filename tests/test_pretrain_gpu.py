"""SURVEY.md 8(f) N4 - the DUET pre-training forward on the GPU (libvlnimagine kernels behind the reference's
``GlocalTextPathCMTPreTraining.forward(batch, task, compute_loss)``) against the golden vectors of the REAL reference
(tests/golden/duet_pretrain.npz, written by oracle/gen_golden.py --model duet_pretrain) and against the CPU oracle on a second,
larger batch: sap logits, mlm prediction scores over the 30522-word vocabulary, mrc logits, both losses.
fp32 check mode 1e-4, 16-bit mode 2e-2 (max-norm relative, identical -inf pattern)."""
import importlib
import json
import os

import pytest
import torch

from conftest import GOLDEN
from parity_utils import TOL, golden, manifest, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env(lib_built):
    synth = importlib.import_module('vln_imagine_b200.synth')
    pretrain = importlib.import_module('vln_imagine_b200.pretrain')
    config = importlib.import_module('vln_imagine_b200.config')
    model = pretrain.GlocalTextPathCMTPreTraining(config.duet_config(config.default_duet_args())).cuda().eval()
    sd = synth.synth_state_dict(manifest('duet_pretrain'), seed=0)
    sd['mlm_head.predictions.decoder.weight'] = sd['bert.embeddings.word_embeddings.weight']        # tied (pretrain_cmt.py:112-117)
    sd['bert.global_encoder.sprel_linear.weight'] = torch.full((1, 1), -0.3)                        # as the golden run: a real GASA bias
    model.load_state_dict(sd)
    return synth, model, sd


def test_parameter_tree_matches_the_reference():
    pretrain = importlib.import_module('vln_imagine_b200.pretrain')
    config = importlib.import_module('vln_imagine_b200.config')
    model = pretrain.GlocalTextPathCMTPreTraining(config.duet_config(config.default_duet_args()))
    names = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert names == json.load(open(os.path.join(GOLDEN, 'duet_pretrain_manifest.json')))
    assert model.mlm_head.predictions.decoder.weight is model.bert.embeddings.word_embeddings.weight


def _run(model, ep):
    with torch.no_grad():
        gl, ll, fl = model(ep, 'sap', compute_loss=False)
        scores = model(ep, 'mlm', compute_loss=False)
        mlm_loss = model(ep, 'mlm', compute_loss=True)
        logits, targets = model(ep, 'mrc', compute_loss=False)
        mrc_loss = model(ep, 'mrc', compute_loss=True)
    return dict(global_logits=gl, local_logits=ll, fused_logits=fl, mlm_scores=scores, mlm_loss=mlm_loss, mrc_logits=logits,
                mrc_loss=mrc_loss)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_pretraining_forward_vs_reference_golden(env, precision):
    synth, model, _ = env
    model.bert.precision = precision
    out = _run(model, synth.to_torch(synth.duet_pretrain_batch()))
    gold = golden('duet_pretrain')
    assert out['mlm_scores'].shape[1] == 30522
    out['mlm_scores'] = out['mlm_scores'][:, ::64]
    for k, v in out.items():
        assert max_rel(v, gold[k]) < TOL[precision], k


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_pretraining_forward_vs_oracle_larger_batch(env, precision):
    from oracle import pretrain_oracle as P
    synth, model, sd = env
    model.bert.precision = precision
    ep = synth.to_torch(synth.duet_pretrain_batch(seed=11, batch=6, n_vp=30, max_steps=6, n_views=36, instr_len=60))
    with torch.no_grad():
        gl, ll, fl = P.forward_sap(sd, ep)
        scores = P.forward_mlm(sd, ep)
        logits, _ = P.forward_mrc(sd, ep)
        mlm, mrc = P.losses(sd, ep)
    ref = dict(global_logits=gl, local_logits=ll, fused_logits=fl, mlm_scores=scores, mlm_loss=mlm, mrc_logits=logits, mrc_loss=mrc)
    out = _run(model, ep)
    for k, v in out.items():
        assert max_rel(v, ref[k]) < TOL[precision], k
    assert torch.equal(out['fused_logits'].cpu().argmax(-1), fl.argmax(-1)) or precision == 'bf16'
