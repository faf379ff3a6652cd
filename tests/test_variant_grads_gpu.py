"""Fine-tuning of the secondary variants on the GPU against gradient fixtures of the REAL reference
(``oracle/gen_golden.py --model hamt_variants --grads`` / ``--model duet_reverie --grads``):

* HAMT-Imagine with the parser-default imagination flags - the ImagineEmbeddings encoder trained, imagination tokens on the vision
  stream, ``act_pred_token='ob_imagine_text'`` (H/models/vilmodel_cmt.py:634-703,1106-1134,1197-1198);
* the margin form of the alignment loss with its backward pass and the ``ob_txt_hist`` action token (:825-856,1195-1196);
* trainable history embeddings (``fix_hist_embedding`` off, :576-618,1036) with the ``ob_hist`` action token;
* DUET-Imagine's REVERIE recipe: object boxes through the panorama encoder and the object-grounding head
  (D/models/vilmodel.py:1096-1131,1220-1225; reverie/agent_obj.py:461-463).

Tolerances: fp32 check mode 1e-3 on every parameter (tests/test_duet_grads_gpu.py); bf16 mode: element median / max, cosine and
norm errors no worse than 1.5 x the figures of the unmodified reference under torch.autocast(bfloat16) on the same step, which the
fixture records next to the fp32 gradients."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from parity_utils import check_gradients, golden, max_rel, to_dev
from parity_utils import LOSS_TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['encvis_imgtxt', 'margin_txthist', 'hist_obhist'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_hamt_variant_gradients(lib_built, tag, precision):
    from oracle.gen_golden import HAMT_GRAD_VARIANTS, hamt_variant_train_step
    from oracle import hamt_oracle as O
    synth = importlib.import_module('vln_imagine_b200.synth')
    hamt = importlib.import_module('vln_imagine_b200.hamt')
    config = importlib.import_module('vln_imagine_b200.config')
    over = HAMT_GRAD_VARIANTS[tag]
    model = hamt.VLNBertCMT(config.default_hamt_args(**over)).cuda().eval()
    net = model.vln_bert
    man = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth.synth_state_dict(man, seed=0))
    net.precision = precision
    net.zero_grad(set_to_none=True)
    ep = to_dev(synth.to_torch(synth.hamt_episode(synth.TINY, 7)))
    hm = O.hist_masks_from_lens(ep['hist_lens'].cpu(), ep['hist_embeds'].shape[1]).cuda()
    loss, ce, aux, logits = hamt_variant_train_step(lambda mode='history', **kw: net(mode, **kw), ep, hm, over)
    loss.backward()
    torch.cuda.synchronize()
    gold = golden('hamt_grads_' + tag)
    with open(os.path.join(GOLDEN, 'hamt_grads_variant_names.json')) as f:
        names = json.load(f)[tag]
    lt = LOSS_TOL[precision]
    assert max_rel(logits, gold['act_logits']) < lt
    for k, v in (('loss', loss), ('ce', ce), ('aux', aux)):
        assert abs(float(v) - float(gold[k])) < lt * abs(float(gold[k])), k
    check_gradients(net, gold, names, precision, 'hamt ' + tag)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_duet_reverie_gradients(lib_built, precision):
    from oracle.gen_golden import duet_reverie_train_step
    synth = importlib.import_module('vln_imagine_b200.synth')
    duet = importlib.import_module('vln_imagine_b200.duet')
    config = importlib.import_module('vln_imagine_b200.config')
    model = duet.VLNBert(config.default_duet_args(dataset='reverie', obj_feat_size=768)).cuda().eval()
    net = model.vln_bert
    with open(os.path.join(GOLDEN, 'duet_reverie_manifest.json')) as f:
        man = json.load(f)
    net.load_state_dict(synth.synth_state_dict(man, seed=0, gasa_stress=True))
    net.precision = precision
    net.zero_grad(set_to_none=True)
    ep = to_dev(synth.to_torch(synth.duet_reverie_episode(synth.TINY, 7)))
    loss, ce, og, aux, nav = duet_reverie_train_step(lambda mode, batch: model(mode, batch), ep)
    loss.backward()
    torch.cuda.synchronize()
    gold = golden('duet_reverie_grads_tiny')
    with open(os.path.join(GOLDEN, 'duet_reverie_grads_names.json')) as f:
        names = json.load(f)
    lt = LOSS_TOL[precision]
    assert max_rel(nav['fused_logits'], gold['fused_logits']) < lt
    assert max_rel(nav['obj_logits'], gold['obj_logits']) < lt
    for k, v in (('loss', loss), ('ce', ce), ('og', og), ('aux', aux)):
        assert abs(float(v) - float(gold[k])) < lt * abs(float(gold[k])), k
    check_gradients(net, gold, names, precision, 'duet reverie')


@pytest.mark.parametrize('tag', ['unfixlang'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_duet_variant_gradients(lib_built, tag, precision):
    """DUET-Imagine with the text encoder trained through the alignment loss as well (fix_lang_inside_cosine_model off, the parser
    default; D/models/vilmodel.py:1256-1262): the noun-phrase token means carry gradients (ragged-mean adjoint = scatter-add of
    dy / len), which changes the gradient of 149 language-encoder parameters against the released recipe"""
    from oracle.gen_golden import DUET_GRAD_VARIANTS, duet_train_step
    synth = importlib.import_module('vln_imagine_b200.synth')
    duet = importlib.import_module('vln_imagine_b200.duet')
    config = importlib.import_module('vln_imagine_b200.config')
    from parity_utils import manifest
    model = duet.VLNBert(config.default_duet_args(**DUET_GRAD_VARIANTS[tag])).cuda().eval()
    net = model.vln_bert
    net.load_state_dict(synth.synth_state_dict(manifest('duet'), seed=0, gasa_stress=True))
    net.precision = precision
    net.zero_grad(set_to_none=True)
    ep = to_dev(synth.to_torch(synth.duet_episode(synth.TINY, 7)))
    loss, ce, aux, nav = duet_train_step(net, ep, lambda mode, batch: model(mode, batch))
    loss.backward()
    torch.cuda.synchronize()
    gold = golden('duet_grads_' + tag)
    with open(os.path.join(GOLDEN, 'duet_grads_variant_names.json')) as f:
        names = json.load(f)[tag]
    lt = LOSS_TOL[precision]
    assert max_rel(nav['fused_logits'], gold['fused_logits']) < lt
    for k, v in (('loss', loss), ('ce', ce), ('aux', aux)):
        assert abs(float(v) - float(gold[k])) < lt * abs(float(gold[k])), k
    check_gradients(net, gold, names, precision, 'duet ' + tag)
