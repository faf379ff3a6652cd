"""Episode prelude + two navigation steps, eager launches (for ncu: the kernels of the LAST step are the ones
summarised in profiles/).  python tools/step_once.py [workload]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'duet_cfg2'
kind, shape, _ = bench.workload(wl)
with torch.no_grad():
    model, ep = bench.build(kind, shape, 0, 'bf16')
    model.use_cuda_graphs = False
    d = bench.device_inputs(kind, model, ep, torch.device('cuda'))
    txt, img2, _ = bench.episode_prelude(kind, model, d)
    step = bench.duet_step if kind == 'duet' else bench.hamt_step
    step(model, d, txt, img2)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only the kernels of this last step are captured
    step(model, d, txt, img2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('ok')
