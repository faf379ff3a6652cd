"""Measure the fastest GEMM tile for every GEMM signature of the benchmark workloads on this GPU and write
vln-imagine_b200/tile_table.json.  Run on a B200:  python tools/tune_tiles.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['VLN_IMAGINE_RETUNE'] = '1'
os.environ['VLN_IMAGINE_AUTOTUNE'] = '1'
import bench  # noqa: E402
from vln_imagine_b200 import ops  # noqa: E402

with torch.no_grad():
    for wl in ('duet_cfg2', 'hamt_cfg3', 'duet_cfg5'):
        kind, shape, _ = bench.workload(wl)
        model, ep = bench.build(kind, shape, 0, 'bf16')
        model.use_cuda_graphs = False
        d = bench.device_inputs(kind, model, ep, torch.device('cuda'))
        txt, img2, _ = bench.episode_prelude(kind, model, d)
        (bench.duet_step if kind == 'duet' else bench.hamt_step)(model, d, txt, img2)
        torch.cuda.synchronize()
        del model
table = dict(ops._TILE_TABLE)            # keep the signatures measured earlier, add / refresh the ones of this run
table.update({ops._key_str(k): v for k, v in ops._TILE_CACHE.items() if v})
path = os.path.join(ROOT, 'vln-imagine_b200', 'tile_table.json')
with open(path, 'w') as f:
    json.dump(table, f, indent=0, sort_keys=True)
print('wrote %d signatures to %s' % (len(table), path))
print(json.dumps(table))
