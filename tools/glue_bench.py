"""A/B of the host side of DeviceGraphMaps on the GPU box: python tools/glue_bench.py [path-to-alternative-graph_map.py]"""
import importlib
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vln_imagine_b200.synth as synth  # noqa: E402

mods = {'current': importlib.import_module('vln_imagine_b200.graph_map')}
for p in sys.argv[1:]:
    spec = importlib.util.spec_from_file_location('vln_imagine_b200.gm_alt', p)
    m = importlib.util.module_from_spec(spec)
    sys.modules['vln_imagine_b200.gm_alt'] = m
    spec.loader.exec_module(m)
    mods[os.path.basename(p)] = m
dev = torch.device('cuda')
world = synth.nav_world(seed=21, n_vp=80, batch=64, steps=7, hidden=768, degree=4)
for st in world:
    st['pano_d'] = torch.from_numpy(st['pano_embeds']).to(dev)
    st['masks_d'] = torch.ones(st['pano_d'].shape[:2], dtype=torch.bool, device=dev)
    st['view_lens_d'] = torch.from_numpy(st['view_lens']).to(dev)
    st['nav_types_d'] = torch.from_numpy(st['nav_types']).to(dev)
for name, gmod in mods.items():
    T = {'init': 0.0, 'nav': 0.0, 'upd': 0.0, 'sync': 0.0}

    def rollout():
        t0 = time.perf_counter()
        gm = gmod.DeviceGraphMaps(world[0]['obs'], dev)
        T['init'] += time.perf_counter() - t0
        for t, st in enumerate(world):
            obs, ended = st['obs'], st['ended']
            gm.set_step_ids(obs, t, ended)
            pin = {'cand_vpids': [[c['viewpointId'] for c in ob['candidate']] for ob in obs], 'view_lens': st['view_lens_d'],
                   'nav_types': st['nav_types_d']}
            t0 = time.perf_counter()
            gm.nav_inputs(obs, st['pano_d'], st['masks_d'], pin, ended)
            T['nav'] += time.perf_counter() - t0
            if t + 1 < len(world):
                t0 = time.perf_counter()
                gm.update_graph(world[t + 1]['obs'], ended)
                T['upd'] += time.perf_counter() - t0
        t0 = time.perf_counter()
        torch.cuda.synchronize()
        T['sync'] += time.perf_counter() - t0
    for _ in range(3):
        rollout()
    for k in T:
        T[k] = 0.0
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        rollout()
    tot = (time.perf_counter() - t0) / n / len(world) * 1e3
    print('%-14s total %.3f ms/step: %s' % (name, tot, {k: round(v / n / len(world) * 1e3, 3) for k, v in T.items()}))
    # device time of the three kernels alone
    gm = gmod.DeviceGraphMaps(world[0]['obs'], dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = world[0]
    pin = {'cand_vpids': [[c['viewpointId'] for c in ob['candidate']] for ob in st['obs']], 'view_lens': st['view_lens_d'], 'nav_types': st['nav_types_d']}
    e0.record()
    gm.nav_inputs(st['obs'], st['pano_d'], st['masks_d'], pin, st['ended'])
    e1.record()
    torch.cuda.synchronize()
    print('               one nav_inputs call, device span %.3f ms' % e0.elapsed_time(e1))
