"""Pin oracle/graph_oracle.py to the REAL reference GraphMap / FloydGraph and write tests/golden/graph_world.npz.

Runs in the build container only (needs /root/reference):  python oracle/gen_graph_golden.py

The reference classes (VLN-DUET/map_nav_src/models/graph_utils.py, numpy only) are imported unmodified and driven
through a synthetic rollout (vln_imagine_b200.synth.nav_world) exactly as r2r/agent.py:386-606 drives them; the agent's
own collate functions (_nav_gmap_variable / _nav_vp_variable, agent.py:98-207) cannot be imported (agent.py needs
MatterSim) and are restated inline here, line by line, on top of the real GraphMap.  The oracle must reproduce every
tensor bit for bit; the fixture holds the reference's outputs.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import graph_oracle as GO  # noqa: E402
from vln_imagine_b200 import synth  # noqa: E402

REF = '/root/reference/VLN-DUET/map_nav_src/models/graph_utils.py'
HIDDEN = 8
WORLD = dict(seed=7, n_vp=24, batch=4, steps=6, n_views=36, hidden=HIDDEN)


def load_reference():
    spec = importlib.util.spec_from_file_location('ref_graph_utils', REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_gmap_variable(gmaps, obs):
    """agent.py:98-171 (enc_full_graph=True, act_visited_nodes=False) on the real GraphMap objects."""
    rows = []
    for i, gmap in enumerate(gmaps):
        visited, unvisited = [], []
        for k in gmap.node_positions.keys():
            (visited if gmap.graph.visited(k) else unvisited).append(k)
        vpids = [None] + visited + unvisited
        emb = [gmap.get_node_embed(vp) for vp in vpids[1:]]
        emb = torch.stack([torch.zeros_like(emb[0])] + emb, 0).numpy()
        pos = gmap.get_pos_fts(obs[i]['viewpoint'], vpids, obs[i]['heading'], obs[i]['elevation'])
        pair = np.zeros((len(vpids), len(vpids)), dtype=np.float32)
        for a in range(1, len(vpids)):
            for b in range(a + 1, len(vpids)):
                pair[a, b] = pair[b, a] = gmap.graph.distance(vpids[a], vpids[b])
        rows.append(dict(vpids=vpids, visited=[0] + [1] * len(visited) + [0] * len(unvisited),
                         step_ids=[gmap.node_step_ids.get(vp, 0) for vp in vpids], emb=emb, pos=pos, pair=pair))
    return rows


def ref_vp_pos_fts(gmaps, obs, n_tokens):
    """agent.py:182-196"""
    out = []
    for i, gmap in enumerate(gmaps):
        cand_vpids = [c['viewpointId'] for c in obs[i]['candidate']]
        cand = gmap.get_pos_fts(obs[i]['viewpoint'], cand_vpids, obs[i]['heading'], obs[i]['elevation'])
        start = gmap.get_pos_fts(obs[i]['viewpoint'], [gmap.start_vp], obs[i]['heading'], obs[i]['elevation'])
        fts = np.zeros((n_tokens, 14), dtype=np.float32)
        fts[:, :7] = start
        fts[1:len(cand) + 1, 7:] = cand
        out.append(fts)
    return np.stack(out, 0)


def main():
    ref = load_reference()
    world = synth.nav_world(**WORLD)
    B = len(world[0]['obs'])
    gmaps = [ref.GraphMap(ob['viewpoint']) for ob in world[0]['obs']]
    states = [GO.GraphState(ob['viewpoint'], HIDDEN) for ob in world[0]['obs']]
    for b, ob in enumerate(world[0]['obs']):
        gmaps[b].update_graph(ob)
        states[b].update_graph(ob['viewpoint'], ob['position'], [(c['viewpointId'], c['position']) for c in ob['candidate']])
    fixture, worst = {}, 0.0
    for t, step in enumerate(world):
        obs, ended = step['obs'], step['ended']
        pano = torch.from_numpy(step['pano_embeds'].copy())
        masks = torch.ones(B, pano.shape[1], dtype=torch.bool)
        # ---- reference (agent.py:461-479)
        for b, gmap in enumerate(gmaps):
            if not ended[b]:
                gmap.node_step_ids[obs[b]['viewpoint']] = t + 1
        avg = torch.sum(pano * masks.unsqueeze(2), 1) / torch.sum(masks, 1, keepdim=True)
        for b, gmap in enumerate(gmaps):
            if not ended[b]:
                gmap.update_node_embed(obs[b]['viewpoint'], avg[b], rewrite=True)
                for j, c in enumerate(obs[b]['candidate']):
                    if not gmap.graph.visited(c['viewpointId']):
                        gmap.update_node_embed(c['viewpointId'], pano[b, j])
        rows = ref_gmap_variable(gmaps, obs)
        vp_pos = ref_vp_pos_fts(gmaps, obs, pano.shape[1] + 1)
        # ---- oracle
        cur = [st.index[ob['viewpoint']] for st, ob in zip(states, obs)]
        cands = [[st.index[c['viewpointId']] for c in ob['candidate']] for st, ob in zip(states, obs)]
        for b, st in enumerate(states):
            if not ended[b]:
                st.step_id[cur[b]] = t + 1
        GO.update_node_embeds(states, cur, cands, step['pano_embeds'], np.ones((B, pano.shape[1]), bool), ended)
        heads = [ob['heading'] for ob in obs]
        elevs = [ob['elevation'] for ob in obs]
        og = GO.nav_gmap_variable(states, cur, heads, elevs)
        ov = GO.nav_vp_variable(states, cur, heads, elevs, step['pano_embeds'], cands, step['view_lens'], step['nav_types'])
        # ---- compare, record the reference's values
        G = og['gmap_masks'].shape[1]
        exp = dict(nodes=np.full((B, G), -1, np.int32), emb=np.zeros((B, G, HIDDEN), np.float32),
                   pos=np.zeros((B, G, 7), np.float32), pair=np.zeros((B, G, G), np.float32),
                   step_ids=np.zeros((B, G), np.int64), visited=np.zeros((B, G), bool), lens=np.zeros((B,), np.int64))
        for b, r in enumerate(rows):
            n = len(r['vpids'])
            assert n == og['gmap_lens'][b]
            names = [None if v < 0 else states[b].names[v] for v in og['gmap_nodes'][b, :n]]
            assert names == r['vpids'], (names, r['vpids'])
            exp['nodes'][b, :n] = og['gmap_nodes'][b, :n]
            exp['emb'][b, :n], exp['pos'][b, :n], exp['pair'][b, :n, :n] = r['emb'], r['pos'], r['pair']
            exp['step_ids'][b, :n], exp['visited'][b, :n], exp['lens'][b] = r['step_ids'], np.array(r['visited'], bool), n
        for k_o, k_e in (('gmap_img_embeds', 'emb'), ('gmap_pos_fts', 'pos'), ('gmap_pair_dists', 'pair'),
                         ('gmap_step_ids', 'step_ids'), ('gmap_visited_masks', 'visited')):
            diff = float(np.abs(og[k_o].astype(np.float64) - exp[k_e].astype(np.float64)).max())
            worst = max(worst, diff)
            assert diff == 0.0, (t, k_o, diff)
        diff = float(np.abs(ov['vp_pos_fts'] - vp_pos).max())
        worst = max(worst, diff)
        assert diff == 0.0, (t, 'vp_pos_fts', diff)
        for k, v in exp.items():
            fixture['t%d_%s' % (t, k)] = v
        fixture['t%d_vp_pos' % t] = vp_pos
        # ---- env step: new observation, update the graph (agent.py:599-604)
        if t + 1 < len(world):
            for b, ob in enumerate(world[t + 1]['obs']):
                if not ended[b]:
                    gmaps[b].update_graph(ob)
                    states[b].update_graph(ob['viewpoint'], ob['position'],
                                           [(c['viewpointId'], c['position']) for c in ob['candidate']])
    out = os.path.join(ROOT, 'tests', 'golden', 'graph_world.npz')
    np.savez_compressed(out, **fixture)
    json.dump({'world': WORLD, 'steps': len(world), 'max_abs_diff_oracle_vs_reference': worst,
               'reference': 'VLN-DUET/map_nav_src/models/graph_utils.py (GraphMap, FloydGraph), unmodified'},
              open(os.path.join(ROOT, 'tests', 'golden', 'graph_oracle_vs_reference.json'), 'w'), indent=1)
    print('wrote %s (%d arrays), oracle vs reference max |diff| = %g' % (out, len(fixture), worst))


if __name__ == '__main__':
    main()
