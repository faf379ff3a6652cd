"""CPU oracle for the per-step graph glue of DUET-Imagine.

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/gen_graph_golden.py and bench.py's cpu_baseline leg; never by
the product (vln-imagine_b200/).  Array-based restatement (numpy, fp64 graph state, fp32 embeddings) of

  * ``FloydGraph`` / ``GraphMap``                       VLN-DUET/map_nav_src/models/graph_utils.py:42-148
  * the node-embedding update of the rollout loop      VLN-DUET/map_nav_src/r2r/agent.py:466-479
  * ``_nav_gmap_variable`` / ``_nav_vp_variable``       VLN-DUET/map_nav_src/r2r/agent.py:98-207

The reference keeps dict-of-dict graphs per episode; here an episode is a set of dense arrays indexed by the order in
which viewpoints entered ``node_positions`` (the order the reference's loops iterate in).  Pinned to the real
``GraphMap`` class by oracle/gen_graph_golden.py (tests/golden/graph_world.npz, max |diff| = 0).
"""
from __future__ import annotations

import math

import numpy as np

MAX_DIST = 30          # graph_utils.py:4
MAX_STEP = 10          # graph_utils.py:5
UNREACHED = 95959595   # graph_utils.py:44 (default distance of the nested defaultdict)


class GraphState:
    """One episode's GraphMap (graph_utils.py:96-107) as dense arrays."""

    def __init__(self, start_vp: str, hidden: int, max_nodes: int = 128):
        self.start_vp = start_vp
        self.index = {}                                           # viewpoint id -> node index (insertion order)
        self.names = []
        n = max_nodes
        self.pos = np.zeros((n, 3), np.float64)
        self.dis = np.full((n, n), float(UNREACHED), np.float64)  # FloydGraph._dis
        self.point = np.full((n, n), -1, np.int32)                # FloydGraph._point ("" == -1)
        self.visited = np.zeros((n,), bool)                       # FloydGraph._visited
        self.esum = np.zeros((n, hidden), np.float32)             # GraphMap.node_embeds[vp][0]
        self.ecnt = np.zeros((n,), np.float32)                    # GraphMap.node_embeds[vp][1]
        self.step_id = np.zeros((n,), np.int64)                   # GraphMap.node_step_ids

    # ---- graph_utils.py:109-115 (update_graph) + :53-70 (add_edge, update)
    def node(self, vp: str, position) -> int:
        i = self.index.get(vp)
        if i is None:
            i = self.index[vp] = len(self.names)
            self.names.append(vp)
        self.pos[i] = position
        return i

    def update_graph(self, vp: str, position, candidates):
        """candidates: list of (viewpointId, position)."""
        k = self.node(vp, position)
        for cvp, cpos in candidates:
            c = self.node(cvp, cpos)
            d = math.sqrt((cpos[0] - position[0]) ** 2 + (cpos[1] - position[1]) ** 2 + (cpos[2] - position[2]) ** 2)
            if d < self.dis[k, c]:                                # add_edge, :53-58
                self.dis[k, c] = self.dis[c, k] = d
                self.point[k, c] = self.point[c, k] = -1
        self.relax(k)
        return k

    def relax(self, k: int):
        """FloydGraph.update (:60-69).  Row / column k cannot change during the pass (dis[k, k] stays UNREACHED), so
        the reference's in-place double loop equals this whole-matrix update."""
        n = len(self.names)
        d = self.dis[:n, :n]
        new = d[:, k:k + 1] + d[k:k + 1, :]
        better = (new < d) & ~np.eye(n, dtype=bool)
        d[better] = new[better]
        self.point[:n, :n][better] = k
        self.visited[k] = True

    def distance(self, x: int, y: int) -> float:               # :47-51
        return 0.0 if x == y else float(self.dis[x, y])

    def path_len(self, x: int, y: int) -> int:                 # len(FloydGraph.path(x, y)), :74-90
        if x == y:
            return 0
        k = int(self.point[x, y])
        if k < 0:
            return 1
        return self.path_len(x, k) + self.path_len(k, y)

    # ---- graph_utils.py:117-129
    def update_node_embed(self, i: int, embed: np.ndarray, rewrite: bool = False):
        if rewrite or self.ecnt[i] == 0:
            self.esum[i] = embed
            self.ecnt[i] = 1
        else:
            self.esum[i] = self.esum[i] + embed
            self.ecnt[i] += 1

    def node_embed(self, i: int) -> np.ndarray:
        return self.esum[i] / self.ecnt[i]

    # ---- graph_utils.py:14-40, 131-148 (calculate_vp_rel_pos_fts, get_angle_fts, get_pos_fts)
    def pos_fts(self, cur: int, nodes, heading: float, elevation: float) -> np.ndarray:
        """nodes: node indices, -1 for the [stop] slot (None in the reference).  -> [len(nodes), 7] float32"""
        ang = np.zeros((len(nodes), 2), np.float64)
        dist = np.zeros((len(nodes), 3), np.float64)
        a = self.pos[cur]
        for j, v in enumerate(nodes):
            if v < 0:
                continue
            b = self.pos[v]
            dx, dy, dz = b[0] - a[0], b[1] - a[1], b[2] - a[2]
            xy = max(math.sqrt(dx * dx + dy * dy), 1e-8)
            xyz = max(math.sqrt(dx * dx + dy * dy + dz * dz), 1e-8)
            h = math.asin(dx / xy)
            if b[1] < a[1]:
                h = math.pi - h
            ang[j] = (h - heading, math.asin(dz / xyz) - elevation)
            dist[j] = (xyz / MAX_DIST, self.distance(cur, v) / MAX_DIST, self.path_len(cur, v) / MAX_STEP)
        ang32 = ang.astype(np.float32)                             # the reference casts the angles before sin / cos
        fts = np.stack([np.sin(ang32[:, 0]), np.cos(ang32[:, 0]), np.sin(ang32[:, 1]), np.cos(ang32[:, 1])], 1)
        return np.concatenate([fts.astype(np.float32), dist.astype(np.float32)], 1)


def update_node_embeds(states, cur_nodes, cand_nodes, pano_embeds, pano_masks, ended):
    """agent.py:466-479.  pano_embeds [B, V, D] fp32, pano_masks [B, V] bool; cand_nodes[b] = node index of each
    candidate view (the first len(cand) views of the panorama, agent.py:57-66)."""
    import torch                                               # same reduction as the reference's torch.sum (agent.py:468-469)
    e, m = torch.from_numpy(np.ascontiguousarray(pano_embeds)), torch.from_numpy(np.ascontiguousarray(pano_masks))
    avg = (torch.sum(e * m.unsqueeze(2), 1) / torch.sum(m, 1, keepdim=True)).numpy()
    for b, st in enumerate(states):
        if ended[b]:
            continue
        st.update_node_embed(cur_nodes[b], avg[b], rewrite=True)
        for j, c in enumerate(cand_nodes[b]):
            if not st.visited[c]:
                st.update_node_embed(c, pano_embeds[b, j])
    return avg


def nav_gmap_variable(states, cur_nodes, headings, elevations):
    """agent.py:98-171 with enc_full_graph (the released setting): [stop] + visited + unvisited nodes, their mean
    embeddings, step ids, 7-d position features, pair distances (raw metres), visited mask; zero padded to the longest."""
    B = len(states)
    rows = []
    for b, st in enumerate(states):
        n = len(st.names)
        vis = [i for i in range(n) if st.visited[i]]
        unv = [i for i in range(n) if not st.visited[i]]
        nodes = [-1] + vis + unv
        emb = np.stack([np.zeros_like(st.esum[0])] + [st.node_embed(i) for i in nodes[1:]], 0)
        pair = np.zeros((len(nodes), len(nodes)), np.float32)
        for i in range(1, len(nodes)):
            for j in range(i + 1, len(nodes)):
                pair[i, j] = pair[j, i] = st.distance(nodes[i], nodes[j])
        rows.append(dict(nodes=nodes, visited=[0] + [1] * len(vis) + [0] * len(unv),
                         step_ids=[0 if i < 0 else int(st.step_id[i]) for i in nodes], emb=emb,
                         pos=st.pos_fts(cur_nodes[b], nodes, headings[b], elevations[b]), pair=pair,
                         no_vp_left=len(unv) == 0))
    G = max(len(r['nodes']) for r in rows)
    D = rows[0]['emb'].shape[1]
    out = dict(gmap_nodes=np.full((B, G), -1, np.int32), gmap_lens=np.array([len(r['nodes']) for r in rows], np.int64),
               gmap_img_embeds=np.zeros((B, G, D), np.float32), gmap_step_ids=np.zeros((B, G), np.int64),
               gmap_pos_fts=np.zeros((B, G, 7), np.float32), gmap_pair_dists=np.zeros((B, G, G), np.float32),
               gmap_visited_masks=np.zeros((B, G), bool), gmap_masks=np.zeros((B, G), bool),
               no_vp_left=[r['no_vp_left'] for r in rows])
    for b, r in enumerate(rows):
        n = len(r['nodes'])
        out['gmap_nodes'][b, :n] = r['nodes']
        out['gmap_img_embeds'][b, :n] = r['emb']
        out['gmap_step_ids'][b, :n] = r['step_ids']
        out['gmap_pos_fts'][b, :n] = r['pos']
        out['gmap_pair_dists'][b, :n, :n] = r['pair']
        out['gmap_visited_masks'][b, :n] = np.array(r['visited'], bool)
        out['gmap_masks'][b, :n] = True
    return out


def nav_vp_variable(states, cur_nodes, headings, elevations, pano_embeds, cand_nodes, view_lens, nav_types):
    """agent.py:173-207: [stop] row in front of the panorama, 14-d position features (start viewpoint | candidate),
    vp_masks, vp_nav_masks."""
    B, V, D = pano_embeds.shape
    vp_img = np.concatenate([np.zeros((B, 1, D), np.float32), pano_embeds], 1)
    pos = np.zeros((B, V + 1, 14), np.float32)
    for b, st in enumerate(states):
        cand = st.pos_fts(cur_nodes[b], list(cand_nodes[b]), headings[b], elevations[b])
        start = st.pos_fts(cur_nodes[b], [st.index[st.start_vp]], headings[b], elevations[b])
        pos[b, :, :7] = start
        pos[b, 1:len(cand) + 1, 7:] = cand
    return dict(vp_img_embeds=vp_img, vp_pos_fts=pos,
                vp_masks=np.arange(V + 1)[None, :] < (np.asarray(view_lens)[:, None] + 1),
                vp_nav_masks=np.concatenate([np.ones((B, 1), bool), np.asarray(nav_types) == 1], 1))
