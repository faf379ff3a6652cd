"""Host-side pieces of the DUET pre-training forward - GROUNDWORK ONLY.

The pre-training model (VLN-DUET/pretrain_src/model/vilmodel.py, pretrain_cmt.py) reuses every block of the navigation model;
what it adds is the per-trajectory aggregation of graph-node features (``GlobalMapEncoder._aggregate_gmap_features``,
vilmodel.py:577-611: a Python triple loop over trajectories, steps and candidates) and the vocabulary-wide MLM head.  This
module holds the index builder that turns the aggregation into ONE segment-mean launch (``vi_gather_mean``) over the
flattened trajectory embeddings; the module that would call it is not built yet (the oracle of the whole forward is:
oracle/pretrain_oracle.py, pinned to the real reference).  tests/test_host_logic.py checks the builder against that oracle
with a numpy segment mean.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def gmap_aggregation_rows(traj_step_lens: Sequence[int], traj_vp_view_lens: Sequence[int], traj_vpids: List[List[str]],
                          traj_cand_vpids: List[List[List[str]]], gmap_vpids: List[list], n_views: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """-> (offsets int32 [B * G + 1], row_idx int32 [n], G) for ``vi_gather_mean`` over ``src`` = the panorama-encoder output of all
    N trajectory steps flattened to [N * n_views + 1, 768] with ONE ZERO ROW appended (index N * n_views).

    Output row b * G + g is the feature of ``gmap_vpids[b][g]``:
      * g = 0 ([stop]) and the padding behind an episode's own nodes: the zero row (a unit segment - an empty segment would be 0 / 0);
      * a visited viewpoint: the mean over the valid views of its LAST panorama in the trajectory (the dict of the reference is
        overwritten by later steps, vilmodel.py:591);
      * an unvisited one: the mean over the candidate views that pointed at it while it was still unvisited (:592-595)."""
    B = len(traj_step_lens)
    G = max(len(v) for v in gmap_vpids)
    N = int(sum(traj_step_lens))
    zero_row = N * n_views
    offsets, rows = [0], []
    step0 = 0
    for b in range(B):
        visited, unvisited = {}, {}
        for t in range(traj_step_lens[b]):
            base = (step0 + t) * n_views
            visited[traj_vpids[b][t]] = [base + j for j in range(int(traj_vp_view_lens[step0 + t]))]
            for j, vp in enumerate(traj_cand_vpids[b][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append(base + j)
        step0 += traj_step_lens[b]
        for g in range(G):
            vp = gmap_vpids[b][g] if g < len(gmap_vpids[b]) else None
            if g == 0 or vp is None:
                rows.append(zero_row)
            elif vp in visited:
                rows.extend(visited[vp])
            else:
                rows.extend(unvisited[vp])         # KeyError: a graph node that was never seen - as in the reference
            offsets.append(len(rows))
    return np.asarray(offsets, np.int32), np.asarray(rows, np.int32), G
