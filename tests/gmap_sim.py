"""Test double for the reference's GraphMap (vlnce_baselines/models/graph_utils.py:133-250): the same attributes
(string-keyed dictionaries, ``graph_nx``, the two all-pairs tables rebuilt by networkx after every update) and the same
mutation pattern — one new node per ``step()``, candidates that are localised onto an existing node (an extra edge: a
loop closure), merged into an existing ghost (its fronts / position list / embedding sum grow) or become a new ghost, the
ghost the agent walked to deleted first — driven by a random walk instead of a simulator.  Test infrastructure: it gives
the stateful packer (etpnav_b200/packing.py: GmapPacker) an evolving map to follow; values are arbitrary, the layout
contract is what is checked (against the stateless flatten, which the reference fixtures pin)."""
import networkx as nx
import numpy as np
import torch


class SimGraphMap:
    def __init__(self, seed, loc_noise=0.5, ghost_aug=0.1, width=768, device="cpu", p_node=0.2, p_ghost=0.25):
        self.rng = np.random.default_rng(seed)
        self.gen = torch.Generator().manual_seed(seed)
        self.loc_noise, self.ghost_aug, self.width, self.device = loc_noise, ghost_aug, width, device
        self.p_node, self.p_ghost = p_node, p_ghost
        self.graph_nx = nx.Graph()
        self.node_pos, self.node_embeds, self.node_stepId = {}, {}, {}
        self.ghost_cnt = 0
        self.ghost_pos, self.ghost_mean_pos, self.ghost_embeds, self.ghost_fronts = {}, {}, {}, {}
        self.ghost_aug_pos = {}
        self.shortest_path = self.shortest_dist = None
        self.cur_vp, self.cur_pos, self.t = None, np.zeros(3), 0

    def _emb(self):
        return torch.randn(self.width, generator=self.gen).to(self.device)

    def _nearest(self, q, table):
        best, bv = 1e9, None
        for v, p in table.items():
            d = float(np.sqrt(((q - p) ** 2).sum()))
            if d < best:
                best, bv = d, v
        return bv if best <= self.loc_noise else None

    def delete_ghost(self, vp):
        for d in (self.ghost_pos, self.ghost_mean_pos, self.ghost_embeds, self.ghost_fronts):
            d.pop(vp)

    def step(self, n_cands=None):
        rng = self.rng
        prev = self.cur_vp
        if self.ghost_pos and prev is not None:       # walk to a ghost: it is removed from the map, a node takes its place
            gv = list(self.ghost_pos)[int(rng.integers(len(self.ghost_pos)))]
            pos = np.asarray(self.ghost_mean_pos[gv]) + rng.normal(0, 0.05, 3)
            prev = self.ghost_fronts[gv][0]
            self.delete_ghost(gv)
        else:
            pos = rng.normal(0, 3, 3)
        self.t += 1
        cur = str(len(self.node_pos))
        self.graph_nx.add_node(cur)
        if prev is not None:
            self.graph_nx.add_edge(prev, cur, weight=float(np.linalg.norm(self.node_pos[prev] - pos)))
        self.node_pos[cur], self.node_embeds[cur], self.node_stepId[cur] = pos, self._emb(), self.t
        for _ in range(int(rng.integers(2, 6)) if n_cands is None else n_cands):
            u = rng.random()
            if u < self.p_node and len(self.node_pos) > 2:     # a candidate that coincides with an older node
                v = list(self.node_pos)[int(rng.integers(len(self.node_pos) - 1))]
                cpos = self.node_pos[v] + rng.normal(0, 0.01, 3)
            elif u < self.p_node + self.p_ghost and self.ghost_mean_pos:     # ... or with an existing ghost
                v = list(self.ghost_mean_pos)[int(rng.integers(len(self.ghost_mean_pos)))]
                cpos = np.asarray(self.ghost_mean_pos[v]) + rng.normal(0, 0.01, 3)
            else:
                cpos = pos + rng.normal(0, 2.5, 3)
            hit = self._nearest(cpos, {k: p for k, p in self.node_pos.items() if k != cur})
            if hit is not None:
                self.graph_nx.add_edge(cur, hit, weight=float(np.linalg.norm(pos - self.node_pos[hit])))
                continue
            gh = self._nearest(cpos, self.ghost_mean_pos)
            if gh is None:
                gh = f"g{self.ghost_cnt}"
                self.ghost_cnt += 1
                self.ghost_pos[gh], self.ghost_mean_pos[gh] = [cpos], cpos
                self.ghost_embeds[gh], self.ghost_fronts[gh] = [self._emb(), 1], [cur]
            else:
                self.ghost_pos[gh].append(cpos)
                self.ghost_mean_pos[gh] = np.mean(self.ghost_pos[gh], axis=0)
                self.ghost_embeds[gh][0] = self.ghost_embeds[gh][0] + self._emb()
                self.ghost_embeds[gh][1] += 1
                self.ghost_fronts[gh].append(cur)
        self.ghost_aug_pos = {}
        for gv, gp in self.ghost_mean_pos.items():
            nz = np.clip(rng.normal(0, self.ghost_aug, 3) * np.array([1.0, 0.0, 1.0]), -self.ghost_aug, self.ghost_aug) \
                if self.ghost_aug else 0.0
            self.ghost_aug_pos[gv] = np.asarray(gp) + nz
        self.shortest_path = dict(nx.all_pairs_dijkstra_path(self.graph_nx))
        self.shortest_dist = dict(nx.all_pairs_dijkstra_path_length(self.graph_nx))
        self.cur_vp, self.cur_pos = cur, pos
        return self

    def pose(self):
        q = self.rng.normal(size=4)
        return self.cur_vp, self.cur_pos, q / np.linalg.norm(q)
