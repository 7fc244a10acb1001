"""CPU: the stateful packer (etpnav_b200/packing.py: GmapPacker, the array-backed mirror of GraphMap that is updated
incrementally) must hand ``etp_gmap_pack`` / ``etp_segment_gather_rows`` exactly what the stateless flatten does — which
tests/test_packing_host_cpu.py pins to the reference's fixtures — at every step of evolving maps: node appends, loop
closures (full table re-read), leaf appends (row / column only), ghost creation / merge / deletion, idle environments,
environments that leave the batch and new episodes."""
import numpy as np
import pytest
import torch

from etpnav_b200 import packing
from tests.gmap_sim import SimGraphMap
from tests.test_packing_cpu import cpu_gmap_names, load
from tests.test_packing_host_cpu import fake_gmaps


def expected_img_tables(gms, n_max):
    table, wt, ptr = [], [], [0]
    for gm in gms:
        k = 1
        ptr.append(len(table))
        for v in gm.node_pos:
            table.append(gm.node_embeds[v].data_ptr()); wt.append(1.0); ptr.append(len(table)); k += 1
        for v in gm.ghost_pos:
            table.append(gm.ghost_embeds[v][0].data_ptr()); wt.append(1.0 / gm.ghost_embeds[v][1]); ptr.append(len(table)); k += 1
        ptr.extend([len(table)] * (n_max - k))
    return (np.asarray(table, dtype=np.int64), np.asarray(ptr, dtype=np.int32), np.arange(len(table), dtype=np.int32),
            np.asarray(wt, dtype=np.float32))


def check_same(pk, gms, poses):
    cur_vp, cur_pos, cur_ori = (list(x) for x in zip(*poses))
    ref = packing.flatten_gmaps(gms, cur_vp, cur_pos, cur_ori)
    got = pk.flatten(gms, cur_vp, cur_pos, cur_ori)
    assert np.array_equal(got[0], ref[0])
    assert got[1].dtype == ref[1].dtype and np.array_equal(got[1], ref[1])
    assert got[2].dtype == ref[2].dtype and np.array_equal(got[2], ref[2])
    assert got[3] == ref[3] and got[4:] == ref[4:]
    # what pack() stages for the device (a second sync of the same state, now with the image rows)
    meta, f64, i32b, table, cptr, cidx, cwt, ok = pk.host_tables(gms, cur_vp, cur_pos, cur_ori)
    assert ok and np.array_equal(meta, ref[0]) and np.array_equal(f64, ref[1]) and np.array_equal(i32b, ref[2])
    for a, b in zip((table, cptr, cidx, cwt), expected_img_tables(gms, ref[4])):
        assert a.dtype == b.dtype and np.array_equal(a, b)


def impls():
    return ["py", "c"] if packing._load_mirror_helper() is not None else ["py"]


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("ghost_aug", [0.0, 0.1])
def test_incremental_mirror_follows_evolving_maps(ghost_aug, impl):
    rng = np.random.default_rng(7)
    gms = [SimGraphMap(100 + e, ghost_aug=ghost_aug, width=8).step() for e in range(6)]
    pk = packing.GmapPacker(device="cpu", width=8, impl=impl)
    assert pk.impl == impl
    closures = leaves = 0
    for t in range(22):
        check_same(pk, gms, [gm.pose() for gm in gms])
        for gm in gms:
            if rng.random() < 0.75:          # the others idle this step (finished-but-kept environments)
                gm.step()
                deg = len(gm.graph_nx[gm.cur_vp])
                closures += deg > 1
                leaves += deg == 1
        if t == 8:
            gms.pop(2)                        # an environment finishes and leaves the batch
        if t == 12:
            gms[1] = SimGraphMap(999, ghost_aug=ghost_aug, width=8).step()   # a new episode in the same slot
        if t == 15:
            gms[0].delete_ghost(next(iter(gms[0].ghost_pos)))                # deletion without an update
    check_same(pk, gms, [gm.pose() for gm in gms])
    assert closures >= 5 and leaves >= 20          # both table paths were exercised


def test_c_helper_is_built_and_default():
    """build() compiles csrc_py/gmap_mirror.c; the packer then takes the C host half by default."""
    assert packing._load_mirror_helper() is not None, "etpnav_b200/_gmap_mirror.so missing: python -m etpnav_b200.build"
    assert packing.GmapPacker(device="cpu").impl == "c"


@pytest.mark.parametrize("name", cpu_gmap_names())
def test_packer_first_call_equals_flatten_on_reference_fixtures(name):
    gms, cur_vp, cur_pos, cur_ori = fake_gmaps(load(name))
    ref = packing.flatten_gmaps(gms, cur_vp, cur_pos, cur_ori)
    for impl in impls():
        got = packing.GmapPacker(device="cpu", impl=impl).flatten(gms, cur_vp, cur_pos, cur_ori)
        for a, b in zip(got[:3], ref[:3]):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        assert got[3:] == ref[3:]


@pytest.mark.parametrize("impl", impls())
def test_packer_falls_back_when_an_embedding_needs_a_gradient(impl):
    gm = SimGraphMap(5, width=8).step().step()
    v = next(iter(gm.node_embeds))
    gm.node_embeds[v] = gm.node_embeds[v].clone().requires_grad_(True)
    pose = [gm.pose()]
    args = [[gm]] + [list(x) for x in zip(*pose)]
    assert not packing.GmapPacker(device="cpu", width=8, impl=impl).host_tables(*args)[-1]   # pack() takes pack_gmap's gather
    with torch.no_grad():
        assert packing.GmapPacker(device="cpu", width=8, impl=impl).host_tables(*args)[-1]


class _GridMap(SimGraphMap):
    """Positions on an integer grid — exact distance ties between alternative paths, zero-length edges between coincident
    nodes — and many extra edges from every new node to older ones: the worst case for the C mirror's selective update of
    the all-pairs tables after a loop closure (only pairs with a route through the new node that is not longer are read
    again; networkx relaxes with a strict `<`, so every other pair keeps distance, path and tie-breaks)."""

    def step(self, n_cands=None):
        import networkx as nx
        rng, prev = self.rng, self.cur_vp
        pos = np.round(rng.normal(0, 2, 3)) if prev is None else np.round(self.node_pos[prev] + rng.integers(-1, 2, 3))
        self.t += 1
        cur = str(len(self.node_pos))
        self.graph_nx.add_node(cur)
        if prev is not None:
            self.graph_nx.add_edge(prev, cur, weight=float(np.linalg.norm(self.node_pos[prev] - pos)))
        self.node_pos[cur], self.node_embeds[cur], self.node_stepId[cur] = pos, self._emb(), self.t
        for v in [v for v in self.node_pos if v != cur]:
            if rng.random() < 0.25:
                self.graph_nx.add_edge(cur, v, weight=float(np.linalg.norm(pos - self.node_pos[v])))
        if rng.random() < 0.7:
            gh = f"g{self.ghost_cnt}"
            self.ghost_cnt += 1
            cp = pos + rng.integers(-2, 3, 3).astype(float)
            self.ghost_pos[gh], self.ghost_mean_pos[gh] = [cp], cp
            self.ghost_embeds[gh], self.ghost_fronts[gh] = [self._emb(), 1], [cur]
        self.ghost_aug_pos = {g: np.asarray(p) for g, p in self.ghost_mean_pos.items()}
        self.shortest_path = dict(nx.all_pairs_dijkstra_path(self.graph_nx))
        self.shortest_dist = dict(nx.all_pairs_dijkstra_path_length(self.graph_nx))
        self.cur_vp, self.cur_pos = cur, pos
        return self


@pytest.mark.skipif("c" not in impls(), reason="C helper not built")
def test_selective_table_update_survives_ties_and_zero_length_edges(monkeypatch):
    monkeypatch.setenv("ETP_PACK_VERIFY", "1")     # the helper re-reads every table after its selective update and raises on a difference
    for seed in range(30):
        gms = [_GridMap(seed * 10 + e, width=8).step() for e in range(3)]
        pk = packing.GmapPacker(device="cpu", width=8, impl="c")
        for t in range(14):
            cur_vp, cur_pos, cur_ori = (list(x) for x in zip(*[gm.pose() for gm in gms]))
            got, ref = pk.flatten(gms, cur_vp, cur_pos, cur_ori), packing.flatten_gmaps(gms, cur_vp, cur_pos, cur_ori)
            assert all(np.array_equal(a, b) for a, b in zip(got[:3], ref[:3])), (seed, t)
            for gm in gms:
                gm.step()
