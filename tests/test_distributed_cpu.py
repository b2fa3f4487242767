"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group, a fake
engine with the same interface as the CUDA one (no compute calls need a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deal_plan_is_an_even_interleave():
    from explicit_hybrid_mpc_b200.partition import deal_plan
    p = deal_plan([5, 3, 0, 1000])
    assert (p.sum(axis=1) == [5, 3, 0, 1000]).all()
    assert p.sum(axis=0).max() - p.sum(axis=0).min() <= 4       # every rank ends up with ~ total/world
    assert (p[3] == 250).all() and list(p[0]) == [2, 1, 1, 1] and list(p[1]) == [0, 1, 1, 1]


def test_rebalance_plan_properties():
    from explicit_hybrid_mpc_b200.partition import rebalance_plan
    rng = np.random.default_rng(0)
    for _ in range(200):
        world = int(rng.integers(2, 9))
        sizes = rng.integers(0, 20000, world) * (rng.random(world) < 0.7)
        plan = rebalance_plan(sizes, 512)
        assert plan.shape == (world, world) and (plan >= 0).all() and (np.diag(plan) == 0).all()
        after = sizes - plan.sum(axis=1) + plan.sum(axis=0)
        assert (after >= 0).all() and after.sum() == sizes.sum()
        if plan.sum():
            # senders stay at or above the mean, receivers do not overshoot it
            mean = sizes.sum() / world
            assert (after[plan.sum(axis=1) > 0] >= np.floor(mean) - 1).all()
            assert (after[plan.sum(axis=0) > 0] <= np.ceil(mean) + 1).all()
    assert rebalance_plan([5000, 4500, 4000, 4800], 512).sum() == 0      # balanced: nothing moves
    assert rebalance_plan([0, 0, 0], 512).sum() == 0
    # one empty rank among full ones is served by everybody above the mean
    p = rebalance_plan([0, 3000, 3200, 2900], 256)
    assert p[:, 0].sum() == 2275 and (p[1:, 0] > 0).all()
    p = rebalance_plan([10000, 0, 50, 3000], 512)
    assert p[0].sum() > 6000 and p[1:].sum() == 0


class FakeEngine:
    """Deterministic stand-in for ``partition._Engine``: every frontier record is a 1-simplex
    [a,b]; a node with depth < its root's target depth splits in two, otherwise it closes."""
    n, nu, ncand = 1, 1, 4

    def __init__(self, oracle, roots, wave=0, max_nodes=0, max_frontier=0, node_budget=0, verbose=0,
                 root_ids=None):
        self.oracle = oracle
        self.opts = type('O', (), dict(wave=wave))()
        roots = np.asarray(roots, dtype=float).reshape(-1, 2, 1)
        ids = list(range(len(roots))) if root_ids is None else [int(i) for i in root_ids]
        self.nodes = []          # dicts
        self.frontier = []
        for r, rid in zip(roots, ids):
            self._new(r, 0, rid, (-1, -1))
            self.frontier.append(len(self.nodes) - 1)
        self.stats = type('S', (), dict(open_left=len(self.frontier), nnodes=len(self.nodes)))()

    def _new(self, verts, depth, rid, origin):
        self.nodes.append(dict(v=np.array(verts, float).reshape(2, 1), depth=depth, rid=rid, left=-1,
                               right=-1, closed=0, origin=origin))

    def geom_width(self):
        return 2 + 2 + 2

    def advance(self, max_waves=0):
        waves = 0
        while self.frontier and (max_waves <= 0 or waves < max_waves):
            batch, self.frontier = self.frontier[-4:], self.frontier[:-4]
            for i in batch:
                nd = self.nodes[i]
                target = 3 + nd['rid'] % 4          # unbalanced sub-trees
                if nd['depth'] >= target:
                    nd['closed'] = 1
                    continue
                a, b = nd['v'][0, 0], nd['v'][1, 0]
                m = (a + b) / 2
                for seg in ([a, m], [m, b]):
                    self._new(seg, nd['depth'] + 1, nd['rid'], (-1, -1))
                nd['left'], nd['right'] = len(self.nodes) - 2, len(self.nodes) - 1
                self.frontier += [len(self.nodes) - 2, len(self.nodes) - 1]
            waves += 1
        self.stats.open_left, self.stats.nnodes = len(self.frontier), len(self.nodes)
        return len(self.frontier)

    def export(self, k):
        ids, self.frontier = self.frontier[:k], self.frontier[k:]
        meta = torch.tensor([[i, 0, self.nodes[i]['depth'], self.nodes[i]['rid'], 0, 0] for i in ids],
                            dtype=torch.int32).reshape(-1, 6)
        geom = torch.tensor([list(self.nodes[i]['v'].ravel()) + [0, 0, 0, 0] for i in ids],
                            dtype=torch.float64).reshape(-1, 6)
        for i in ids:
            self.nodes[i]['closed'] = 2
        self.stats.open_left = len(self.frontier)
        return meta, geom, torch.zeros(len(ids), self.ncand, dtype=torch.float32)

    def import_(self, meta, geom, lb, origin_rank):
        for e in range(meta.shape[0]):
            self._new(geom[e, :2].numpy(), int(meta[e, 2]), int(meta[e, 3]),
                      (int(origin_rank[e]), int(meta[e, 0])))
            self.frontier.append(len(self.nodes) - 1)
        self.stats.open_left, self.stats.nnodes = len(self.frontier), len(self.nodes)

    def fetch(self, with_origin=False):
        N = self.nodes
        arr = dict(left=np.array([x['left'] for x in N], np.int32), right=np.array([x['right'] for x in N], np.int32),
                   root_id=np.array([x['rid'] for x in N], np.int32), depth=np.array([x['depth'] for x in N], np.int32),
                   cand=np.zeros(len(N), np.int32), closed=np.array([x['closed'] for x in N], np.uint8),
                   vertices=np.array([x['v'] for x in N]).reshape(-1, 2, 1), vcost=np.zeros((len(N), 2)),
                   vinput=np.zeros((len(N), 2, 1)))
        if with_origin:
            arr['origin'] = np.array([x['origin'] for x in N], np.int32).reshape(-1, 2)
        return arr

    def stats_dict(self):
        return dict(nnodes=len(self.nodes), nleaves=sum(1 for x in self.nodes if x['closed'] == 1))

    def close(self):
        pass


def _worker(rank, world, port, q, kw):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from explicit_hybrid_mpc_b200 import partition as P
    oracle = type('Orc', (), dict(device=torch.device('cpu'), n=1, nu=1, ncand=4,
                                  pb=type('PB', (), dict(u_scale=np.ones(1)))()))()
    nroots = 7
    roots = np.array([[[float(i)], [float(i) + 1.0 + 0.1 * i]] for i in range(nroots)])
    res, info = P.partition_distributed(oracle, roots, waves_per_round=1, engine_cls=FakeEngine,
                                        fill_depth=3, **kw)
    if rank == 0:
        q.put((dict(left=res.left, right=res.right, root_id=res.root_id, depth=res.depth,
                    closed=res.closed, vertices=res.vertices), info['rounds']))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('kw', [dict(signal=True), dict(signal=True, deal_depth=2), dict(signal=False),
                                dict(signal=True, deal_depth=2, world=4)],
                         ids=['flagged', 'flagged+deal', 'fixed-rounds', 'flagged+deal-world4'])
def test_partition_distributed_gloo_world2(kw):
    kw = dict(kw)
    world = kw.pop('world', 2)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kw)) for r in range(world)]
    for p in procs:
        p.start()
    out, rounds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: the same fake partition on one rank
    ref = FakeEngine(None, np.array([[[float(i)], [float(i) + 1.0 + 0.1 * i]] for i in range(7)]))
    ref.advance(0)
    r = ref.fetch()
    assert len(out['left']) == len(r['left'])                 # exported duplicates were removed
    leaves_o = np.nonzero(out['left'] < 0)[0]
    leaves_r = np.nonzero(r['left'] < 0)[0]
    assert len(leaves_o) == len(leaves_r) and (out['closed'][leaves_o] == 1).all()
    key = lambda a, idx: sorted((int(a['root_id'][i]), int(a['depth'][i]), float(a['vertices'][i][0, 0]),
                                 float(a['vertices'][i][1, 0])) for i in idx)
    assert key(out, leaves_o) == key(r, leaves_r)             # identical leaf sets
    # every non-leaf has two valid children, every node except the roots exactly one parent
    inner = np.nonzero(out['left'] >= 0)[0]
    ch = np.concatenate([out['left'][inner], out['right'][inner]])
    assert len(set(ch.tolist())) == len(ch) == len(out['left']) - 7
    assert rounds >= 1                                        # the rebalance path was exercised


def test_kuhn_triangulation_tiles_the_box():
    import math
    from explicit_hybrid_mpc_b200.tools import kuhn_triangulation, simplex_volume
    lo, hi = np.array([-1.0, -0.087, -1.0, -0.35]), np.array([1.0, 0.087, 1.0, 0.35])
    T = kuhn_triangulation(lo, hi)
    assert T.shape == (math.factorial(4), 5, 4)
    assert abs(sum(simplex_volume(S) for S in T) - np.prod(hi - lo)) <= 1e-12 * np.prod(hi - lo)
    # random points fall in exactly one simplex (up to boundaries)
    rng = np.random.default_rng(0)
    for x in lo + rng.random((200, 4)) * (hi - lo):
        inside = 0
        for S in T:
            al = np.linalg.solve(np.column_stack([v - S[0] for v in S[1:]]), x - S[0])
            inside += (al.min() >= -1e-12) and (al.sum() <= 1 + 1e-12)
        assert inside == 1
    assert kuhn_triangulation(-np.ones(8), np.ones(8), limit=1024).shape == (1024, 9, 8)


def test_overall_status_fields(tmp_path):
    """status.txt / statistics 'overall' block (lib/scheduler.py:293-333) from a node store."""
    from explicit_hybrid_mpc_b200.partition import PartitionResult
    eng = FakeEngine(None, np.array([[[0.0], [1.0]], [[1.0], [3.0]]]))
    eng.advance(0)
    arr = eng.fetch()
    res = PartitionResult(arr, dict(nleaves=int((arr['closed'] == 1).sum()), open_left=0, device_ms=12.0,
                                    nsolves=5, nfail=0), None)
    o = res.write_status(str(tmp_path / 'status.txt'))
    assert abs(o['volume_filled_total'] - 3.0) < 1e-12 and abs(o['volume_filled_frac'] - 1.0) < 1e-12
    assert o['simplex_count_total'] == int((arr['closed'] == 1).sum()) and o['num_tasks_in_queue'] == 0
    txt = open(str(tmp_path / 'status.txt')).read()
    assert 'volume filled (total [%]): 1.0000e+02' in txt and 'simplex_count: %d' % o['simplex_count_total'] in txt


def test_stitch_resolves_forwarding_chains():
    """A frontier record may be handed on more than once (rank 0 -> 1 -> 2): the node it stood for
    must end up as ONE node of the merged forest, carrying the sub-tree grown by the last owner."""
    from explicit_hybrid_mpc_b200.partition import stitch

    def part(rows):
        # rows: (left, right, root_id, depth, closed, origin_rank, origin_node, a, b)
        n = len(rows)
        return dict(left=np.array([r[0] for r in rows], np.int32), right=np.array([r[1] for r in rows], np.int32),
                    root_id=np.array([r[2] for r in rows], np.int32), depth=np.array([r[3] for r in rows], np.int32),
                    cand=np.zeros(n, np.int32), closed=np.array([r[4] for r in rows], np.uint8),
                    vertices=np.array([[[r[7]], [r[8]]] for r in rows], float), vcost=np.zeros((n, 2)),
                    vinput=np.zeros((n, 2, 1)), origin=np.array([[r[5], r[6]] for r in rows], np.int32).reshape(n, 2))
    # rank 0: root [0,4] split into [0,2] (kept, closed) and [2,4] (exported -> closed=2)
    p0 = part([(1, 2, 0, 0, 0, -1, -1, 0, 4), (-1, -1, 0, 1, 1, -1, -1, 0, 2), (-1, -1, 0, 1, 2, -1, -1, 2, 4)])
    # rank 1: imported (0,2), immediately exported again
    p1 = part([(-1, -1, 0, 1, 2, 0, 2, 2, 4)])
    # rank 2: imported (1,0), split into two closed leaves
    p2 = part([(1, 2, 0, 1, 0, 1, 0, 2, 4), (-1, -1, 0, 2, 1, -1, -1, 2, 3), (-1, -1, 0, 2, 1, -1, -1, 3, 4)])
    m = stitch([p0, p1, p2])
    assert len(m['left']) == 5                                   # 7 stored nodes - 2 aliases
    root = int(np.nonzero(m['depth'] == 0)[0][0])
    kids = [int(m['left'][root]), int(m['right'][root])]
    assert sorted(float(m['vertices'][k][0, 0]) for k in kids) == [0.0, 2.0]
    inner = [k for k in kids if m['left'][k] >= 0][0]            # [2,4] continues on rank 2
    leaves = sorted((float(m['vertices'][c][0, 0]), float(m['vertices'][c][1, 0]))
                    for c in (m['left'][inner], m['right'][inner]))
    assert leaves == [(2.0, 3.0), (3.0, 4.0)]
    assert (m['closed'][m['left'] < 0] == 1).all() and not (m['closed'] == 2).any()


def test_record_pack_roundtrip():
    """frontier records travel as one packed byte row each (meta | geom | bounds); one-row buffers
    were the trap: a one-row slice is 'contiguous' whatever its row stride."""
    from explicit_hybrid_mpc_b200.partition import _pack_records, _unpack_records
    for k in (0, 1, 2, 7):
        meta = torch.arange(k * 6, dtype=torch.int32).view(k, 6)
        geom = torch.randn(k, 30, dtype=torch.float64)
        lb = torch.randn(k, 3125)
        buf = _pack_records(meta, geom, lb)
        assert buf.shape == (k, 24 + 8 * 30 + 4 * 3125)
        m, g, l = _unpack_records(buf, 30, 3125)
        assert torch.equal(m, meta) and torch.equal(g, geom) and torch.equal(l, lb)
