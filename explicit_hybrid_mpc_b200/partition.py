"""
The partition driver: replaces the reference's ``scheduler.main()`` / ``worker.main()`` pair
(``lib/scheduler.py:693-708``, ``lib/worker.py:463-466``) by calls into the device-resident
frontier loop (``ehm_partition_*``) and rebuilds the reference's ``Tree`` / ``NodeData``
objects (``lib/tree.py``) from the flat node store, so ``tree.pkl`` consumers keep working.

* :func:`partition` -- one GPU, one call.
* :func:`partition_distributed` -- one process per GPU (``torch.distributed``): the root
  simplices are dealt to the ranks, every rank runs its own frontier loop, and every few waves
  the ranks compare frontier depths (one all-gather of a scalar) and, only when they diverge,
  move surplus frontier records with an all-to-all over NVLink -- the device analogue of the
  reference's worker -> scheduler -> worker task hand-over (``lib/worker.py:187-239``).
"""
import ctypes as C
import time

import numpy as np

from . import _capi
from .tree import NodeData, Tree

_FIELDS = ('left', 'right', 'root_id', 'depth', 'cand', 'closed', 'vertices', 'vcost', 'vinput')


class PartitionResult:
    """Flat node arrays of one partition run + run statistics."""

    def __init__(self, arrays, stats, problem):
        self.__dict__.update(arrays)
        self.stats = stats
        self.problem = problem

    @property
    def nnodes(self):
        return len(self.left)

    def leaf_ids(self):
        return np.nonzero(self.left < 0)[0]

    def root_nodes(self):
        """node id of every root simplex, indexed by root_id"""
        ids = np.nonzero(self.depth == 0)[0]
        out = np.full(int(self.root_id.max()) + 1, -1, dtype=np.int64)
        out[self.root_id[ids]] = ids
        return out

    def subtree(self, node_id, tree_cls=Tree, data_cls=NodeData):
        """Reference ``Tree`` of the subtree rooted at node_id (iterative, depth-safe)."""
        pb = self.problem

        def data(i):
            if self.cand[i] >= 0:
                d = data_cls(vertices=self.vertices[i].copy(),
                             commutation=pb.commutation_from_index(int(self.cand[i])),
                             vertex_costs=self.vcost[i].copy(),
                             vertex_inputs=self.vinput[i].copy())
            else:
                d = data_cls(vertices=self.vertices[i].copy())
            d.is_epsilon_suboptimal = bool(self.closed[i] == 1)
            return d
        root = tree_cls(data(node_id), top=False)
        stack = [(root, node_id)]
        while stack:
            t, i = stack.pop()
            if self.left[i] >= 0:
                t.grow(data(self.left[i]), data(self.right[i]))
                stack.append((t.left, self.left[i]))
                stack.append((t.right, self.right[i]))
        return root

    def graft(self, initial_tree):
        """Graft the partitioned subtrees onto the leaves of the initial triangulation tree
        (what ``build_tree`` does with the ``branch_<location>.pkl`` files,
        ``lib/scheduler.py:644-691``).  Leaves are matched in depth-first order, which is the
        order :func:`roots_of` extracted them in."""
        from .tools import leaves
        rn = self.root_nodes()
        for r, (_, leaf) in enumerate(leaves(initial_tree)):
            self.subtree(int(rn[r])).copy(leaf)
            leaf.top = False
        initial_tree.top = True
        return initial_tree


def roots_of(initial_tree):
    """Root simplices (nroots,n+1,n) = the open leaves of the initial tree in depth-first
    order (``Scheduler.populate_queue``, ``lib/scheduler.py:620-642``)."""
    from .tools import leaves
    return np.array([leaf.data.vertices for _, leaf in leaves(initial_tree)], dtype=np.float64)


class _Engine:
    """Thin owner of one ``ehm_partition`` handle."""

    def __init__(self, oracle, roots, wave=0, max_nodes=0, max_frontier=0, node_budget=0, verbose=0,
                 root_ids=None):
        self.oracle = oracle
        self.n, self.nu, self.ncand = oracle.n, oracle.nu, oracle.ncand
        roots = np.ascontiguousarray(roots, dtype=np.float64).reshape(-1, self.n + 1, self.n)
        self.opts = _capi.PartitionOpts(int(wave), int(max_nodes), int(max_frontier),
                                        int(node_budget), int(verbose))
        self.stats = _capi.PartitionStats()
        self.h = C.c_void_p()
        self.L = _capi.lib()
        _capi.check(self.L.ehm_partition_create(
            oracle._p.handle, roots.shape[0], roots.ctypes.data_as(C.POINTER(C.c_double)),
            C.byref(self.opts), C.byref(self.h)))
        if root_ids is not None and roots.shape[0] > 0:
            ids = np.ascontiguousarray(root_ids, dtype=np.int32)
            assert ids.shape[0] == roots.shape[0]
            _capi.check(self.L.ehm_partition_set_root_ids(self.h, roots.shape[0], ids.ctypes.data))

    def advance(self, max_waves=0):
        _capi.check(self.L.ehm_partition_advance(self.h, int(max_waves), C.byref(self.stats)))
        return int(self.stats.open_left)

    def frontier_size(self):
        return int(self.stats.open_left)

    def geom_width(self):
        n, nu = self.n, self.nu
        return (n + 1) * n + (n + 1) + (n + 1) * nu

    def export(self, k):
        """pop the k oldest frontier records into fresh CUDA tensors (meta, geom, lb)"""
        import torch
        dev = self.oracle.device
        meta = torch.empty(k, 6, dtype=torch.int32, device=dev)
        geom = torch.empty(k, self.geom_width(), dtype=torch.float64, device=dev)
        lb = torch.empty(k, self.ncand, dtype=torch.float32, device=dev)
        _capi.check(self.L.ehm_partition_export(self.h, int(k), meta.data_ptr(), geom.data_ptr(),
                                                lb.data_ptr()))
        self.stats.open_left -= k
        return meta, geom, lb

    def import_(self, meta, geom, lb, origin_rank):
        k = int(meta.shape[0])
        if k == 0:
            return
        _capi.check(self.L.ehm_partition_import(self.h, k, meta.contiguous().data_ptr(),
                                                geom.contiguous().data_ptr(),
                                                lb.contiguous().data_ptr(),
                                                origin_rank.contiguous().data_ptr()))
        self.stats.open_left += k

    def fetch(self, with_origin=False):
        nn, n, nu = int(self.stats.nnodes), self.n, self.nu
        arr = dict(left=np.empty(nn, np.int32), right=np.empty(nn, np.int32),
                   root_id=np.empty(nn, np.int32), depth=np.empty(nn, np.int32),
                   cand=np.empty(nn, np.int32), closed=np.empty(nn, np.uint8),
                   vertices=np.empty((nn, n + 1, n)), vcost=np.empty((nn, n + 1)),
                   vinput=np.empty((nn, n + 1, nu)))
        _capi.check(self.L.ehm_partition_fetch(self.h, *[arr[k].ctypes.data for k in _FIELDS]))
        arr['vinput'] = arr['vinput'] * self.oracle.pb.u_scale
        if with_origin:
            org = np.empty((nn, 2), np.int32)
            _capi.check(self.L.ehm_partition_fetch_origin(self.h, org.ctypes.data))
            arr['origin'] = org
        return arr

    def stats_dict(self):
        return {k: (list(getattr(self.stats, k)) if k.startswith('stage_') else getattr(self.stats, k))
                for k, _ in _capi.PartitionStats._fields_}

    def close(self):
        if self.h:
            self.L.ehm_partition_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition(oracle, roots, wave=0, max_nodes=0, max_frontier=0, node_budget=0, verbose=0):
    """Run ECC + L-CSS to convergence on the device.

    oracle : :class:`explicit_hybrid_mpc_b200.oracle.Oracle`
    roots  : (nroots, n+1, n) array of root simplices (HOST memory; copied to the device,
             the resulting node store is copied back -- both inside this call).
    """
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    assert roots.ndim == 3 and roots.shape[1:] == (oracle.n + 1, oracle.n)
    t0 = time.perf_counter()
    eng = _Engine(oracle, roots, wave, max_nodes, max_frontier, node_budget, verbose)
    try:
        eng.advance(0)
        arr = eng.fetch()
        st = eng.stats_dict()
    finally:
        eng.close()
    st['wall_s'] = time.perf_counter() - t0
    return PartitionResult(arr, st, oracle.pb)


# --------------------------------------------------------------------------------------------
# multi-GPU: sharded frontier + all-to-all rebalance
# --------------------------------------------------------------------------------------------
def shard_roots(roots, rank, world):
    """Deal root simplices round-robin, largest volume first (deterministic)."""
    from .tools import simplex_volume
    order = np.argsort([-simplex_volume(r) for r in roots], kind='stable')
    return np.sort(order[rank::world])


def rebalance_plan(sizes, fill, slack=1.5):
    """Who sends how many frontier records to whom.  sizes: frontier depth per rank; fill: the
    depth a rank needs to keep its GPU busy (one wave).  Returns an integer matrix send[i][j].
    Nothing moves unless some rank is starving (depth < fill) while another holds more than
    slack x the mean (and more than it needs itself).  Deterministic, identical on all ranks."""
    sizes = np.asarray(sizes, dtype=np.int64)
    world = len(sizes)
    send = np.zeros((world, world), dtype=np.int64)
    total = int(sizes.sum())
    if total == 0 or world == 1:
        return send
    mean = total / world
    starving = sizes < min(fill, mean)
    rich = (sizes > slack * mean) & (sizes > fill)
    if not (starving.any() and rich.any()):
        return send
    target = int(np.floor(mean))
    surplus = np.maximum(sizes - max(target, 1), 0)
    deficit = np.maximum(target - sizes, 0)
    surplus[~rich] = 0
    j = 0
    for i in np.argsort(-surplus, kind='stable'):
        s = int(surplus[i])
        while s > 0 and j < world:
            order = np.argsort(-deficit, kind='stable')
            r = int(order[0])
            if deficit[r] <= 0:
                break
            k = min(s, int(deficit[r]))
            send[i, r] += k
            s -= k
            deficit[r] -= k
            j += 0
        surplus[i] = s
    return send


def stitch(parts):
    """Merge per-rank node arrays into one forest: a node exported by rank a (closed==2) is
    replaced by the node that imported it (origin == (a, node)) on another rank.
    parts: list (indexed by rank) of dicts of node arrays incl. 'origin'."""
    off = np.cumsum([0] + [len(p['left']) for p in parts])
    out = {k: np.concatenate([p[k] for p in parts]) for k in _FIELDS}
    org = np.concatenate([p['origin'] for p in parts])
    for r, p in enumerate(parts):                       # local -> global child indices
        sl = slice(off[r], off[r + 1])
        for k in ('left', 'right'):
            v = out[k][sl]
            out[k][sl] = np.where(v >= 0, v + off[r], -1)
    alias = np.arange(off[-1])
    imported = np.nonzero(org[:, 0] >= 0)[0]
    for y in imported:                                   # exported node x continues as y
        x = off[org[y, 0]] + org[y, 1]
        alias[x] = y
    # resolve chains (a record may be forwarded more than once)
    for x in np.nonzero(alias != np.arange(off[-1]))[0]:
        y = alias[x]
        while alias[y] != y:
            y = alias[y]
        alias[x] = y
    for k in ('left', 'right'):
        v = out[k]
        out[k] = np.where(v >= 0, alias[np.maximum(v, 0)], -1).astype(np.int32)
    keep = np.ones(off[-1], dtype=bool)
    moved = np.nonzero(alias != np.arange(off[-1]))[0]
    # a root simplex that was itself exported keeps its identity through the alias
    keep[moved] = False
    # depth-0 nodes that were exported: make the importer carry depth 0 (it already does)
    new_id = np.cumsum(keep) - 1
    res = {k: out[k][keep] for k in _FIELDS}
    for k in ('left', 'right'):
        v = res[k]
        res[k] = np.where(v >= 0, new_id[np.maximum(v, 0)], -1).astype(np.int32)
    return res


def partition_distributed(oracle, roots, group=None, waves_per_round=2, wave=0, max_nodes=0,
                          max_frontier=0, verbose=0, engine_cls=_Engine, gather=True,
                          fill_depth=512):
    """Partition ``roots`` with one process per GPU (see module docstring).

    Returns (PartitionResult or None, info): rank 0 gets the stitched result when ``gather``;
    ``info`` has the per-rank statistics needed by bench.py (leaves, device seconds, bytes
    exchanged, number of rebalance rounds)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = oracle.device
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    mine = shard_roots(roots, rank, world)
    eng = engine_cls(oracle, roots[mine], wave, max_nodes, max_frontier, 0, verbose, root_ids=mine)
    fill = fill_depth
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    rounds = moved = 0
    xch_s = 0.0
    t0 = time.perf_counter()
    ev0 = ev1 = None
    if dev.type == 'cuda':
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.default_stream(dev))
    try:
        while True:
            mysize = eng.advance(waves_per_round)
            me = torch.tensor([mysize], dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sizes, me, group=group)
            sz = sizes.cpu().numpy()
            if int(sz.sum()) == 0:
                break
            plan = rebalance_plan(sz, fill)
            if plan.sum() == 0:
                continue
            tx = time.perf_counter()
            nsend = plan[rank]
            k = int(nsend.sum())
            meta, geom, lb = eng.export(k) if k else (
                torch.empty(0, 6, dtype=torch.int32, device=dev),
                torch.empty(0, eng.geom_width(), dtype=torch.float64, device=dev),
                torch.empty(0, eng.ncand, dtype=torch.float32, device=dev))
            nrecv = plan[:, rank]
            kr = int(nrecv.sum())
            rmeta = torch.empty(kr, 6, dtype=torch.int32, device=dev)
            rgeom = torch.empty(kr, eng.geom_width(), dtype=torch.float64, device=dev)
            rlb = torch.empty(kr, eng.ncand, dtype=torch.float32, device=dev)
            ins, outs = [int(x) for x in nsend], [int(x) for x in nrecv]
            dist.all_to_all_single(rmeta, meta, outs, ins, group=group)
            dist.all_to_all_single(rgeom, geom, outs, ins, group=group)
            dist.all_to_all_single(rlb, lb, outs, ins, group=group)
            origin = torch.repeat_interleave(torch.arange(world, dtype=torch.int32, device=dev),
                                             torch.tensor(outs, device=dev))
            eng.import_(rmeta, rgeom, rlb, origin)
            if dev.type == 'cuda':
                torch.cuda.synchronize(dev)
            xch_s += time.perf_counter() - tx
            rounds += 1
            moved += k
        loop_ms = None
        if ev0 is not None:
            ev1.record(torch.cuda.default_stream(dev))
            ev1.synchronize()
            loop_ms = ev0.elapsed_time(ev1)   # device clock, idle gaps between rounds included
        arr = eng.fetch(with_origin=True)
        st = eng.stats_dict()
    finally:
        eng.close()
    wall = time.perf_counter() - t0
    info = dict(rank=rank, world=world, wall_s=wall, exchange_s=xch_s, rounds=rounds,
                records_sent=moved, stats=st, loop_device_ms=loop_ms)
    result = None
    if gather:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(arr, parts, dst=0, group=group)
        if rank == 0:
            merged = stitch(parts)
            result = PartitionResult(merged, dict(st), oracle.pb)
    return result, info
