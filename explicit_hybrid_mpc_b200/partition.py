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
import os
import time

import numpy as np

from . import _capi
from .tree import NodeData, Tree

_FIELDS = ('left', 'right', 'root_id', 'depth', 'cand', 'closed', 'vertices', 'vcost', 'vinput')


class PartitionResult:
    """Flat node arrays of one partition run + run statistics."""

    def __init__(self, arrays, stats, problem):
        arrays = dict(arrays)
        # vertex inputs may arrive as the engine's scaled inputs uhat ('vinput_hat'): the physical
        # inputs u = u_scale * uhat are formed on first access (a pass over up to tens of GB)
        self._vinput_hat = arrays.pop('vinput_hat', None)
        self._vinput = arrays.pop('vinput', None)
        self.__dict__.update(arrays)
        self.stats = stats
        self.problem = problem

    @property
    def vinput(self):
        if self._vinput is None:
            self._vinput = self._vinput_hat * self.problem.u_scale
        return self._vinput

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

    def overall_status(self, total_volume=None):
        """The ``overall`` block of the reference's status / statistics files
        (``lib/scheduler.py:293-305``) computed from the node store and the device counters:
        volume filled (closed leaves), simplex count, queue length, time elapsed."""
        from .tools import simplex_volume
        leaves = self.leaf_ids()
        closed = leaves[self.closed[leaves] == 1]
        vol = float(sum(simplex_volume(self.vertices[i]) for i in closed))
        if total_volume is None:
            total_volume = float(sum(simplex_volume(self.vertices[i]) for i in np.nonzero(self.depth == 0)[0]))
        st = self.stats
        return dict(num_proc_active=0, num_proc_failed=0,
                    num_tasks_in_queue=int(st.get('open_left', 0)),
                    volume_filled_total=vol,
                    volume_filled_frac=vol / total_volume if total_volume else 0.0,
                    simplex_count_total=int(len(closed)),
                    time_elapsed=float(st.get('wall_s', st.get('device_ms', 0.0) * 1e-3)),
                    algorithms_running=[], time_active_total=float(st.get('device_ms', 0.0)) * 1e-3,
                    time_idle_total=0.0, eta=0 if not st.get('open_left', 0) else None,
                    max_depth=int(self.depth.max()), oracle_solves=int(st.get('nsolves', 0)),
                    oracle_failures=int(st.get('nfail', 0)))

    def statistics_record(self, total_volume=None, ranks=None):
        """One record of the reference's ``statistics.pkl`` (``lib/scheduler.py:293-308``):
        ``dict(overall=..., process=[...])``.  ``process`` has one entry per GPU rank (``ranks``: the
        per-rank stats dicts of a multi-GPU run, default this result's own) with the fields of the
        reference's worker status (``lib/worker.py:31-116``) that have a device counterpart."""
        o = self.overall_status(total_volume)
        o['scheduler_looprate'] = dict(publisher=0.0, dispatcher=0.0, collector=0.0)   # no scheduler threads here
        procs = []
        for st in (ranks if ranks is not None else [self.stats]):
            ms = list(st.get('stage_ms', [0.0] * 8))
            procs.append(dict(status='idle', algorithm=None, current_branch='', current_location='',
                              volume_filled_total=o['volume_filled_total'] if ranks is None else None,
                              volume_filled_current=0.0, simplex_count_total=int(st.get('nleaves', 0)),
                              simplex_count_current=0, time_active_total=float(st.get('device_ms', 0.0)) * 1e-3,
                              time_active_current=0.0, time_idle=0.0, time_ecc=ms[7] * 1e-3,
                              time_lcss=sum(ms[:7]) * 1e-3, oracle_solves=int(st.get('nsolves', 0)),
                              oracle_failures=int(st.get('nfail', 0))))
        return dict(overall=o, process=procs)

    def append_statistics(self, path, total_volume=None, ranks=None):
        """Append one record to a ``statistics.pkl``-style file (the reference appends a pickle per
        publisher tick, ``lib/scheduler.py:306-308``; ``post_process.py:99-128`` reads them back one by one)."""
        import pickle
        rec = self.statistics_record(total_volume, ranks)
        with open(path, 'ab') as f:
            pickle.dump(rec, f)
        return rec

    def write_status(self, path, total_volume=None):
        """``status.txt`` in the reference's layout (``lib/scheduler.py:311-333``, overall part)."""
        o = self.overall_status(total_volume)
        with open(path, 'w') as f:
            f.write('\n'.join([
                '# overall',
                'number of processes active: %d' % o['num_proc_active'],
                'number of processes failed: %d' % o['num_proc_failed'],
                'number of tasks queue: %d' % o['num_tasks_in_queue'],
                'volume filled (total [%%]): %.4e' % (o['volume_filled_frac'] * 100.),
                'simplex_count: %d' % o['simplex_count_total'],
                'time elapsed [s]: %d' % o['time_elapsed'],
                'time active (total for all processes [s]): %.0f' % o['time_active_total'],
                'time idle (total for all processes [s]): %.0f' % o['time_idle_total'],
                'ETA [s]: %s' % str(o['eta']),
            ]) + '\n\n')
        return o

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


def leafset_digest(left, root_id, vertices, nroots):
    """Per root simplex: (number of leaves, sha256 of the lexicographically sorted leaf vertex sets,
    raw fp64 bytes) -- independent of node numbering, so any two engines that build the same
    partition agree on it bit for bit."""
    import hashlib
    left, root_id = np.asarray(left), np.asarray(root_id)
    V = np.asarray(vertices, dtype=np.float64)
    counts, digests = [], []
    for r in range(nroots):
        L = V[(left < 0) & (root_id == r)]
        L = L.reshape(len(L), -1)
        L = L[np.lexsort(L.T[::-1])] if len(L) else L
        counts.append(int(len(L)))
        digests.append(hashlib.sha256(np.ascontiguousarray(L).tobytes()).hexdigest())
    return counts, digests


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

    def advance(self, max_waves=0, budget_ms=0.0):
        _capi.check(self.L.ehm_partition_advance_for(self.h, int(max_waves), float(budget_ms),
                                                     C.byref(self.stats)))
        return int(self.stats.open_left)

    def frontier_size(self):
        """frontier entries that :meth:`export` can hand over (root simplices still waiting for
        their ECC pass are open work but not exportable)"""
        return int(self.stats.frontier)

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
        self.stats.frontier -= k
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
        self.stats.frontier += k

    def fetch(self, with_origin=False, pinned=False):
        """Copy the node store to host arrays.  pinned=True: into page-locked buffers cached on the
        oracle (grow-only, REUSED by the next fetch on the same oracle -- the arrays are views that
        the next partition overwrites) and without the u_scale pass: 'vinput_hat' holds the scaled
        inputs, :class:`PartitionResult` scales on access.  D2H of a 27 GB store: 1.2 s instead of 7.6 s."""
        nn, n, nu = int(self.stats.nnodes), self.n, self.nu
        shapes = dict(left=((nn,), np.int32), right=((nn,), np.int32), root_id=((nn,), np.int32),
                      depth=((nn,), np.int32), cand=((nn,), np.int32), closed=((nn,), np.uint8),
                      vertices=((nn, n + 1, n), np.float64), vcost=((nn, n + 1), np.float64),
                      vinput=((nn, n + 1, nu), np.float64))
        if pinned:
            import torch
            pool = self.oracle.__dict__.setdefault('_pinned_pool', {})
            arr = {}
            for k, (shp, dt) in shapes.items():
                need = int(np.prod(shp)) * np.dtype(dt).itemsize
                buf = pool.get(k)
                if buf is None or buf.numel() < need:
                    pool[k] = buf = torch.empty(max(need, 1) + need // 4, dtype=torch.uint8, pin_memory=True)
                arr[k] = buf.numpy()[:need].view(dt).reshape(shp)
        else:
            arr = {k: np.empty(shp, dt) for k, (shp, dt) in shapes.items()}
        _capi.check(self.L.ehm_partition_fetch(self.h, *[arr[k].ctypes.data for k in _FIELDS]))
        if pinned:
            arr['vinput_hat'] = arr.pop('vinput')
        else:
            arr['vinput'] *= self.oracle.pb.u_scale          # in place: the store can be tens of GB
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


def partition(oracle, roots, wave=0, max_nodes=0, max_frontier=0, node_budget=0, verbose=0,
              status_path=None, status_every=8, pinned=False):
    """Run ECC + L-CSS to convergence on the device.

    oracle : :class:`explicit_hybrid_mpc_b200.oracle.Oracle`
    roots  : (nroots, n+1, n) array of root simplices (HOST memory; copied to the device,
             the resulting node store is copied back -- both inside this call).
    status_path : optional file rewritten every ``status_every`` waves with the live counters the
             reference's scheduler publishes in ``status.txt`` (``lib/scheduler.py:311-333``): closed
             simplices, queue length, time elapsed -- read from the 104-byte control block, no
             extra device traffic.
    pinned : fetch into page-locked host buffers cached on the oracle (see ``_Engine.fetch``): the
             result's arrays are valid until the next ``partition(..., pinned=True)`` on this oracle.
    """
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    assert roots.ndim == 3 and roots.shape[1:] == (oracle.n + 1, oracle.n)
    t0 = time.perf_counter()
    eng = _Engine(oracle, roots, wave, max_nodes, max_frontier, node_budget, verbose)
    try:
        if status_path is None:
            eng.advance(0)
        else:
            while True:
                left = eng.advance(max(1, int(status_every)))
                _write_live_status(status_path, eng.stats, time.perf_counter() - t0)
                if left == 0 or (node_budget > 0 and int(eng.stats.nnodes) >= node_budget):
                    break
        arr = eng.fetch(pinned=pinned)
        st = eng.stats_dict()
    finally:
        eng.close()
    st['wall_s'] = time.perf_counter() - t0
    return PartitionResult(arr, st, oracle.pb)


def _write_live_status(path, stats, elapsed):
    with open(path, 'w') as f:
        f.write('\n'.join([
            '# overall',
            'number of tasks queue: %d' % int(stats.open_left),
            'simplex_count: %d' % int(stats.nleaves),
            'nodes: %d' % int(stats.nnodes),
            'max depth: %d' % int(stats.max_depth),
            'oracle solves: %d' % int(stats.nsolves),
            'time elapsed [s]: %d' % int(elapsed),
            'time active (total for all processes [s]): %.0f' % (float(stats.device_ms) * 1e-3),
        ]) + '\n')


def partition_checkpoint(oracle, roots, path, node_budget, wave=0, max_nodes=0, max_frontier=0,
                         verbose=0):
    """Run the partition until about ``node_budget`` nodes exist, then write a checkpoint
    (``.npz``): the node store so far and every open frontier record (geometry, vertex costs
    and inputs, certified per-candidate bounds).  The reference checkpoints by pickling
    ``tree.pkl`` / ``branch_*.pkl`` and re-queueing unfinished leaves on restart
    (``lib/scheduler.py:388-391,620-642``).  Returns the number of open records (0 = the run
    finished and the checkpoint holds the complete tree)."""
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    eng = _Engine(oracle, roots, wave, max_nodes, max_frontier, node_budget, verbose)
    try:
        left = eng.advance(0)
        while left > eng.frontier_size():          # root simplices still waiting for their ECC pass
            left = eng.advance(1)
        k = eng.frontier_size()
        meta, geom, lb = eng.export(k) if k else (None, None, None)
        arr = eng.fetch(with_origin=True)
        st = eng.stats_dict()
    finally:
        eng.close()
    out = {('node_' + f): arr[f] for f in arr}
    if k:
        out.update(meta=meta.cpu().numpy(), geom=geom.cpu().numpy(), lb=lb.cpu().numpy())
    out['open'] = np.int64(k)
    out['nleaves'] = np.int64(st['nleaves'])
    np.savez_compressed(path, **out)
    return k


def partition_resume(oracle, path, wave=0, max_nodes=0, max_frontier=0, verbose=0):
    """Continue a :func:`partition_checkpoint` to convergence; returns the complete
    :class:`PartitionResult` (identical to an uninterrupted run)."""
    import torch
    z = np.load(path)
    first = {f[5:]: z[f] for f in z.files if f.startswith('node_')}
    k = int(z['open'])
    if k == 0:
        merged = stitch([first])
        return PartitionResult(merged, dict(nleaves=int(z['nleaves'])), oracle.pb)
    dev = oracle.device
    eng = _Engine(oracle, np.zeros((0, oracle.n + 1, oracle.n)), wave, max_nodes, max_frontier, 0, verbose)
    try:
        eng.import_(torch.as_tensor(z['meta'], device=dev), torch.as_tensor(z['geom'], device=dev),
                    torch.as_tensor(z['lb'], device=dev),
                    torch.zeros(k, dtype=torch.int32, device=dev))       # origin "rank" 0 = the checkpoint
        eng.advance(0)
        second = eng.fetch(with_origin=True)
        st = eng.stats_dict()
    finally:
        eng.close()
    st['nleaves'] = int(st['nleaves']) + int(z['nleaves'])
    return PartitionResult(stitch([first, second]), st, oracle.pb)


# --------------------------------------------------------------------------------------------
# multi-GPU: sharded frontier + all-to-all rebalance
# --------------------------------------------------------------------------------------------
def shard_roots(roots, rank, world):
    """Deal root simplices round-robin, largest volume first (deterministic)."""
    from .tools import simplex_volume
    order = np.argsort([-simplex_volume(r) for r in roots], kind='stable')
    return np.sort(order[rank::world])


def rebalance_plan(sizes, fill, slack=2.0):
    """Who sends how many frontier records to whom.  sizes: exportable frontier depth per rank;
    fill: the depth a rank needs to keep its GPU busy.  Returns an integer matrix send[i][j].
    Nothing moves unless some rank is starving (depth < fill and below the mean) while the
    deepest frontier is more than slack x the shallowest; then every rank above the mean gives
    its surplus (largest first) to the starving ranks, up to the mean.  Deterministic,
    identical on all ranks."""
    sizes = np.asarray(sizes, dtype=np.int64)
    world = len(sizes)
    send = np.zeros((world, world), dtype=np.int64)
    total = int(sizes.sum())
    if total == 0 or world == 1:
        return send
    mean = total / world
    target = int(np.floor(mean))
    starving = sizes < min(fill, mean)
    if not starving.any() or sizes.max() <= slack * max(int(sizes.min()), 1) or target < 1:
        return send
    surplus = np.maximum(sizes - target, 0)
    deficit = np.where(starving, np.maximum(target - sizes, 0), 0)
    for i in np.argsort(-surplus, kind='stable'):
        s = int(surplus[i])
        for r in np.argsort(-deficit, kind='stable'):
            if s <= 0:
                break
            k = min(s, int(deficit[r]))
            if k <= 0:
                break
            send[i, r] += k
            s -= k
            deficit[r] -= k
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


def deal_plan(sizes):
    """The one-off shuffle that ends the ramp-up: every rank spreads ALL of its frontier records
    evenly over all ranks (its own share included), so that each rank ends up with an
    interleaved sample of every root's sub-trees.  send[i][j] = records rank i gives to rank j."""
    sizes = np.asarray(sizes, dtype=np.int64)
    world = len(sizes)
    send = np.zeros((world, world), dtype=np.int64)
    for i, k in enumerate(sizes):
        # record e of rank i goes to rank (e + i) % world
        dest = (np.arange(int(k)) + i) % world
        send[i] = np.bincount(dest, minlength=world)
    return send


_CALLS = [0]


class _ShmFlags:
    """Starvation flags of one partition job in shared host memory (/dev/shm): all ranks of the
    contract's single node map the same file, so polling after every wave is a local memory read
    instead of a TCP round trip to the rendezvous store (round 1).  One int32 counter per meeting."""

    SLOTS = 1 << 16

    def __init__(self, name, rank, group):
        import torch.distributed as dist
        self.path = os.path.join('/dev/shm', name)
        if rank == 0:
            np.zeros(self.SLOTS, dtype=np.int32).tofile(self.path)
        dist.barrier(group=group)
        self.a = np.memmap(self.path, dtype=np.int32, mode='r+', shape=(self.SLOTS,))
        self.rank = rank

    def raised(self, meeting):
        return int(self.a[meeting % self.SLOTS]) != 0

    def raise_(self, meeting):
        self.a[meeting % self.SLOTS] = 1

    def close(self, group):
        import torch.distributed as dist
        dist.barrier(group=group)
        del self.a
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def _pack_records(meta, geom, lb):
    """one byte buffer per record: meta int32[6] | geom f64[ng] | lb f32[ncand]"""
    import torch
    k = meta.shape[0]
    widths = (24, 8 * geom.shape[1], 4 * lb.shape[1])
    if k == 0:
        return torch.empty(0, sum(widths), dtype=torch.uint8, device=meta.device)
    return torch.cat([t.contiguous().view(torch.uint8).view(k, w) for t, w in zip((meta, geom, lb), widths)],
                     dim=1).contiguous()


def _unpack_records(buf, ng, ncand):
    import torch
    k = buf.shape[0]
    a, b = 24, 24 + 8 * ng
    def dense(t):   # a fresh row-major copy (.contiguous() / .clone() keep the row stride of a one-row slice)
        out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
        out.copy_(t)
        return out
    meta = dense(buf[:, :a]).view(torch.int32).view(k, 6)
    geom = dense(buf[:, a:b]).view(torch.float64).view(k, ng)
    lb = dense(buf[:, b:]).view(torch.float32).view(k, ncand)
    return meta, geom, lb


def _gather_arrays(arr, rank, world, group, dev):
    """Node arrays of every rank -> rank 0, as tensors (one uneven all-to-all per field over the
    process group's backend; round 1 pickled the arrays through gather_object)."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([len(arr['left'])], dtype=torch.int64, device=dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = [int(x) for x in sizes.tolist()]
    parts = [dict() for _ in range(world)] if rank == 0 else None
    for key in list(_FIELDS) + ['origin']:
        a = np.ascontiguousarray(arr[key])
        row = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        src = torch.from_numpy(a.reshape(len(a), row) if a.ndim > 1 else a.reshape(len(a), 1)).to(dev)
        out_split = [s if rank == 0 else 0 for s in sizes]
        in_split = [len(a)] + [0] * (world - 1)
        dst = torch.empty((sum(out_split), row), dtype=src.dtype, device=dev)
        dist.all_to_all_single(dst, src, out_split, in_split, group=group)
        if rank == 0:
            host = dst.cpu().numpy()
            off = 0
            for r in range(world):
                parts[r][key] = host[off:off + sizes[r]].reshape((sizes[r],) + a.shape[1:])
                off += sizes[r]
        del src, dst
    return parts


def _default_store():
    import torch.distributed as dist
    try:
        return dist.distributed_c10d._get_default_store()
    except Exception:
        return None


def partition_distributed(oracle, roots, group=None, waves_per_round=1, wave=0, max_nodes=0,
                          max_frontier=0, verbose=0, engine_cls=_Engine, gather=True,
                          fill_depth=256, round_ms=0.0, deal_depth=512, signal=True, pinned=False):
    """Partition ``roots`` with one process per GPU (see module docstring).

    Every rank runs its own frontier loop on its share of the root simplices.  Ranks meet
    (all-gather of two integers) only when one of them says so: a rank whose frontier falls
    below ``fill_depth`` raises a flag in the process group's key-value store, the others see it
    at their next wave boundary (``signal=True``; the device analogue of a reference worker
    reporting idle, lib/worker.py:419-461).  At a meeting, frontier records move with
    all-to-alls over NCCL: once as a full interleaved shuffle when the ramp-up has produced
    ``deal_depth`` records per rank (:func:`deal_plan`), afterwards only from ranks holding a
    surplus to starving ranks (:func:`rebalance_plan`).  With ``signal=False`` the ranks meet
    every ``waves_per_round`` waves / ``round_ms`` milliseconds instead.

    ``gather=False`` leaves every rank's share of the node store in that rank's host memory (the
    per-rank dump of SURVEY 8e(3); ``info['arrays']``), ``pinned`` as in :func:`partition`.

    Returns (PartitionResult or None, info): rank 0 gets the stitched result when ``gather``;
    ``info`` has the per-rank statistics needed by bench.py (leaves, device seconds, bytes
    exchanged, number of exchange rounds)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = oracle.device
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    mine = shard_roots(roots, rank, world)
    eng = engine_cls(oracle, roots[mine], wave, max_nodes, max_frontier, 0, verbose, root_ids=mine)
    _CALLS[0] += 1
    flags = None
    if signal and world > 1:
        # one file per job: the rendezvous port makes the name unique per launch on this node
        name = 'ehm_b200_%s_%d' % (os.environ.get('MASTER_PORT', 'p'), _CALLS[0])
        try:
            flags = _ShmFlags(name, rank, group)
        except OSError:
            flags = None
    store = _default_store() if (signal and flags is None) else None
    prefix = 'ehm_b200/%d/' % _CALLS[0]
    fill = fill_depth
    sizes = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    rounds = meetings = moved = 0
    dealt = world == 1
    xch_s = 0.0
    t0 = time.perf_counter()
    ev0 = ev1 = None
    if dev.type == 'cuda':
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.default_stream(dev))

    def frontier_of(e):
        return e.frontier_size() if hasattr(e, 'frontier_size') else int(e.stats.open_left)

    try:
        while True:
            if flags is not None:
                while True:
                    mysize = eng.advance(waves_per_round)
                    if mysize < fill or flags.raised(meetings):
                        break
                flags.raise_(meetings)
            elif store is not None:
                key = prefix + str(meetings)
                while True:
                    mysize = eng.advance(waves_per_round)
                    if mysize < fill or store.add(key, 0) > 0:
                        break
                store.add(key, 1)
            elif round_ms > 0:
                mysize = eng.advance(waves_per_round, round_ms)
            else:
                mysize = eng.advance(waves_per_round)
            meetings += 1
            me = torch.tensor([[mysize, frontier_of(eng)]], dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sizes, me, group=group)
            sz = sizes.cpu().numpy()
            if int(sz[:, 0].sum()) == 0:
                break
            fr = sz[:, 1]
            shuffle = False
            if not dealt and int(fr.sum()) >= deal_depth * world and (sz[:, 0] == fr).all():
                plan, dealt, shuffle = deal_plan(fr), True, True
            else:
                plan = rebalance_plan(fr, fill)
            if plan.sum() == 0:
                continue
            tx = time.perf_counter()
            nsend = plan[rank]
            k = int(nsend.sum())
            meta, geom, lb = eng.export(k) if k else (
                torch.empty(0, 6, dtype=torch.int32, device=dev),
                torch.empty(0, eng.geom_width(), dtype=torch.float64, device=dev),
                torch.empty(0, eng.ncand, dtype=torch.float32, device=dev))
            if shuffle and k:
                # interleave: record e goes to rank (e + rank) % world
                order = torch.argsort((torch.arange(k, device=dev) + rank) % world, stable=True)
                meta, geom, lb = meta[order].contiguous(), geom[order].contiguous(), lb[order].contiguous()
            nrecv = plan[:, rank]
            kr = int(nrecv.sum())
            ins, outs = [int(x) for x in nsend], [int(x) for x in nrecv]
            # one packed byte record per frontier entry, one all-to-all (round 1: three)
            sbuf = _pack_records(meta, geom, lb)
            rbuf = torch.empty(kr, sbuf.shape[1], dtype=torch.uint8, device=dev)
            dist.all_to_all_single(rbuf, sbuf, outs, ins, group=group)
            rmeta, rgeom, rlb = _unpack_records(rbuf, eng.geom_width(), eng.ncand)
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
        arr = eng.fetch(with_origin=True, pinned=True) if pinned else eng.fetch(with_origin=True)
        st = eng.stats_dict()
    finally:
        eng.close()
    wall = time.perf_counter() - t0
    info = dict(rank=rank, world=world, wall_s=wall, exchange_s=xch_s, rounds=rounds,
                meetings=meetings, records_sent=moved, stats=st, loop_device_ms=loop_ms,
                arrays=None if gather else arr)
    if flags is not None:
        flags.close(group)
    result = None
    if gather:
        if 'vinput_hat' in arr:
            arr['vinput'] = arr.pop('vinput_hat') * oracle.pb.u_scale
        parts = _gather_arrays(arr, rank, world, group, dev)
        if rank == 0:
            merged = stitch(parts)
            result = PartitionResult(merged, dict(st), oracle.pb)
    return result, info
