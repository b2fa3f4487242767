"""Multi-GPU path on real GPUs (skipped on single-GPU boxes): two ranks over NCCL partition
the same root set with frontier rebalancing; the stitched forest must equal the single-GPU one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, sel, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.oracle import Oracle
    from explicit_hybrid_mpc_b200.partition import partition, partition_distributed
    roots = np.load(os.path.join(ROOT, 'tests', 'golden', 'pendulum_roots.npy'))[sel]
    orc = Oracle(InvertedPendulumOnCart(), 19.21220798973866, 1.0)
    # small waves + low fill depth so that the rebalance path is exercised on a small job
    res, info = partition_distributed(orc, roots, waves_per_round=1, wave=64, fill_depth=8)
    if rank == 0:
        single = partition(orc, roots)
        np.savez(out_path, rounds=info['rounds'],
                 **{'d_' + k: getattr(res, k) for k in ('left', 'right', 'root_id', 'depth', 'cand', 'closed', 'vertices', 'vcost', 'vinput')},
                 **{'s_' + k: getattr(single, k) for k in ('left', 'right', 'root_id', 'depth', 'cand', 'closed', 'vertices', 'vcost', 'vinput')})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_partition_equals_single(tmp_path):
    import torch.multiprocessing as mp
    from test_gpu_tree_parity import compare_trees
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'trees.npz')
    sel = [22, 44, 8, 14]
    mp.spawn(_worker, args=(2, port, sel, out), nprocs=2, join=True)
    z = np.load(out)
    d = {k[2:]: z[k] for k in z.files if k.startswith('d_')}
    s1 = {k[2:]: z[k] for k in z.files if k.startswith('s_')}
    # root ids are positions in `sel` on both sides
    assert int(z['rounds']) >= 1, 'rebalance path not exercised'
    assert len(d['left']) == len(s1['left'])
    compare_trees(d, s1, len(sel))
