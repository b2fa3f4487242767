"""Thread scaling of the CPU oracle's partition on this host (diagnostic for bench.py's CPU legs)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))


def main():
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    mpc = InvertedPendulumOnCart()
    roots = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'pendulum_roots.npy'))
    cnt = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                                      'pendulum_root_leafcounts.json')))['leaves_per_root']
    order = [int(i) for i in np.argsort(cnt, kind='stable')]
    orc = CpuOracle(mpc.problem, 19.21220798973866, 1.0, mpc.theta_scale)
    print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
    for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us', '/sys/fs/cgroup/cpu/cpu.cfs_period_us'):
        if os.path.exists(f):
            print(f, open(f).read().strip())
    try:
        print(open('/proc/cpuinfo').read().split('model name')[1].split('\n')[0])
    except Exception:
        pass
    sample = order[:int(sys.argv[1]) if len(sys.argv) > 1 else 10]
    for nt in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '8,16,32,64,128').split(',')]:
        t0 = time.perf_counter()
        r = orc.partition(roots[sample], nthreads=nt)
        dt = time.perf_counter() - t0
        print('threads %4d  %7.2f s  %8.0f solves/s  %7.1f leaves/s' % (nt, dt, r['nsolves'] / dt, r['nleaves'] / dt), flush=True)


if __name__ == '__main__':
    main()
