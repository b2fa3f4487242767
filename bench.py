#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 explicit-hybrid-MPC partitioning engine.

Metric (BASELINE.json): epsilon-converged leaf simplices partitioned per second.
Workload (BASELINE.json configs[1]): inverted pendulum (lib/examples.py:120-163), N=5,
abs_frac=0.25 / eps_r=1.0 (the reference's own pairing, lib/post_process.py:334-338), the 66
root simplices of tests/golden/pendulum_roots.npy.  One "step" = one complete ECC + L-CSS
partition of that parameter set (synthetic in the sense that the root simplices are a
committed fixture; the oracle problems are the reference's pendulum MPC).

  value : leaves / device time of the partition loop (CUDA events around the whole loop,
          candidate tables already resident in HBM), max over ranks.
  e2e   : the same through the public Python API `partition(oracle, roots)` with HOST buffers:
          H2D of the roots and D2H of the whole node store inside the timed region.
  --impl reference : the reference algorithm's CPU restatement (oracle/, MOSEK and MPI are not
          installable here) on the box's host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ABS_FRAC, EPS_R = 0.25, 1.0
EPS_A_FIXTURE = 19.21220798973866   # max_v P_theta(0.25 v).J, recomputed on the GPU at start-up
METRIC = 'simplices partitioned/sec (eps-converged leaves)'
# bounded CPU sample: the cheapest root simplices of the workload (by leaf count), as many as keep
# the host threads busy for ~12 s; the CPU oracle shares the work of ALL sampled roots over all
# threads (shared LIFO task queue)


def _clock_sampler(stop, out, gpu_index):
    q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not stop.is_set():
        try:
            r = subprocess.run(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + q,
                                '--format=csv,noheader,nounits'], capture_output=True, text=True,
                               timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                out.append(r.stdout.strip().split(', '))
        except Exception:
            pass
        stop.wait(0.2)


def _clocks_summary(samples):
    if not samples:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
    sm = sorted(int(s[1]) for s in samples if s[1].isdigit())
    mx = max(int(s[2]) for s in samples if s[2].isdigit())
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [nm for k, nm in enumerate(names)
               if any(len(s) > 5 + k and s[5 + k].strip() == 'Active' for s in samples)]
    return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=reasons,
                samples=len(samples))


def _host_cores():
    """Host threads this process can really use: the affinity mask, capped by the cgroup CPU
    quota (the GPU boxes report 128 CPUs but run under a 16-CPU quota; more threads than the
    quota only adds contention: 16 threads 554k solves/s, 128 threads 337k)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        q = open('/sys/fs/cgroup/cpu.max').read().split()
        if q[0] != 'max':
            n = min(n, max(1, int(float(q[0]) / float(q[1]) + 0.5)))
    except Exception:
        try:
            quota = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
            period = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
            if quota > 0:
                n = min(n, max(1, int(quota / period + 0.5)))
        except Exception:
            pass
    return n


def _roots():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'pendulum_roots.npy'))


def _cpu_sample(threads=None, seconds=12.0):
    """Bounded sample of the workload for the CPU legs: the cheapest root simplices (by leaf
    count of the committed full-size result) whose estimated work keeps `threads` host threads
    busy for about `seconds`.  The rate comes from a calibration run on the four cheapest roots
    (it under-estimates: their small trees cannot feed many threads for long), hence the 1.5."""
    p = os.path.join(ROOT, 'tests', 'golden', 'pendulum_root_leafcounts.json')
    if not os.path.exists(p):
        return [22, 23, 24, 25]
    meta = json.load(open(p))
    cnt = meta['leaves_per_root']
    order = [int(i) for i in np.argsort(cnt, kind='stable')]
    if threads is None:
        return order[:6]
    cal = cpu_leg(threads, order[:4])
    budget = 1.5 * cal['value'] * seconds          # leaves
    sample, tot = [], 0
    for i in order:
        if len(sample) >= 6 and tot + cnt[i] > budget:
            break
        sample.append(i)
        tot += cnt[i]
    return sample


def cpu_leg(threads, sample):
    """Time the CPU oracle (kind 'port') on a bounded sample of the workload."""
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    mpc = InvertedPendulumOnCart()
    orc = CpuOracle(mpc.problem, EPS_A_FIXTURE, EPS_R, mpc.theta_scale)
    roots = _roots()[sample]
    t0 = time.perf_counter()
    out = orc.partition(roots, nthreads=threads)
    dt = time.perf_counter() - t0
    return dict(leaves=out['nleaves'], seconds=dt, solves=out['nsolves'], threads=threads,
                value=out['nleaves'] / dt, solves_per_s=out['nsolves'] / dt)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = _host_cores()
    sample = _cpu_sample(threads)
    vals, ms = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_leg(threads, sample if i >= args.warmup else sample[:1])
        if i >= args.warmup:
            vals.append(r)
            ms.append(r['seconds'] * 1e3)
    leaves = sum(v['leaves'] for v in vals)
    secs = sum(v['seconds'] for v in vals)
    value = leaves / secs
    line = dict(metric=METRIC, value=value, unit='leaves/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=float(np.mean(ms)), higher_is_better=True,
                scaling='strong', vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload='pendulum N=5 abs_frac=0.25 eps_r=1.0, 66 Delaunay root simplices',
                            note='CPU restatement of the reference algorithm (MOSEK/CVXPY/mpi4py '
                                 'not installable); bounded sample of root simplices %s' % sample),
                cpu_baseline=dict(value=value, unit='leaves/s', cores=threads, kind='port',
                                  sample='%d cheapest root simplices (sized by a calibration run), %d leaves, %.1f s/step'
                                         % (len(sample), vals[-1]['leaves'], vals[-1]['seconds']),
                                  solves_per_s=sum(v['solves'] for v in vals) / secs),
                e2e=dict(value=value, unit='leaves/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--wave', type=int, default=0, help='frontier entries per wave (0: library default)')
    ap.add_argument('--round-ms', type=float, default=0.0,
                    help='N>1: meet every so many ms instead of on a starvation flag (0: flag)')
    ap.add_argument('--fill', type=int, default=256, help='N>1: frontier depth below which a rank asks for work')
    ap.add_argument('--deal', type=int, default=512, help='N>1: frontier depth per rank at which the one-off shuffle happens')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from explicit_hybrid_mpc_b200 import _capi
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.oracle import Oracle
    from explicit_hybrid_mpc_b200.partition import partition

    mpc = InvertedPendulumOnCart()
    roots_all = _roots()
    # eps_a exactly as examples.create_oracle (lib/examples.py:42-46): max over the vertices of
    # abs_frac * Theta of the MICP optimal cost -- 16 batched P_theta calls on the device
    import itertools
    orc = Oracle(mpc, 1.0, 1.0)
    box = np.array(list(itertools.product(*[(-s, s) for s in mpc.theta_scale])))
    eps_a = max(orc.P_theta(ABS_FRAC * v)[2] for v in box)
    assert abs(eps_a - EPS_A_FIXTURE) <= 1e-6 * EPS_A_FIXTURE, eps_a
    orc.set_eps(eps_a, EPS_R)
    from explicit_hybrid_mpc_b200.partition import partition_distributed
    roots = roots_all           # world == 1: everything; world > 1: sharded inside the driver

    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    def step():
        flush.fill_(1)                      # evict L2 between timed iterations
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world == 1:
            res = partition(orc, roots, wave=args.wave)
            info = dict(exchange_s=0.0, rounds=0, records_sent=0)
        else:
            # one partition job over all ranks: sharded frontier + all-to-all rebalance; the
            # stitched tree is gathered to rank 0 inside the timed region (end-to-end)
            res, info = partition_distributed(orc, roots, gather=True, wave=args.wave or 1024,
                                              waves_per_round=1 if args.round_ms <= 0 else 0,
                                              round_ms=args.round_ms, signal=args.round_ms <= 0,
                                              fill_depth=args.fill, deal_depth=args.deal)
            if res is None:
                res = type('R', (), dict(stats=info['stats'], nnodes=info['stats']['nnodes']))()
            else:
                res.stats = info['stats']
        torch.cuda.synchronize()
        return res, time.perf_counter() - t0, info

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    samples, stop = [], threading.Event()
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, local), daemon=True)
    th.start()
    l0 = _capi.lib().ehm_kernel_launches()
    t_begin = time.perf_counter()
    dev_ms, walls, results, xch = [], [], [], []
    for _ in range(args.steps):
        res, wall, info = step()
        # device time: one GPU = CUDA events around the loop inside the library; N GPUs = CUDA
        # events around the whole sharded loop of this rank (waits for the other ranks included)
        dev_ms.append(info['loop_device_ms'] if world > 1 else res.stats['device_ms'])
        walls.append(wall)
        results.append(res)
        xch.append(info)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_total = time.perf_counter() - t_begin
    launches = _capi.lib().ehm_kernel_launches() - l0
    stop.set()
    th.join(timeout=2)

    res = results[-1]
    leaves = sum(r.stats['nleaves'] for r in results)
    nn = res.stats['nnodes']
    n, nu = 4, 1
    d2h = nn * (5 * 4 + 1 + 8 * ((n + 1) * n + (n + 1) + (n + 1) * nu))
    h2d = roots.nbytes // world
    tot = torch.tensor([float(leaves), float(sum(r.stats['nsolves'] for r in results)),
                        float(launches)], dtype=torch.float64, device='cuda')
    mx = torch.tensor([sum(dev_ms) * 1e-3, sum(walls)], dtype=torch.float64, device='cuda')
    busy = torch.tensor([results[-1].stats['device_ms']], dtype=torch.float64, device='cuda')
    busy_all = torch.zeros(world, dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_gather_into_tensor(busy_all, busy)
    else:
        busy_all = busy
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    leaves_all, solves_all, launches_all = tot.tolist()
    dev_s, wall_s = mx.tolist()
    value = leaves_all / dev_s
    e2e = leaves_all / wall_s

    # ---- roofline of the dominant kernel (k_lcss_eval: batched fp64 IPM) --------------------
    st = res.stats
    pk = __import__('ctypes').c_double()
    _capi.check(_capi.lib().ehm_fp64_peak(__import__('ctypes').byref(pk)))
    hb = __import__('ctypes').c_double()
    _capi.check(_capi.lib().ehm_hbm_peak(1 << 30, __import__('ctypes').byref(hb)))
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    hbm_peak, hbm_src = 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    if os.path.exists(peaks_path):
        try:
            hbm_peak = float(json.load(open(peaks_path))['hbm_gbs'])
            hbm_src = 'MEASURED_PEAKS.json'
        except Exception:
            pass
    stage_names = ['k_pop', 'k_lcss_ref', 'k_lcss_items', 'k_lcss_eval', 'k_lcss_select',
                   'k_lcss_points', 'k_lcss_final', 'ecc_phase']
    sm = st['stage_ms']
    dom = int(np.argmax(sm))
    solve_ms = sm[1] + sm[3] + sm[5] + sm[7]
    solve_fl = st['stage_flops'][1] + st['stage_flops'][3] + st['stage_flops'][5] + st['stage_flops'][7]
    ach = st['stage_flops'][3] / (sm[3] * 1e-3) / 1e12 if sm[3] > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r1_final_traffic.json')
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
        except Exception:
            pass
    roofline = dict(bound='fp64', kernel='k_lcss_eval (warp-per-problem IPM)', achieved=ach,
                    peak=pk.value, unit='TFLOP/s', frac=ach / pk.value if pk.value else None,
                    traffic=traffic, traffic_unit='dram bytes of ONE saturated launch (not per step)',
                    traffic_source=traffic_src,
                    note='fp64 FMA-pipe bound (SURVEY 8d): peak = in-run DFMA microbenchmark '
                         '(ehm_fp64_peak); achieved = algorithmic flops of all k_lcss_eval launches '
                         'of the last step / their summed CUDA-event time; all IPM kernels together: '
                         '%.2f TFLOP/s over %.1f ms' % (solve_fl / (solve_ms * 1e-3) / 1e12, solve_ms),
                    stage_ms={stage_names[k]: round(sm[k], 3) for k in range(8)},
                    dominant_stage=stage_names[dom])
    pop_gbs = st['stage_bytes'][0] / (sm[0] * 1e-3) / 1e9 if sm[0] > 0 else 0.0
    roofline_hbm = dict(bound='hbm', kernel='k_pop (frontier record stream)', achieved=pop_gbs,
                        peak=hbm_peak, unit='GB/s', frac=pop_gbs / hbm_peak, traffic=None,
                        peak_source=hbm_src, in_run_copy_gbs=hb.value)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = _host_cores()
        sample = _cpu_sample(threads)
        c = cpu_leg(threads, sample)
        cpu = dict(value=c['value'], unit='leaves/s', cores=threads, kind='port',
                   sample='%d cheapest root simplices of the same workload (%d leaves, %.1f s)'
                          % (len(sample), c['leaves'], c['seconds']),
                   solves_per_s=c['solves_per_s'])

    # secondary: online evaluation of the produced control law, not timed above
    online = None
    if world == 1:
        try:
            from explicit_hybrid_mpc_b200.mpc_library import ExplicitMPC
            empc = ExplicitMPC(res)
            nq = 1 << 21
            X = (torch.rand(nq, 4, dtype=torch.float64, device='cuda') * 2 - 1) * torch.tensor(
                mpc.theta_scale, device='cuda')
            empc.evaluate_batch(X[:1024])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            U_, leaf_ = empc.evaluate_batch(X)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            online = dict(kernel='k_locate', queries=nq, ms=ms, evals_per_s=nq / (ms * 1e-3),
                          tree_nodes=int(len(empc.flat['left'])),
                          note='ExplicitMPC.evaluate_batch on the partition of this run; the '
                               'reference evaluates one state in 0.15-0.5 ms on a CPU '
                               '(lib/post_process.py:331-338)')
            empc.close()
        except Exception as ex:   # never let the secondary metric break the headline line
            online = dict(error=repr(ex))

    line = dict(
        metric=METRIC, value=value, unit='leaves/s', n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=dev_s * 1e3 / args.steps, higher_is_better=True,
        scaling='strong', vs_baseline=None, dtype='f64', data='synthetic',
        config=dict(workload='pendulum N=5 abs_frac=0.25 eps_r=1.0, 66 Delaunay root simplices '
                             '(BASELINE.json configs[1])',
                    eps_a=eps_a, leaves_per_step=leaves_all / args.steps,
                    nodes_per_step=nn if world == 1 else None,
                    parallelism=('single GPU' if world == 1 else 'frontier sharded over %d ranks: roots dealt '
                                 'round-robin; ranks meet (all-gather of depths) when one raises a starvation flag%s; '
                                 'NCCL all-to-all of frontier records: one interleaved shuffle after the ramp-up, '
                                 'then surplus -> starving ranks (%d meetings, %d exchanges, %d records sent by '
                                 'rank 0 in the last step)'
                                 % (world, '' if args.round_ms <= 0 else ' / every %.0f ms' % args.round_ms,
                                    xch[-1].get('meetings', 0), xch[-1]['rounds'], xch[-1]['records_sent'])),
                    wave=(args.wave or 4096) if world == 1 else (args.wave or 1024),
                    l2='flushed between timed iterations (256 MiB write)'),
        e2e=dict(value=e2e, unit='leaves/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
        oracle_solves_per_s=solves_all / dev_s,
        gpu_launches=int(launches_all),
        clocks=_clocks_summary(samples),
        roofline=roofline, roofline_hbm=roofline_hbm, cpu_baseline=cpu,
        online_evaluation=online,
        rank_busy_ms_last_step=[round(x, 1) for x in busy_all.tolist()],
        wall_s_timed_region=t_total)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
