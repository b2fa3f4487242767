#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 explicit-hybrid-MPC partitioning engine.

Metric (BASELINE.json): epsilon-converged leaf simplices partitioned per second.  One "step" = one
complete ECC + L-CSS partition of the workload's parameter set (root simplices are committed /
deterministic fixtures; the oracle problems are the workload's MPC problem).

Workloads (``--config``, numbering of SURVEY.md section 8d):
  2  BASELINE configs[1]: inverted pendulum (lib/examples.py:120-163), N=5, abs_frac 0.25 / eps_r 1.0
     (the reference's own pairing, lib/post_process.py:334-338), 66 Delaunay root simplices.
     The partition is checked leaf for leaf against the CPU oracle's full-size fixture.
  3  BASELINE configs[2]: the same pendulum at abs_frac 0.03 / eps_r 0.05 (eps_a = 0.27937 by
     examples.create_oracle), root simplex 22: 114 076 leaves (the whole Theta at this tolerance is ~6e7).
  5  BASELINE configs[4]: synthetic n=8 mp-LP, 4 modes, first 1024 Kuhn simplices of [-1,1]^8,
     eps tuned to >= 10^6 leaves -- the scaling workload (a job long enough to shard over 8 GPUs).
Default: config 5 at every N (the configuration the metric "at 1/2/4/8 B200" is quoted on), with the
configs[1] measurement, its roofline and its parity check reported in the same line under
"pendulum_config2" at N=1.

  value : leaves / device time of the partition loop (CUDA events around the whole loop,
          candidate tables already resident in HBM), max over ranks.
  e2e   : the same through the public Python API `partition(oracle, roots)` with HOST buffers:
          H2D of the roots and D2H of the whole node store inside the timed region.
  --impl reference : the reference algorithm's CPU restatement (oracle/, MOSEK and MPI are not
          installable here) on the box's host cores, bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'simplices partitioned/sec (eps-converged leaves)'
EPS_A_CONFIG2 = 19.21220798973866   # max_v P_theta(0.25 v).J (lib/examples.py:42-46), recomputed on the GPU
EPS_A_CONFIG3 = 0.2793700817913687   # the same for abs_frac 0.03 (tests/golden/tools_golden.npz)
FP64_NOMINAL_TFLOPS = 37.0          # B200 data sheet (HGX B200: 296 TFLOP/s fp64 per 8 GPUs)


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, config):
        self.config = config
        if config in (2, 3):
            from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
            self.mpc = InvertedPendulumOnCart()
            allroots = np.load(os.path.join(ROOT, 'tests', 'golden', 'pendulum_roots.npy'))
            if config == 2:
                self.abs_frac, self.eps_a, self.eps_r = 0.25, EPS_A_CONFIG2, 1.0
                self.root_ids = list(range(len(allroots)))
                self.name = ('pendulum N=5 abs_frac=0.25 eps_r=1.0, 66 Delaunay root simplices '
                             '(BASELINE.json configs[1])')
            else:
                self.abs_frac, self.eps_a, self.eps_r = 0.03, EPS_A_CONFIG3, 0.05
                self.root_ids = [22]
                self.name = ('pendulum N=5 abs_frac=0.03 eps_r=0.05 (eps_a=0.27937), root simplex 22 of the 66: '
                             '114 076 leaves, depth 29 (BASELINE.json configs[2]; the whole Theta is ~6e7 leaves at this tolerance)')
            self.roots = allroots[self.root_ids]
            self.max_nodes = 1 << 22
            self.wave_multi = 1024
            self.max_frontier = 0
        elif config == 5:
            from explicit_hybrid_mpc_b200.mpc_library import SyntheticMPLP
            from explicit_hybrid_mpc_b200.tools import kuhn_triangulation
            self.mpc = SyntheticMPLP()
            self.abs_frac, self.eps_a, self.eps_r = None, 0.038, 0.0114
            self.roots = kuhn_triangulation(-np.ones(8), np.ones(8), limit=1024)
            self.root_ids = list(range(1024))
            self.name = ('synthetic n=8 mp-LP (16 inputs, 80 rows, 4 modes, seed 20260922), first 1024 Kuhn '
                         'simplices of [-1,1]^8, eps_a=0.038 eps_r=0.0114: ~7e6 leaves (BASELINE.json configs[4])')
            self.max_nodes = 40 << 20
            self.wave_multi = 4096
            self.max_frontier = 1 << 21
        else:
            raise SystemExit('unknown --config %r (2, 3 or 5)' % config)
        self.n, self.nu = self.mpc.problem.n_x, self.mpc.problem.n_u

    def node_bytes(self):
        n, nu = self.n, self.nu
        return 5 * 4 + 1 + 8 * ((n + 1) * n + (n + 1) + (n + 1) * nu)

    def cpu_oracle(self):
        from oracle.oracle_cpu import CpuOracle
        return CpuOracle(self.mpc.problem, self.eps_a, self.eps_r, self.mpc.theta_scale)

    def cpu_sample(self, threads=None, seconds=12.0):
        """Bounded sample of the workload for the CPU legs.  config 2: the cheapest root simplices
        by the leaf count of the oracle's full-size fixture, as many as keep `threads` host threads busy
        for about `seconds` (rate from a calibration run on the four cheapest: it under-estimates,
        their small trees cannot feed many threads for long, hence the 1.5).  config 3: root simplex 22;
        config 5: every 128th root simplex (the cost per root varies by orders of magnitude over the box)."""
        if self.config == 5:
            return list(range(0, 1024, 128))
        if self.config == 3:
            return [0]          # index into self.roots: root simplex 22
        p = os.path.join(ROOT, 'tests', 'golden', 'pendulum_full_leafsets.json')
        cnt = json.load(open(p))['leaves_per_root']
        order = [int(i) for i in np.argsort(cnt, kind='stable')]
        if threads is None:
            return order[:6]
        cal = cpu_leg(self, threads, order[:4])
        budget = 1.5 * cal['value'] * seconds          # leaves
        sample, tot = [], 0
        for i in order:
            if len(sample) >= 6 and tot + cnt[i] > budget:
                break
            sample.append(i)
            tot += cnt[i]
        return sample


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


def cpu_leg(wl, threads, sample):
    """Time the CPU oracle (kind 'port') on a bounded sample (indices into wl.roots)."""
    orc = wl.cpu_oracle()
    roots = wl.roots[sample]
    t0 = time.perf_counter()
    out = orc.partition(roots, nthreads=threads)
    dt = time.perf_counter() - t0
    return dict(leaves=out['nleaves'], seconds=dt, solves=out['nsolves'], threads=threads,
                value=out['nleaves'] / dt, solves_per_s=out['nsolves'] / dt, nfail=out['nfail'])


def _sample_text(wl, sample, r):
    ids = [wl.root_ids[i] for i in sample]
    return ('root simplices %s of the same workload (%d leaves, %.1f s on %d threads)'
            % (ids if len(ids) <= 40 else '%d..%d' % (ids[0], ids[-1]), r['leaves'], r['seconds'], r['threads']))


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = Workload(args.config)
    threads = _host_cores()
    sample = wl.cpu_sample(threads)
    vals, ms = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_leg(wl, threads, sample if i >= args.warmup else sample[:1])
        if i >= args.warmup:
            vals.append(r)
            ms.append(r['seconds'] * 1e3)
    leaves = sum(v['leaves'] for v in vals)
    secs = sum(v['seconds'] for v in vals)
    value = leaves / secs
    line = dict(metric=METRIC, value=value, unit='leaves/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=float(np.mean(ms)), higher_is_better=True,
                scaling='strong', vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload=wl.name, sample=[wl.root_ids[i] for i in sample],
                            note='CPU restatement of the reference ALGORITHM (oracle/: condensed problems, row '
                                 'pre-solve, certified pruning -- far faster than the reference\'s CVXPY->MOSEK path, '
                                 '~10 solves/s/core by lib/post_process.py:328-330, which is not installable here); '
                                 'each step partitions the bounded sample of root simplices listed in "sample"; the '
                                 'GPU arm prints the same sample under "same_sample"'),
                cpu_baseline=dict(value=value, unit='leaves/s', cores=threads, kind='port',
                                  sample=_sample_text(wl, sample, vals[-1]),
                                  solves_per_s=sum(v['solves'] for v in vals) / secs),
                e2e=dict(value=value, unit='leaves/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def _peaks():
    from explicit_hybrid_mpc_b200 import _capi
    pk, hb = ctypes.c_double(), ctypes.c_double()
    _capi.check(_capi.lib().ehm_fp64_peak(ctypes.byref(pk)))
    _capi.check(_capi.lib().ehm_hbm_peak(1 << 30, ctypes.byref(hb)))
    hbm_peak, hbm_src = 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            hbm_peak = float(json.load(open(p))['hbm_gbs'])
            hbm_src = 'MEASURED_PEAKS.json'
        except Exception:
            pass
    return pk.value, hb.value, hbm_peak, hbm_src


STAGES = ['k_pop', 'k_lcss_ref', 'k_lcss_items', 'k_lcss_eval', 'k_lcss_select', 'k_lcss_points',
          'k_lcss_final', 'ecc_phase']


def _roofline(st, fp64_peak, config):
    """fp64 roofline of the dominant IPM kernel of the last step."""
    sm, fl = st['stage_ms'], st['stage_flops']
    solve = [1, 3, 5, 7]
    dom = max(solve, key=lambda k: sm[k])
    ach = fl[dom] / (sm[dom] * 1e-3) / 1e12 if sm[dom] > 0 else 0.0
    solve_ms = sum(sm[k] for k in solve)
    solve_fl = sum(fl[k] for k in solve)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(str(config))
            if tj:
                traffic, traffic_src = tj['dram_bytes_per_launch'], tj['source']
        except Exception:
            pass
    return dict(bound='fp64', kernel='%s (warp-per-problem interior-point solver)' % STAGES[dom], achieved=ach,
                peak=fp64_peak, unit='TFLOP/s', frac=ach / fp64_peak if fp64_peak else None,
                frac_of_nominal=ach / FP64_NOMINAL_TFLOPS, nominal_peak=FP64_NOMINAL_TFLOPS,
                peak_source='in-run DFMA microbenchmark ehm_fp64_peak (MEASURED_PEAKS.json has no fp64 entry); '
                            'nominal 37 TFLOP/s = B200 data sheet; DMMA alone measured 37.2 TFLOP/s '
                            '(scripts/ubench/dmma_bench.cu, profiles/README.md)',
                traffic=traffic, traffic_unit='dram bytes of ONE saturated launch (not per step)',
                traffic_source=traffic_src,
                note='fp64 FMA-pipe bound (SURVEY 8d): achieved = algorithmic flops of all launches of the '
                     'dominant kernel in the last step / their summed CUDA-event time; all IPM kernels together: '
                     '%.2f TFLOP/s over %.1f ms' % (solve_fl / (solve_ms * 1e-3) / 1e12 if solve_ms else 0.0, solve_ms),
                stage_ms={STAGES[k]: round(sm[k], 3) for k in range(8)}, dominant_stage=STAGES[dom])


def _check_config2(res):
    """The partition of configs[1] against the CPU oracle's full-size fixture: leaves per root and
    sha256 of the sorted leaf vertex sets (scripts/make_full_leafsets.py)."""
    from explicit_hybrid_mpc_b200.partition import leafset_digest
    fx = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'pendulum_full_leafsets.json')))
    cnt, dig = leafset_digest(res.left, res.root_id, res.vertices, fx['nroots'])
    ok = cnt == fx['leaves_per_root'] and dig == fx['sha256_per_root']
    return dict(fixture='tests/golden/pendulum_full_leafsets.json (CPU oracle, all 66 roots)',
                leaves=int(sum(cnt)), fixture_leaves=int(fx['nleaves']), leaf_sets_bit_exact=bool(ok))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--config', type=int, default=5, help='workload: 2, 3 or 5 (see module docstring)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the configs[1] block of the default run')
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
    from explicit_hybrid_mpc_b200.oracle import Oracle
    from explicit_hybrid_mpc_b200.partition import partition, partition_distributed

    wl = Workload(args.config)
    orc = Oracle(wl.mpc, 1.0, 1.0)
    eps_a = wl.eps_a
    if wl.abs_frac is not None:
        # eps_a exactly as examples.create_oracle (lib/examples.py:42-46): max over the vertices of
        # abs_frac * Theta of the MICP optimal cost -- one batched P_theta call on the device
        import itertools
        box = np.array(list(itertools.product(*[(-s, s) for s in wl.mpc.theta_scale])))
        _, Jb, _ = orc.P_theta_batch(wl.abs_frac * box)
        eps_a = float(Jb.max().item())
        assert abs(eps_a - wl.eps_a) <= 1e-6 * wl.eps_a, eps_a
    orc.set_eps(eps_a, wl.eps_r)
    roots = wl.roots

    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    def step(o=orc, rts=roots, max_nodes=wl.max_nodes):
        flush.fill_(1)                      # evict L2 between timed iterations
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world == 1:
            res = partition(o, rts, wave=args.wave, max_nodes=max_nodes, max_frontier=wl.max_frontier, pinned=True)
            info = dict(exchange_s=0.0, rounds=0, records_sent=0)
        else:
            # one partition job over all ranks: sharded frontier + all-to-all rebalance; every rank
            # copies its share of the node store to its own host buffers inside the timed region
            # (the per-rank dump of SURVEY 8e(3); stitching to one forest is tests/test_gpu_distributed.py)
            res, info = partition_distributed(o, rts, gather=False, pinned=True, wave=args.wave or wl.wave_multi,
                                              waves_per_round=1 if args.round_ms <= 0 else 0,
                                              round_ms=args.round_ms, signal=args.round_ms <= 0,
                                              fill_depth=args.fill, deal_depth=args.deal,
                                              max_nodes=max_nodes, max_frontier=wl.max_frontier)
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
        if len(results) > 1:
            results[-2] = type('R', (), dict(stats=results[-2].stats))()   # keep only the statistics
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_total = time.perf_counter() - t_begin
    launches = _capi.lib().ehm_kernel_launches() - l0
    stop.set()
    th.join(timeout=2)

    res = results[-1]
    leaves = sum(r.stats['nleaves'] for r in results)
    nfail = sum(r.stats['nfail'] for r in results)
    nn = res.stats['nnodes']
    d2h = nn * wl.node_bytes()
    h2d = roots.nbytes // world
    tot = torch.tensor([float(leaves), float(sum(r.stats['nsolves'] for r in results)),
                        float(launches), float(nfail)], dtype=torch.float64, device='cuda')
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
    leaves_all, solves_all, launches_all, nfail_all = tot.tolist()
    dev_s, wall_s = mx.tolist()
    value = leaves_all / dev_s
    e2e = leaves_all / wall_s

    fp64_peak, hbm_inrun, hbm_peak, hbm_src = _peaks()
    st = res.stats
    roofline = _roofline(st, fp64_peak, wl.config)
    sm = st['stage_ms']
    pop_gbs = st['stage_bytes'][0] / (sm[0] * 1e-3) / 1e9 if sm[0] > 0 else 0.0
    roofline_hbm = dict(bound='hbm', kernel='k_pop (frontier record stream)', achieved=pop_gbs,
                        peak=hbm_peak, unit='GB/s', frac=pop_gbs / hbm_peak, traffic=None,
                        peak_source=hbm_src, in_run_copy_gbs=hbm_inrun)

    parity = None
    if wl.config == 2 and world == 1:
        parity = _check_config2(res)
        assert parity['leaf_sets_bit_exact'], parity
    if nfail_all:   # reported in config.nfail; a failed sub-problem is skipped for one visit, never inherited
        print('WARNING: %d numerical failures in the timed partitions' % nfail_all, file=sys.stderr)

    cpu, same_sample = None, None
    if not args.no_cpu_baseline and world == 1:
        threads = _host_cores()
        sample = wl.cpu_sample(threads)
        c = cpu_leg(wl, threads, sample)
        cpu = dict(value=c['value'], unit='leaves/s', cores=threads, kind='port',
                   sample=_sample_text(wl, sample, c), solves_per_s=c['solves_per_s'],
                   note='CPU restatement of the reference algorithm with all of this repository\'s condensing, '
                        'pre-solve and pruning (not MOSEK: ~10 solves/s/core, lib/post_process.py:328-330)')
        # the GPU on exactly the sample the CPU legs partition: a same-work ratio
        rs = partition(orc, roots[sample], wave=args.wave, max_nodes=wl.max_nodes, max_frontier=wl.max_frontier)
        rs2 = partition(orc, roots[sample], wave=args.wave, max_nodes=wl.max_nodes, max_frontier=wl.max_frontier)
        g = rs2.stats['nleaves'] / (rs2.stats['device_ms'] * 1e-3)
        same_sample = dict(sample=[wl.root_ids[i] for i in sample], gpu_leaves=int(rs2.stats['nleaves']),
                           cpu_leaves=int(c['leaves']), gpu_leaves_per_s=g, cpu_leaves_per_s=c['value'],
                           gpu_ms=rs2.stats['device_ms'], cpu_s=c['seconds'], ratio=g / c['value'],
                           same_leaf_count=bool(rs.stats['nleaves'] == c['leaves']),
                           note='identical root simplices on both sides (the sample is too small to fill the GPU: '
                                'the full-workload number is the headline)')

    # secondary: BASELINE configs[1] (tree-equality config) in the same line of the default run
    secondary = None
    if wl.config == 5 and world == 1 and not args.no_secondary:
        w2 = Workload(2)
        o2 = Oracle(w2.mpc, w2.eps_a, w2.eps_r)
        rr = None
        ms2 = []
        for i in range(2 + 2):
            flush.fill_(1)
            torch.cuda.synchronize()
            rr = partition(o2, w2.roots)
            if i >= 2:
                ms2.append(rr.stats['device_ms'])
        par = _check_config2(rr)
        secondary = dict(workload=w2.name, value=rr.stats['nleaves'] / (np.mean(ms2) * 1e-3), unit='leaves/s',
                         ms_per_step=float(np.mean(ms2)), steps=2, warmup=2, leaves_per_step=int(rr.stats['nleaves']),
                         nfail=int(rr.stats['nfail']), parity=par,
                         oracle_solves_per_s=rr.stats['nsolves'] / (rr.stats['device_ms'] * 1e-3),
                         solves_per_leaf=rr.stats['nsolves'] / max(rr.stats['nleaves'], 1),
                         roofline=_roofline(rr.stats, fp64_peak, 2))
        del o2, rr

    # secondary: online evaluation of the produced control law, not timed above
    online = None
    if world == 1 and wl.config in (2, 3):
        try:
            from explicit_hybrid_mpc_b200.mpc_library import ExplicitMPC
            empc = ExplicitMPC(res)
            nq = 1 << 21
            X = (torch.rand(nq, wl.n, dtype=torch.float64, device='cuda') * 2 - 1) * torch.tensor(
                wl.mpc.theta_scale, device='cuda')
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
        config=dict(workload=wl.name, survey_config=wl.config,
                    eps_a=eps_a, eps_r=wl.eps_r, leaves_per_step=leaves_all / args.steps,
                    nodes_per_step=nn if world == 1 else None, nfail=int(nfail_all),
                    parallelism=('single GPU' if world == 1 else 'frontier sharded over %d ranks: roots dealt '
                                 'round-robin; ranks meet (all-gather of depths) when one raises a starvation flag%s; '
                                 'NCCL all-to-all of frontier records: one interleaved shuffle after the ramp-up, '
                                 'then surplus -> starving ranks (%d meetings, %d exchanges, %d records sent by '
                                 'rank 0 in the last step)'
                                 % (world, '' if args.round_ms <= 0 else ' / every %.0f ms' % args.round_ms,
                                    xch[-1].get('meetings', 0), xch[-1]['rounds'], xch[-1]['records_sent'])),
                    wave=(args.wave or 4096) if world == 1 else (args.wave or wl.wave_multi),
                    l2='flushed between timed iterations (256 MiB write)'),
        e2e=dict(value=e2e, unit='leaves/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
        oracle_solves_per_s=solves_all / dev_s, solves_per_leaf=solves_all / max(leaves_all, 1.0),
        gpu_launches=int(launches_all),
        clocks=_clocks_summary(samples),
        roofline=roofline, roofline_hbm=roofline_hbm, cpu_baseline=cpu, same_sample=same_sample,
        parity=parity, pendulum_config2=secondary,
        online_evaluation=online,
        rank_busy_ms_last_step=[round(x, 1) for x in busy_all.tolist()],
        wall_s_timed_region=t_total)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
