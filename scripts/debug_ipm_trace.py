"""Debug aid: find sub-problems where the CUDA solver and the CPU oracle disagree and print the
CUDA solver's per-iteration trace (needs a GPU)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from explicit_hybrid_mpc_b200 import _capi  # noqa: E402
from explicit_hybrid_mpc_b200._capi import SX_TQP, SX_ALIVE  # noqa: E402
from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart  # noqa: E402
from explicit_hybrid_mpc_b200.oracle import Oracle  # noqa: E402
from oracle.oracle_cpu import CpuOracle  # noqa: E402

mpc = InvertedPendulumOnCart()
eps = (19.21220798973866, 1.0)
gpu, cpu = Oracle(mpc, *eps), CpuOracle(mpc.problem, *eps, mpc.theta_scale)
roots = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'pendulum_roots.npy'))
L = _capi.lib()
L.ehm_debug_trace.argtypes = [C.c_int]
L.ehm_debug_fetch.argtypes = [C.c_void_p]
R = roots[0]
d, vx = cpu.V_R(R)
Vr = np.array([v[1] for v in vx])
sub, t_c = cpu.bar_E_delta_R(R, Vr, return_tstar=True)
val, sol, st = gpu.simplex_batch(np.repeat(R[None], gpu.ncand, 0), np.repeat(Vr[None], gpu.ncand, 0),
                                 gpu._all, SX_TQP, True)
val, st = val.cpu().numpy(), st.cpu().numpy()
bad = [c for c in range(gpu.ncand) if (st[c] == 1) != (not np.isnan(t_c[c])) or
       (st[c] == 1 and abs(val[c] - t_c[c]) > 1e-6 * max(1, abs(t_c[c])))]
print('status counts', {k: int((st == k).sum()) for k in (-1, 0, 1)}, 'cpu alive', int((~np.isnan(t_c)).sum()))
print('mismatching candidates:', bad[:20], len(bad))
for c in bad[:3]:
    print('--- cand', c, 'gpu st', st[c], 'gpu val', val[c], 'cpu t', t_c[c])
    for mode, ph in ((SX_ALIVE, True), (SX_TQP, False)):
        L.ehm_debug_trace(1)
        v1, _, s1 = gpu.simplex_batch(R[None], Vr[None], np.array([c], np.int32), mode, ph)
        tr = np.zeros(8 * 128)
        L.ehm_debug_fetch(tr.ctypes.data)
        L.ehm_debug_trace(0)
        tr = tr.reshape(128, 8)
        n = int(np.max(np.nonzero(tr[:, 6])[0])) + 1 if tr[:, 6].any() else 0
        print(' mode', mode, 'status', int(s1[0]), 'val', float(v1[0]))
        for row in tr[:min(n, 90)]:
            print('   it %2d rp %.3e mu %.3e obj %.10g rho*rd0 %.3e lmax %.3e m %d nv %d' % tuple(
                [int(row[0])] + list(row[1:6]) + [int(row[6]), int(row[7])]))
