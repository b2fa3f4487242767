"""
Drop-in mirror of the reference's ``Oracle`` (``lib/oracle.py:18-474``) backed by the
batched sm_100a interior-point kernels of ``libehm_b200.so``.

Same constructor and the same six public methods with the reference's return conventions
(``None`` for "no solution", lists of ``(u0, J, t_solve)`` for vertex results, 0/1 float
commutation vectors).  Every method is one or two batched kernel launches over all
``K**N`` commutation candidates; MOSEK's branch-and-bound is replaced by explicit
enumeration with the canonical selection rules of DESIGN.md.  These stand-alone calls exist
for API compatibility (``lib/worker.py`` could run unmodified on top of them); the fast path
is :func:`explicit_hybrid_mpc_b200.partition.partition`, which keeps the whole ECC/L-CSS
loop on the device.
"""
import numpy as np
import torch

from . import _capi
from ._capi import SX_ALIVE, SX_MINV, SX_TQP


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Oracle:
    def __init__(self, mpc, eps_a, eps_r, device=None):
        """mpc: object with ``.problem`` (HybridMPCProblem), ``.n_x``, ``.n_u``, ``.N``,
        ``.delta_size`` -- e.g. :class:`mpc_library.InvertedPendulumOnCart`."""
        if not torch.cuda.is_available():
            raise RuntimeError('explicit_hybrid_mpc_b200.Oracle needs a CUDA device (sm_100a); '
                               'there is no CPU fallback')
        self.mpc = mpc
        self.eps_a, self.eps_r = float(eps_a), float(eps_r)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        torch.cuda.set_device(self.device)
        pb = mpc.problem
        self.pb = pb
        ts = getattr(mpc, 'theta_scale', None)
        if ts is None:
            ts = np.ones(pb.n_x)
        self._p = _capi.Problem(pb, eps_a, eps_r, ts)
        self.n, self.nu, self.ncand = pb.n_x, pb.n_u, self._p.ncand
        self._all = torch.arange(self.ncand, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ helpers
    def set_eps(self, eps_a, eps_r):
        self.eps_a, self.eps_r = float(eps_a), float(eps_r)
        self._p.set_eps(eps_a, eps_r)

    def _dev(self, a, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def _cand_index(self, delta):
        if np.isscalar(delta) or isinstance(delta, (int, np.integer)):
            return int(delta)
        return self.pb.index_from_commutation(delta)

    def point_batch(self, theta, cand):
        """P_theta_delta for B (theta, candidate) pairs. theta (B,n) and cand (B,) may be
        numpy or CUDA tensors.  Returns CUDA tensors (J, u0hat, status)."""
        th = theta if torch.is_tensor(theta) else self._dev(theta)
        cd = cand if torch.is_tensor(cand) else self._dev(cand, torch.int32)
        th = th.contiguous().view(-1, self.n)
        B = th.shape[0]
        J = torch.empty(B, dtype=torch.float64, device=self.device)
        u0 = torch.zeros(B, self.nu, dtype=torch.float64, device=self.device)
        st = torch.empty(B, dtype=torch.int32, device=self.device)
        _capi.check(_capi.lib().ehm_qp_theta_delta_batch(
            self._p.handle, B, th.data_ptr(), cd.data_ptr(), J.data_ptr(), u0.data_ptr(),
            st.data_ptr(), _stream()))
        return J, u0, st

    def simplex_batch(self, R, Vr, cand, mode, with_phase1=True):
        """Simplex sub-problems for B (R, candidate) pairs; see ``ehm_simplex_batch``."""
        Rt = R if torch.is_tensor(R) else self._dev(R)
        cd = cand if torch.is_tensor(cand) else self._dev(cand, torch.int32)
        Rt = Rt.contiguous().view(-1, self.n + 1, self.n)
        B = Rt.shape[0]
        Vt = None
        if Vr is not None:
            Vt = (Vr if torch.is_tensor(Vr) else self._dev(Vr)).contiguous().view(B, self.n + 1)
        val = torch.empty(B, dtype=torch.float64, device=self.device)
        sol = torch.zeros(B, self.n + 1, dtype=torch.float64, device=self.device)
        st = torch.empty(B, dtype=torch.int32, device=self.device)
        _capi.check(_capi.lib().ehm_simplex_batch(
            self._p.handle, B, Rt.data_ptr(), Vt.data_ptr() if Vt is not None else None,
            cd.data_ptr(), int(mode), int(bool(with_phase1)), val.data_ptr(), sol.data_ptr(),
            st.data_ptr(), _stream()))
        return val, sol, st

    def _vertex_problems(self, R, cands):
        """P_theta_delta at every vertex of R for every candidate in cands (tensor)."""
        nc = cands.numel()
        Rt = self._dev(R)
        th = Rt.unsqueeze(0).expand(nc, self.n + 1, self.n).reshape(-1, self.n)
        cd = cands.view(-1, 1).expand(nc, self.n + 1).reshape(-1).contiguous()
        J, u0, st = self.point_batch(th, cd)
        return (J.view(nc, self.n + 1), u0.view(nc, self.n + 1, self.nu),
                st.view(nc, self.n + 1))

    def _tstar_all(self, R, Vr):
        """t*_c for every alive candidate on R (NaN elsewhere) + [beta*, V*]."""
        Rt = self._dev(R).view(1, self.n + 1, self.n).expand(self.ncand, -1, -1).contiguous()
        Vt = self._dev(Vr).view(1, self.n + 1).expand(self.ncand, -1).contiguous()
        val, sol, st = self.simplex_batch(Rt, Vt, self._all, SX_TQP, True)
        t = torch.where(st == 1, val, torch.full_like(val, float('nan')))
        return t, sol

    # ------------------------------------------------------------------ reference API
    def P_theta(self, theta, check_feasibility=False):
        """``lib/oracle.py:104-139``: argmin over commutations (ties -> smallest index)."""
        th = self._dev(theta).view(1, self.n).expand(self.ncand, -1).contiguous()
        J, u0, st = self.point_batch(th, self._all)
        ok = st == 1
        if check_feasibility:
            return bool(ok.any().item())
        if not bool(ok.any().item()):
            return None, None, np.inf, 0.0
        Jm = torch.where(ok, J, torch.full_like(J, float('inf')))
        c = int(torch.argmin(Jm).item())       # first minimum = smallest index
        return (self.pb.u_scale * u0[c].cpu().numpy(), self.pb.commutation_from_index(c),
                float(J[c].item()), 0.0)

    def P_theta_delta(self, theta, delta, check_feasibility=False):
        """``lib/oracle.py:141-173``."""
        c = self._cand_index(delta)
        J, u0, st = self.point_batch(np.asarray(theta, dtype=float).reshape(1, -1),
                                     np.array([c], dtype=np.int32))
        ok = int(st[0].item()) == 1
        if check_feasibility:
            return ok
        if not ok:
            return None, np.inf, 0.0
        return self.pb.u_scale * u0[0].cpu().numpy(), float(J[0].item()), 0.0

    def V_R(self, R):
        """``lib/oracle.py:175-218``.  Canonical rule: among candidates feasible (and
        convergent) at all n+1 vertices the one minimising the sum of vertex costs."""
        J, u0, st = self._vertex_problems(R, self._all)
        ok = (st == 1).all(dim=1)
        if not bool(ok.any().item()):
            return None, None
        tot = torch.where(ok, J.sum(dim=1), torch.full((self.ncand,), float('inf'),
                                                       dtype=torch.float64, device=self.device))
        c = int(torch.argmin(tot).item())
        Jc, uc = J[c].cpu().numpy(), u0[c].cpu().numpy()
        return (self.pb.commutation_from_index(c),
                [(self.pb.u_scale * uc[i], float(Jc[i]), 0.0) for i in range(self.n + 1)])

    def bar_E_delta_R(self, R, V_delta_R):
        """``lib/oracle.py:285-309``: True iff no commutation satisfies the two
        eps-improvement inequalities anywhere in R."""
        t, _ = self._tstar_all(R, V_delta_R)
        return not bool((t <= 0.0).any().item())

    def bar_D_delta_R(self, R, V_delta_R, delta_ref):
        """``lib/oracle.py:311-414`` (+ ``in_variability_ball`` :220-283)."""
        cref = self._cand_index(delta_ref)
        Rn = np.asarray(R, dtype=float)
        Vr = np.asarray(V_delta_R, dtype=float)
        t, sol = self._tstar_all(Rn, Vr)
        # delta_ref is feasible at every vertex by construction: no phase-1 for it
        tr, solr, str_ = self.simplex_batch(Rn[None], Vr[None], np.array([cref], np.int32),
                                            SX_TQP, False)
        tref = float(tr[0].item()) if int(str_[0].item()) == 1 else float('inf')
        tt = t.clone()
        tt[cref] = float('nan')
        better = (tt <= 0.0) & (tt < tref)
        idx = torch.nonzero(better).flatten()
        none = (None, None, None, None)
        if idx.numel() == 0:
            return none
        order = idx[torch.argsort(tt[idx], stable=True)]
        J, u0, st = self._vertex_problems(Rn, order.to(torch.int32))
        ok = (st == 1).all(dim=1)
        if not bool(ok.any().item()):
            return none
        k = int(torch.nonzero(ok).flatten()[0].item())
        c = int(order[k].item())
        mv, _, sm = self.simplex_batch(Rn[None], None, np.array([cref], np.int32), SX_MINV, False)
        if int(sm[0].item()) != 1:
            return none
        s = sol[c].cpu().numpy()
        theta_star = Rn[0] + (Rn[1:] - Rn[0]).T.dot(s[:self.n])
        var_small = (Vr.max() - float(mv[0].item())) < max(self.eps_a, self.eps_r * s[self.n])
        Jc, uc = J[k].cpu().numpy(), u0[k].cpu().numpy()
        return (self.pb.commutation_from_index(c), theta_star,
                [(self.pb.u_scale * uc[i], float(Jc[i]), 0.0) for i in range(self.n + 1)],
                bool(var_small))

    def in_variability_ball(self, R, V_delta_R, delta_ref, delta_star, theta_star):
        """``lib/oracle.py:220-283``."""
        cref = self._cand_index(delta_ref)
        Rn = np.asarray(R, dtype=float)
        mv, _, sm = self.simplex_batch(Rn[None], None, np.array([cref], np.int32), SX_MINV, False)
        if int(sm[0].item()) != 1:
            raise RuntimeError('min_over_simplex_for_this_delta failed')
        V = self.P_theta_delta(theta_star, delta_star)[1]
        return (np.max(V_delta_R) - float(mv[0].item())) < max(self.eps_a, self.eps_r * V)
