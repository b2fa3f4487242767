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

    # ------------------------------------------------------------------ batched reference calls
    def _i32(self, B):
        return torch.empty(B, dtype=torch.int32, device=self.device)

    def P_theta_batch(self, theta):
        """``ehm_p_theta_batch``: theta (B,n) -> CUDA tensors (cand int32 (B,), J (B,), u0hat (B,nu))."""
        th = (theta if torch.is_tensor(theta) else self._dev(theta)).contiguous().view(-1, self.n)
        B = th.shape[0]
        cand, st = self._i32(B), self._i32(B)
        J = torch.empty(B, dtype=torch.float64, device=self.device)
        u0 = torch.zeros(B, self.nu, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().ehm_p_theta_batch(self._p.handle, B, th.data_ptr(), cand.data_ptr(),
                                                  J.data_ptr(), u0.data_ptr(), st.data_ptr(), _stream()))
        return cand, J, u0

    def V_R_batch(self, R):
        """``ehm_v_r_batch``: R (B,n+1,n) -> (cand (B,), Jv (B,n+1), Uvhat (B,n+1,nu))."""
        Rt = (R if torch.is_tensor(R) else self._dev(R)).contiguous().view(-1, self.n + 1, self.n)
        B = Rt.shape[0]
        cand, st = self._i32(B), self._i32(B)
        Jv = torch.empty(B, self.n + 1, dtype=torch.float64, device=self.device)
        Uv = torch.zeros(B, self.n + 1, self.nu, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().ehm_v_r_batch(self._p.handle, B, Rt.data_ptr(), cand.data_ptr(),
                                              Jv.data_ptr(), Uv.data_ptr(), st.data_ptr(), _stream()))
        return cand, Jv, Uv

    def bar_E_batch(self, R, Vr):
        """``ehm_bar_e_batch``: (R (B,n+1,n), Vr (B,n+1)) -> (converged int32 (B,), tmin (B,))."""
        Rt = (R if torch.is_tensor(R) else self._dev(R)).contiguous().view(-1, self.n + 1, self.n)
        B = Rt.shape[0]
        Vt = (Vr if torch.is_tensor(Vr) else self._dev(Vr)).contiguous().view(B, self.n + 1)
        conv = self._i32(B)
        tmin = torch.empty(B, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().ehm_bar_e_batch(self._p.handle, B, Rt.data_ptr(), Vt.data_ptr(),
                                                conv.data_ptr(), tmin.data_ptr(), _stream()))
        return conv, tmin

    def bar_D_batch(self, R, Vr, cand_ref):
        """``ehm_bar_d_batch`` -> (cand_star (B,), theta_star (B,n), Jv (B,n+1), Uvhat (B,n+1,nu),
        var_small (B,), status (B,))."""
        Rt = (R if torch.is_tensor(R) else self._dev(R)).contiguous().view(-1, self.n + 1, self.n)
        B = Rt.shape[0]
        Vt = (Vr if torch.is_tensor(Vr) else self._dev(Vr)).contiguous().view(B, self.n + 1)
        cr = (cand_ref if torch.is_tensor(cand_ref) else self._dev(cand_ref, torch.int32)).contiguous().view(B)
        cs, vs, st = self._i32(B), self._i32(B), self._i32(B)
        ths = torch.zeros(B, self.n, dtype=torch.float64, device=self.device)
        Jv = torch.zeros(B, self.n + 1, dtype=torch.float64, device=self.device)
        Uv = torch.zeros(B, self.n + 1, self.nu, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib().ehm_bar_d_batch(self._p.handle, B, Rt.data_ptr(), Vt.data_ptr(), cr.data_ptr(),
                                                cs.data_ptr(), ths.data_ptr(), Jv.data_ptr(), Uv.data_ptr(),
                                                vs.data_ptr(), st.data_ptr(), _stream()))
        return cs, ths, Jv, Uv, vs, st

    # ------------------------------------------------------------------ reference API
    def P_theta(self, theta, check_feasibility=False):
        """``lib/oracle.py:104-139``: argmin over commutations (ties -> smallest index)."""
        cand, J, u0 = self.P_theta_batch(np.asarray(theta, dtype=float).reshape(1, -1))
        c = int(cand[0].item())
        if check_feasibility:
            return c >= 0
        if c < 0:
            return None, None, np.inf, 0.0
        return (self.pb.u_scale * u0[0].cpu().numpy(), self.pb.commutation_from_index(c),
                float(J[0].item()), 0.0)

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
        cand, Jv, Uv = self.V_R_batch(np.asarray(R, dtype=float)[None])
        c = int(cand[0].item())
        if c < 0:
            return None, None
        Jc, uc = Jv[0].cpu().numpy(), Uv[0].cpu().numpy()
        return (self.pb.commutation_from_index(c),
                [(self.pb.u_scale * uc[i], float(Jc[i]), 0.0) for i in range(self.n + 1)])

    def bar_E_delta_R(self, R, V_delta_R):
        """``lib/oracle.py:285-309``: True iff no commutation satisfies the two
        eps-improvement inequalities anywhere in R."""
        conv, _ = self.bar_E_batch(np.asarray(R, dtype=float)[None], np.asarray(V_delta_R, dtype=float)[None])
        return bool(conv[0].item())

    def bar_D_delta_R(self, R, V_delta_R, delta_ref):
        """``lib/oracle.py:311-414`` (+ ``in_variability_ball`` :220-283)."""
        cref = self._cand_index(delta_ref)
        cs, ths, Jv, Uv, vs, st = self.bar_D_batch(np.asarray(R, dtype=float)[None],
                                                   np.asarray(V_delta_R, dtype=float)[None],
                                                   np.array([cref], dtype=np.int32))
        c = int(cs[0].item())
        if c < 0:
            return None, None, None, None
        Jc, uc = Jv[0].cpu().numpy(), Uv[0].cpu().numpy()
        return (self.pb.commutation_from_index(c), ths[0].cpu().numpy(),
                [(self.pb.u_scale * uc[i], float(Jc[i]), 0.0) for i in range(self.n + 1)],
                bool(vs[0].item()))

    def in_variability_ball(self, R, V_delta_R, delta_ref, delta_star, theta_star):
        """``lib/oracle.py:220-283``."""
        cref = self._cand_index(delta_ref)
        Rn = np.asarray(R, dtype=float)
        mv, _, sm = self.simplex_batch(Rn[None], None, np.array([cref], np.int32), SX_MINV, False)
        if int(sm[0].item()) != 1:
            raise RuntimeError('min_over_simplex_for_this_delta failed')
        V = self.P_theta_delta(theta_star, delta_star)[1]
        return (np.max(V_delta_R) - float(mv[0].item())) < max(self.eps_a, self.eps_r * V)
