"""
ctypes binding of ``libehm_b200.so`` (C-ABI in ``include/ehm.h``).

There is no CPU path: importing this module without the built library, or creating a
problem without a CUDA device, raises.  PyTorch is used only as the owner of device
buffers / streams (``tensor.data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EHM_LIB_PATH') or os.path.join(_HERE, 'libehm_b200.so')   # override: kernel experiments

EHM_OK, EHM_EINVAL, EHM_ECUDA, EHM_ENOMEM, EHM_EINFEASIBLE, EHM_ECAPACITY, EHM_ENUMERIC = (
    0, -1, -2, -3, -4, -5, -6)
SX_ALIVE, SX_TQP, SX_MINV = 0, 1, 2

EXPORTS = [
    'ehm_problem_create', 'ehm_problem_destroy', 'ehm_set_eps', 'ehm_num_candidates',
    'ehm_last_error', 'ehm_kernel_launches', 'ehm_qp_theta_delta_batch', 'ehm_simplex_batch',
    'ehm_split_batch', 'ehm_partition_run', 'ehm_partition_fetch', 'ehm_partition_free',
    'ehm_read_counters', 'ehm_fp64_peak', 'ehm_hbm_peak', 'ehm_partition_create',
    'ehm_partition_advance', 'ehm_partition_advance_for', 'ehm_locator_create', 'ehm_locator_eval',
    'ehm_locator_destroy', 'ehm_partition_export', 'ehm_partition_import',
    'ehm_partition_fetch_origin', 'ehm_partition_set_root_ids', 'ehm_last_failure',
    'ehm_debug_trace', 'ehm_debug_fetch', 'ehm_p_theta_batch', 'ehm_v_r_batch', 'ehm_bar_e_batch',
    'ehm_bar_d_batch',
]


class EhmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libehm_b200 error %d: %s' % (code, msg))
        self.code = code


class ProblemDesc(C.Structure):
    _fields_ = ([(k, C.c_int) for k in ('n', 'nu', 'N', 'K', 'p')] +
                [(k, C.POINTER(C.c_double)) for k in
                 ('A', 'B', 'w', 'C', 'D', 'e', 'Q', 'P', 'R', 'ru', 'theta_scale')] +
                [('eps_a', C.c_double), ('eps_r', C.c_double)])


class PartitionOpts(C.Structure):
    _fields_ = [('wave', C.c_int), ('max_nodes', C.c_int64), ('max_frontier', C.c_int64),
                ('node_budget', C.c_int64), ('verbose', C.c_int)]


class PartitionStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ('nnodes', 'nleaves', 'nsolves', 'niters', 'nfail',
                                         'nwaves', 'nlaunches', 'open_left', 'frontier')] + \
               [('max_depth', C.c_int), ('device_ms', C.c_double),
                ('stage_ms', C.c_double * 8), ('stage_flops', C.c_double * 8),
                ('stage_bytes', C.c_double * 8)]


_lib = None


def lib():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and 'EHM_LIB_PATH' not in os.environ:
            # a fresh checkout (the .so is not in git): build it in place once, exactly as
            # __graft_entry__.build() does.  Still no CPU fallback: if nvcc is absent this raises.
            import subprocess
            try:
                subprocess.check_call(['make', '-s', '-C', os.path.join(os.path.dirname(__file__), 'csrc')])
            except Exception as ex:
                raise ImportError('%s is missing and could not be built (%r): run `python -c "import '
                                  '__graft_entry__ as g; g.build()"` (nvcc, sm_100a). There is no CPU '
                                  'fallback.' % (LIB_PATH, ex))
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s is missing: run `python -c "import __graft_entry__ as g; '
                              'g.build()"` (nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
        L.ehm_problem_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
        L.ehm_problem_destroy.argtypes = [vp]
        L.ehm_set_eps.argtypes = [vp, C.c_double, C.c_double]
        L.ehm_num_candidates.argtypes = [vp]
        L.ehm_last_error.restype = C.c_char_p
        L.ehm_kernel_launches.restype = C.c_int64
        L.ehm_qp_theta_delta_batch.argtypes = [vp, C.c_int, dp, ip, dp, dp, ip, vp]
        L.ehm_simplex_batch.argtypes = [vp, C.c_int, dp, dp, ip, C.c_int, C.c_int, dp, dp, ip, vp]
        L.ehm_split_batch.argtypes = [C.c_int, C.c_int, dp, dp, dp, ip, vp]
        L.ehm_read_counters.argtypes = [vp] + [C.POINTER(C.c_int64)] * 3
        L.ehm_fp64_peak.argtypes = [C.POINTER(C.c_double)]
        L.ehm_hbm_peak.argtypes = [C.c_int64, C.POINTER(C.c_double)]
        L.ehm_partition_run.argtypes = [vp, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(PartitionOpts), C.POINTER(vp),
                                        C.POINTER(PartitionStats)]
        L.ehm_partition_fetch.argtypes = [vp] + [vp] * 9
        L.ehm_partition_free.argtypes = [vp]
        L.ehm_partition_create.argtypes = [vp, C.c_int, C.POINTER(C.c_double),
                                           C.POINTER(PartitionOpts), C.POINTER(vp)]
        L.ehm_partition_advance.argtypes = [vp, C.c_int64, C.POINTER(PartitionStats)]
        L.ehm_partition_advance_for.argtypes = [vp, C.c_int64, C.c_double, C.POINTER(PartitionStats)]
        L.ehm_locator_create.argtypes = [C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, C.c_int32,
                                         C.POINTER(C.c_void_p)]
        L.ehm_locator_eval.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
        L.ehm_locator_destroy.argtypes = [vp]
        L.ehm_partition_export.argtypes = [vp, C.c_int, vp, vp, vp]
        L.ehm_partition_import.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.ehm_partition_fetch_origin.argtypes = [vp, vp]
        L.ehm_partition_set_root_ids.argtypes = [vp, C.c_int, vp]
        L.ehm_last_failure.argtypes = [vp, vp]
        L.ehm_p_theta_batch.argtypes = [vp, C.c_int, dp, ip, dp, dp, ip, vp]
        L.ehm_v_r_batch.argtypes = [vp, C.c_int, dp, ip, dp, dp, ip, vp]
        L.ehm_bar_e_batch.argtypes = [vp, C.c_int, dp, dp, ip, dp, vp]
        L.ehm_bar_d_batch.argtypes = [vp, C.c_int, dp, dp, ip, ip, dp, dp, dp, ip, ip, vp]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise EhmError(rc, lib().ehm_last_error().decode())


def _hp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Problem:
    """Owns an ``ehm_problem`` (device-resident candidate tables)."""

    def __init__(self, problem, eps_a, eps_r, theta_scale):
        pb = problem
        self.problem = pb
        keep = [np.ascontiguousarray(x, dtype=np.float64) for x in
                (pb.A, pb.B, pb.w, pb.C, pb.D, pb.e, pb.Q, pb.P, pb.R, pb.ru, theta_scale)]
        desc = ProblemDesc(pb.n_x, pb.n_u, pb.N, pb.K, pb.p, *[_hp(a) for a in keep],
                           float(eps_a), float(eps_r))
        h = C.c_void_p()
        check(lib().ehm_problem_create(C.byref(desc), C.byref(h)))
        self.handle = h
        self.ncand = lib().ehm_num_candidates(h)

    def set_eps(self, eps_a, eps_r):
        check(lib().ehm_set_eps(self.handle, float(eps_a), float(eps_r)))

    def counters(self):
        it, so, fa = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().ehm_read_counters(self.handle, C.byref(it), C.byref(so), C.byref(fa)))
        return dict(iters=it.value, solves=so.value, fails=fa.value)

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            lib().ehm_problem_destroy(h)
            self.handle = None
