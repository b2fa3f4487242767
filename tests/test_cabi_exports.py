"""The C-ABI library loads and exports every symbol include/ehm.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import __graft_entry__ as ge
    ge.build()
    from explicit_hybrid_mpc_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'ehm.h')).read()
    names = set(re.findall(r'\b(ehm_[a-z_0-9]+)\s*\(', hdr))
    assert len(names) >= 12
    L = ctypes.CDLL(_capi.LIB_PATH)
    for nm in sorted(names):
        assert hasattr(L, nm), 'missing export ' + nm


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        return
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.oracle import Oracle
    import pytest
    with pytest.raises(RuntimeError):
        Oracle(InvertedPendulumOnCart(), 1.0, 1.0)


def test_header_is_plain_c():
    """include/ehm.h is the drop-in boundary: it must compile as C (no CUDA or C++ types)."""
    import subprocess
    hdr = os.path.join(ROOT, 'include', 'ehm.h')
    subprocess.check_call(['gcc', '-fsyntax-only', '-x', 'c', '-std=c99', '-Wall', '-Werror', hdr])


def test_binding_covers_every_export():
    """The ctypes binding names every symbol of the header (no hasattr guards)."""
    from explicit_hybrid_mpc_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'ehm.h')).read()
    names = set(re.findall(r'\b(ehm_[a-z_0-9]+)\s*\(', hdr))
    assert names <= set(_capi.EXPORTS), sorted(names - set(_capi.EXPORTS))
