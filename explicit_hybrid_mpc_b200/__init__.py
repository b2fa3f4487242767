"""B200-native explicit hybrid MPC partitioning engine (drop-in for the oracle / worker /
scheduler / tree hot path of dmalyuta/explicit_hybrid_mpc).  See DESIGN.md."""
from . import mpc_library, tools, tree  # noqa: F401

__all__ = ['mpc_library', 'tools', 'tree', 'oracle', 'partition']
