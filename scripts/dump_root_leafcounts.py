"""Leaves per root simplex of the bench workload -> tests/golden/pendulum_root_leafcounts.json
(used by bench.py to pick a bounded, cheap CPU-baseline sample).  Needs a GPU."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart  # noqa: E402
from explicit_hybrid_mpc_b200.oracle import Oracle  # noqa: E402
from explicit_hybrid_mpc_b200.partition import partition  # noqa: E402

root = os.path.join(os.path.dirname(__file__), '..')
roots = np.load(os.path.join(root, 'tests', 'golden', 'pendulum_roots.npy'))
orc = Oracle(InvertedPendulumOnCart(), 19.21220798973866, 1.0)
res = partition(orc, roots)
leaves = res.leaf_ids()
cnt = np.bincount(res.root_id[leaves], minlength=len(roots)).tolist()
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, 'tests', 'golden', 'pendulum_root_leafcounts.json')
json.dump(dict(eps_a=19.21220798973866, eps_r=1.0, leaves_per_root=cnt, stats=res.stats), open(out, 'w'))
print(sum(cnt), cnt)
