"""INTEGRATION.md level 1: the reference's Worker recursion driven through the six PUBLIC methods of
the GPU ``Oracle`` (P_theta, V_R, bar_E_delta_R, bar_D_delta_R, P_theta_delta + the tools / tree
mirrors) -- the calls ``lib/worker.py:241-417`` makes, in its order, with its return conventions.

The reference's own files cannot travel to the GPU box (tests must not read /root/reference), so the
control flow of ``Worker.ecc`` / ``Worker.lcss`` is restated below line for line (iteratively: no
recursion limit); that the restatement equals the reference's code is pinned on the CPU side, where
``tests/golden/make_worker_golden.py`` ran the UNMODIFIED ``lib/worker.py`` to produce
``worker_tree_golden.npz``.  Here: walk over the GPU Oracle == that golden == the fused device loop."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def reference_walk(oracle, vertices):
    """Worker.ecc / Worker.lcss (lib/worker.py:241-417) with offload_child_computation short-circuited
    to "do it yourself" (what the reference does when no worker is idle, :230-239).  Left child first,
    as the recursion does; the tree does not depend on the order."""
    from explicit_hybrid_mpc_b200.tools import split_along_longest_edge
    from explicit_hybrid_mpc_b200.tree import NodeData, Tree
    top = Tree(NodeData(vertices=np.array(vertices, dtype=float)))
    todo = [(top, 'ecc')]
    while todo:
        node, alg = todo.pop()
        if alg == 'ecc':                                                     # worker.py:262-291
            c_R = np.average(node.data.vertices, axis=0)
            if not oracle.P_theta(theta=c_R, check_feasibility=True):
                raise RuntimeError('STOP, Theta contains infeasible regions')
            delta_hat, vx = oracle.V_R(node.data.vertices)
            if delta_hat is None:
                S_1, S_2 = split_along_longest_edge(node.data.vertices)[:2]
                node.grow(NodeData(vertices=S_1), NodeData(vertices=S_2))
                todo += [(node.right, 'ecc'), (node.left, 'ecc')]
            else:
                Nvx = node.data.vertices.shape[0]
                node.data = NodeData(vertices=node.data.vertices, commutation=delta_hat,
                                     vertex_costs=np.array([vx[i][1] for i in range(Nvx)]),
                                     vertex_inputs=np.array([vx[i][0] for i in range(Nvx)]))
                todo.append((node, 'lcss'))
            continue
        d = node.data                                                        # worker.py:366-417
        if oracle.bar_E_delta_R(R=d.vertices, V_delta_R=d.vertex_costs):
            d.is_epsilon_suboptimal = True
            continue
        delta_star, theta_star, new_vx, varies_little = oracle.bar_D_delta_R(
            R=d.vertices, V_delta_R=d.vertex_costs, delta_ref=d.commutation)
        feasible = delta_star is not None
        if not feasible:
            delta_star, new_costs, new_inputs = d.commutation, d.vertex_costs, d.vertex_inputs
        else:
            Nvx = d.vertices.shape[0]
            new_costs = np.array([new_vx[i][1] for i in range(Nvx)])
            new_inputs = np.array([new_vx[i][0] for i in range(Nvx)])
        if feasible and varies_little:
            d.commutation, d.vertex_costs, d.vertex_inputs = delta_star, new_costs, new_inputs
            todo.append((node, 'lcss'))
            continue
        S_1, S_2, v_idx = split_along_longest_edge(d.vertices)
        v_mid = S_1[v_idx[0]]
        u_mid, V_mid = oracle.P_theta_delta(theta=v_mid, delta=delta_star)[:2]   # update_vertex_costs :332-365
        in1, in2, c1, c2 = new_inputs.copy(), new_inputs.copy(), new_costs.copy(), new_costs.copy()
        in1[v_idx[0]], in2[v_idx[1]] = u_mid, u_mid
        c1[v_idx[0]], c2[v_idx[1]] = V_mid, V_mid
        node.grow(NodeData(vertices=S_1, commutation=delta_star, vertex_costs=c1, vertex_inputs=in1),
                  NodeData(vertices=S_2, commutation=delta_star, vertex_costs=c2, vertex_inputs=in2))
        todo += [(node.right, 'lcss'), (node.left, 'lcss')]
    return top


def _flatten(tops, problem):
    """reference-style Tree objects -> flat node arrays keyed like the device store"""
    out = dict(left=[], right=[], root_id=[], depth=[], cand=[], closed=[], vertices=[], vcost=[], vinput=[])
    for r, top in enumerate(tops):
        stack = [(top, 0, None, 0)]
        while stack:
            t, depth, parent, side = stack.pop()
            i = len(out['left'])
            d = t.data
            has = hasattr(d, 'commutation') and d.commutation is not None
            nv = d.vertices.shape[0]
            out['left'].append(-1); out['right'].append(-1); out['root_id'].append(r); out['depth'].append(depth)
            out['cand'].append(problem.index_from_commutation(np.asarray(d.commutation)) if has else -1)
            out['closed'].append(1 if getattr(d, 'is_epsilon_suboptimal', False) else 0)
            out['vertices'].append(np.asarray(d.vertices, dtype=float))
            out['vcost'].append(np.asarray(d.vertex_costs, dtype=float) if has else np.zeros(nv))
            out['vinput'].append(np.asarray(d.vertex_inputs, dtype=float).reshape(nv, -1) if has
                                 else np.zeros((nv, problem.n_u)))
            if parent is not None:
                out['right' if side else 'left'][parent] = i
            if not t.is_leaf():
                stack.append((t.right, depth + 1, i, 1))
                stack.append((t.left, depth + 1, i, 0))
    return {k: np.array(v) for k, v in out.items()}


def test_reference_walk_over_gpu_oracle(pendulum, pendulum_roots, gpu_oracle):
    """roots 22 + 44 (the golden's): ~1000 nodes, ~2100 calls of the six-method API, each one or more
    batched kernel launches over all 3125 commutations."""
    from conftest import GOLDEN
    from explicit_hybrid_mpc_b200.partition import partition
    from test_gpu_tree_parity import _gpu_arrays, compare_trees
    from test_oracle_golden import check_against_worker_golden
    z = np.load(os.path.join(GOLDEN, 'worker_tree_golden.npz'))
    roots = pendulum_roots[z['root_ids']]
    walk = _flatten([reference_walk(gpu_oracle, R) for R in roots], pendulum.problem)
    check_against_worker_golden(walk, z, len(roots), cost_rtol=1e-6, input_atol=2e-3)
    fused = partition(gpu_oracle, roots, wave=128)
    compare_trees(_gpu_arrays(fused), walk, len(roots))
