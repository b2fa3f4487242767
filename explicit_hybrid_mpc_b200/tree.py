"""
Binary partition tree and node payload with the attribute names, constructor signatures and
leaf convention of the reference's ``lib/tree.py:12-95`` -- a node is a leaf iff it has no
``left`` attribute -- so that a tree built here can be pickled for, and walked by, code written
against the reference (``ExplicitMPC``, ``lib/mpc_library.py:662-792``; ``build_tree``,
``lib/scheduler.py:644-691``).  The engine itself never touches these objects: it works on the
flat node store and converts at the boundary (``partition.PartitionResult.subtree``).
"""
import time

_OPTIONAL_FIELDS = ('commutation', 'vertex_costs', 'vertex_inputs')


class NodeData:
    """Payload of one cell.  ``commutation`` / ``vertex_costs`` / ``vertex_inputs`` exist only
    once a commutation has been assigned (the reference tests ``hasattr`` on them)."""

    def __init__(self, vertices, commutation=None, vertex_costs=None, vertex_inputs=None):
        given = dict(zip(_OPTIONAL_FIELDS, (commutation, vertex_costs, vertex_inputs)))
        self.__dict__.update({k: v for k, v in given.items() if v is not None})
        self.vertices = vertices
        self.is_epsilon_suboptimal = False
        self.timestamp = time.time()


class Tree:
    """Node of the binary tree: ``data`` (NodeData or None), ``top`` flag, and -- for inner nodes
    only -- the ``left`` / ``right`` children."""

    def __init__(self, data, top=True):
        self.data, self.top = data, top

    def is_leaf(self):
        return 'left' not in self.__dict__

    def grow(self, left, right):
        """Attach two children carrying the payloads ``left`` and ``right``."""
        self.left, self.right = (Tree(payload, top=False) for payload in (left, right))

    def copy(self, node):
        """Make ``node`` a shallow copy of this node (payload, flag and child links)."""
        for key in ('data', 'top') + (() if self.is_leaf() else ('left', 'right')):
            setattr(node, key, getattr(self, key))

    # -- conveniences the reference does not have ---------------------------------------------
    def walk(self, location=''):
        """Yield ``(location, node)`` depth first, left before right (iterative)."""
        stack = [(self, location)]
        while stack:
            node, loc = stack.pop()
            yield loc, node
            if not node.is_leaf():
                stack.append((node.right, loc + '1'))
                stack.append((node.left, loc + '0'))

    def count(self):
        """(number of nodes, number of leaves)"""
        n = leaves = 0
        for _, node in self.walk():
            n += 1
            leaves += node.is_leaf()
        return n, leaves
