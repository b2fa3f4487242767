"""
Binary tree + node data container, API- and pickle-shape-compatible with the reference's
``lib/tree.py:12-95`` (same class names, attributes and leaf convention: a node is a leaf
iff it has no ``left`` attribute), so a tree produced here can be consumed by the
reference's ``ExplicitMPC`` (``lib/mpc_library.py:662-792``) and ``build_tree`` consumers.
"""
import time


class NodeData:
    def __init__(self, vertices, commutation=None, vertex_costs=None, vertex_inputs=None):
        self.timestamp = time.time()
        self.vertices = vertices
        self.is_epsilon_suboptimal = False
        if commutation is not None:
            self.commutation = commutation
        if vertex_costs is not None:
            self.vertex_costs = vertex_costs
        if vertex_inputs is not None:
            self.vertex_inputs = vertex_inputs


class Tree:
    def __init__(self, data, top=True):
        self.data = data
        self.top = top

    def copy(self, node):
        node.data = self.data
        node.top = self.top
        if not self.is_leaf():
            node.left = self.left
            node.right = self.right

    def grow(self, left, right):
        self.left = Tree(left, top=False)
        self.right = Tree(right, top=False)

    def is_leaf(self):
        return not hasattr(self, 'left')
