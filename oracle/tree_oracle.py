"""CPU restatement of the reference's tree step (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows /root/reference/snacc/distmatrix_to_tree.py:9-45 and misc.py:20-25 with scipy itself (the reference module
cannot be imported here: it needs matplotlib at import time, :3): metrify, squareform, linkage(method='average'),
to_tree + the get_newick recursion.
"""
import numpy as np
import scipy.cluster.hierarchy
import scipy.spatial.distance


def metrify(D):
    """misc.py:20-25"""
    D_sym = 0.5 * (D + D.T)
    np.fill_diagonal(D_sym, 0.)
    return D_sym


def hierarchical(D_sym):
    """distmatrix_to_tree.py:9-15"""
    return scipy.cluster.hierarchy.linkage(scipy.spatial.distance.squareform(D_sym), method="average")


def get_newick(node, newick, parentdist, leaf_names):
    """distmatrix_to_tree.py:23-40"""
    if node.is_leaf():
        return "%s:%.2f%s" % (leaf_names[node.id], parentdist - node.dist, newick)
    if len(newick) > 0:
        newick = "):%.2f%s" % (parentdist - node.dist, newick)
    else:
        newick = ");"
    newick = get_newick(node.get_left(), newick, node.dist, leaf_names)
    newick = get_newick(node.get_right(), ",%s" % (newick), node.dist, leaf_names)
    return "(%s" % (newick)


def newick(linkage, leaf_names):
    """distmatrix_to_tree.py:43-45"""
    tree = scipy.cluster.hierarchy.to_tree(linkage, False)
    return get_newick(tree, "", tree.dist, leaf_names)
