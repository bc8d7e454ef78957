"""Mirror of the reference's ``snacc/distmatrix_to_tree.py`` with the clustering on the GPU.

Same names and argument meaning: ``hierarchical(D_sym)`` (reference :9-15, scipy average linkage -> here
``Engine.upgma``, the ``upgma_kernel`` of libsnacc_b200.so), ``get_newick`` / ``write_newick`` (:23-45, the same
"%s:%.2f" recursion, restated without scipy's ``to_tree``), ``main(csv_file, plt_file, nwk_file)`` (:48-55).  The
dendrogram PNG (``plot_hierarchical``, :18-21) needs matplotlib, which is optional here: ``plt_file=None`` skips it.
"""
import threading

import numpy as np

from .engine import Engine
from .misc import metrify, read_dist_values_names

_lock = threading.Lock()
_engine = None


def _shared_engine():
    global _engine
    with _lock:
        if _engine is None:
            _engine = Engine(0)
        return _engine


def hierarchical(D_sym, engine=None):
    """linkage matrix of UPGMA on the symmetric matrix D_sym (reference :9-15), computed on the device"""
    return (engine or _shared_engine()).upgma(np.asarray(D_sym, dtype=np.float64), metrify=False)


def linkage_from_distances(D, engine=None):
    """metrify + UPGMA in one device call (what ``main`` does in two steps, reference :50-53)"""
    return (engine or _shared_engine()).upgma(np.asarray(D, dtype=np.float64), metrify=True)


def newick_from_linkage(Z, leaf_names):
    """the string reference get_newick(to_tree(Z), "", root.dist, leaf_names) builds (:23-45): children in linkage order
    (left = first id), branch lengths "%.2f" of parent height minus own height"""
    Z = np.asarray(Z, dtype=np.float64)
    n = Z.shape[0] + 1
    left = {n + i: int(Z[i, 0]) for i in range(n - 1)}
    right = {n + i: int(Z[i, 1]) for i in range(n - 1)}
    height = {n + i: float(Z[i, 2]) for i in range(n - 1)}
    root = 2 * n - 2

    def rec(node, newick, parentdist):
        if node < n:
            return "%s:%.2f%s" % (leaf_names[node], parentdist - 0.0, newick)
        if len(newick) > 0:
            newick = "):%.2f%s" % (parentdist - height[node], newick)
        else:
            newick = ");"
        newick = rec(left[node], newick, height[node])
        newick = rec(right[node], ",%s" % (newick), height[node])
        return "(%s" % (newick)

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 4 * n + 100))
    try:
        return rec(root, "", height[root])
    finally:
        sys.setrecursionlimit(old)


def write_newick(linkage, leaf_names, nwk_file):
    with open(nwk_file, "w+") as f:
        f.write(newick_from_linkage(linkage, leaf_names))


def plot_hierarchical(linkage, plt_file, n, labels=None):
    import matplotlib.pyplot as plt          # optional dependency, as in the reference (:3)
    import scipy.cluster
    plt.switch_backend("agg")
    plt.figure(figsize=(min(3 + 0.1 * n, 13), 4))
    scipy.cluster.hierarchy.dendrogram(linkage, labels=labels)
    plt.savefig(plt_file, bbox_inches="tight", dpi=300)


def main(csv_file, plt_file, nwk_file):
    leaf_names, D = read_dist_values_names(csv_file)
    D_sym = metrify(D)
    n, _ = D_sym.shape
    linkage = hierarchical(D_sym)
    if plt_file is not None:
        plot_hierarchical(linkage, plt_file, n, labels=leaf_names)
    write_newick(linkage, leaf_names, nwk_file)
