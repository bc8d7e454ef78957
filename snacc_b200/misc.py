"""CSV helpers mirroring the reference's ``snacc/misc.py:5-25`` (readers used by the skew transform)."""
import numpy as np
import pandas as pd


def read_dist(csv_file):
    return pd.read_csv(csv_file, index_col=0, sep=None, engine="python").values


def read_dist_dataframe(csv_file):
    return pd.read_csv(csv_file, index_col=0, sep=None, engine="python")


def read_dist_values_names(csv_file):
    df = pd.read_csv(csv_file, index_col=0, sep=None, engine="python")
    return df.index.values, df.values


def metrify(D):
    """symmetrise and zero the diagonal (reference misc.py:20-25)"""
    D_sym = 0.5 * (D + D.T)
    np.fill_diagonal(D_sym, 0.)
    return D_sym
