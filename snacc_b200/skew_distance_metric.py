"""Bijective transforms of a finished distance matrix; mirrors ``snacc/skew_distance_metric.py:7-19``.

Kept on the host in numpy: these are element-wise float64 libm calls on an N x N matrix,
and CUDA's log/atanh are not bit-identical to numpy's.
"""
import numpy as np

from .misc import read_dist_dataframe


def f_ln(x):
    return -np.log(1 - x)


def f_inv(x):
    return x / (1 - x)


def f_arctanh(x):
    return np.arctanh(x)


def main(csv_input, function, csv_output):
    D = read_dist_dataframe(csv_input)
    D_new = function(D)
    D_new.to_csv(csv_output, index=False)
