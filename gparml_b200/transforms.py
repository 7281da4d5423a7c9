"""Host-side positive-parameter transforms of the optimiser's flat vector
(``supporting_functions.py:125-168``): softplus for parameters bounded by (0, None), identity
otherwise.  These act on the M*Q + Q + 2 global parameters only; the per-point variance
transform runs on the device inside prep_points / embed_finish."""
import sys

import numpy as np

lim_val = -np.log(sys.float_info.epsilon)


def transform(b, x):
    if b == (0, None):
        assert -lim_val < x < lim_val
        return np.log(1 + np.exp(x))
    return x


def transform_back(b, x):
    if b == (0, None):
        assert sys.float_info.epsilon < x < lim_val
        return np.log(-1 + np.exp(x))
    return x


def transform_grad(b, x):
    if b == (0, None):
        assert -lim_val < x < lim_val
        return 1 / (np.exp(-x) + 1)
    return 1


def transformVar(x):
    x = np.asarray(x)
    assert np.all(-lim_val < x) and np.all(x < lim_val)
    return np.log(1 + np.exp(x))


def transformVar_back(x):
    x = np.asarray(x)
    assert np.all(sys.float_info.epsilon < x) and np.all(x < lim_val)
    return np.log(-1 + np.exp(x))


def transformVar_grad(x):
    x = np.asarray(x)
    assert np.all(-lim_val < x) and np.all(x < lim_val)
    return 1 / (np.exp(-x) + 1)


def PCA(Y, input_dim):
    """One-off initialisation (supporting_functions.py:102-121): principal components by SVD of
    the centred data, each scaled to unit standard deviation."""
    U, s, Vt = np.linalg.svd(Y - Y.mean(axis=0), full_matrices=False)
    X = U[:, :input_dim].copy()
    X /= X.std(axis=0)
    return X
