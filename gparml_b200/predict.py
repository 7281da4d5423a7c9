"""Python-3 replay of the reference's test-time latent inference (``predict.py:19-144``; the
reference file is Python-2-only).  A caller of the hot path: for test
outputs ``Y_test`` it optimises the variational mean / variance of the test points with the
trained model's statistics held fixed, by adding the test points' statistics to the stored
``accumulated_statistics_*_f.npy`` and evaluating the same bound and per-point gradients.

Every number comes from the CUDA library through the ``partial_terms`` mirror; the reference's
per-call construction of a new ``partial_terms`` object (predict.py:122) is replaced by one
device context that lives for the whole optimisation.
"""
import glob
import os

import numpy

from . import transforms as sp
from .partial_terms import partial_terms as _PartialTerms
from .scg_adapted import SCG_adapted

_state = {}


def _load(name):
    return numpy.load(name)


def setup(options_, Y_test_, device=0):
    """predict.py:26-35: load the trained globals and the accumulated statistics of the final
    ('f') evaluation and build the device context."""
    options = dict(options_)
    Y_test = numpy.ascontiguousarray(numpy.atleast_2d(Y_test_), dtype=numpy.float64)
    gs = {}
    for key in ('Z', 'sf2', 'alpha', 'beta'):
        gs[key] = _load(options['statistics'] + '/global_statistics_' + key + '_f.npy')
    acc = {}
    for key in ('sum_YYT', 'sum_exp_K_mi_K_im', 'sum_exp_K_miY', 'sum_exp_K_ii', 'sum_KL'):
        acc[key] = _load(options['statistics'] + '/accumulated_statistics_' + key + '_f.npy')
    close()
    pt = _PartialTerms(gs['Z'], float(numpy.squeeze(gs['sf2'])), numpy.atleast_1d(numpy.squeeze(gs['alpha'])),
                       float(numpy.squeeze(gs['beta'])), options['M'], options['Q'], options['N'], options['D'],
                       update_global_statistics=True, device=device)
    shape = (Y_test.shape[0], options['Q'])
    bounds = [(None, None)] * int(numpy.prod(shape)) + [(0, None)] * int(numpy.prod(shape))
    _state.update(dict(options=options, Y_test=Y_test, global_statistics=gs, accumulated_statistics=acc, pt=pt,
                       shape=shape, bounds=bounds))
    return _state


def close():
    pt = _state.pop('pt', None)
    if pt is not None:
        pt.close()
    _state.clear()


def likelihood_and_gradient(flat_array, iteration=0, step_size=0):
    """predict.py:116-144."""
    s = _state
    n = len(flat_array) // 2
    t = numpy.array([sp.transform(b, x) for b, x in zip(s['bounds'], flat_array)])
    X_mu_, X_S_ = t[:n].reshape(s['shape']), t[n:].reshape(s['shape'])
    pt, acc = s['pt'], s['accumulated_statistics']
    pt.set_data(s['Y_test'], X_mu_, X_S_, is_set_statistics=True)
    new = pt.get_local_statistics()
    pt.set_local_statistics(acc['sum_YYT'] + new['sum_YYT'], acc['sum_exp_K_mi_K_im'] + new['sum_exp_K_mi_K_im'],
                            acc['sum_exp_K_miY'] + new['exp_K_miY'], acc['sum_exp_K_ii'] + new['sum_exp_K_ii'],
                            acc['sum_KL'] + new['KL'])
    likelihood = pt.logmarglik()
    gradient = numpy.concatenate((pt.grad_X_mu().flatten(), pt.grad_X_S().flatten()))
    gradient = numpy.array([g * sp.transform_grad(b, x) for b, x, g in zip(s['bounds'], flat_array, gradient)])
    return -1 * likelihood, -1 * gradient


def test(options_, Y_test_, mask=None, is_random_init=False, random_iterations=100, random_restarts=100, device=0):
    """predict.py:19-111: returns [X_mu, X_S, likelihood] for the test points."""
    s = setup(options_, Y_test_, device=device)
    options, Y_test, shape = s['options'], s['Y_test'], s['shape']
    Z = s['global_statistics']['Z']
    if is_random_init:
        X_mu = numpy.repeat(numpy.atleast_2d(Z[numpy.random.randint(options['M'])]), shape[0], axis=0)
    else:
        # nearest training output's embedding (predict.py:44-65)
        import scipy.spatial
        random_restarts = 0
        if mask is None:
            mask = list(range(Y_test.shape[1]))
        Y_dists = numpy.full(shape[0], numpy.inf)
        X_mu = numpy.zeros(shape)
        for file_name in sorted(glob.glob(options['input'] + '/*')):
            Y = numpy.genfromtxt(file_name, delimiter=',')
            if Y.ndim == 1:
                Y = numpy.atleast_2d(Y).T
            X = _load(options['embeddings'] + '/' + os.path.basename(file_name) + '.embedding.npy')
            tree = scipy.spatial.cKDTree(Y[:, mask], leafsize=100)
            dist, ind = tree.query(Y_test[:, mask], k=1, distance_upper_bound=6)
            for i in range(shape[0]):
                if dist[i] < Y_dists[i]:
                    Y_dists[i] = dist[i]
                    X_mu[i, :] = X[ind[i], :]
    X_S = numpy.clip(numpy.ones(shape) * 0.5 + 0.01 * numpy.random.randn(*shape), 0.001, 1)

    def run(X_mu0):
        x0 = numpy.concatenate((X_mu0.flatten(), X_S.flatten()))
        x0 = numpy.array([sp.transform_back(b, x) for b, x in zip(s['bounds'], x0)])
        # fixed_embeddings=True: the optimiser's "globals" are the test embeddings (predict.py:79-80)
        x = SCG_adapted(likelihood_and_gradient, x0, options['embeddings'], fixed_embeddings=True, display=False,
                        maxiters=random_iterations)
        t = numpy.array([sp.transform(b, y) for b, y in zip(s['bounds'], x[0])])
        h = len(t) // 2
        return [t[:h].reshape(shape), t[h:].reshape(shape), -x[1][-1]]

    best = run(X_mu)
    if is_random_init:
        for _ in range(random_restarts):
            cand = run(numpy.repeat(numpy.atleast_2d(Z[numpy.random.randint(options['M'])]), shape[0], axis=0))
            if cand[-1] > best[-1]:
                best = cand
    return best
