"""Python-3 replay of the reference driver's evaluation protocol
(``parallel_GPLVM.py:78-369``; the reference file is Python-2-only).  It is a *caller* of the
hot path: option handling, the flat parameter vector with its softplus transforms, the
per-evaluation sequence cache -> statistics_MR -> global statistics/derivatives ->
embeddings_MR, timing dictionary and the final ``'f'`` evaluation are kept as in the
reference so that the ``b200_MapReduce`` backend is exercised exactly the way
``local_MapReduce`` is.  ``options['parallel']`` selects the backend ('b200' here).

Launched as ``python -m torch.distributed.run --nproc-per-node G -m gparml_b200.parallel_GPLVM ...`` (or by
calling :func:`main` from every rank) the same code runs one process per GPU: every rank maps its own input
files, the backend all-reduces the packed statistics, and all ranks step the same optimiser with identical
scalars (``b200_MapReduce`` module docstring).  Rank 0 owns the statistics folder.
"""
import os
import pickle
import time

import numpy

from . import transforms as sp
from .scg_adapted import SCG_adapted

options = {}
map_reduce = None
time_acc = {
    'time_acc_statistics_map_reduce': [], 'time_acc_statistics_mapper': [], 'time_acc_statistics_reducer': [],
    'time_acc_calculate_global_statistics': [], 'time_acc_embeddings_MR': [], 'time_acc_embeddings_MR_mapper': [],
}

DEFAULTS = dict(parallel='b200', iterations=5, M=2, Q=2, D=4, init='PCA', load=False, keep=False,
                fixed_embeddings=False, fixed_beta=False, optimiser='SCG_adapted', drop_out_fraction=0,
                local_no_pool=False, b200_write_files=True)


def default_options(**kw):
    o = dict(DEFAULTS)
    o.update(kw)
    for need in ('input', 'embeddings', 'statistics', 'tmp'):
        if need not in o:
            raise ValueError("option %r is required (parallel_GPLVM.py:479-496)" % need)
        if not os.path.isdir(o[need]):
            raise IOError("folder %s does not exist" % o[need])
    return o


def main(opt_param):
    """parallel_GPLVM.py:78-131."""
    global options, map_reduce
    options = opt_param
    if options['parallel'] == 'b200':
        from . import b200_MapReduce
        map_reduce = b200_MapReduce
    else:
        raise Exception("backend %r is not available here (b200 only)" % options['parallel'])
    options = map_reduce.init(options)
    options, global_statistics = init_statistics(map_reduce, options)
    x0 = flatten_global_statistics(options, global_statistics)
    x0 = numpy.array([sp.transform_back(b, x) for b, x in zip(options['flat_global_statistics_bounds'], x0)])
    rank = map_reduce.dist_info(options)[0]
    if options['optimiser'] != 'SCG_adapted':
        raise Exception("only the SCG_adapted optimiser is replayed here")
    x_opt = SCG_adapted(likelihood_and_gradient, x0, options['embeddings'], options['fixed_embeddings'],
                        display=options.get('display', True), maxiters=options['iterations'], xtol=0, ftol=0, gtol=0)
    flat_array = x_opt[0]
    options['iteration'] = len(x_opt[1]) - 1
    if rank == 0:
        clean(options)
    likelihood_and_gradient(flat_array, 'f')          # parallel_GPLVM.py:120: the checkpoint evaluation
    map_reduce.flush(options)                         # every rank writes the embeddings of its own shards
    if rank == 0:
        with open(options['statistics'] + '/time_acc.obj', 'wb') as f:
            pickle.dump(time_acc, f)
        with open(options['statistics'] + '/nlml_acc.obj', 'wb') as f:
            pickle.dump(x_opt[1], f)
        with open(options['statistics'] + '/time_acc_SCG_adapted.obj', 'wb') as f:
            pickle.dump(x_opt[4], f)
    return x_opt


def init_statistics(map_reduce, options):
    """parallel_GPLVM.py:134-216 (Z from k-means of the first embeddings + 0.05 noise)."""
    options['global_statistics_names'] = {
        'Z': (options['M'], options['Q']), 'sf2': (1, 1), 'alpha': (1, options['Q']), 'beta': (1, 1)}
    options['accumulated_statistics_names'] = [
        'sum_YYT', 'sum_exp_K_mi_K_im', 'sum_exp_K_miY', 'sum_exp_K_ii', 'sum_KL',
        'sum_d_exp_K_miY_d_Z', 'sum_d_exp_K_mi_K_im_d_Z', 'sum_d_exp_K_miY_d_alpha',
        'sum_d_exp_K_mi_K_im_d_alpha', 'sum_d_exp_K_ii_d_sf2', 'sum_d_exp_K_miY_d_sf2',
        'sum_d_exp_K_mi_K_im_d_sf2']
    options['partial_derivatives_names'] = ['F', 'dF_dsum_exp_K_ii', 'dF_dKmm', 'dF_dsum_exp_K_miY',
                                            'dF_dsum_exp_K_mi_K_im']
    options['cache_names'] = ['Kmm', 'Kmm_inv']
    rank, world, _ = map_reduce.dist_info(options) if hasattr(map_reduce, 'dist_info') else (0, 1, None)
    if not options['load'] and rank == 0:
        names = sorted(os.listdir(options['input'] + '/'))[::world]      # the shards rank 0 maps (all of them in one process)
        fid = 0
        embeddings = map_reduce.load(options['embeddings'] + '/' + names[fid] + '.embedding.npy')
        while embeddings.shape[0] < options['M']:
            fid += 1
            embeddings = numpy.concatenate(
                (embeddings, map_reduce.load(options['embeddings'] + '/' + names[fid] + '.embedding.npy')))
        if embeddings.shape[1] != options['Q']:
            raise Exception('Given Q does not equal existing embedding data dimensions!')
        if options.get('b200_device_init', True) and hasattr(map_reduce, 'kmeans'):
            Z = map_reduce.kmeans(options, options['M'])      # same algorithm, one device pass per iteration
        else:
            import scipy.cluster.vq as cl
            Z = cl.kmeans(embeddings, options['M'])[0]
        missing = options['M'] - Z.shape[0]
        if missing > 0:
            Z = numpy.concatenate((Z, embeddings[:missing]))
        Z = Z + numpy.random.randn(options['M'], options['Q']) * 0.05
        global_statistics = {'Z': Z, 'sf2': numpy.array([[1.0]]), 'alpha': numpy.ones((1, options['Q'])),
                             'beta': numpy.array([[1.0]])}
    elif not options['load']:
        global_statistics = None                                          # rank 0's initial values arrive below
    else:
        global_statistics = {}
        for key in options['global_statistics_names']:
            global_statistics[key] = map_reduce.load(options['statistics'] + '/global_statistics_' + key + '_f.npy')
    if world > 1:
        global_statistics = map_reduce._bcast(global_statistics, options)
    bounds = {'Z': [(None, None)] * (options['M'] * options['Q']), 'sf2': [(0, None)],
              'alpha': [(0, None)] * options['Q'], 'beta': [(0, None)]}
    flat = []
    for key in options['global_statistics_names']:
        flat = flat + bounds[key]
    options['flat_global_statistics_bounds'] = flat
    options['flat_positive'] = numpy.array([b == (0, None) for b in flat])      # vectorised transforms below
    return options, global_statistics


def _transform_vec(options, x):
    """supporting_functions.py:131-136 over the flat vector: softplus where the bound is (0, None)."""
    pos = options['flat_positive']
    assert numpy.all(numpy.abs(x[pos]) < sp.lim_val)
    out = numpy.array(x, dtype=float)
    out[pos] = numpy.log(1 + numpy.exp(x[pos]))
    return out


def _transform_grad_vec(options, x):
    """supporting_functions.py:143-148: sigmoid where the bound is (0, None), 1 elsewhere."""
    pos = options['flat_positive']
    out = numpy.ones(len(x))
    out[pos] = 1 / (numpy.exp(-x[pos]) + 1)
    return out


def likelihood_and_gradient(flat_array, iteration, step_size=0):
    """parallel_GPLVM.py:222-279: returns (-F, -grad * transform_grad)."""
    global options, map_reduce, time_acc
    flat_array = numpy.asarray(flat_array, dtype=float)
    flat_t = _transform_vec(options, flat_array)
    global_statistics = rebuild_global_statistics(options, flat_t)
    options['i'] = iteration
    options['step_size'] = step_size
    rank, world, _ = map_reduce.dist_info(options) if hasattr(map_reduce, 'dist_info') else (0, 1, None)
    if rank == 0:
        clean(options)
    want_files = options.get('b200_write_files', True) or iteration == 'f'      # the 'f' checkpoint is always written
    # the reference's file protocol (cache -> statistics_MR -> partial_terms -> embeddings_MR through .npy files)
    # in one process; several ranks must stay bit-identical, so they all take the file-less evaluation and rank 0
    # writes the same files from the device state afterwards
    write = want_files and world == 1 and options.get('b200_write_files', True)
    if write:
        for key in global_statistics:
            map_reduce.save(options['statistics'] + '/global_statistics_' + key + '_' + str(options['i']) + '.npy',
                            global_statistics[key])
    t0 = time.time()
    if write:
        map_reduce.cache(options, global_statistics)
        files, mapper_time, reducer_time = map_reduce.statistics_MR(options)
        t1 = time.time()
        partial_derivatives, accumulated, partial_terms = calculate_global_statistics(
            options, global_statistics, files, map_reduce)
        gradient = calculate_global_derivatives(options, partial_derivatives, accumulated, global_statistics, partial_terms)
        partial_terms.close()
        likelihood = partial_derivatives['F']
        t2 = time.time()
        embeddings_time = []
        if not options['fixed_embeddings']:
            embeddings_time = map_reduce.embeddings_MR(options)
        t3 = time.time()
    else:
        likelihood, g = map_reduce.fast_evaluation(options, global_statistics)
        gradient = {'Z': g['Z'], 'sf2': numpy.array([[g['sf2']]]), 'alpha': g['alpha'].reshape(1, -1),
                    'beta': numpy.array([[g['beta']]])}
        mapper_time, reducer_time, embeddings_time = [], [], []
        t1 = t2 = t3 = time.time()
        if want_files:
            map_reduce.write_evaluation_files(options, global_statistics)
    time_acc['time_acc_statistics_map_reduce'] += [t1 - t0]
    time_acc['time_acc_statistics_mapper'] += [mapper_time]
    time_acc['time_acc_statistics_reducer'] += [reducer_time]
    time_acc['time_acc_calculate_global_statistics'] += [t2 - t1]
    if not options['fixed_embeddings']:
        time_acc['time_acc_embeddings_MR'] += [t3 - t2]
        time_acc['time_acc_embeddings_MR_mapper'] += [embeddings_time]
    gradient = flatten_global_statistics(options, gradient) * _transform_grad_vec(options, flat_array)
    return -1 * likelihood, -1 * gradient


def flatten_global_statistics(options, global_statistics):
    """parallel_GPLVM.py:286-290, in the key order of global_statistics_names."""
    return numpy.concatenate([numpy.asarray(global_statistics[k], dtype=float).flatten()
                              for k in options['global_statistics_names']])


def rebuild_global_statistics(options, flat_array):
    """parallel_GPLVM.py:292-299."""
    out, start = {}, 0
    for key, shape in options['global_statistics_names'].items():
        size = shape[0] * shape[1]
        out[key] = flat_array[start:start + size].reshape(shape)
        start += size
    return out


def calculate_global_statistics(options, global_statistics, accumulated_statistics_files, map_reduce):
    """parallel_GPLVM.py:302-334."""
    accumulated = {}
    for statistic, file_name in accumulated_statistics_files:
        accumulated[statistic] = map_reduce.load(file_name)
    partial_terms = map_reduce.load_partial_terms(options, global_statistics)
    map_reduce.load_cache(options, partial_terms)
    partial_terms.set_local_statistics(accumulated['sum_YYT'], accumulated['sum_exp_K_mi_K_im'],
                                       accumulated['sum_exp_K_miY'], accumulated['sum_exp_K_ii'], accumulated['sum_KL'])
    partial_derivatives = {
        'F': partial_terms.logmarglik(), 'dF_dsum_exp_K_ii': partial_terms.dF_dexp_K_ii(),
        'dF_dsum_exp_K_miY': partial_terms.dF_dexp_K_miY(),
        'dF_dsum_exp_K_mi_K_im': partial_terms.dF_dexp_K_mi_K_im(), 'dF_dKmm': partial_terms.dF_dKmm()}
    for key in partial_derivatives:
        map_reduce.save(options['statistics'] + '/partial_derivatives_' + key + '_' + str(options['i']) + '.npy',
                        partial_derivatives[key])
    return partial_derivatives, accumulated, partial_terms


def calculate_global_derivatives(options, pd, acc, global_statistics, partial_terms):
    """parallel_GPLVM.py:336-369."""
    grad_Z = partial_terms.grad_Z(pd['dF_dKmm'], partial_terms.dKmm_dZ(), pd['dF_dsum_exp_K_miY'],
                                  acc['sum_d_exp_K_miY_d_Z'], pd['dF_dsum_exp_K_mi_K_im'], acc['sum_d_exp_K_mi_K_im_d_Z'])
    grad_alpha = partial_terms.grad_alpha(pd['dF_dKmm'], partial_terms.dKmm_dalpha(), pd['dF_dsum_exp_K_miY'],
                                          acc['sum_d_exp_K_miY_d_alpha'], pd['dF_dsum_exp_K_mi_K_im'],
                                          acc['sum_d_exp_K_mi_K_im_d_alpha'])
    grad_sf2 = partial_terms.grad_sf2(pd['dF_dKmm'], partial_terms.dKmm_dsf2(), pd['dF_dsum_exp_K_ii'],
                                      acc['sum_d_exp_K_ii_d_sf2'], pd['dF_dsum_exp_K_miY'], acc['sum_d_exp_K_miY_d_sf2'],
                                      pd['dF_dsum_exp_K_mi_K_im'], acc['sum_d_exp_K_mi_K_im_d_sf2'])
    gradient = {'Z': grad_Z, 'sf2': numpy.array([[grad_sf2]]), 'alpha': numpy.asarray(grad_alpha).reshape(1, -1)}
    if not options['fixed_beta']:
        gradient['beta'] = numpy.array([[partial_terms.grad_beta()]])
    else:
        gradient['beta'] = numpy.zeros((1, 1))
    return gradient


def clean(options):
    """parallel_GPLVM.py:373-404."""
    if options['keep'] or options['i'] == 'f':
        return
    groups = (('global_statistics_', options['global_statistics_names']),
              ('accumulated_statistics_', options['accumulated_statistics_names']),
              ('partial_derivatives_', options['partial_derivatives_names']), ('cache_', options['cache_names']))
    for prefix, names in groups:
        for key in names:
            for it in (-1, options['i'] - 1, options['i']):
                map_reduce.remove(options['statistics'] + '/' + prefix + key + '_' + str(it) + '.npy')


def _parse_args(argv=None):
    """The reference's command line (parallel_GPLVM.py:407-516), reduced to the options this replay supports."""
    import argparse
    ap = argparse.ArgumentParser(description="Bayesian GPLVM / sparse GP on B200 (replay of parallel_GPLVM.py)")
    ap.add_argument("-i", "--input", required=True, help="folder of input shards (one CSV or .npy file per shard)")
    ap.add_argument("-e", "--embeddings", required=True, help="folder of the embeddings / variances / local gradients")
    ap.add_argument("-s", "--statistics", help="statistics folder (default: <tmp>/statistics)")
    ap.add_argument("--tmp", default="/tmp")
    ap.add_argument("-p", "--parallel", default="b200", choices=["b200"])
    ap.add_argument("-T", "--iterations", type=int, default=5)
    ap.add_argument("-M", type=int, default=2, dest="M")
    ap.add_argument("-Q", type=int, default=2, dest="Q")
    ap.add_argument("-D", type=int, default=4, dest="D")
    ap.add_argument("--init", default="PCA", choices=["PCA", "random"])
    ap.add_argument("--load", action="store_true", help="resume from the 'f' checkpoint in the statistics / embeddings folders")
    ap.add_argument("-k", "--keep", action="store_true", help="keep the per-iteration files")
    ap.add_argument("--fixed_embeddings", action="store_true")
    ap.add_argument("--fixed_beta", action="store_true")
    ap.add_argument("--drop_out_fraction", type=float, default=0)
    ap.add_argument("--no_files", action="store_true", help="no per-evaluation files (only the 'f' checkpoint)")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args(argv)
    o = dict(DEFAULTS)
    o.update(input=a.input, embeddings=a.embeddings, tmp=a.tmp, statistics=a.statistics or os.path.join(a.tmp, "statistics"),
             parallel=a.parallel, iterations=a.iterations, M=a.M, Q=a.Q, D=a.D, init=a.init, load=a.load, keep=a.keep,
             fixed_embeddings=a.fixed_embeddings, fixed_beta=a.fixed_beta, drop_out_fraction=a.drop_out_fraction,
             b200_write_files=not a.no_files, display=not a.quiet)
    return o


if __name__ == '__main__':
    opts = _parse_args()
    if int(os.environ.get("RANK", "0")) != 0:
        opts['display'] = False
    for d in ('statistics',):
        os.makedirs(opts[d], exist_ok=True)
    main(default_options(**opts))
