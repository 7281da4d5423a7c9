"""Python-3 replay of the reference's adapted Scaled Conjugate Gradient driver
(``scg_adapted.py:78-336``; the reference file is Python-2-only and cannot be imported).

This is a *caller* of the hot path, not part of it: it is kept so that an optimisation can be
run end to end against the new backend with the reference's evaluation protocol
(SURVEY.md 3.3): one ``f_and_gradf(x, iteration, step_size)`` callback, the per-point part of
the parameter vector living with the shards, and every inner product completed by the
local-state module (``scg_adapted_local_MapReduce`` in the reference,
``gparml_b200.scg_adapted_b200_MapReduce`` here; any module with the same 12 functions works).
"""
import sys
import traceback

import numpy as np
from numpy.linalg import LinAlgError

_allowed_failures = 100


class _Failures(object):
    count = 0


def safe_f_and_grad_f(f_and_gradf, x, iteration=0, step_size=0, *optargs):
    """scg_adapted.py:46-76: numerical failures become f = inf, grad = ones."""
    try:
        f, gradf = f_and_gradf(x, iteration, step_size, *optargs)
        _Failures.count = 0
    except (LinAlgError, ZeroDivisionError, ValueError, Warning, AssertionError) as e:
        if _Failures.count >= _allowed_failures:
            print("Too many errors...")
            raise e
        _Failures.count += 1
        tb = traceback.extract_tb(sys.exc_info()[2])[-1]
        print("An error occurred on line %s in filename %s" % (tb[1], tb[0]))
        print("Increasing failed count (%d) and returning nlml inf" % _Failures.count)
        f = np.inf
        gradf = np.ones(x.shape)
    return f, gradf


def _report(width, display, fnow, current_grad, beta, iteration):
    if display:
        print("{0:>0{w}g}  {1:> 12e}  {2:> 12e}  {3:> 12e}".format(iteration, float(fnow), float(beta),
                                                                   float(current_grad), w=width))
        sys.stdout.flush()


def SCG_adapted(f_and_gradf, x, tmp_folder, fixed_embeddings=False, optargs=(), maxiters=500, max_f_eval=500,
                display=True, xtol=None, ftol=None, gtol=None, local_ops=None):
    """Returns (x, flog, function_eval, status, local_ops.time_acc) like scg_adapted.py:336."""
    if local_ops is None:
        from . import scg_adapted_b200_MapReduce as local_ops
    xtol = 1e-6 if xtol is None else xtol
    ftol = 1e-6 if ftol is None else ftol
    gtol = 1e-5 if gtol is None else gtol
    sigma0 = 1.0e-4
    local = not fixed_embeddings

    fold, gradnew = safe_f_and_grad_f(f_and_gradf, x, 0, 0, *optargs)
    assert fold != float("inf")
    function_eval = 1
    fnow = fold
    gradold = gradnew.copy()
    d = -gradnew
    if local:
        local_ops.embeddings_set_grads(tmp_folder)
    current_grad = np.dot(gradnew, gradnew)
    if local:
        current_grad += local_ops.embeddings_get_grads_current_grad(tmp_folder)

    success, nsuccess = True, 0
    beta, betamin, betamax = 1.0, 1.0e-60, 1.0e100
    status = "Not converged"
    flog = [fold]
    iteration = 0
    width = len(str(maxiters))
    if display:
        print(" {0:{w}s}   {1:11s}    {2:11s}    {3:11s}".format("I", "F", "Scale", "|g|", w=width))
        print("Starting optimisation for %d iterations" % maxiters)

    while iteration < maxiters:
        if success:                                                # scg_adapted.py:145-167
            mu = np.dot(d, gradnew)
            if local:
                mu += local_ops.embeddings_get_grads_mu(tmp_folder)
            if mu >= 0:
                d = -gradnew
                if local:
                    local_ops.embeddings_set_grads_reset_d(tmp_folder)
                mu = np.dot(d, gradnew)
                if local:
                    mu += local_ops.embeddings_get_grads_mu(tmp_folder)
            kappa = np.dot(d, d)
            if local:
                kappa += local_ops.embeddings_get_grads_kappa(tmp_folder)
            sigma = sigma0 / np.sqrt(kappa)
            gplus = safe_f_and_grad_f(f_and_gradf, x + sigma * d, -1, sigma, *optargs)[1]
            theta = np.dot(d, gplus - gradnew)
            if local:
                theta += local_ops.embeddings_get_grads_theta(tmp_folder)
            theta = theta * np.sqrt(kappa) / sigma0

        delta = theta + beta * kappa                               # scg_adapted.py:187-192
        if delta <= 0:
            delta = beta * kappa
            beta = beta - theta / kappa
        alpha = -mu / delta

        xnew = x + alpha * d                                       # scg_adapted.py:203-230
        fnew, gradcand = safe_f_and_grad_f(f_and_gradf, xnew, iteration + 1, alpha, *optargs)
        function_eval += 1
        if function_eval >= max_f_eval:
            status = "Maximum number of function evaluations exceeded"
            break
        Delta = 2.0 * (fnew - fold) / (alpha * mu)
        if Delta >= 0.0:
            success = True
            nsuccess += 1
            x = xnew
            if local:
                local_ops.embeddings_set_grads_update_X(tmp_folder, alpha)
            fnow = fnew
        else:
            success = False
            fnow = fold

        flog.append(fnow)
        iteration += 1
        _report(width, display, fnow, current_grad, beta, iteration)

        if success:                                                # scg_adapted.py:248-283
            max_alpha_d = np.max(np.abs(alpha * d))
            if local:
                max_alpha_d = max(max_alpha_d, local_ops.embeddings_get_grads_max_d(tmp_folder, alpha))
            if max_alpha_d < xtol or np.abs(fnew - fold) < ftol:
                status = "converged"
                break
            gradold = gradnew
            if local:
                local_ops.embeddings_set_grads_update_grad_old(tmp_folder)
            gradnew = gradcand
            if local:
                local_ops.embeddings_set_grads_update_grad_new(tmp_folder)
            current_grad = np.dot(gradnew, gradnew)
            if local:
                current_grad += local_ops.embeddings_get_grads_current_grad(tmp_folder)
            fold = fnew
            if current_grad <= gtol:
                status = "converged"
                break

        if Delta < 0.25:                                           # scg_adapted.py:286-289
            beta = min(4.0 * beta, betamax)
        if Delta > 0.75:
            beta = max(0.5 * beta, betamin)

        if nsuccess == x.size:                                     # scg_adapted.py:299-314
            d = -gradnew
            nsuccess = 0
        elif success:
            Gamma = (np.dot(gradold, gradnew) - current_grad) / mu
            if local:
                Gamma += local_ops.embeddings_get_grads_gamma(tmp_folder) / mu
            d = Gamma * d - gradnew
            if local:
                local_ops.embeddings_set_grads_update_d(tmp_folder, Gamma)
    else:
        status = "maxiter exceeded"

    if display:
        _report(width, display, fnow, current_grad, beta, iteration)
        print(status)
    return x, flog, function_eval, status, getattr(local_ops, "time_acc", {})
