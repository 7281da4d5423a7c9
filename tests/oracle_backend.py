"""Test helper: the oracle hosted behind the same two interfaces the product implements
(an ``f_and_gradf(x, iteration, step_size)`` callback and a local-state module), so that the
py3 SCG driver can be run against it and compared with the CUDA backend."""
import numpy as np

from oracle import gparml_oracle as O


class OracleBackend(object):
    def __init__(self, shards, M, Q, fixed_embeddings=False, evaluate=None):
        # shards: list of dict(Y, X_mu, X_S) -- copies are taken
        self.st = [dict(Y=s["Y"].copy(), X_mu=s["X_mu"].copy(), X_S=s["X_S"].copy()) for s in shards]
        self.M, self.Q, self.fixed = M, Q, fixed_embeddings
        self.evaluate = evaluate or O.evaluate
        self.time_acc = {}

    def f_and_gradf(self, x, iteration, step_size=0):
        M, Q = self.M, self.Q
        pos = x.copy()
        pos[M * Q:] = O.softplus(x[M * Q:])
        Z = pos[:M * Q].reshape(M, Q)
        sf2, alpha, beta = pos[M * Q], pos[M * Q + 1:M * Q + 1 + Q], pos[M * Q + 1 + Q]
        shards = [dict(Y=s["Y"], X_mu=s["X_mu"], X_S=s["X_S"], d=s.get("d")) for s in self.st]
        res = self.evaluate(shards, Z, sf2, alpha, beta, step_size=step_size, fixed_embeddings=self.fixed)
        for s, g in zip(self.st, res["grad_latest"]):
            s["latest"] = g
        return O.objective_and_flat_gradient(res, x, M, Q)

    # local-state module surface (scg_adapted_local_MapReduce.py)
    def embeddings_set_grads(self, folder): O.scg_set_grads(self.st)
    def embeddings_get_grads_mu(self, folder): return O.scg_get_mu(self.st)
    def embeddings_get_grads_kappa(self, folder): return O.scg_get_kappa(self.st)
    def embeddings_get_grads_theta(self, folder): return O.scg_get_theta(self.st)
    def embeddings_get_grads_current_grad(self, folder): return O.scg_get_current_grad(self.st)
    def embeddings_get_grads_gamma(self, folder): return O.scg_get_gamma(self.st)
    def embeddings_get_grads_max_d(self, folder, alpha): return O.scg_get_max_d(self.st, alpha)
    def embeddings_set_grads_reset_d(self, folder): O.scg_reset_d(self.st)
    def embeddings_set_grads_update_d(self, folder, gamma): O.scg_update_d(self.st, gamma)
    def embeddings_set_grads_update_X(self, folder, alpha): O.scg_update_X(self.st, alpha)
    def embeddings_set_grads_update_grad_old(self, folder): O.scg_update_grad_old(self.st)
    def embeddings_set_grads_update_grad_new(self, folder): O.scg_update_grad_new(self.st)
