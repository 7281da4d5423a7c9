"""Drop-in mirror of the reference's arithmetic object ``partial_terms.partial_terms``
(``/root/reference/partial_terms.py:15``): same constructor, methods, argument meaning,
public attributes and error behaviour, with every array computed by the sm_100a library
through the C ABI (``include/gparml_b200.h``).  There is no numpy arithmetic here beyond
moving arrays and multiplying two scalars; without the CUDA library the class cannot be
constructed.

Differences a maintainer should know about (also in INTEGRATION.md):

* ``set_global_statistics(Kmm, Kmm_inv)`` stores the arrays for attribute compatibility
  (``.Kmm``, ``.Kmm_inv``) but the device always derives Kmm and its inverse from the
  current ``Z`` / ``hyp`` itself (it is the same function of them, ``local_MapReduce.py:
  383-394``), by Cholesky instead of LU.
* the per-point tensors ``exp_K_mi_K_im`` (n x M x M) are never materialised; the four
  ``d*_dZ`` / ``d*_dalpha`` tensors of the last ``set_data(..., is_set_statistics=True)``
  are kept instead, which is all the reference's callers read.
"""
import numpy as np

from . import _lib
from .engine import ShardContext


class ArdHypers(object):
    """kernels.ArdHypers (kernels.py:11-35): sf = sqrt(sf2), ard = alpha ** -0.5."""

    def __init__(self, D, sf=1.0, ll=1.0, ard=None):
        self.D = D
        self.sf = sf
        if ard is None:
            self.ard = np.ones(D) * ll
        else:
            self.ard = np.atleast_1d(np.array(ard, dtype=np.float64).squeeze())
            assert self.ard.ndim == 1

    @property
    def ll(self):
        if np.all(self.ard == self.ard[0]):
            return self.ard[0]
        raise ValueError("RBF kernel is not isotropic")

    @ll.setter
    def ll(self, value):
        self.ard = np.ones(self.D) * value


_LOCAL_FIVE = ("sum_YYT", "sum_exp_K_mi_K_im", "exp_K_miY", "sum_exp_K_ii", "KL")


class partial_terms(object):
    def __init__(self, Z, sf2, alpha, beta, M, Q, N, D, update_global_statistics=True, device=0):
        self.Z = Z
        self.M, self.Q, self.N, self.D = int(M), int(Q), int(N), int(D)
        self.beta = beta
        self.hyp = ArdHypers(self.Q, sf=float(sf2) ** 0.5, ard=np.asarray(alpha, dtype=np.float64) ** -0.5)
        self._ctx = ShardContext(self.M, self.Q, self.D, self.N, device=device)
        self._pushed = None          # globals last sent to the device
        self._stats_sig = None       # statistics last sent to / produced on the device
        self._glob = None            # cached result of the device global step
        self._local_named = None     # the 12 statistics of the last set_data(True)
        self._have_data = False
        self.local_N = 0
        if update_global_statistics:
            self.update_global_statistics()

    # ------------------------------------------------------------------ plumbing
    def close(self):
        self._ctx.close()

    def _alpha(self):
        return np.asarray(self.hyp.ard, dtype=np.float64) ** -2

    def _push_globals(self):
        Z = np.ascontiguousarray(np.asarray(self.Z, dtype=np.float64).reshape(self.M, self.Q))
        cur = (Z.tobytes(), float(self.hyp.sf) ** 2, self._alpha().tobytes(), float(self.beta))
        if cur != self._pushed:
            self._ctx.set_globals(Z, cur[1], self._alpha(), cur[3])
            self._pushed = cur
            self._glob = None
            return True
        return False

    def _current_five(self):
        return (float(self.sum_YYT), np.asarray(self.sum_exp_K_mi_K_im, dtype=np.float64),
                np.asarray(self.exp_K_miY, dtype=np.float64), float(self.sum_exp_K_ii), float(self.KL))

    def _sync_stats(self):
        """Make the device's packed buffer hold the statistics currently in the attributes."""
        five = self._current_five()
        sig = (five[0], five[1].tobytes(), five[2].tobytes(), five[3], five[4])
        if sig != self._stats_sig:
            self._ctx.set_stats_named({"sum_YYT": five[0], "sum_exp_K_mi_K_im": five[1], "sum_exp_K_miY": five[2],
                                       "sum_exp_K_ii": five[3], "sum_KL": five[4]})
            self._stats_sig = sig
            self._glob = None

    def _global(self):
        changed = self._push_globals()
        if changed and self._have_data:
            pass            # statistics in the attributes are the caller's responsibility, as in the reference
        self._sync_stats()
        if self._glob is None:
            F, grad = self._ctx.global_step()
            self._glob = {"F": F, "grad": grad}
        return self._glob

    # ------------------------------------------------------------------ data / statistics
    def set_data(self, Y, X_mu, X_S, is_set_statistics=True):
        """partial_terms.py:38-52.  X_S is the positive variance."""
        self.Y = Y
        self.X_mu = X_mu
        self.X_S = X_S
        self.local_N = np.asarray(X_mu).shape[0]
        assert np.all(np.asarray(X_S) >= 0.0)                      # kernel_exp.py:29
        self._ctx.upload_shard(Y, X_mu, X_S, positive_variance=True)
        self._have_data = True
        self._glob = None
        if is_set_statistics:
            self.update_local_statistics()

    def update_local_statistics(self):
        """partial_terms.py:74-87 (and the derivative tensors of :162-205, :256-284)."""
        self._push_globals()
        self._ctx.statistics()
        named = self._ctx.stats_named()
        self._local_named = named
        self.sum_YYT = named["sum_YYT"]
        self.sum_exp_K_mi_K_im = named["sum_exp_K_mi_K_im"]
        self.exp_K_miY = named["sum_exp_K_miY"]
        self.sum_exp_K_ii = named["sum_exp_K_ii"]
        self.KL = named["sum_KL"]
        five = self._current_five()
        self._stats_sig = (five[0], five[1].tobytes(), five[2].tobytes(), five[3], five[4])
        self._glob = None

    def set_local_statistics(self, sum_YYT, sum_exp_K_mi_K_im, exp_K_miY, sum_exp_K_ii, KL):
        """partial_terms.py:54-61."""
        self.sum_YYT = sum_YYT
        self.sum_exp_K_mi_K_im = sum_exp_K_mi_K_im
        self.exp_K_miY = exp_K_miY
        self.sum_exp_K_ii = sum_exp_K_ii
        self.KL = KL
        self._glob = None

    def get_local_statistics(self):
        """partial_terms.py:63-68."""
        return {"sum_YYT": self.sum_YYT, "sum_exp_K_mi_K_im": self.sum_exp_K_mi_K_im, "exp_K_miY": self.exp_K_miY,
                "sum_exp_K_ii": self.sum_exp_K_ii, "KL": self.KL}

    def set_global_statistics(self, Kmm, Kmm_inv):
        """partial_terms.py:70-72."""
        self.Kmm = Kmm
        self.Kmm_inv = Kmm_inv

    def update_global_statistics(self):
        """partial_terms.py:89-95."""
        self._push_globals()
        self._ctx.update_global_statistics()
        self.Kmm = self._ctx.download(_lib.A_KMM, (self.M, self.M))
        self.Kmm_inv = self._ctx.download(_lib.A_KMM_INV, (self.M, self.M))
        self._glob = None

    @property
    def exp_K_mi(self):
        """kernel_exp.calc_expect_K_mi (kernel_exp.py:51-82): (n, M)."""
        self._push_globals()
        return self._ctx.download(_lib.A_PSI1, (self.local_N, self.M))

    @property
    def Kmm_plus_op_inv(self):
        self._global()
        return self._ctx.download(_lib.A_A_INV, (self.M, self.M))

    # ------------------------------------------------------------------ partial gradients of F
    def dF_dKmm(self):
        """partial_terms.py:102-113."""
        self._global()
        return self._ctx.download(_lib.A_DF_DKMM, (self.M, self.M))

    def dF_dexp_K_miY(self):
        """partial_terms.py:115-121."""
        self._global()
        return self._ctx.download(_lib.A_DF_DPSI1Y, (self.M, self.D))

    def dF_dexp_K_mi_K_im(self):
        """partial_terms.py:123-131."""
        self._global()
        return self._ctx.download(_lib.A_DF_DPSI2, (self.M, self.M))

    def dF_dexp_K_ii(self):
        """partial_terms.py:133-138: a product of two scalars and a constant."""
        return -0.5 * self.beta * self.D

    # ------------------------------------------------------------------ Z
    def dKmm_dZ(self):
        """partial_terms.py:146-160."""
        self._push_globals()
        self._ctx.update_global_statistics()
        return self._ctx.kmm_derivative(0)

    def _local(self, key):
        if self._local_named is None:
            if not self._have_data:
                raise ValueError("set_data must be called first")
            saved = None
            if hasattr(self, "sum_YYT"):
                saved = self.get_local_statistics()
            self.update_local_statistics()
            if saved is not None:       # set_data(False) + set_local_statistics(): keep the caller's sums
                self.set_local_statistics(saved["sum_YYT"], saved["sum_exp_K_mi_K_im"], saved["exp_K_miY"],
                                          saved["sum_exp_K_ii"], saved["KL"])
        return self._local_named[key]

    def dexp_K_miY_dZ(self):
        """partial_terms.py:162-188."""
        return self._local("sum_d_exp_K_miY_d_Z")

    def dexp_K_mi_K_im_dZ(self):
        """partial_terms.py:190-205."""
        return self._local("sum_d_exp_K_mi_K_im_d_Z")

    def grad_Z(self, dF_dKmm, dKmm_dZ, dF_dexp_K_miY, dexp_K_miY_dZ, dF_dexp_K_mi_K_im, dexp_K_mi_K_im_dZ):
        """partial_terms.py:207-240."""
        return self._ctx.grad_contract(0, dF_dKmm, dKmm_dZ, dF_dexp_K_miY, dexp_K_miY_dZ, dF_dexp_K_mi_K_im,
                                       dexp_K_mi_K_im_dZ)

    # ------------------------------------------------------------------ alpha
    def dKmm_dalpha(self):
        """partial_terms.py:247-254."""
        self._push_globals()
        self._ctx.update_global_statistics()
        return self._ctx.kmm_derivative(1)

    def dexp_K_miY_dalpha(self):
        """partial_terms.py:256-271."""
        return self._local("sum_d_exp_K_miY_d_alpha")

    def dexp_K_mi_K_im_dalpha(self):
        """partial_terms.py:273-284."""
        return self._local("sum_d_exp_K_mi_K_im_d_alpha")

    def grad_alpha(self, dF_dKmm, dKmm_dalpha, dF_dexp_K_miY, dexp_K_miY_dalpha, dF_dexp_K_mi_K_im,
                   dexp_K_mi_K_im_dalpha):
        """partial_terms.py:286-299."""
        return self._ctx.grad_contract(1, dF_dKmm, dKmm_dalpha, dF_dexp_K_miY, dexp_K_miY_dalpha,
                                       dF_dexp_K_mi_K_im, dexp_K_mi_K_im_dalpha)

    # ------------------------------------------------------------------ sf2
    def dKmm_dsf2(self):
        """partial_terms.py:306-308."""
        self._push_globals()
        self._ctx.update_global_statistics()
        return self._ctx.kmm_derivative(2)

    def dexp_K_miY_dsf2(self):
        """partial_terms.py:310-312."""
        self._push_globals()
        self._sync_stats()
        return self._ctx.stats_named(("sum_d_exp_K_miY_d_sf2",))["sum_d_exp_K_miY_d_sf2"]

    def dexp_K_mi_K_im_dsf2(self):
        """partial_terms.py:314-316."""
        self._push_globals()
        self._sync_stats()
        return self._ctx.stats_named(("sum_d_exp_K_mi_K_im_d_sf2",))["sum_d_exp_K_mi_K_im_d_sf2"]

    def dexp_K_ii_dsf2(self):
        """partial_terms.py:318-320."""
        return self.local_N

    def grad_sf2(self, dF_dKmm, dKmm_dsf2, dF_dexp_K_ii, dexp_K_ii_dsf2, dF_dexp_K_miY, dexp_K_miY_dsf2,
                 dF_dexp_K_mi_K_im, dexp_K_mi_K_im_dsf2):
        """partial_terms.py:322-333."""
        mat = self._ctx.grad_contract(2, dF_dKmm, dKmm_dsf2, dF_dexp_K_miY, dexp_K_miY_dsf2, dF_dexp_K_mi_K_im,
                                      dexp_K_mi_K_im_dsf2)
        return float(mat[0]) + float(dF_dexp_K_ii) * float(dexp_K_ii_dsf2)

    # ------------------------------------------------------------------ beta, X, bound
    def grad_beta(self):
        """partial_terms.py:340-360."""
        return self._global()["grad"]["beta"]

    def _embed(self):
        if not self._have_data:
            raise ValueError("set_data must be called first")
        g = self._global()
        if getattr(self, "_embed_for", None) is g:
            return                      # grad_X_mu / grad_X_S of the same state: one map serves both
        self._ctx.embedding_grads()
        self._embed_for = g

    def grad_X_mu(self):
        """partial_terms.py:367-398."""
        self._embed()
        return self._ctx.download(_lib.A_GRAD_X_MU, (self.local_N, self.Q))

    def grad_X_S(self):
        """partial_terms.py:400-431."""
        self._embed()
        return self._ctx.download(_lib.A_GRAD_X_S, (self.local_N, self.Q))

    def logmarglik(self):
        """partial_terms.py:436-473."""
        return self._global()["F"]
