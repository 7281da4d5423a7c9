"""Times the device-side initialisation at a BASELINE shape and the reference's
host algorithms on a bounded sample.  Usage: python tools/init_time.py [N M Q D]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gparml_b200 import _lib, init_device  # noqa: E402
from gparml_b200.engine import ShardContext  # noqa: E402

N, M, Q, D = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1000000, 100, 10, 10)
rng = np.random.default_rng(0)
Y = rng.standard_normal((N, Q)) @ rng.standard_normal((Q, D)) + 0.1 * rng.standard_normal((N, D))
c = ShardContext(M, Q, D, N)
t = time.perf_counter(); c.upload_outputs(Y); c.synchronize(); t_up = time.perf_counter() - t
for rep in range(2):
    t = time.perf_counter(); init_device.pca([c]); c.synchronize(); t_pca = time.perf_counter() - t
t = time.perf_counter(); init_device.random_variances([c], 1); c.synchronize(); t_var = time.perf_counter() - t
X = c.download(_lib.A_X_MU, (N, Q))
guess = X[rng.choice(N, M, replace=False)]
c.kmeans_step(guess)
t = time.perf_counter(); c.kmeans_step(guess); t_km = time.perf_counter() - t
t = time.perf_counter(); book, dist = init_device.kmeans_from_guess([c], guess); t_lloyd = time.perf_counter() - t
print("device: upload %.1f ms, PCA %.2f ms, variance draw %.2f ms, one k-means pass %.2f ms, Lloyd run %.1f ms (distortion %.4f)"
      % (t_up * 1e3, t_pca * 1e3, t_var * 1e3, t_km * 1e3, t_lloyd * 1e3, dist))
ns = min(N, 200000)
t = time.perf_counter(); Z = np.linalg.svd(Y[:ns] - Y[:ns].mean(axis=0), full_matrices=False); t_svd = time.perf_counter() - t
import scipy.cluster.vq as cl  # noqa: E402
nk = min(N, 125000)
t = time.perf_counter(); cl.kmeans(X[:nk], guess); t_sk = time.perf_counter() - t
print("host: numpy SVD of %d rows %.1f ms (linear in N), scipy Lloyd run from the same guess on %d rows %.1f ms" % (ns, t_svd * 1e3, nk, t_sk * 1e3))
c.close()
