/* TEST INFRASTRUCTURE ONLY -- plain C restatement of the two per-shard maps of
 * GParML's variational-bound path, used as the parity checker at sizes where the
 * numpy oracle (oracle/gparml_oracle.py, one Python iteration per point like the
 * reference) would take minutes.  The product never links or calls this file.
 *
 * Pinning: tests/test_oracle.py checks this library against the numpy oracle and
 * against the golden vectors generated from the live reference (tests/golden).
 *
 * The formulas are written as the reference writes them (the closed forms in
 * kernel_exp.py and the per-point derivative expressions in partial_terms.py),
 * NOT in the refactored form the CUDA kernels use, so that agreement between
 * the two is meaningful.  All citations are relative to /root/reference.
 *
 * Build:  gcc -O2 -fopenmp -shared -fPIC -o libgparml_oracle.so gparml_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Psi1[n, m]: kernel_exp.py:80 */
static double psi1_nm(int Q, const double *z, const double *mu, const double *S,
                      double sf2, const double *alpha)
{
    double prod = 1.0, arg = 0.0;
    for (int q = 0; q < Q; ++q) {
        double den = S[q] * alpha[q] + 1.0;
        double d = z[q] - mu[q];
        prod *= sqrt(den);
        arg += d * d * alpha[q] / den;
    }
    return sf2 / prod * exp(-0.5 * arg);
}

/* Psi2_n[m, m']: kernel_exp.py:142-146 */
static double psi2_nmm(int Q, const double *za, const double *zb, const double *mu,
                       const double *S, double sf2, const double *alpha)
{
    double prod = 1.0, t1 = 0.0, t2 = 0.0;
    for (int q = 0; q < Q; ++q) {
        double den = 2.0 * alpha[q] * S[q] + 1.0;
        double dz = za[q] - zb[q];
        double dm = mu[q] - 0.5 * za[q] - 0.5 * zb[q];
        prod *= den;
        t1 += alpha[q] * dz * dz;
        t2 += alpha[q] * dm * dm / den;
    }
    return sf2 * sf2 / sqrt(prod) * exp(-0.25 * t1 - t2);
}

/* statistics map for one shard.
 *   psi2   (M, M)      partial_terms.py:45-48,79
 *   psi1y  (M, D)      kernel_exp.py:13-49
 *   d1_dZ  (M, Q, D)   partial_terms.py:162-188
 *   d2_dZ  (M, Q, M)   partial_terms.py:190-205
 *   d1_da  (Q, M, D)   partial_terms.py:256-271
 *   d2_da  (Q, M, M)   partial_terms.py:273-284
 *   scal[0] = sum_n y_n.y_n (partial_terms.py:40), scal[1] = KL (:83-87; 0 if all S == 0)
 */
void oracle_shard_stats(long n, int M, int Q, int D,
                        const double *Y, const double *mu, const double *S,
                        const double *Z, double sf2, const double *alpha,
                        double *psi2, double *psi1y, double *d1_dZ, double *d2_dZ,
                        double *d1_da, double *d2_da, double *scal)
{
    memset(psi2, 0, sizeof(double) * M * M);
    memset(psi1y, 0, sizeof(double) * M * D);
    memset(d1_dZ, 0, sizeof(double) * M * Q * D);
    memset(d2_dZ, 0, sizeof(double) * M * Q * M);
    memset(d1_da, 0, sizeof(double) * Q * M * D);
    memset(d2_da, 0, sizeof(double) * Q * M * M);

    double yyt = 0.0, kl = 0.0;
    int all_zero = 1;
    for (long i = 0; i < n; ++i) {
        for (int d = 0; d < D; ++d) yyt += Y[i * D + d] * Y[i * D + d];
        for (int q = 0; q < Q; ++q) if (S[i * Q + q] != 0.0) all_zero = 0;
    }
    if (!all_zero) {
        for (long i = 0; i < n; ++i) {
            double t = 0.0;
            for (int q = 0; q < Q; ++q) {
                double s = S[i * Q + q], m = mu[i * Q + q];
                t += s - log(s) + m * m;
            }
            kl += t - Q;
        }
        kl *= 0.5;
    }
    scal[0] = yyt;
    scal[1] = kl;

#pragma omp parallel for schedule(dynamic, 1)
    for (int a = 0; a < M; ++a) {
        const double *za = Z + (long)a * Q;
        for (long i = 0; i < n; ++i) {
            const double *mi = mu + i * Q, *si = S + i * Q, *yi = Y + i * D;
            /* Psi1 side */
            double p1 = psi1_nm(Q, za, mi, si, sf2, alpha);
            for (int d = 0; d < D; ++d) psi1y[a * D + d] += p1 * yi[d];
            for (int q = 0; q < Q; ++q) {
                double den = alpha[q] * si[q] + 1.0;
                double wz = p1 * alpha[q] * ((mi[q] - za[q]) / den);            /* :185 */
                double r = (mi[q] - za[q]) / den;
                double wa = -0.5 * p1 * (r * r + si[q] / den);                  /* :265 */
                for (int d = 0; d < D; ++d) {
                    d1_dZ[((long)a * Q + q) * D + d] += wz * yi[d];
                    d1_da[((long)q * M + a) * D + d] += wa * yi[d];
                }
            }
            /* Psi2 side */
            for (int b = 0; b < M; ++b) {
                const double *zb = Z + (long)b * Q;
                double p2 = psi2_nmm(Q, za, zb, mi, si, sf2, alpha);
                psi2[(long)a * M + b] += p2;
                for (int q = 0; q < Q; ++q) {
                    double den = 2.0 * alpha[q] * si[q] + 1.0;
                    double dz = za[q] - zb[q];
                    double sm = 2.0 * mi[q] - za[q] - zb[q];
                    d2_dZ[((long)a * Q + q) * M + b] +=
                        p2 * (-0.5 * alpha[q] * dz + 0.5 * alpha[q] * sm / den);        /* :201-203 */
                    double r = sm / den;
                    d2_da[((long)q * M + a) * M + b] +=
                        p2 * (-0.25 * dz * dz - 0.25 * r * r - si[q] / den);            /* :279-282 */
                }
            }
        }
    }
}

/* embeddings map for one shard: grad_X_mu (partial_terms.py:367-398) and grad_X_S
 * (:400-431) given G1 = dF/dPsi1Y (M, D) and G2 = dF/dPsi2 (M, M).  Positive-domain
 * gradients; the softplus chain (local_MapReduce.py:358) is applied by the caller. */
void oracle_embed_grads(long n, int M, int Q, int D,
                        const double *Y, const double *mu, const double *S,
                        const double *Z, double sf2, const double *alpha,
                        const double *G1, const double *G2,
                        double *g_mu, double *g_S)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double *mi = mu + i * Q, *si = S + i * Q, *yi = Y + i * D;
        double *gm = g_mu + i * Q, *gs = g_S + i * Q;
        for (int q = 0; q < Q; ++q) {
            gm[q] = -mi[q];                                   /* :385 */
            gs[q] = -0.5 * (1.0 - 1.0 / si[q]);               /* :418 */
        }
        for (int a = 0; a < M; ++a) {
            const double *za = Z + (long)a * Q;
            double p1 = psi1_nm(Q, za, mi, si, sf2, alpha);
            double b = 0.0;                                   /* sum_d G1[a,d] y_d */
            for (int d = 0; d < D; ++d) b += G1[a * D + d] * yi[d];
            for (int q = 0; q < Q; ++q) {
                double d1 = alpha[q] * si[q] + 1.0;
                double dm = mi[q] - za[q];
                gm[q] += b * p1 * (-alpha[q] * dm / d1);                                   /* :388-390 */
                double r = alpha[q] * dm / d1;
                gs[q] += b * p1 * (0.5 * r * r - 0.5 * (alpha[q] / d1));                  /* :421-423 */
            }
            for (int c = 0; c < M; ++c) {
                const double *zc = Z + (long)c * Q;
                double p2 = psi2_nmm(Q, za, zc, mi, si, sf2, alpha) * G2[(long)a * M + c];
                for (int q = 0; q < Q; ++q) {
                    double d2 = 2.0 * alpha[q] * si[q] + 1.0;
                    double sm = 2.0 * mi[q] - za[q] - zc[q];
                    gm[q] += p2 * (-alpha[q] * sm / d2);                                   /* :393 */
                    double r = alpha[q] * sm / (2.0 * d2);
                    gs[q] += p2 * (2.0 * r * r - alpha[q] / d2);                           /* :425-426 */
                }
            }
        }
    }
}

/* Psi1 matrix (n, M): kernel_exp.py:51-82 */
void oracle_psi1(long n, int M, int Q, const double *mu, const double *S, const double *Z,
                 double sf2, const double *alpha, double *out)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        for (int a = 0; a < M; ++a)
            out[i * M + a] = psi1_nm(Q, Z + (long)a * Q, mu + i * Q, S + i * Q, sf2, alpha);
}
