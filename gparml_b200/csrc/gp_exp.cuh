// Table-driven double-precision exp for the Psi kernels.
//
//   exp(x) = 2^m * 2^(j/N) * exp(r),   k = round(x * N/ln2) = N m + j,   r = x - k ln2/N
//   N = 256 (default): |r| <= ln2/512 = 1.35e-3, degree-3 near-minimax polynomial, max relative error 1.8e-14
//                      -> 7 FP64-pipe instructions (3 DFMA/DADD reduction, 3 DFMA polynomial, 1 DMUL by the table entry)
//   N = 64  (GP_EXP_LOG2_TAB 6, psi2.cu): |r| <= ln2/128 = 5.4e-3, degree-4 Taylor polynomial (truncation 3.9e-14) -> 8
// instead of the 18 of libdevice's exp() (round 1: 32 entries, degree 5, 9 instructions).  The 2^m scaling and the
// table index are integer-pipe work and the table read is one 8-byte shared-memory load (tools/gen_exp_table.py wrote
// gp_exp_table8.inc, correctly rounded entries).  Measured on B200 at c3 (tools/tune.py): embed_psi2x 18.98 ms with
// 64 entries / 8 instructions, 18.67 with 256 / 7, 18.73 with 1024 entries and the degree-3 Taylor polynomial;
// psi2x_stats 21.24 / 22.27 / 22.90 -- there the shorter polynomial no longer hides the table read and the larger
// tables add bank conflicts, so psi2.cu keeps the 64-entry form.
// Relative error <= ~4e-14 + |x| * 1.1e-16 (argument reduction with a single rounded ln2/N), far inside the 1e-9
// parity budget on the summed statistics.
//
// Domain: any x <= 700 including -inf (Psi values are bounded by sf^2 resp. sf^4).  The argument is first
// clamped to >= -745.25 by gp_exp_clamp -- an unsigned integer min on the high word (negative doubles order
// like their bit patterns), one integer-pipe instruction, no FP64 slot -- because k = round(x N/ln2) is read
// from the low word of t and would wrap for very negative x (an exponent of -1e8 is what un-normalised
// regression inputs give: alpha = 1, inputs spanning 1e4, kernel_exp.py:143-146 then calls np.exp(-1e8) = 0).
// x < -709 returns ~4e-308..9e-308 instead of a denormal / zero (m is clamped), which contributes nothing
// to any sum.
#pragma once

#ifndef GP_EXP_LOG2_TAB
#define GP_EXP_LOG2_TAB 8
#endif
#define GP_EXP_TAB (1 << GP_EXP_LOG2_TAB)
#define GP_EXP_SHIFT 6755399441055744.0           // 1.5 * 2^52: the low word of fma(x, SCALE, SHIFT) is round(x SCALE)

#if GP_EXP_LOG2_TAB == 8
// 256 entries, |r| <= ln2/512 = 1.35e-3: degree-3 near-minimax polynomial (Chebyshev interpolation refined on the max
// relative error: 1.8e-14; the Taylor coefficients would give 1.4e-13)
#define GP_EXP_SCALE 369.3299304675746           // 256 / ln2
#define GP_EXP_NEG_STEP -0.0027076061740622863    // -ln2 / 256
#define GP_EXP_C3 0.16666670155072924
#define GP_EXP_C2 0.5000000762607874
#define GP_EXP_C1 0.9999999999999927
#define GP_EXP_C0 0.9999999999999827
#define GP_EXP_POLY_HEAD(r) fma(fma((r), GP_EXP_C3, GP_EXP_C2), (r), GP_EXP_C1)
#define GP_EXP_POLY_STEPS 3
static __device__ const double gp_exp_table_const[GP_EXP_TAB] = {
#include "gp_exp_table8.inc"
};
#elif GP_EXP_LOG2_TAB == 6
#define GP_EXP_SCALE 92.33248261689366            // 64 / ln2
#define GP_EXP_NEG_STEP -0.010830424696249145     // -ln2 / 64
#define GP_EXP_POLY_HEAD(r) fma(fma(fma((r), 1.0 / 24.0, 1.0 / 6.0), (r), 0.5), (r), 1.0)
#define GP_EXP_POLY_STEPS 4
#define GP_EXP_C1 1.0
#define GP_EXP_C0 1.0
// 2^(j/64), j = 0 .. 63, correctly rounded
#define GP_EXP_TABLE_VALUES                                                                                          \
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138,                          \
    1.0556451783605572, 1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841,            \
    1.1143867425958924, 1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775,           \
    1.1763969916502812, 1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,                \
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096,             \
    1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905,            \
    1.383909881963832, 1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467,              \
    1.460917794180647, 1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,            \
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543,             \
    1.6280274218573478, 1.645755478153965, 1.6636765803267364, 1.681792830507429, 1.7001063537185235,             \
    1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072,            \
    1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,              \
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951

static __device__ const double gp_exp_table_const[GP_EXP_TAB] = {GP_EXP_TABLE_VALUES};
#else
#error "GP_EXP_LOG2_TAB must be 6 or 8"
#endif

// copy the table into shared memory (call from all threads, then __syncthreads())
__device__ __forceinline__ void gp_exp_load_table(double *tab_smem)
{
    for (int i = threadIdx.x; i < GP_EXP_TAB; i += blockDim.x) tab_smem[i] = gp_exp_table_const[i];
}

// x >= -745.25 for every x (positive x, and x >= -745.25, pass through bit-identically)
__device__ __forceinline__ double gp_exp_clamp(double x)
{
    const unsigned hi = (unsigned)__double2hiint(x);
    return __hiloint2double((int)(hi < 0xC0874A00u ? hi : 0xC0874A00u), __double2loint(x));
}

// sign_word: 0 or 0x80000000, XORed into the sign bit of the result (integer pipe), i.e.
// +-exp(x) without a multiplication
__device__ __forceinline__ double gp_exp_signed(double x, const double *tab_smem, int sign_word);

__device__ __forceinline__ double gp_exp(double x, const double *tab_smem) { return gp_exp_signed(x, tab_smem, 0); }

__device__ __forceinline__ double gp_exp_signed(double x, const double *tab_smem, int sign_word)
{
    x = gp_exp_clamp(x);
    const double t = fma(x, GP_EXP_SCALE, GP_EXP_SHIFT);
    const int k = __double2loint(t);
    const double kd = t - GP_EXP_SHIFT;
    const double r = fma(kd, GP_EXP_NEG_STEP, x);
    const double p = fma(GP_EXP_POLY_HEAD(r), r, GP_EXP_C0);
    int m = k >> GP_EXP_LOG2_TAB;
    m = m < -1021 ? -1021 : m;
    const double s = tab_smem[k & (GP_EXP_TAB - 1)] * p;      // in [1, 2.03)
    return __hiloint2double((__double2hiint(s) + (m << 20)) ^ sign_word, __double2loint(s));
}
