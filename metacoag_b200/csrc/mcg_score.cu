// Coverage featurisation, bin profiles and the composition x coverage scoring kernels
// (SURVEY.md section 8a rows 5-12).
//
//   reference arithmetic: matching_utils.py:21-66 (normpdf, get_tetramer_distance,
//   get_comp_probability, get_cov_probability), the aggregation rules of
//   matching_utils.py:117-155 / 368-398 (A), label_prop_utils.py:54-86 (B) and
//   label_prop_utils.py:308-345 (C), feature_utils.py:217-235 (get_bin_profiles).
//
// Everything is float64 on the CUDA cores (the reference is float64, K = 136 is small,
// and parity is 1e-9 relative on the probabilities), so the roofline is the FP64 pipe.
//
// pair_kernel: a CTA owns a 64-row tile of the "A" side (contigs to score) and walks the
// "B" side (bin centroids, or member contigs for rules A/B) in 64-row tiles.  Both tiles
// are staged in shared memory transposed ([k][row], XOR-swizzled so that the transposing
// store and the strided reads are both conflict-free); each of the 256 threads owns a
// 4 x 4 block of pairs (rows ti+16e, columns tj+16f), so every shared-memory operand is
// reused four times from registers.  The Poisson product over samples is kept strictly
// in sample order per pair (the reference's running product decides the
// prob_product > 0 branch through underflow), with log(c) and lgamma(c+1) hoisted per
// (row, sample) by cov_terms_kernel.
//
// Rule C on large batches (every long contig against every bin), by default a branch and bound:
// the all-pairs distances as a tensor-core BOUND with a stated margin (mcg_tc.cu: d2a, plus the
// smallest lower bound per row and 64-bin tile); bound_pre_kernel (one lane per contig: nearest
// centroid, its weight W, the distance threshold that follows; it lists the single candidate of
// the contigs whose other bins are all beyond that threshold); bound_kernel (one warp per
// remaining contig: second probe, candidate list); cand_score_kernel (float64 d^2 and weight of
// exactly the listed pairs); cand_argmin_kernel (first minimum).  Same bits as the all-pairs
// pass, which is dist2_kernel (the float64 d^2 matrix on the FP64 tensor path) + prob_kernel:
// taken when the weight matrix is wanted, with MCG_PRUNE=0, or when the list overflows.
#include "mcg_internal.cuh"
#include <cstdlib>

namespace {

constexpr int kTile = 64;          // rows per side per CTA tile
constexpr int kPairThreads = 256;
constexpr int kSampleChunk = 10;   // samples staged per pass

// matching_utils.py:12-15
struct GaussConst {
    double two_var_intra, denom_intra;   // 2*sd^2, sd*sqrt(2*pi)
    double mu_inter, two_var_inter, denom_inter;
    double ln10;
};
__constant__ GaussConst c_gauss;

// ---- CPython's math.lgamma for x > 0 (Modules/mathmodule.c: Lanczos approximation with
// the lanczos13m53 rational, g = 6.024680040776729583740234375), restated; the operation
// order is the one oracle/mcg_oracle.c pins against math.lgamma. ---------------------------
__constant__ double c_lanczos_num[13] = {
    23531376880.410759688572007674451636754734846804940,
    42919803642.649098768957899047001988850926355848959,
    35711959237.355668049440185451547166705960488635843,
    17921034426.037209699919755754458931112671403265390,
    6039542586.3520280050642916443072979210699388420708,
    1439720407.3117216736632230727949123939715485786772,
    248874557.86205415651146038641322942321632125127801,
    31426415.585400194380614231628318205362874684987640,
    2876370.6289353724412254090516208496135991145378768,
    186056.26539522349504029498971604569928220784236328,
    8071.6720023658162106380029022722506138218516325024,
    210.82427775157934587250973392071336271166969580291,
    2.5066282746310002701649081771338373386264310793408};
__constant__ double c_lanczos_den[13] = {0.0,        39916800.0, 120543840.0, 150917976.0,
                                         105258076.0, 45995730.0, 13339535.0,  2637558.0,
                                         357423.0,    32670.0,    1925.0,      66.0,
                                         1.0};

__device__ double lgamma_cpython(double x) {
    const double g = 6.024680040776729583740234375;
    if (!(x > 0.0) || isinf(x)) return lgamma(x);   // outside the path's domain
    if (x == floor(x) && x <= 2.0) return 0.0;
    if (x < 1e-20) return -log(x);
    double num = 0.0, den = 0.0;
    if (x < 5.0) {
        for (int i = 12; i >= 0; --i) {
            num = __dadd_rn(__dmul_rn(num, x), c_lanczos_num[i]);
            den = __dadd_rn(__dmul_rn(den, x), c_lanczos_den[i]);
        }
    } else {
        // CPython divides by x at every step; multiplying by 1/x with an FMA differs from
        // that by a couple of ulp of the sums (checked against math.lgamma at 5e-15)
        const double rx = __ddiv_rn(1.0, x);
        for (int i = 0; i < 13; ++i) {
            num = fma(num, rx, c_lanczos_num[i]);
            den = fma(den, rx, c_lanczos_den[i]);
        }
    }
    double r = __dsub_rn(log(__ddiv_rn(num, den)), g);
    const double t = __dsub_rn(log(__dsub_rn(__dadd_rn(x, g), 0.5)), 1.0);
    r = __dadd_rn(r, __dmul_rn(__dsub_rn(x, 0.5), t));
    return r;
}

// matching_utils.py:21-25
__device__ __forceinline__ double normpdf_dev(double x, double mean, double two_var,
                                              double denom) {
    const double dx = __dsub_rn(x, mean);
    const double num = exp(__ddiv_rn(-__dmul_rn(dx, dx), two_var));
    return __ddiv_rn(num, denom);
}

// matching_utils.py:36-39.  NaN when both Gaussians underflow (the reference raises
// ZeroDivisionError there); the caller turns that into MCG_ERR_DOMAIN.
__device__ __noinline__ double comp_prob_dev(double d) {
    const double gi = normpdf_dev(d, 0.0, c_gauss.two_var_intra, c_gauss.denom_intra);
    const double ge = normpdf_dev(d, c_gauss.mu_inter, c_gauss.two_var_inter,
                                  c_gauss.denom_inter);
    return __ddiv_rn(gi, __dadd_rn(gi, ge));
}

// ---- fast exp / log for the hot loop -----------------------------------------------------------
// exp(x) = 2^(k/16) * e^r with k = round(16 x / ln 2), |r| <= ln2/32: a 16-entry table of
// 2^(j/16) (one 128-byte shared-memory wavefront, so 32 lanes with 32 different arguments
// never conflict) times a degree-7 Taylor polynomial (truncation 1.2e-18), the exponent added
// as an integer.  ~11 FP64 instructions against ~28 for the library exp(); < 2 ulp.
// Valid for |x| < 700 (result stays a normal number); callers guarantee the range.
//
// log(x) = e ln2 - log(rcp_j) + log1p(m rcp_j - 1): j = top six mantissa bits, rcp_j ~ 1/c_j
// with c_j the midpoint of the j-th mantissa interval, |m rcp_j - 1| <= 2^-7, degree-7 series.
// Valid for normal positive x; ~13 FP64 instructions against ~45 for the library log().
__device__ double g_exp_table[16];
__device__ double g_log_table[128];   // [0..63] rcp_j, [64..127] -log(rcp_j)

struct MathConst {
    double inv_ln2_16, magic, ln2_16_hi, ln2_16_lo;   // exp range reduction
    double e7, e6, e5, e4, e3;                        // 1/5040 .. 1/6 (0.5 is an immediate)
    double ln2_hi, ln2_lo;                            // log: e * ln2 in two parts
    double l7, l6, l5, l4, l3;                        // 1/7, -1/6, 1/5, -1/4, 1/3
    double log_floor;                                 // log(1e-10), matching_utils.py:14
    double inv_two_var_intra, inv_denom_intra, inv_two_var_inter, inv_denom_inter, inv_ln10;
    double log_inv_denom_intra;                       // -log(sd_intra * sqrt(2 pi))
    double sd_ratio;                                  // sd_intra / sd_inter
    double log_sd_ratio;
    // log rho = qa d^2 + qb d + qc (bound_kernel): 4 qa, qb, qc, 1 / (2 qa)
    double prune_4a, prune_b, prune_c, prune_inv_2a;
};
__constant__ MathConst c_math;

__device__ __forceinline__ double exp_core(double x, const double *__restrict__ table) {
    const double kd = fma(x, c_math.inv_ln2_16, c_math.magic);
    const int ki = __double2loint(kd);
    const double kf = kd - c_math.magic;
    double r = fma(kf, c_math.ln2_16_hi, x);
    r = fma(kf, c_math.ln2_16_lo, r);
    double q = fma(r, c_math.e7, c_math.e6);
    q = fma(q, r, c_math.e5);
    q = fma(q, r, c_math.e4);
    q = fma(q, r, c_math.e3);
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);
    const double t = table[ki & 15];
    const double res = fma(t, em1, t);
    const int hi = __double2hiint(res) + ((ki & ~15) << 16);   // += (k >> 4) << 20
    return __hiloint2double(hi, __double2loint(res));
}

// exp(x) for x in (-746, -690]: the value is scaled up by 2^1000 (an exponent-field add, exact)
// and one multiplication by 2^-1000 then rounds it onto the denormal grid, the single rounding
// a correctly rounded exp() performs there.
__device__ __forceinline__ double exp_gradual(double x, const double *__restrict__ table) {
    const double kd = fma(x, c_math.inv_ln2_16, c_math.magic);
    const int ki = __double2loint(kd);
    const double kf = kd - c_math.magic;
    double r = fma(kf, c_math.ln2_16_hi, x);
    r = fma(kf, c_math.ln2_16_lo, r);
    double q = fma(r, c_math.e7, c_math.e6);
    q = fma(q, r, c_math.e5);
    q = fma(q, r, c_math.e4);
    q = fma(q, r, c_math.e3);
    q = fma(q, r, 0.5);
    const double em1 = fma(q, r * r, r);
    const double t = table[ki & 15];
    const double res = fma(t, em1, t);
    const int hi = __double2hiint(res) + (((ki >> 4) + 1000) << 20);
    return __hiloint2double(hi, __double2loint(res)) * 0x1p-1000;
}

__device__ __noinline__ double log_slow(double x) { return log(x); }

__device__ __forceinline__ double log_core(double x, const double *__restrict__ table) {
    int hi = __double2hiint(x);
    double bias = 0.0;
    if (hi < 0x00100000) {   // denormal (or 0 / negative, which go to the library)
        if (!(x > 0.0)) return log_slow(x);
        x *= 0x1p64;
        bias = -64.0;
        hi = __double2hiint(x);
    } else if (hi >= 0x7ff00000) {
        return log_slow(x);   // inf, nan
    }
    const int j = (hi >> 14) & 63;
    const double e = static_cast<double>((hi >> 20) - 1023) + bias;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double r = fma(m, table[j], -1.0);
    double q = fma(r, c_math.l7, c_math.l6);
    q = fma(q, r, c_math.l5);
    q = fma(q, r, c_math.l4);
    q = fma(q, r, c_math.l3);
    q = fma(q, r, -0.5);
    const double l1p = fma(q, r * r, r);
    const double head = fma(e, c_math.ln2_hi, table[64 + j]);
    return head + fma(e, c_math.ln2_lo, l1p);
}

// One sample of matching_utils.py:48-57: max(exp(c1*log(c2) - lgamma(c1+1) - c2), 1e-10).  The
// floor is applied to the exponent (exp is monotonic), so no exp argument leaves [-23.03, ~0].
__device__ __forceinline__ double poisson_exponent(double c1, double lg1, double c2,
                                                   double logc2) {
    const double x = __dsub_rn(__dsub_rn(__dmul_rn(c1, logc2), lg1), c2);
    return x < c_math.log_floor ? c_math.log_floor : x;
}
// sum += max(x, L) - L for x = c1*log(c2) - lgamma(c1+1) - c2 and L = log 1e-10, given
// m1 = -(lgamma(c1+1) + L): one FMA, one add, and the clamp as a test of the sign bit
// (x < L  <=>  x - L < 0), which keeps it off the FP64 pipe.  The FMA differs from the
// reference's rounding of x by an ulp of x (~4e-15), far inside the 1e-9 bar.
__device__ __forceinline__ void add_poisson_excess(double &sum, double c1, double m1, double c2,
                                                   double logc2) {
    const double d = fma(c1, logc2, m1) - c2;
    // sign-bit test + predicated add (the compiler would otherwise add and select twice)
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\t@p add.f64 %0, %0, %2;\n\t}"
        : "+d"(sum) : "r"(__double2hiint(d)), "d"(d));
}

__device__ __forceinline__ double poisson_term(double c1, double lg1, double c2, double logc2,
                                               const double *__restrict__ table) {
    return exp_core(poisson_exponent(c1, lg1, c2, logc2), table);
}

// prob_comp from the squared distance (matching_utils.py:21-39).  Fast path: table exp and
// reciprocal multiplies (a few ulp, far inside the 1e-9 bar).  When a Gaussian is about to
// underflow (exponent below -700) the reference's own operation sequence is used instead, so
// that the denormal / exact-zero results that decide prob_product > 0 come out the same way.
__device__ __forceinline__ double comp_prob_fast(double d2, const double *__restrict__ table) {
    const double d = sqrt(d2);
    const double dx = __dsub_rn(d, c_gauss.mu_inter);
    const double ai = -d2 * c_math.inv_two_var_intra;
    const double ae = -(dx * dx) * c_math.inv_two_var_inter;
    if (!(ae >= -700.0) || !(ai <= 0.0)) return comp_prob_dev(d);   // d > 1.3 or NaN
    // exp() underflows to exactly 0 below -745.14: prob_comp is then 0 / (0 + ge) = 0
    if (ai < -745.2) return 0.0;
    const double ge = exp_core(ae, table) * c_math.inv_denom_inter;
    double gi;
    if (ai >= -700.0) gi = exp_core(ai, table) * c_math.inv_denom_intra;
    else gi = __ddiv_rn(exp_gradual(ai, table), c_gauss.denom_intra);   // the reference's two roundings
    return __ddiv_rn(gi, gi + ge);
}

__device__ __noinline__ double exp_slow(double x) { return exp(x); }

// log(prob_comp), or NaN when prob_comp is exactly 0 (an underflowed intra Gaussian,
// d >~ 0.21).  Fast path: log gi is its exponent minus log(sd sqrt(2 pi)), so one table log of
// gi + ge replaces the division and the log of the quotient.
// log(prob_comp), or NaN when prob_comp is exactly 0 (an underflowed intra Gaussian,
// d >~ 0.21).  Fast path: log gi is its exponent minus log(sd sqrt(2 pi)), so one table log of
// gi + ge replaces the division and the log of the quotient.
__device__ __forceinline__ double log_comp_prob(double d2, const double *__restrict__ etab,
                                                const double *__restrict__ ltab,
                                                bool *domain_error) {
    const double kNone = __longlong_as_double(0x7ff8000000000000LL);
    const double d = sqrt(d2);
    const double dx = __dsub_rn(d, c_gauss.mu_inter);
    const double ai = -d2 * c_math.inv_two_var_intra;
    const double ae = -(dx * dx) * c_math.inv_two_var_inter;
    if (!(ae >= -700.0) || !(ai <= 0.0)) {   // d > 1.3 or NaN: the reference's own sequence
        const double pc = comp_prob_dev(d);
        if (pc != pc) *domain_error = true;
        if (!(pc > 0.0)) return kNone;
        return log_core(pc, ltab);
    }
    if (ai < -745.2) return kNone;           // exp() underflows to exactly 0
    const double ge = exp_core(ae, etab) * c_math.inv_denom_inter;
    if (ai >= -700.0) {
        const double gi = exp_core(ai, etab) * c_math.inv_denom_intra;
        return (ai + c_math.log_inv_denom_intra) - log_core(gi + ge, ltab);
    }
    // the intra Gaussian is (nearly) denormal: keep the reference's roundings
    const double gi = __ddiv_rn(exp_gradual(ai, etab), c_gauss.denom_intra);
    const double pc = __ddiv_rn(gi, gi + ge);
    if (!(pc > 0.0)) return kNone;
    return log_core(pc, ltab);
}

// log of a normal positive double (no range checks).
__device__ __forceinline__ double log_normal(double x, const double *__restrict__ table) {
    const int hi = __double2hiint(x);
    const int j = (hi >> 14) & 63;
    const double e = static_cast<double>((hi >> 20) - 1023);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double r = fma(m, table[j], -1.0);
    double q = fma(r, c_math.l7, c_math.l6);
    q = fma(q, r, c_math.l5);
    q = fma(q, r, c_math.l4);
    q = fma(q, r, c_math.l3);
    q = fma(q, r, -0.5);
    const double l1p = fma(q, r * r, r);
    const double head = fma(e, c_math.ln2_hi, table[64 + j]);
    return head + fma(e, c_math.ln2_lo, l1p);
}

// The lanes log_comp_prob_warp cannot take through its straight-line path: an intra
// Gaussian that underflows to 0 (NaN marker), is (nearly) denormal, or d > 1.3 / NaN.
__device__ __noinline__ double log_comp_prob_special(double d, double ai, double ae,
                                                      const double *etab, const double *ltab,
                                                      int32_t *status) {
    const double kNone = __longlong_as_double(0x7ff8000000000000LL);
    if (!(ae >= -700.0) || !(ai <= 0.0)) {   // the reference's own sequence
        const double pc = comp_prob_dev(d);
        if (pc != pc) atomicOr(status, 1);
        return (pc > 0.0) ? log_core(pc, ltab) : kNone;
    }
    if (ai < -745.2) return kNone;           // exp() underflows to exactly 0
    const double ge = exp_core(ae, etab) * c_math.inv_denom_inter;
    const double gi = __ddiv_rn(exp_gradual(ai, etab), c_gauss.denom_intra);
    const double pc = __ddiv_rn(gi, gi + ge);
    return (pc > 0.0) ? log_core(pc, ltab) : kNone;
}

// log(prob_comp) for a whole warp.  prob_comp = gi / (gi + ge) = 1 / (1 + rho) with
// rho = (sd_i / sd_e) exp(ae - ai), so log prob_comp = -log(1 + rho): one table exp and one
// table log.  ae - ai >= -1.96 (its minimum, at d = 0), so rho >= 0.021 and 1 + rho loses
// nothing; the fast path covers ai, ae >= -700 (rho <= 1e303).  The few lanes outside that range
// -- an intra Gaussian that is denormal or exactly 0, d > 1.3, NaN -- are fixed up under one
// warp-level branch with the reference's own operation sequence.  NaN = prob_comp is 0.
__device__ __forceinline__ double log_comp_prob_warp(double d2, const double *__restrict__ etab,
                                                     const double *__restrict__ ltab,
                                                     int32_t *status) {
    const double d = sqrt(d2);
    const double dx = d - c_gauss.mu_inter;
    const double ai = -d2 * c_math.inv_two_var_intra;
    const double ae = -(dx * dx) * c_math.inv_two_var_inter;
    const bool special = !(ai >= -700.0) || !(ae >= -700.0) || !(ai <= 0.0);
    // rho > 2^62 once ae - ai > 44: log(1 + rho) and log(rho) = log(sd_i / sd_e) + (ae - ai) are
    // then the same double.  That is every pair whose profiles are further apart than ~0.05;
    // a pass whose 32 lanes are all there skips the table exp and log.  A lane takes the
    // closed form whatever its neighbours do, so a pair's weight does not depend on the batch
    // it is scored in.
    const double diff = ae - ai;
    const bool big = !special && diff > 44.0;
    double out = -(diff + c_math.log_sd_ratio);
    if (!__all_sync(0xffffffffu, big || special)) {
        const double rho = exp_core((special || big) ? 0.0 : diff, etab) * c_math.sd_ratio;
        const double slow = -log_normal(1.0 + rho, ltab);
        if (!big) out = slow;
    }
    if (__any_sync(0xffffffffu, special)) {
        if (special) out = log_comp_prob_special(d, ai, ae, etab, ltab, status);
    }
    return out;
}

// ---- coverage terms --------------------------------------------------------------------
// status bit 2: a coverage value is NaN or infinite (the fast exp path assumes finite inputs).
__global__ void cov_terms_kernel(double *__restrict__ cov, double *__restrict__ logc,
                                 double *__restrict__ lgam, long long count,
                                 int32_t *__restrict__ status) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    double c = cov[i];
    if (!(fabs(c) <= DBL_MAX)) atomicOr(status, 2);
    if (c < 0.0001) c = 0.0001;   // feature_utils.py:152-153
    cov[i] = c;
    logc[i] = log(c);
    lgam[i] = lgamma_cpython(__dadd_rn(c, 1.0));
}

// ||row||^2 of a [rows,136] table.  It goes through the same DMMA k-step sequence as the u.v of
// pair_kernel / dist2_kernel, so that for two bitwise-equal rows |u|^2 + |v|^2 - 2 u.v is exactly
// 0, as the reference's difference form is: near d = 0 the inter-species Gaussian is linear in
// d = sqrt(d^2), and a d^2 of 1e-17 instead of 0 would already move prob_comp by 4e-9.
// One warp per 8 rows; the diagonal of the 8 x 8 product holds the norms.
// With `counts` given, the rows are first produced from the scan counts
// (prof = counts / max(1, windows), feature_utils.py:70, IEEE division) and written out.
__global__ void row_norms_kernel(const double *__restrict__ prof_in,
                                 const uint32_t *__restrict__ counts,
                                 const ContigMeta *__restrict__ meta, double *__restrict__ prof_out,
                                 long long rows, double *__restrict__ norm) {
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long row = warp * 8 + g;
    if (warp * 8 >= rows) return;
    const long long r = row < rows ? row : rows - 1;
    double denom = 1.0;
    if (counts) {
        const int nwin = meta[r].nwin;
        denom = static_cast<double>(nwin > 1 ? nwin : 1);
    }
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < MCG_NPROF; k0 += 4) {
        double a;
        if (counts) {
            a = __ddiv_rn(static_cast<double>(counts[r * MCG_NPROF + k0 + t]), denom);
            if (row < rows) prof_out[r * MCG_NPROF + k0 + t] = a;
        } else {
            a = prof_in[r * MCG_NPROF + k0 + t];
        }
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(a));
    }
    // lane (g, t) holds C[g][2t] and C[g][2t+1]
    if (row < rows) {
        if (g == 2 * t) norm[row] = c0;
        if (g == 2 * t + 1) norm[row] = c1;
    }
}

// ---- the pair kernel ---------------------------------------------------------------------
// A CTA (8 warps) owns 64 "A" rows (contigs to score) and walks the "B" rows (bin centroids, or
// member contigs for rules A/B) in tiles of 64.
//
//   distance  u.v for the 64 x 64 pairs as an FP64 GEMM on mma.sync.m8n8k4 (DMMA): the FP64 units
//             deliver the same FLOP rate through DMMA as through DFMA (measured 37 TFLOP/s
//             either way, and no more when mixed), but one DMMA issues 256 FMAs, so the
//             contraction stops being issue / shared-load bound.  Warp w holds the 16 x 32
//             block (rows 16*(w/2).., columns 32*(w%2)..) as 2 x 4 accumulator fragments;
//             both operand tiles sit in shared memory as [row][k] with a pitch of 36 doubles,
//             which makes every fragment load hit 16 distinct bank pairs twice (the minimum
//             for 256 bytes).  k is staged in chunks of 32 with 16-byte cp.async, two chunks
//             in flight, and the first chunk of the next B tile flies during phase 2.
//             d^2 = |u|^2 + |v|^2 - 2 u.v from precomputed norms.
//   phase 2   every lane then owns 2 rows x 8 columns (its accumulator fragments) and runs the
//             Poisson terms over the samples for 4 pairs at a time, then the epilogue.
//
// 112 KB of shared memory per CTA and <= 128 registers: two CTAs share an SM.
constexpr int kKC = 24;                       // k-chunk (6 DMMA k-steps); 136 = 5 * 24 + 16
constexpr int kPitch = kKC + 4;               // pitch = 12 mod 16 doubles: conflict-free fragments
constexpr int kChunkDoubles = kTile * kPitch; // one side of one stage
constexpr int kCovDoubles = kSampleChunk * 3 * kTile;   // one side's coverage terms
constexpr int kPairSmemDoubles = 4 * kChunkDoubles + 3 * kCovDoubles + 3 * kTile + 16 + 128 +
                                 2 * kTile;
constexpr size_t kPairSmem = sizeof(double) * kPairSmemDoubles + sizeof(long long) * 3 * kTile +
                             sizeof(int) * 2 * kTile;

__device__ __forceinline__ void cp_async16(double *dst, const double *src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(double *dst, const double *src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// dst[r][kk] <- prof[row(r)][k0 + kk], kk < kc, asynchronously, 16 bytes a time: four
// threads per row, thread (r, p0) takes pieces p0, p0 + 4, ...  Rows past the end are zero.
__device__ __forceinline__ void stage_prof_chunk(double *dst, const double *__restrict__ prof,
                                                 const long long *rows, int k0, int kc) {
    const int r = threadIdx.x >> 2;
    const long long row = rows[r];
    double *d = dst + r * kPitch;
    const double *src = prof + row * MCG_NPROF + k0;
    for (int kk = 2 * (threadIdx.x & 3); kk < kc; kk += 8) {
        if (row >= 0) cp_async16(d + kk, src + kk);
        else { d[kk] = 0.0; d[kk + 1] = 0.0; }
    }
}

// dst[q][term][r] for samples [s0, s0+sc); rows past the end get c = 1 (log 0, lgamma 0).
__device__ __forceinline__ void load_cov_chunk(double *dst, const ScoreSide &s,
                                               const long long *rows, int S, int s0, int sc) {
    for (int idx = threadIdx.x; idx < kTile * sc; idx += kPairThreads) {
        const int r = idx / sc;
        const int q = idx - r * sc;
        double c = 1.0, lc = 0.0, lg = 0.0;
        const long long row = rows[r];
        if (row >= 0) {
            const long long off = row * S + s0 + q;
            c = s.cov[off];
            lc = s.logc[off];
            lg = s.lgam[off];
        }
        dst[(q * 3 + 0) * kTile + r] = c;
        dst[(q * 3 + 1) * kTile + r] = lc;
        dst[(q * 3 + 2) * kTile + r] = lg;
    }
}

// The same for the B side of the *next* tile, without stalling: 8-byte cp.async, four
// threads per row.
__device__ __forceinline__ void stage_cov_async(double *dst, const ScoreSide &s,
                                                const long long *rows, int S) {
    const int r = threadIdx.x >> 2;
    const long long row = rows[r];
    for (int q = threadIdx.x & 3; q < S; q += 4) {
        double *d = dst + q * 3 * kTile + r;
        if (row >= 0) {
            const long long off = row * S + q;
            cp_async8(d, s.cov + off);
            cp_async8(d + kTile, s.logc + off);
            cp_async8(d + 2 * kTile, s.lgam + off);
        } else {
            d[0] = 1.0; d[kTile] = 0.0; d[2 * kTile] = 0.0;
        }
    }
}

__device__ __forceinline__ void tile_rows(long long *rows, double *norms, const ScoreSide &s,
                                          long long r0, bool want_norm) {
    const int tid = threadIdx.x;
    if (tid < kTile) {
        long long row = -1;
        if (r0 + tid < s.n) row = s.index ? s.index[s.first + r0 + tid] : (s.first + r0 + tid);
        rows[tid] = row;
        norms[tid] = (want_norm && row >= 0) ? s.norm[row] : 0.0;
    }
}

// RULE_C: rule-C weights / argmin (label_prop_utils.py:308-345), else raw per-pair outputs.
// LOGSUM (S <= 30): the product of the floored Poisson terms is carried as the sum of their
// floored exponents, prod_s max(e^x_s, 1e-10) = exp(sum_s max(x_s, log 1e-10)): one exp per
// pair instead of 2 S.  With S <= 30 the sum stays above -691, so neither running product can
// underflow in the reference either and the two forms agree to ~S ulp.  For S > 30 the
// reference's products reach the denormal range, where its sequential rounding (and the
// exact zero that decides prob_product > 0) can only be reproduced by multiplying in order.
template <bool RULE_C, bool LOGSUM>
__global__ void __launch_bounds__(kPairThreads, 2)
pair_kernel(ScoreSide A, ScoreSide B, int S, ScoreOut out, int32_t *__restrict__ status,
            const int *__restrict__ n_rows_dev, long long part_stride) {
    // a query count that only the device knows (the redo list of the guarded rule-C path)
    if (n_rows_dev) {
        const long long nn = *n_rows_dev;
        if (nn < A.n) A.n = nn;
        if (static_cast<long long>(blockIdx.x) * kTile >= A.n) return;
    }
    extern __shared__ __align__(16) double sm[];
    double *stage = sm;                                   // [2 stages][A, B][64][36]
    double *CA = stage + 4 * kChunkDoubles;
    double *CB2 = CA + kCovDoubles;                       // [2] B-side terms, double-buffered
    double *NA = CB2 + 2 * kCovDoubles;
    double *NB = NA + kTile;                              // [2][64]
    double *ET = NB + 2 * kTile;                          // 16
    double *LT = ET + 16;                                 // 128
    double *best_w_s = LT + 128;                          // [2][64] cross-warp argmin
    long long *rowA = reinterpret_cast<long long *>(best_w_s + 2 * kTile);
    long long *rowB = rowA + kTile;                       // [2][64]
    int *best_i_s = reinterpret_cast<int *>(rowB + 2 * kTile);   // [2][64]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    const int row0 = 16 * wr + g;          // this lane's rows: row0, row0 + 8
    const int col0 = 32 * wc + 2 * t;      // this lane's columns: col0 + 8 j + {0, 1}
    const long long a0 = static_cast<long long>(blockIdx.x) * kTile;
    const long long n_btiles = (B.n + kTile - 1) / kTile;
    const bool single_chunk = S <= kSampleChunk;
    constexpr int kChunks = (MCG_NPROF + kKC - 1) / kKC;   // 6: five of 24 and one of 16

    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    tile_rows(rowA, NA, A, a0, true);
    long long bt = blockIdx.y;
    if (bt < n_btiles) tile_rows(rowB, NB, B, bt * kTile, true);
    __syncthreads();
    if (single_chunk) load_cov_chunk(CA, A, rowA, S, 0, S);
    if (bt < n_btiles) {   // chunk 0 and the coverage terms of the first tile
        stage_prof_chunk(stage, A.prof, rowA, 0, kKC);
        stage_prof_chunk(stage + kChunkDoubles, B.prof, rowB, 0, kKC);
        if (single_chunk) stage_cov_async(CB2, B, rowB, S);
    }
    cp_async_commit();

    double best_w[2];
    int best_i[2];
    best_w[0] = best_w[1] = 0.0;
    best_i[0] = best_i[1] = -1;
    bool domain_error = false;

    for (int it = 0; bt < n_btiles; bt += gridDim.y, ++it) {
        const int cur = it & 1;
        const long long b0 = bt * kTile;
        const long long *rows_b = rowB + cur * kTile;
        const double *norm_b = NB + cur * kTile;
        const long long bt_next = bt + gridDim.y;
        // 8-column blocks of this warp that hold real B rows (warp-uniform)
        const long long left = B.n - b0 - 32 * wc;
        const int njv = left <= 0 ? 0 : (left >= 32 ? 4 : static_cast<int>((left + 7) >> 3));
        // rows of the next tile (its chunk 0 is staged below, during this tile's phase 2)
        if (bt_next < n_btiles) tile_rows(rowB + (cur ^ 1) * kTile, NB + (cur ^ 1) * kTile, B,
                                          bt_next * kTile, true);
        double *CB = CB2 + (single_chunk ? cur : 0) * kCovDoubles;

        // ---- u.v on the FP64 tensor path (matching_utils.py:28-29), k in five chunks ------
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
            if (c + 1 < kChunks) {
                double *nxt = stage + ((c + 1) & 1) * 2 * kChunkDoubles;
                const int kc = (c + 1 == kChunks - 1) ? MCG_NPROF - (kChunks - 1) * kKC : kKC;
                stage_prof_chunk(nxt, A.prof, rowA, (c + 1) * kKC, kc);
                stage_prof_chunk(nxt + kChunkDoubles, B.prof, rows_b, (c + 1) * kKC, kc);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();   // chunk c has landed for everyone
            const double *Ak = stage + (c & 1) * 2 * kChunkDoubles + row0 * kPitch + t;
            const double *Bk = stage + (c & 1) * 2 * kChunkDoubles + kChunkDoubles +
                               (32 * wc + g) * kPitch + t;
            const int ksteps = (c == kChunks - 1) ? (MCG_NPROF - (kChunks - 1) * kKC) / 4 : kKC / 4;
            if (njv > 0) {
#pragma unroll 2
                for (int ks = 0; ks < ksteps; ++ks) {
                    const double a_lo = Ak[4 * ks];
                    const double a_hi = Ak[8 * kPitch + 4 * ks];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < njv) {
                            const double b = Bk[8 * j * kPitch + 4 * ks];
                            dmma(acc[0][j][0], acc[0][j][1], a_lo, b);
                            dmma(acc[1][j][0], acc[1][j][1], a_hi, b);
                        }
                    }
                }
            }
            __syncthreads();   // stage c & 1 may be refilled
        }
        if (bt_next < n_btiles) {   // chunk 0 of the next tile flies during phase 2
            stage_prof_chunk(stage, A.prof, rowA, 0, kKC);
            stage_prof_chunk(stage + kChunkDoubles, B.prof, rowB + (cur ^ 1) * kTile, 0, kKC);
            if (single_chunk)
                stage_cov_async(CB2 + (cur ^ 1) * kCovDoubles, B, rowB + (cur ^ 1) * kTile, S);
        }
        cp_async_commit();
        // d^2 = |u|^2 + |v|^2 - 2 u.v, clamped at 0
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double d2 = fma(-2.0, acc[i][j][h],
                                          NA[row0 + 8 * i] + norm_b[col0 + 8 * j + h]);
                    acc[i][j][h] = d2 > 0.0 ? d2 : 0.0;
                }

        // ---- 4 pairs at a time (2 rows x 2 adjacent columns): Poisson terms, epilogue ------
        // with several sample chunks the loop holds barriers, so every warp runs the trip
        // count of the fullest warp (wc = 0) and idles through the blocks it does not have
        const long long left0 = B.n - b0;
        const int nj_loop = single_chunk ? njv
                                         : (left0 >= 32 ? 4 : static_cast<int>((left0 + 7) >> 3));
#pragma unroll 1
        for (int j = 0; j < nj_loop; ++j) {
            const bool live = j < njv;
            const int col = col0 + 8 * j;
            double p1[4], p2[4];   // pair index = 2 * i + h
#pragma unroll
            for (int e = 0; e < 4; ++e) { p1[e] = LOGSUM ? 0.0 : 1.0; p2[e] = p1[e]; }
            for (int s0 = 0; s0 < S; s0 += kSampleChunk) {
                const int sc = min(kSampleChunk, S - s0);
                if (!single_chunk) {
                    __syncthreads();
                    load_cov_chunk(CA, A, rowA, S, s0, sc);
                    load_cov_chunk(CB, B, rows_b, S, s0, sc);
                    __syncthreads();
                }
                if (!live) continue;
#pragma unroll 2
                for (int q = 0; q < sc; ++q) {
                    const double *ca = CA + q * 3 * kTile + row0;
                    const double2 bc = *reinterpret_cast<const double2 *>(CB + q * 3 * kTile + col);
                    const double2 bl =
                        *reinterpret_cast<const double2 *>(CB + (q * 3 + 1) * kTile + col);
                    const double2 bg =
                        *reinterpret_cast<const double2 *>(CB + (q * 3 + 2) * kTile + col);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double ac = ca[8 * i];
                        const double al = ca[kTile + 8 * i];
                        const double ag = ca[2 * kTile + 8 * i];
                        // cov1 = the A row, cov2 = the B row (matching_utils.py:48-54)
                        if (LOGSUM) {
                            p1[2 * i] += poisson_exponent(ac, ag, bc.x, bl.x);
                            p2[2 * i] += poisson_exponent(bc.x, bg.x, ac, al);
                            p1[2 * i + 1] += poisson_exponent(ac, ag, bc.y, bl.y);
                            p2[2 * i + 1] += poisson_exponent(bc.y, bg.y, ac, al);
                        } else {
                            p1[2 * i] = __dmul_rn(p1[2 * i], poisson_term(ac, ag, bc.x, bl.x, ET));
                            p2[2 * i] = __dmul_rn(p2[2 * i], poisson_term(bc.x, bg.x, ac, al, ET));
                            p1[2 * i + 1] =
                                __dmul_rn(p1[2 * i + 1], poisson_term(ac, ag, bc.y, bl.y, ET));
                            p2[2 * i + 1] =
                                __dmul_rn(p2[2 * i + 1], poisson_term(bc.y, bg.y, ac, al, ET));
                        }
                    }
                }
            }
            if (!live) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = e >> 1, h = e & 1;
                const long long ar = a0 + row0 + 8 * i;
                const long long bc = b0 + col + h;
                // j is a runtime index: pick acc[i][j][h] without making acc[] addressable
                const double d2 = j == 0 ? acc[i][0][h] : j == 1 ? acc[i][1][h]
                                : j == 2 ? acc[i][2][h] : acc[i][3][h];
                if (ar >= A.n || bc >= B.n) continue;
                const double pmin = p1[e] < p2[e] ? p1[e] : p2[e];   // log or value of prob_cov
                double lp = MCG_MAX_WEIGHT, pc = 0.0, pv = 0.0;
                bool valid = false;
                if (RULE_C && LOGSUM) {
                    // only -(log10 pc + log10 pv) is needed: log pc = ai - log(sd sqrt(2 pi)) -
                    // log(gi + ge) avoids the division, log pv is the exponent sum itself
                    const double log_pc = log_comp_prob(d2, ET, LT, &domain_error);
                    if (log_pc == log_pc) {   // NaN = prob_comp is exactly 0
                        const double tsum = log_pc + pmin;
                        // prob_comp * prob_cov > 0.0 in double arithmetic: certain above
                        // e^-744, impossible below e^-746 (2^-1075 = e^-745.13)
                        if (tsum > -744.0) valid = true;
                        else if (tsum >= -746.0)
                            valid = __dmul_rn(exp_slow(log_pc), exp_slow(pmin)) > 0.0;
                        if (valid) lp = -tsum * c_math.inv_ln10;
                    }
                } else {
                    pc = comp_prob_fast(d2, ET);
                    pv = LOGSUM ? exp_core(pmin, ET) : pmin;
                    if (pc != pc) domain_error = true;
                    valid = __dmul_rn(pc, pv) > 0.0;
                    if (valid) {
                        // -(math.log(pc, 10) + math.log(pv, 10)): CPython's two-argument log is
                        // log(x) / log(10) per term, so two divisions, then the sum (rules A / B
                        // feed hard comparisons: <= w_intra, > w_inter, first-minimum argmin)
                        const double log_pv = LOGSUM ? pmin : log_core(pv, LT);
                        lp = -(__ddiv_rn(log_core(pc, LT), c_gauss.ln10) +
                               __ddiv_rn(log_pv, c_gauss.ln10));
                    }
                }
                const long long o = ar * B.n + bc;
                if (RULE_C) {
                    // label_prop_utils.py:331-337: log_prob stays 0 unless the product is
                    // positive, and an exact 0 becomes MAX_WEIGHT
                    const double w = (valid && lp != 0.0) ? lp : MCG_MAX_WEIGHT;
                    if (out.weights) out.weights[o] = w;
                    const int bi = static_cast<int>(bc);
                    if (best_i[i] < 0 || w < best_w[i] || (w == best_w[i] && bi < best_i[i])) {
                        best_w[i] = w;
                        best_i[i] = bi;
                    }
                } else {
                    if (out.pc) out.pc[o] = pc;
                    if (out.pv) out.pv[o] = pv;
                    if (out.lp) out.lp[o] = lp;
                }
            }
        }
        __syncthreads();   // CB / the other rowB slot are rewritten at the top of the next tile
    }
    cp_async_wait<0>();

    if (domain_error) atomicOr(status, 1);

    if (RULE_C) {
        // first index of the minimum: over the 4 lanes of a row, then over the two warps
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double w = best_w[i];
            int bi = best_i[i];
#pragma unroll
            for (int m = 2; m >= 1; m >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, w, m);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
                if (oi >= 0 && (bi < 0 || ow < w || (ow == w && oi < bi))) { w = ow; bi = oi; }
            }
            if (t == 0) {
                best_w_s[wc * kTile + row0 + 8 * i] = w;
                best_i_s[wc * kTile + row0 + 8 * i] = bi;
            }
        }
        __syncthreads();
        if (tid < kTile && a0 + tid < A.n) {
            double w = best_w_s[tid];
            int bi = best_i_s[tid];
            const double ow = best_w_s[kTile + tid];
            const int oi = best_i_s[kTile + tid];
            if (oi >= 0 && (bi < 0 || ow < w || (ow == w && oi < bi))) { w = ow; bi = oi; }
            // label_prop_utils.py:342-345: None when the minimum is MAX_WEIGHT
            const bool none = (bi < 0) || (w == MCG_MAX_WEIGHT);
            // part_stride > 0: the B tiles are spread over grid.y and every slice leaves its
            // own (argmin, minimum) for scatter_redo_kernel to combine
            const long long o = a0 + tid + part_stride * blockIdx.y;
            if (out.argmin) out.argmin[o] = none ? -1 : bi;
            if (out.wmin) out.wmin[o] = none ? MCG_MAX_WEIGHT : w;
        }
    }
}

// ---- rule C at scale: distance GEMM + probability kernel ---------------------------------------
// For the big rule-C batches (every long contig against every bin centroid, S <= 30) the work
// is split in two kernels with different shapes:
//   dist2_kernel  the d^2 matrix by the DMMA path of pair_kernel (same staging, same fragment
//                 layout), written to a scratch matrix in HBM (8 bytes per pair, read back once);
//   prob_kernel   one warp = one contig against 32 bins per pass.  The contig's coverage terms
//                 are warp-uniform (broadcast reads), the bins' terms sit in shared memory for
//                 the whole kernel, there is no barrier after the prologue and the kernel needs
//                 <= 64 registers, so 32 warps per SM hide the FP64 latencies that the
//                 two-CTA fused kernel exposes.
// The arithmetic per pair is the same as pair_kernel<RULE_C, LOGSUM>.
template <int DUMMY>
__global__ void __launch_bounds__(kPairThreads, 3)
dist2_kernel(ScoreSide A, ScoreSide B, double *__restrict__ d2out, long long ldc,
             unsigned long long *__restrict__ tilemin, long long n_btiles_all,
             const int *__restrict__ gate_total, int gate_cap) {
    // behind the tensor-core bound this kernel only runs when the candidate list overflowed and
    // prob_kernel has to score every pair after all
    if (gate_total && *gate_total <= gate_cap) return;
    extern __shared__ __align__(16) double sm[];
    double *stage = sm;                                   // [2 stages][A, B][64][pitch]
    double *NA = stage + 4 * kChunkDoubles;
    double *NB = NA + kTile;                              // [2][64]
    long long *rowA = reinterpret_cast<long long *>(NB + 2 * kTile);
    long long *rowB = rowA + kTile;                       // [2][64]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    const int row0 = 16 * wr + g;
    const int col0 = 32 * wc + 2 * t;
    const long long a0 = static_cast<long long>(blockIdx.x) * kTile;
    const long long n_btiles = (B.n + kTile - 1) / kTile;
    constexpr int kChunks = (MCG_NPROF + kKC - 1) / kKC;

    tile_rows(rowA, NA, A, a0, true);
    long long bt = blockIdx.y;
    if (bt < n_btiles) tile_rows(rowB, NB, B, bt * kTile, true);
    __syncthreads();
    if (bt < n_btiles) {
        stage_prof_chunk(stage, A.prof, rowA, 0, kKC);
        stage_prof_chunk(stage + kChunkDoubles, B.prof, rowB, 0, kKC);
    }
    cp_async_commit();

    for (int it = 0; bt < n_btiles; bt += gridDim.y, ++it) {
        const int cur = it & 1;
        const long long b0 = bt * kTile;
        const long long *rows_b = rowB + cur * kTile;
        const double *norm_b = NB + cur * kTile;
        const long long bt_next = bt + gridDim.y;
        const long long left = B.n - b0 - 32 * wc;
        const int njv = left <= 0 ? 0 : (left >= 32 ? 4 : static_cast<int>((left + 7) >> 3));
        if (bt_next < n_btiles) tile_rows(rowB + (cur ^ 1) * kTile, NB + (cur ^ 1) * kTile, B,
                                          bt_next * kTile, true);
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
            if (c + 1 < kChunks) {
                double *nxt = stage + ((c + 1) & 1) * 2 * kChunkDoubles;
                const int kc = (c + 1 == kChunks - 1) ? MCG_NPROF - (kChunks - 1) * kKC : kKC;
                stage_prof_chunk(nxt, A.prof, rowA, (c + 1) * kKC, kc);
                stage_prof_chunk(nxt + kChunkDoubles, B.prof, rows_b, (c + 1) * kKC, kc);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const double *Ak = stage + (c & 1) * 2 * kChunkDoubles + row0 * kPitch + t;
            const double *Bk = stage + (c & 1) * 2 * kChunkDoubles + kChunkDoubles +
                               (32 * wc + g) * kPitch + t;
            const int ksteps = (c == kChunks - 1) ? (MCG_NPROF - (kChunks - 1) * kKC) / 4 : kKC / 4;
            if (njv == 4 && ksteps == kKC / 4) {
                // the common case, fully unrolled: all fragment loads of the chunk can be in
                // flight ahead of the DMMAs
#pragma unroll
                for (int ks = 0; ks < kKC / 4; ++ks) {
                    const double a_lo = Ak[4 * ks];
                    const double a_hi = Ak[8 * kPitch + 4 * ks];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double b = Bk[8 * j * kPitch + 4 * ks];
                        dmma(acc[0][j][0], acc[0][j][1], a_lo, b);
                        dmma(acc[1][j][0], acc[1][j][1], a_hi, b);
                    }
                }
            } else if (njv > 0) {
#pragma unroll 2
                for (int ks = 0; ks < ksteps; ++ks) {
                    const double a_lo = Ak[4 * ks];
                    const double a_hi = Ak[8 * kPitch + 4 * ks];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < njv) {
                            const double b = Bk[8 * j * kPitch + 4 * ks];
                            dmma(acc[0][j][0], acc[0][j][1], a_lo, b);
                            dmma(acc[1][j][0], acc[1][j][1], a_hi, b);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (bt_next < n_btiles) {
            stage_prof_chunk(stage, A.prof, rowA, 0, kKC);
            stage_prof_chunk(stage + kChunkDoubles, B.prof, rowB + (cur ^ 1) * kTile, 0, kKC);
        }
        cp_async_commit();
        // d^2 = |u|^2 + |v|^2 - 2 u.v, clamped at 0; the scratch matrix is padded to whole tiles
        unsigned long long rmin[2] = {~0ull, ~0ull};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < njv) {
                    double2 v;
                    const double n_a = NA[row0 + 8 * i];
                    v.x = fma(-2.0, acc[i][j][0], n_a + norm_b[col0 + 8 * j]);
                    v.y = fma(-2.0, acc[i][j][1], n_a + norm_b[col0 + 8 * j + 1]);
                    v.x = v.x > 0.0 ? v.x : 0.0;
                    v.y = v.y > 0.0 ? v.y : 0.0;
                    *reinterpret_cast<double2 *>(d2out + (a0 + row0 + 8 * i) * ldc + b0 + col0 +
                                                 8 * j) = v;
                    if (tilemin) {
                        // smallest d^2 of the row in this 64-bin tile, as ordered bits (d^2 >= 0;
                        // a NaN counts as 0 so that its tile is never skipped); real bins only
                        const long long c = b0 + col0 + 8 * j;
                        const unsigned long long kx =
                            v.x == v.x ? static_cast<unsigned long long>(__double_as_longlong(v.x)) : 0ull;
                        const unsigned long long ky =
                            v.y == v.y ? static_cast<unsigned long long>(__double_as_longlong(v.y)) : 0ull;
                        if (c < B.n && kx < rmin[i]) rmin[i] = kx;
                        if (c + 1 < B.n && ky < rmin[i]) rmin[i] = ky;
                    }
                }
            }
        if (tilemin) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                unsigned long long k = rmin[i];
#pragma unroll
                for (int m = 1; m <= 2; m <<= 1) {   // the four lanes of a row
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
                    k = o < k ? o : k;
                }
                if (t == 0 && k != ~0ull && a0 + row0 + 8 * i < A.n)
                    atomicMin(tilemin + (a0 + row0 + 8 * i) * n_btiles_all + bt, k);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}
constexpr size_t kDistSmem = sizeof(double) * (4 * kChunkDoubles + 3 * kTile) +
                             sizeof(long long) * 3 * kTile;

constexpr int kProbThreads = 256;
// prob_kernel keeps one row of 3 S doubles per bin, padded to an odd pitch so that the 32 lanes
// of a warp (32 different bins, same sample) hit 16 distinct bank pairs twice
static inline int bin_pitch(int S) { return (3 * S) | 1; }
constexpr int kProbWarps = kProbThreads / 32;

constexpr int kProbRows = 2;   // contigs per warp iteration: the bins' terms are loaded once for both

// Redo list of the guarded (S > 30) path: contigs whose answer may depend on a pair whose
// Poisson product leaves the normal range; they are re-scored by the in-order pair_kernel.
struct RedoList {
    int *count;           // [1]
    long long *rows;      // table row of the contig
    long long *pos;       // where its result goes (index into ScoreOut::argmin / wmin)
};
// Gate between the two ways to finish rule C: the candidate kernels run when the list
// bound_kernel wrote holds every candidate (total <= cap), prob_kernel over all pairs otherwise.
struct CandGate {
    const int *total;   // nullptr: no candidate list, prob_kernel always runs
    int cap;
};
constexpr double kGuardLog = -680.0;      // log prob_cov below this: the pair is deferred
constexpr double kGuardWeight = 295.0;    // 295 ln 10 = 679.3 < 680: no deferred pair can beat it
constexpr int kDeferBit = 1 << 30;

// bins [g0, g0 + ncols) of the B side: BT[col][pitch] = (c, log c, m) x S; warp buffers AT.
//
// GUARD (S > 30).  The log-sum form is exact only while the reference's running products stay
// normal.  A pair whose smaller log product is below kGuardLog is *deferred*: its weight is
// either MAX_WEIGHT or above 680 / ln 10 = 295.3, so it cannot be the argmin of a contig that
// has a non-deferred weight below kGuardWeight.  Only contigs without such a weight go to the
// redo list.
//
// WARPS per CTA: 8, or 16 for sample counts whose terms leave room for one CTA per SM only
// (S ~ 100): twice the warps next to the same shared bin terms.
template <bool GUARD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (GUARD || WARPS > 8) ? 1 : 4)
prob_kernel(ScoreSide A, ScoreSide B, int S, const double *__restrict__ d2, long long ldc,
            long long g0, int ncols, int ncols_pad, int pitch, ScoreOut out,
            int32_t *__restrict__ status, RedoList redo, CandGate gate) {
    // with a candidate list (bound_kernel) this kernel only runs when the list overflowed
    if (gate.total && *gate.total <= gate.cap) return;
    extern __shared__ __align__(16) double sm[];
    double *ET = sm;                                  // 16
    double *LT = ET + 16;                             // 128
    double *AT = LT + 128;                            // [8 warps][2 rows][S][4]
    const int at_row = 4 * S;
    double *BT = AT + WARPS * kProbRows * at_row;   // [ncols][pitch]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    for (int idx = tid; idx < ncols * S; idx += WARPS * 32) {
        const int c = idx / S, q = idx - c * S;
        const long long row = B.index ? B.index[B.first + g0 + c] : (B.first + g0 + c);
        const long long off = row * S + q;
        const double bc = B.cov[off], bl = B.logc[off];
        const double bm = -(B.lgam[off] + c_math.log_floor);
        BT[c * pitch + 3 * q + 0] = bc;
        BT[c * pitch + 3 * q + 1] = bl;
        BT[c * pitch + 3 * q + 2] = bm;
    }
    __syncthreads();

    double *at = AT + warp * kProbRows * at_row;
    const long long n_warps = static_cast<long long>(gridDim.x) * WARPS;
    const long long n_iter = (A.n + kProbRows - 1) / kProbRows;
    const bool last_group = g0 + ncols >= B.n;
    for (long long it = static_cast<long long>(blockIdx.x) * WARPS + warp; it < n_iter;
         it += n_warps) {
        const long long r0 = it * kProbRows;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kProbRows; ++k) {
            // a missing second row repeats the first one; its results are not stored
            const long long r = r0 + k < A.n ? r0 + k : r0;
            const long long row = A.index ? A.index[A.first + r] : (A.first + r);
            for (int q = lane; q < S; q += 32) {
                const long long off = row * S + q;
                at[k * at_row + 4 * q + 0] = A.cov[off];
                at[k * at_row + 4 * q + 1] = A.logc[off];
                at[k * at_row + 4 * q + 2] = -(A.lgam[off] + c_math.log_floor);
            }
        }
        __syncwarp();
        double best_w[kProbRows];
        int best_i[kProbRows];
        const double *d2row[kProbRows];
        double dd_next[kProbRows];
        bool deferred[kProbRows];
#pragma unroll
        for (int k = 0; k < kProbRows; ++k) {
            const long long r = r0 + k < A.n ? r0 + k : r0;
            best_w[k] = 0.0;
            best_i[k] = -1;
            deferred[k] = false;
            if (g0 > 0) {   // continue from the previous bin group
                best_w[k] = out.wmin[A.first + r];
                int v = out.argmin[A.first + r];
                if (GUARD) {   // (argmin + 1) | deferred << 30
                    deferred[k] = (v & kDeferBit) != 0;
                    v = (v & (kDeferBit - 1)) - 1;
                }
                best_i[k] = v;
            }
            d2row[k] = d2 + r * ldc + g0;
            dd_next[k] = d2row[k][lane < ncols ? lane : 0];
        }
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            const int col = c0 + lane;
            const bool live = col < ncols;
            const int nxt = col + 32 < ncols ? col + 32 : 0;
            double log_pc[kProbRows];
            bool any_pc = false;
#pragma unroll
            for (int k = 0; k < kProbRows; ++k) {
                const double dd = dd_next[k];
                dd_next[k] = d2row[k][nxt];   // the next pass's distance
                log_pc[k] = log_comp_prob_warp(dd, ET, LT, status);
                any_pc = any_pc || (log_pc[k] == log_pc[k]);   // NaN: prob_comp is exactly 0
            }
            double tsum[kProbRows];
            bool defer_pair[kProbRows];
#pragma unroll
            for (int k = 0; k < kProbRows; ++k) { tsum[k] = log_pc[k]; defer_pair[k] = false; }
            // when prob_comp is 0 for every pair of the pass the coverage terms are not needed
            if (__any_sync(0xffffffffu, any_pc)) {
                double s1[kProbRows], s2[kProbRows];
#pragma unroll
                for (int k = 0; k < kProbRows; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
                const double *bt = BT + (live ? col : 0) * pitch;
                const double *ap = at;
#pragma unroll 5
                for (int q = 0; q < S; ++q) {
                    const double bc = bt[0], bl = bt[1], bm = bt[2];
#pragma unroll
                    for (int k = 0; k < kProbRows; ++k) {
                        const double ac = ap[k * at_row], al = ap[k * at_row + 1],
                                     am = ap[k * at_row + 2];
                        add_poisson_excess(s1[k], ac, am, bc, bl);
                        add_poisson_excess(s2[k], bc, bm, ac, al);
                    }
                    ap += 4;
                    bt += 3;
                }
#pragma unroll
                for (int k = 0; k < kProbRows; ++k) {
                    // log prob_cov = min over the two directions of sum_s max(x_s, L)
                    const double pmin = fma(static_cast<double>(S), c_math.log_floor,
                                            s1[k] < s2[k] ? s1[k] : s2[k]);
                    tsum[k] = log_pc[k] + pmin;   // stays NaN when prob_comp is 0
                    if (GUARD) defer_pair[k] = live && pmin < kGuardLog && log_pc[k] == log_pc[k];
                }
            }
#pragma unroll
            for (int k = 0; k < kProbRows; ++k) {
                // prob_comp * prob_cov > 0.0 in double arithmetic: certain above e^-744,
                // impossible below e^-746 (2^-1075 = e^-745.13); in between, multiply
                bool valid = tsum[k] > -744.0;
                const bool edge = !valid && tsum[k] >= -746.0;
                if (__any_sync(0xffffffffu, edge)) {
                    if (edge) valid = __dmul_rn(exp_slow(log_pc[k]),
                                                exp_slow(tsum[k] - log_pc[k])) > 0.0;
                }
                const double lp = -tsum[k] * c_math.inv_ln10;
                const double w = (valid && lp != 0.0) ? lp : MCG_MAX_WEIGHT;
                if (GUARD && defer_pair[k]) {
                    deferred[k] = true;   // not a candidate here; see kGuardWeight
                } else if (live) {
                    const int bi = static_cast<int>(g0) + col;
                    if (out.weights && r0 + k < A.n)
                        out.weights[(A.first + r0 + k) * B.n + bi] = w;
                    if (best_i[k] < 0 || w < best_w[k]) { best_w[k] = w; best_i[k] = bi; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kProbRows; ++k) {
            // first index of the minimum over the 32 lanes
            double bw = best_w[k];
            int bi = best_i[k];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, bw, m);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
                if (oi >= 0 && (bi < 0 || ow < bw || (ow == bw && oi < bi))) { bw = ow; bi = oi; }
            }
            const bool def = GUARD && __any_sync(0xffffffffu, deferred[k]);
            if (lane == 0 && r0 + k < A.n) {
                // label_prop_utils.py:342-345: None when the minimum is MAX_WEIGHT
                const bool none = (bi < 0) || (bw == MCG_MAX_WEIGHT);
                if (last_group) {
                    if (out.argmin) out.argmin[A.first + r0 + k] = none ? -1 : bi;
                    if (out.wmin) out.wmin[A.first + r0 + k] = none ? MCG_MAX_WEIGHT : bw;
                    if (def && (none || !(bw < kGuardWeight))) {
                        const int slot = atomicAdd(redo.count, 1);
                        const long long r = r0 + k;
                        redo.rows[slot] = A.index ? A.index[A.first + r] : (A.first + r);
                        redo.pos[slot] = A.first + r;
                    }
                } else {   // carry the running minimum to the next bin group
                    out.argmin[A.first + r0 + k] = GUARD ? ((bi + 1) | (def ? kDeferBit : 0)) : bi;
                    out.wmin[A.first + r0 + k] = bw;
                }
            }
        }
    }
}


// ---- rule C as a branch and bound --------------------------------------------------------------
// The argmin of rule C (label_prop_utils.py:340-345) needs very few of the B weights of a contig.
// Two facts about w = -(log prob_comp + log prob_cov) / ln 10:
//   * -log prob_comp = log(1 + rho) >= log rho = qa d^2 + qb d + qc with rho = (sd_i / sd_e)
//     exp(ae - ai): a parabola that grows with the tetramer distance d (for d > ~0.05 the two are
//     the same double, see log_comp_prob_warp; where the intra Gaussian leaves the normal range
//     the weight is above 300 or MAX_WEIGHT);
//   * log prob_cov = min(A, B) <= A = sum_s max(x_s, L), and x_s = c1 log c2 - lgamma(c1 + 1) - c2 is
//     largest at c2 = c1: whatever the bin, A <= U = sum_s max(c1 log c1 - lgamma(c1 + 1) - c1, L),
//     a number of the contig alone.
// So with W the weight of ANY bin of the contig (here: its nearest centroid), every bin with
// parabola(d) > W ln 10 + U has a weight strictly above W and is neither the argmin nor a
// first-index tie.  bound_kernel (one warp per contig) scores the nearest centroid, turns W into
// a d^2 threshold and lists the bins under it, in bin order, as the contig's candidates (~1 of
// 200 on the config-2 set); cand_score_kernel scores the listed pairs with exactly the
// arithmetic of prob_kernel (one thread per pair) and cand_argmin_kernel takes the first
// minimum per contig, with the same None / deferral rules.  A contig without a usable W
// (nothing within d < 0.17, or W >= kPruneMaxWeight) lists every bin.  Margins of 1e-9 relative on W, U
// and d are far above the rounding of either side.  When the list would exceed its capacity
// (a quarter of all pairs) the candidate kernels leave and prob_kernel scores every pair.
constexpr double kPruneMaxWeight = 287.0;
struct CandList {
    int *total;          // [1] candidates of all contigs (keeps counting beyond cap)
    int cap;
    int *start, *count;  // [A.n] range of a contig
    int2 *pair;          // [cap] (row of the chunk, bin)
    double *w;           // [cap] weight; NaN = deferred (GUARD)
};

constexpr int kHeadSamples = 8;    // samples of the second, coverage-aware test of bound_kernel
constexpr int kMaxSlabs = 64;      // 32-bin slabs per contig the candidate path takes (B <= 2048)

// (c, log c) of the first kHeadSamples samples of every bin, sample-major, so that the 32 lanes
// of bound_kernel (32 consecutive bins) read them coalesced: head[(2 q + {0,1}) * ldh + bin].
// headf: the same pairs as float2 [q * ldh + bin] for the second-probe SEARCH of bound_kernel,
// which only chooses a bin (any choice is valid) and waits for this table.
__global__ void bin_head_kernel(ScoreSide B, int S, int nq, long long ldh,
                                double *__restrict__ head, float2 *__restrict__ headf) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= B.n * nq) return;
    const long long j = i / nq;
    const int q = static_cast<int>(i - j * nq);
    const long long brow = B.index ? B.index[B.first + j] : (B.first + j);
    const double c = B.cov[brow * S + q], l = B.logc[brow * S + q];
    head[(2 * q) * ldh + j] = c;
    head[(2 * q + 1) * ldh + j] = l;
    headf[q * ldh + j] = make_float2(static_cast<float>(c), static_cast<float>(l));
}

__device__ __forceinline__ void ld_v4(double (&v)[4], const double *p) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

// What bound_kernel needs to know about a contig before it looks at bins: the weight W of its
// nearest centroid, the distance threshold that follows from it, and the coverage ceilings.
struct BoundPre {
    double w, thr2, u_tail, cov_upper;
    int jm;
    int done;   // the pre-pass listed this contig's candidates itself (see bound_pre_kernel)
};

__device__ __forceinline__ double prune_threshold(double wv, double cov_upper) {
    // parabola(d) > W ln 10 + U  <=>  d > (-qb + sqrt(qb^2 + 4 qa target)) / (2 qa)
    const double target = wv * (1.0 + 1e-9) * c_gauss.ln10 + cov_upper - c_math.prune_c;
    double disc = fma(c_math.prune_4a, target, c_math.prune_b * c_math.prune_b);
    if (!(disc > 0.0)) disc = 0.0;
    double d = (sqrt(disc) - c_math.prune_b) * c_math.prune_inv_2a;
    d = d > 0.0 ? d * (1.0 + 1e-9) : 0.0;
    return d * d * (1.0 + 1e-15);
}

// One LANE per contig (tensor-core distances): nearest centroid from the tile minima and the 64
// d2a of that tile, U and its tail, W = the weight of the nearest centroid at the upper end of
// its distance interval, the threshold.  In bound_kernel this is ~700 warp instructions per
// contig that are the same in all 32 lanes (ncu source page, r02h); here it is ~20.  The sums
// run in sample order instead of a shuffle tree: W is a bound used with a 1e-9 margin, not a
// result.
//
// A contig whose every OTHER bin is already beyond that threshold by the distance alone -- the
// minima of the other tiles and the second smallest d2a of the nearest tile, the same one-sided
// comparisons bound_kernel makes -- has at most one candidate, its nearest centroid: the lane
// runs bound_kernel's tests on that one bin and writes the list entry itself (`done`), and
// bound_kernel never sees the contig (it walks the compacted list `todo` of the others).
__global__ void __launch_bounds__(128)
bound_pre_kernel(ScoreSide A, ScoreSide B, int S, const float *__restrict__ d2, long long ldc,
                 const float *__restrict__ mrow, const unsigned long long *__restrict__ tilemin,
                 BoundPre *__restrict__ pre, int32_t *__restrict__ status, CandList cl,
                 int *__restrict__ todo) {   // todo[0] = count, todo[1 ..] = contigs left for bound_kernel
    __shared__ double ET[16];
    __shared__ double LT[128];
    const int tid = threadIdx.x;
    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    __syncthreads();
    const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + tid;
    const bool live = r < A.n;
    const long long rr = live ? r : 0;
    const int nb = static_cast<int>(B.n);
    const int nbt = (nb + 63) >> 6;
    const int nq = S < kHeadSamples ? S : kHeadSamples;
    const double sl = static_cast<double>(S) * c_math.log_floor;
    // nearest centroid (first index on ties): the tile with the smallest minimum, then its bins
    unsigned long long bk = ~0ull;
    int bt = 0;
    for (int t = 0; t < nbt; ++t) {
        const unsigned long long key = tilemin[rr * nbt + t];
        if (key < bk) { bk = key; bt = t; }
    }
    const float *row = d2 + rr * ldc + 64 * bt;
    float dmf = INFINITY, second = INFINITY;   // smallest and second smallest d2a of the tile
    int jm = 64 * bt;
    bool odd = false;                          // a NaN in the tile: bound_kernel keeps such a bin
    auto look = [&](float v, int j) {
        if (v < dmf) { second = dmf; dmf = v; jm = j; }
        else if (v < second) second = v;
        if (v != v) odd = true;
    };
#pragma unroll 4
    for (int i = 0; i < 64; i += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(row + i);   // padding reads +inf
        look(v.x, 64 * bt + i);
        look(v.y, 64 * bt + i + 1);
        look(v.z, 64 * bt + i + 2);
        look(v.w, 64 * bt + i + 3);
    }
    if (jm >= nb) { jm = 0; dmf = INFINITY; }
    const double mg = static_cast<double>(mrow[rr]);
    const double dd = static_cast<double>(dmf) + mg;
    const bool ok = dd < 0.03;   // d < 0.17: prob_comp normal with room to spare
    const long long arow = A.index ? A.index[A.first + rr] : (A.first + rr);
    const long long brow = B.index ? B.index[B.first + jm] : (B.first + jm);
    const double *ac_p = A.cov + arow * S, *al_p = A.logc + arow * S, *ag_p = A.lgam + arow * S;
    const double *bc_p = B.cov + brow * S, *bl_p = B.logc + brow * S, *bg_p = B.lgam + brow * S;
    double u = 0.0, u_tail = 0.0, s1 = 0.0, s2 = 0.0, e_head = 0.0;
    auto sample = [&](int q, double ac, double al, double ag, double bc, double bl, double bg) {
        const double am = -(ag + c_math.log_floor);
        double e = fma(ac, al, am) - ac;   // the most this sample can add
        e = e > 0.0 ? e : 0.0;
        u += e;
        if (q >= nq) u_tail += e;
        if (q == nq) e_head = s1;          // the head excess of bound_kernel's second test
        add_poisson_excess(s1, ac, am, bc, bl);
        add_poisson_excess(s2, bc, -(bg + c_math.log_floor), ac, al);
    };
    if ((S & 3) == 0) {
#pragma unroll 1
        for (int h = 0; h < S; h += 4) {
            double ac[4], al[4], ag[4], bc[4], bl[4], bg[4];
            ld_v4(ac, ac_p + h); ld_v4(al, al_p + h); ld_v4(ag, ag_p + h);
            ld_v4(bc, bc_p + h); ld_v4(bl, bl_p + h); ld_v4(bg, bg_p + h);
#pragma unroll
            for (int k = 0; k < 4; ++k) sample(h + k, ac[k], al[k], ag[k], bc[k], bl[k], bg[k]);
        }
    } else if ((S & 1) == 0) {
        // rows of an even S are 16-byte aligned: two samples per load
#pragma unroll 2
        for (int h = 0; h < S; h += 2) {
            const double2 ac = *reinterpret_cast<const double2 *>(ac_p + h);
            const double2 al = *reinterpret_cast<const double2 *>(al_p + h);
            const double2 ag = *reinterpret_cast<const double2 *>(ag_p + h);
            const double2 bc = *reinterpret_cast<const double2 *>(bc_p + h);
            const double2 bl = *reinterpret_cast<const double2 *>(bl_p + h);
            const double2 bg = *reinterpret_cast<const double2 *>(bg_p + h);
            sample(h, ac.x, al.x, ag.x, bc.x, bl.x, bg.x);
            sample(h + 1, ac.y, al.y, ag.y, bc.y, bl.y, bg.y);
        }
    } else {
#pragma unroll 2
        for (int q = 0; q < S; ++q) sample(q, ac_p[q], al_p[q], ag_p[q], bc_p[q], bl_p[q], bg_p[q]);
    }
    double cov_upper = fma(u, 1.0 + 1e-9, sl);   // U, rounded up
    if (cov_upper > 0.0) cov_upper = 0.0;
    const double lpc = log_comp_prob_warp(ok ? dd : 0.01, ET, LT, status);
    const double w = ok ? -(lpc + sl + (s1 < s2 ? s1 : s2)) * c_math.inv_ln10 : INFINITY;
    if (S <= nq) e_head = s1;
    if (live) {
        BoundPre out;
        out.w = w;
        out.thr2 = (w < kPruneMaxWeight) ? prune_threshold(w, cov_upper) : INFINITY;
        out.u_tail = u_tail;
        out.cov_upper = cov_upper;
        out.jm = jm;
        // every other bin beyond the threshold by the distance alone (lower ends)?
        bool alone = out.thr2 < INFINITY && !odd &&
                     fmax(static_cast<double>(second) - mg, 0.0) > out.thr2;
        for (int t = 0; t < nbt && alone; ++t) {
            if (t == bt) continue;
            const unsigned long long key = tilemin[rr * nbt + t];
            const double lb = __longlong_as_double(static_cast<long long>(
                key <= 0x7ff0000000000000ull ? key : 0x7ff0000000000000ull));
            if (!(lb > out.thr2)) alone = false;
        }
        out.done = alone ? 1 : 0;
        pre[r] = out;
        if (!alone) todo[1 + atomicAdd(todo, 1)] = static_cast<int>(r);
        if (alone) {
            // bound_kernel's tests for the one bin left
            const double raw = static_cast<double>(dmf);
            const double dlo = fmax(raw - mg, 0.0);
            bool keep = !(dlo > out.thr2);
            if (dlo > 0.0405 && raw + mg < 1.5) keep = false;
            const double slack = w * (1.0 + 1e-9) * c_gauss.ln10 + sl + u_tail * (1.0 + 1e-9) -
                                 c_math.prune_c;
            if (fma(c_math.prune_4a * 0.25, dlo, -e_head * (1.0 + 1e-9)) > slack) keep = false;
            int at = 0;
            if (keep) {
                at = atomicAdd(cl.total, 1);
                if (at < cl.cap) cl.pair[at] = make_int2(static_cast<int>(r), jm);
            }
            cl.start[r] = at;
            cl.count[r] = keep ? 1 : 0;
        }
    }
}

// D2T = double: d2 is dist2_kernel's exact matrix (mrow = nullptr).  D2T = float: d2 is the
// tensor-core estimate d2a of mcg_tc.cu and mrow[r] the margin within which every d2a of row r
// lies around the true d^2.  Every use of a distance below is one-sided, so it takes the side of
// the interval that keeps the bound valid: ruling a bin OUT compares the LOWER end
// (d2a - margin), and the weight W of a probed bin -- which only has to be AT LEAST the weight
// of some bin -- is evaluated at the UPPER end (the weight grows with the distance).
// PRE: W, the threshold and the coverage ceilings of every contig come from bound_pre_kernel.
template <typename D2T, bool PRE>
__global__ void __launch_bounds__(256)
bound_kernel(ScoreSide A, ScoreSide B, int S, const D2T *__restrict__ d2, long long ldc,
             const float *__restrict__ mrow, const double *__restrict__ head, long long ldh,
             const unsigned long long *__restrict__ tilemin, CandList cl,
             int32_t *__restrict__ status, const BoundPre *__restrict__ pre,
             const float2 *__restrict__ headf, const int *__restrict__ todo) {
    __shared__ double ET[16];
    __shared__ double LT[128];
    __shared__ double HA[8][2 * kHeadSamples];   // per warp: (c, m) of the contig's first samples
    __shared__ float HAF[8][2 * kHeadSamples];   // the same in float, for the second-probe search
    __shared__ unsigned MASK[8][kMaxSlabs];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    __syncthreads();
    const int nq = S < kHeadSamples ? S : kHeadSamples;
    const double sl = static_cast<double>(S) * c_math.log_floor;
    const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    // PRE: the contigs the pre-pass left (it listed the single candidate of the others itself)
    const long long n_work = PRE ? todo[0] : A.n;
    for (long long it = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + wid; it < n_work;
         it += n_warps) {
        const long long r = PRE ? todo[1 + it] : it;
        const D2T *row = d2 + r * ldc;
        const double mg = mrow ? static_cast<double>(mrow[r]) : 0.0;
        // dist2_kernel left the smallest d^2 of the row in every 64-bin tile (ordered bits):
        // lane t holds tile t; tiles whose minimum is beyond a threshold are never read
        const int nb = static_cast<int>(B.n);
        const int nbt = (nb + 63) >> 6;
        const unsigned long long tile_key = lane < nbt ? tilemin[r * nbt + lane] : ~0ull;
        const double tile_d2 = __longlong_as_double(static_cast<long long>(
            tile_key <= 0x7ff0000000000000ull ? tile_key : 0x7ff0000000000000ull));
        const long long arow = A.index ? A.index[A.first + r] : (A.first + r);
        double dm = INFINITY, u_tail = 0.0, cov_upper = 0.0, w_pre = INFINITY, thr_pre = INFINITY;
        int jm = 0;
        __syncwarp();
        if (PRE) {
            const BoundPre bp = pre[r];
            w_pre = bp.w; thr_pre = bp.thr2; u_tail = bp.u_tail; cov_upper = bp.cov_upper; jm = bp.jm;
            if (lane < nq) {
                const double hc = A.cov[arow * S + lane];
                const double hm = -(A.lgam[arow * S + lane] + c_math.log_floor);
                HA[wid][2 * lane] = hc;
                HA[wid][2 * lane + 1] = hm;
                HAF[wid][2 * lane] = static_cast<float>(hc);
                HAF[wid][2 * lane + 1] = static_cast<float>(hm);
            }
            __syncwarp();
        } else {
        // nearest centroid (first index on ties): the tile with the smallest minimum, then its bins
        unsigned long long bk = tile_key;
        int bt_near = lane;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, m);
            const int ot = __shfl_xor_sync(0xffffffffu, bt_near, m);
            if (ok < bk || (ok == bk && ot < bt_near)) { bk = ok; bt_near = ot; }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 64 * bt_near + 32 * h + lane;
            if (j < nb) {
                const double v = static_cast<double>(row[j]);
                if (v < dm) { dm = v; jm = j; }
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, dm, m);
            const int oj = __shfl_xor_sync(0xffffffffu, jm, m);
            if (od < dm || (od == dm && oj < jm)) { dm = od; jm = oj; }
        }
        // the contig's side: U (and its tail beyond the head samples), head terms for the
        // coverage-aware test
        double u = 0.0;   // lanes over samples
        for (int q = lane; q < S; q += 32) {
            const double ac = A.cov[arow * S + q], al = A.logc[arow * S + q];
            const double am = -(A.lgam[arow * S + q] + c_math.log_floor);
            double e = fma(ac, al, am) - ac;   // the most this sample can add
            e = e > 0.0 ? e : 0.0;
            u += e;
            if (q >= nq) u_tail += e;
            if (q < nq) {
                HA[wid][2 * q] = ac; HA[wid][2 * q + 1] = am;
                HAF[wid][2 * q] = static_cast<float>(ac); HAF[wid][2 * q + 1] = static_cast<float>(am);
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            u += __shfl_xor_sync(0xffffffffu, u, m);
            u_tail += __shfl_xor_sync(0xffffffffu, u_tail, m);
        }
        __syncwarp();
        cov_upper = fma(u, 1.0 + 1e-9, sl);   // U, rounded up
        if (cov_upper > 0.0) cov_upper = 0.0;
        }
        // weight of one bin, by the whole warp (NaN: prob_comp is 0 or not a plain number)
        auto probe = [&](int j, double dd) -> double {
            if (!(dd < 0.03)) return INFINITY;   // d < 0.17: prob_comp normal with room to spare
            const long long brow = B.index ? B.index[B.first + j] : (B.first + j);
            double s1 = 0.0, s2 = 0.0;
            for (int q = lane; q < S; q += 32) {
                const double ac = A.cov[arow * S + q], al = A.logc[arow * S + q];
                const double am = -(A.lgam[arow * S + q] + c_math.log_floor);
                const double bc = B.cov[brow * S + q], bl = B.logc[brow * S + q];
                const double bm = -(B.lgam[brow * S + q] + c_math.log_floor);
                add_poisson_excess(s1, ac, am, bc, bl);
                add_poisson_excess(s2, bc, bm, ac, al);
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, m);
                s2 += __shfl_xor_sync(0xffffffffu, s2, m);
            }
            const double lpc = log_comp_prob_warp(dd, ET, LT, status);
            return -(lpc + sl + (s1 < s2 ? s1 : s2)) * c_math.inv_ln10;
        };
        double w = PRE ? w_pre : probe(jm, dm + mg);
        auto threshold = [&](double wv) -> double { return prune_threshold(wv, cov_upper); };
        const int n_slabs = static_cast<int>((B.n + 31) >> 5);
        double thr2 = PRE ? thr_pre : ((w < kPruneMaxWeight) ? threshold(w) : INFINITY);
        if (!(thr2 < 0.045 * 0.045) && nq > 0) {
            // The nearest centroid gave no bound, or a loose one (short contigs sit as close to
            // a neighbour's centroid as to their own).  Second probe: the bin with the smallest
            // LOWER bound of its weight from the distance and the head samples.
            double lb = INFINITY;
            int jb = 0;
            const unsigned tiles2 = __ballot_sync(0xffffffffu, lane < nbt && !(tile_d2 > thr2));
            for (unsigned tm = tiles2; tm; tm &= tm - 1)   // only the tiles with a bin that near
            for (int sb = 2 * (__ffs(tm) - 1), h = 0; h < 2 && sb < n_slabs; ++h, ++sb) {
                const int j = 32 * sb + lane;
                if (j < B.n) {
                    const double dd = fmax(static_cast<double>(row[j]) - mg, 0.0);
                    // float terms from the float2 table: this pass only CHOOSES the bin to probe
                    // (any choice is valid) and it waits for the table (ncu r02o: 39 % of the
                    // kernel's samples at S = 50): half the bytes, one load per sample
                    float e1 = 0.f;
                    const float2 *hf = headf + j;
#pragma unroll 4
                    for (int q = 0; q < nq; ++q) {
                        const float2 b = __ldg(hf + q * ldh);
                        const float d = fmaf(HAF[wid][2 * q], b.y, HAF[wid][2 * q + 1]) - b.x;
                        e1 += d > 0.f ? d : 0.f;
                    }
                    const double v = fma(c_math.prune_4a * 0.25, dd, -static_cast<double>(e1));
                    if (v < lb) { lb = v; jb = j; }
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double ol = __shfl_xor_sync(0xffffffffu, lb, m);
                const int oj = __shfl_xor_sync(0xffffffffu, jb, m);
                if (ol < lb || (ol == lb && oj < jb)) { lb = ol; jb = oj; }
            }
            if (lb < INFINITY && jb != jm) {
                const double w2 = probe(jb, static_cast<double>(row[jb]) + mg);
                if (w2 < w || !(w == w)) w = w2;
                if (w < kPruneMaxWeight) thr2 = threshold(w);
            }
        }
        // second test: out when  qa d^2 - (head excess)  >  slack, i.e. when
        // (qa d^2 + qc) - (S L + head excess + tail bound) > W ln 10  (the qb d term of the
        // parabola, >= 0, is dropped: no sqrt per pair)
        double slack = INFINITY;
        if (w < kPruneMaxWeight)
            slack = w * (1.0 + 1e-9) * c_gauss.ln10 + sl + u_tail * (1.0 + 1e-9) - c_math.prune_c;
        // the candidates, in bin order; a NaN distance stays in
        int count = 0;
        const unsigned tiles = __ballot_sync(0xffffffffu, lane < nbt && !(tile_d2 > thr2));
        __syncwarp();
        for (unsigned tm = tiles; tm; tm &= tm - 1)   // the other tiles: every bin beyond thr2
        for (int sb = 2 * (__ffs(tm) - 1), h = 0; h < 2 && sb < n_slabs; ++h, ++sb) {
            const int j = 32 * sb + lane;
            const double raw = j < nb ? static_cast<double>(row[j]) : INFINITY;
            const double dd = fmax(raw - mg, 0.0);       // lower end (a NaN stays in: dd = 0)
            bool keep = j < nb && !(dd > thr2);
            // prob_comp is exactly 0 from d = 0.2003 on (exp underflow, log_comp_prob_special):
            // the weight is MAX_WEIGHT, which never wins and means None when nothing else is
            // left.  Beyond d^2 = 1.5 the reference's own sequence decides (domain errors): kept.
            if (dd > 0.0405 && raw + mg < 1.5) keep = false;
            if (slack < INFINITY && __any_sync(0xffffffffu, keep)) {
                double e1 = 0.0;
                const double *hp = head + (j < B.n ? j : 0);
#pragma unroll 4
                for (int q = 0; q < nq; ++q) {
                    const double bc = hp[(2 * q) * ldh], bl = hp[(2 * q + 1) * ldh];
                    add_poisson_excess(e1, HA[wid][2 * q], HA[wid][2 * q + 1], bc, bl);
                }
                const double lhs = fma(c_math.prune_4a * 0.25, dd, -e1 * (1.0 + 1e-9));
                if (lhs > slack) keep = false;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) MASK[wid][sb] = mask;
            count += __popc(mask);
        }
        int base = 0;
        if (lane == 0) {
            base = atomicAdd(cl.total, count);
            cl.start[r] = base;
            cl.count[r] = count;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        if (base + count <= cl.cap) {
            int at = base;
            for (unsigned tm = tiles; tm && at < base + count; tm &= tm - 1)
            for (int sb = 2 * (__ffs(tm) - 1), h = 0; h < 2 && sb < n_slabs; ++h, ++sb) {
                const unsigned mask = MASK[wid][sb];
                if (mask == 0u) continue;
                if ((mask >> lane) & 1u)
                    cl.pair[at + __popc(mask & ((1u << lane) - 1u))] =
                        make_int2(static_cast<int>(r), 32 * sb + lane);
                at += __popc(mask);
            }
        }
    }
}

// One thread per listed pair: the weight prob_kernel computes for it, operation for operation
// (log_comp_prob_warp, excess sums in sample order, the e^-744 / e^-746 window, the GUARD
// deferral); NaN marks a deferred pair.
// The float64 d^2 of one pair with the bits dist2_kernel gives it: mma.sync.m8n8k4.f64 is a chain
// of FMAs over k in ascending order (experiments/dmma_fma_chain.cu: 262 144 of 262 144 outputs
// bit-equal on the B200), the norms are the DMMA norms of row_norms_kernel, and the epilogue is
// dist2_kernel's fma(-2, u.v, |u|^2 + |v|^2) clamped at 0.
// N x 4 slots of the chain: all 2 N loads (32 bytes each; rows are 32-byte aligned, 136 x 8 =
// 34 x 32) are in flight before the first FMA needs one -- a lane walks its own two rows, and
// with one load per FMA group the walk paid an L2 round trip per 32 bytes
template <int N>
__device__ __forceinline__ double fma_chain(const double *u, const double *v, double acc) {
    double a[N][4], b[N][4];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ld_v4(a[i], u + 4 * i);
        ld_v4(b[i], v + 4 * i);
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc = fma(a[i][q], b[i][q], acc);
    return acc;
}

__device__ __forceinline__ double exact_d2(const ScoreSide &A, const ScoreSide &B, long long arow,
                                           long long brow) {
    const double *u = A.prof + arow * MCG_NPROF, *v = B.prof + brow * MCG_NPROF;
    double acc = fma_chain<9>(u, v, 0.0);          // k = 0 .. 35
    acc = fma_chain<9>(u + 36, v + 36, acc);       //     36 .. 71
    acc = fma_chain<8>(u + 72, v + 72, acc);       //     72 .. 103
    acc = fma_chain<8>(u + 104, v + 104, acc);     //    104 .. 135
    const double d2 = fma(-2.0, acc, A.norm[arow] + B.norm[brow]);
    return d2 > 0.0 ? d2 : 0.0;
}

// EXACT: d2 is nullptr (the distances came from the tensor-core bound) and every listed pair gets
// its float64 d^2 here.
template <bool GUARD, bool EXACT>
__global__ void __launch_bounds__(256)
cand_score_kernel(ScoreSide A, ScoreSide B, int S, const double *__restrict__ d2, long long ldc,
                  CandList cl, int32_t *__restrict__ status) {
    __shared__ double ET[16];
    __shared__ double LT[128];
    const int tid = threadIdx.x;
    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    __syncthreads();
    const int total = *cl.total;
    if (total > cl.cap) return;
    const int stride = gridDim.x * blockDim.x;
    const int padded = (total + 31) & ~31;   // whole warps: log_comp_prob_warp votes
    for (int i = blockIdx.x * blockDim.x + tid; i < padded; i += stride) {
        const bool live = i < total;
        const int2 pr = live ? cl.pair[i] : make_int2(0, 0);
        const long long arow = A.index ? A.index[A.first + pr.x] : (A.first + pr.x);
        const long long brow = B.index ? B.index[B.first + pr.y] : (B.first + pr.y);
        const double dd = EXACT ? exact_d2(A, B, arow, brow) : d2[pr.x * ldc + pr.y];
        const double log_pc = log_comp_prob_warp(dd, ET, LT, status);
        double tsum = log_pc;
        bool defer = false;
        if (log_pc == log_pc) {
            const double *ac_p = A.cov + arow * S, *al_p = A.logc + arow * S, *ag_p = A.lgam + arow * S;
            const double *bc_p = B.cov + brow * S, *bl_p = B.logc + brow * S, *bg_p = B.lgam + brow * S;
            double s1 = 0.0, s2 = 0.0;
            if ((S & 3) == 0) {
                // four samples per 32-byte load (sm_100: ld.global.v4.f64; rows of S % 4 == 0 are
                // 32-byte aligned): one sector per lane and request instead of a quarter of one
#pragma unroll 1
                for (int h = 0; h < S; h += 4) {
                    double ac[4], al[4], ag[4], bc[4], bl[4], bg[4];
                    ld_v4(ac, ac_p + h); ld_v4(al, al_p + h); ld_v4(ag, ag_p + h);
                    ld_v4(bc, bc_p + h); ld_v4(bl, bl_p + h); ld_v4(bg, bg_p + h);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        add_poisson_excess(s1, ac[u], -(ag[u] + c_math.log_floor), bc[u], bl[u]);
                        add_poisson_excess(s2, bc[u], -(bg[u] + c_math.log_floor), ac[u], al[u]);
                    }
                }
            } else if ((S & 1) == 0) {
                // two samples per 16-byte load (rows of an even S are 16-byte aligned): half the
                // L1TEX requests of the per-lane row walks that bound this kernel; same order
                const double2 *ac2 = reinterpret_cast<const double2 *>(ac_p);
                const double2 *al2 = reinterpret_cast<const double2 *>(al_p);
                const double2 *ag2 = reinterpret_cast<const double2 *>(ag_p);
                const double2 *bc2 = reinterpret_cast<const double2 *>(bc_p);
                const double2 *bl2 = reinterpret_cast<const double2 *>(bl_p);
                const double2 *bg2 = reinterpret_cast<const double2 *>(bg_p);
#pragma unroll 2
                for (int h = 0; h < S / 2; ++h) {
                    const double2 ac = ac2[h], al = al2[h], ag = ag2[h];
                    const double2 bc = bc2[h], bl = bl2[h], bg = bg2[h];
                    add_poisson_excess(s1, ac.x, -(ag.x + c_math.log_floor), bc.x, bl.x);
                    add_poisson_excess(s2, bc.x, -(bg.x + c_math.log_floor), ac.x, al.x);
                    add_poisson_excess(s1, ac.y, -(ag.y + c_math.log_floor), bc.y, bl.y);
                    add_poisson_excess(s2, bc.y, -(bg.y + c_math.log_floor), ac.y, al.y);
                }
            } else {
#pragma unroll 2
                for (int q = 0; q < S; ++q) {
                    const double ac = ac_p[q], al = al_p[q], am = -(ag_p[q] + c_math.log_floor);
                    const double bc = bc_p[q], bl = bl_p[q], bm = -(bg_p[q] + c_math.log_floor);
                    add_poisson_excess(s1, ac, am, bc, bl);
                    add_poisson_excess(s2, bc, bm, ac, al);
                }
            }
            const double pmin = fma(static_cast<double>(S), c_math.log_floor, s1 < s2 ? s1 : s2);
            tsum = log_pc + pmin;
            if (GUARD) defer = pmin < kGuardLog;
        }
        bool valid = tsum > -744.0;
        if (!valid && tsum >= -746.0)
            valid = __dmul_rn(exp_slow(log_pc), exp_slow(tsum - log_pc)) > 0.0;
        const double lp = -tsum * c_math.inv_ln10;
        const double w = (valid && lp != 0.0) ? lp : MCG_MAX_WEIGHT;
        if (live) cl.w[i] = defer ? __longlong_as_double(0x7ff8000000000000LL) : w;
    }
}

// First minimum of a contig's candidates and the bookkeeping of prob_kernel's last bin group.
// One lane per contig walks a short list (the common case: 1-20 candidates); a contig that is
// far from every centroid lists every bin (hundreds), and that one serial walk used to set the
// kernel's time (77 us at config 3 against 6 us at config 2): lists beyond 32 entries are taken
// by the whole warp, one after the other, lanes striding over the list.  (weight, position) is
// compared lexicographically, which is the sequential loop's "first position of the smallest
// weight".
template <bool GUARD>
__global__ void __launch_bounds__(256)
cand_argmin_kernel(ScoreSide A, CandList cl, ScoreOut out, RedoList redo) {
    if (*cl.total > cl.cap) return;
    const int lane = threadIdx.x & 31;
    const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool live = r < A.n;
    const int lo = live ? cl.start[r] : 0, cnt = live ? cl.count[r] : 0;
    const int kNone = 0x7fffffff;
    double bw = INFINITY;
    int bp = kNone;              // position of the best candidate; none yet
    bool def = false;
    const bool wide = cnt > 32;
    if (!wide)
        for (int i = lo; i < lo + cnt; ++i) {
            const double w = cl.w[i];
            if (w != w) { def = true; continue; }
            if (bp == kNone || w < bw) { bw = w; bp = i; }
        }
    for (unsigned todo = __ballot_sync(0xffffffffu, wide); todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const int lo_s = __shfl_sync(0xffffffffu, lo, src), hi_s = lo_s + __shfl_sync(0xffffffffu, cnt, src);
        double w1 = INFINITY;
        int p1 = kNone;
        bool d1 = false;
        for (int i = lo_s + lane; i < hi_s; i += 32) {
            const double w = cl.w[i];
            if (w != w) { d1 = true; continue; }
            if (p1 == kNone || w < w1) { w1 = w; p1 = i; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double ow = __shfl_xor_sync(0xffffffffu, w1, m);
            const int op = __shfl_xor_sync(0xffffffffu, p1, m);
            if (op != kNone && (p1 == kNone || ow < w1 || (ow == w1 && op < p1))) { w1 = ow; p1 = op; }
        }
        d1 = __any_sync(0xffffffffu, d1);
        if (lane == src) { bw = w1; bp = p1; def = d1; }
    }
    if (!live) return;
    const int bi = bp == kNone ? -1 : cl.pair[bp].y;
    const bool none = (bi < 0) || (bw == MCG_MAX_WEIGHT);   // label_prop_utils.py:342-345
    if (out.argmin) out.argmin[A.first + r] = none ? -1 : bi;
    if (out.wmin) out.wmin[A.first + r] = none ? MCG_MAX_WEIGHT : bw;
    if (GUARD && def && (none || !(bw < kGuardWeight))) {
        const int slot = atomicAdd(redo.count, 1);
        redo.rows[slot] = A.index ? A.index[A.first + r] : (A.first + r);
        redo.pos[slot] = A.first + r;
    }
}

// get_tetramer_distance for all pairs in the reference's difference form (diagnostic output of
// mcg_pair_scores; the scorer itself uses the norm / dot-product form).
__global__ void pair_distance_kernel(ScoreSide A, ScoreSide B, double *__restrict__ dist) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= A.n * B.n) return;
    const long long qa = i / B.n, qb = i - qa * B.n;
    const double *u = A.prof + (A.index ? A.index[A.first + qa] : A.first + qa) * MCG_NPROF;
    const double *v = B.prof + (B.index ? B.index[B.first + qb] : B.first + qb) * MCG_NPROF;
    double acc = 0.0;
    for (int k = 0; k < MCG_NPROF; ++k) {
        const double d = u[k] - v[k];
        acc = fma(d, d, acc);
    }
    dist[i] = sqrt(acc);
}

// Rules A / B over a raw lp matrix [nq, nm]: members of bin b are the columns
// seg_offsets[b] .. seg_offsets[b+1].  The sum runs in member order like the reference,
// so MAX + finite = MAX and MAX + MAX = inf come out the same way.
__global__ void aggregate_members_kernel(const double *__restrict__ lp, long long nq,
                                         long long nm, const int64_t *__restrict__ seg_offsets,
                                         long long nseg, int rule,
                                         double *__restrict__ weights) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= nq * nseg) return;
    const long long q = i / nseg, b = i - q * nseg;
    const long long lo = seg_offsets[b], hi = seg_offsets[b + 1];
    double sum = 0.0;
    long long valid = 0;
    for (long long j = lo; j < hi; ++j) {
        const double v = lp[q * nm + j];
        if (v != MCG_MAX_WEIGHT) ++valid;
        sum = __dadd_rn(sum, v);
    }
    double w;
    if (rule == MCG_RULE_A)
        w = (sum != INFINITY) ? __ddiv_rn(sum, static_cast<double>(hi - lo)) : MCG_MAX_WEIGHT;
    else
        w = (sum != INFINITY && valid != 0) ? __ddiv_rn(sum, static_cast<double>(valid))
                                            : MCG_MAX_WEIGHT;
    weights[i] = w;
}

// Rules A / B for LISTED (query, bin) items.  label_prop_utils.py:54-86 scores a candidate only
// against the bins its walk reached (a handful of the B bins), so at scale the dense
// [queries, bins] form above does hundreds of times the work the reference does.  One warp per
// item walks the bin's members in order: squared distance in the reference's difference form
// (matching_utils.py:28-29; lanes over the 136 slots), prob_comp, the Poisson terms with one
// sample per lane and the running sums / products taken in sample order through shuffles
// (sums of floored exponents for S <= 30, in-order products beyond, exactly like pair_kernel),
// the reference's -(log(pc)/log 10 + log(pv)/log 10), and the member sum with MAX + finite = MAX,
// MAX + MAX = inf.
__global__ void __launch_bounds__(256)
member_items_kernel(ScoreSide T, const int64_t *__restrict__ item_query,
                    const int32_t *__restrict__ item_bin, long long n_items,
                    const int64_t *__restrict__ members, const int64_t *__restrict__ seg_offsets,
                    int S, int rule, double *__restrict__ weights, int32_t *__restrict__ status) {
    __shared__ double ET[16];
    __shared__ double LT[128];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 16) ET[tid] = g_exp_table[tid];
    if (tid < 128) LT[tid] = g_log_table[tid];
    __syncthreads();
    const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
    const bool logsum = S <= 30;
    for (long long it = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (tid >> 5);
         it < n_items; it += n_warps) {
        const long long q = item_query[it];
        const int b = item_bin[it];
        const long long lo = seg_offsets[b], hi = seg_offsets[b + 1];
        const double *u = T.prof + q * MCG_NPROF;
        double ur[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) ur[i] = (lane + 32 * i < MCG_NPROF) ? u[lane + 32 * i] : 0.0;
        double sum = 0.0;
        long long n_valid = 0;
        bool domain = false;
        for (long long j = lo; j < hi; ++j) {
            const long long m = members[j];
            const double *v = T.prof + m * MCG_NPROF;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const double d = ur[i] - ((lane + 32 * i < MCG_NPROF) ? v[lane + 32 * i] : 0.0);
                acc = fma(d, d, acc);
            }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, w);
            const double pc = comp_prob_fast(acc, ET);
            if (pc != pc) domain = true;
            // cov1 = the query, cov2 = the member (label_prop_utils.py:66-68)
            double p1 = logsum ? 0.0 : 1.0, p2 = p1;
            for (int s0 = 0; s0 < S; s0 += 32) {
                const int s = s0 + lane;
                double t1 = 0.0, t2 = 0.0;
                if (s < S) {
                    const double ac = T.cov[q * S + s], al = T.logc[q * S + s], ag = T.lgam[q * S + s];
                    const double bc = T.cov[m * S + s], bl = T.logc[m * S + s], bg = T.lgam[m * S + s];
                    t1 = poisson_exponent(ac, ag, bc, bl);
                    t2 = poisson_exponent(bc, bg, ac, al);
                    if (!logsum) { t1 = exp_core(t1, ET); t2 = exp_core(t2, ET); }
                }
                const int sc = S - s0 < 32 ? S - s0 : 32;
                for (int k = 0; k < sc; ++k) {
                    const double a1 = __shfl_sync(0xffffffffu, t1, k);
                    const double a2 = __shfl_sync(0xffffffffu, t2, k);
                    if (logsum) { p1 = __dadd_rn(p1, a1); p2 = __dadd_rn(p2, a2); }
                    else { p1 = __dmul_rn(p1, a1); p2 = __dmul_rn(p2, a2); }
                }
            }
            const double pmin = p1 < p2 ? p1 : p2;
            const double pv = logsum ? exp_core(pmin, ET) : pmin;
            double lp = MCG_MAX_WEIGHT;
            if (__dmul_rn(pc, pv) > 0.0) {
                const double log_pv = logsum ? pmin : log_core(pv, LT);
                lp = -(__ddiv_rn(log_core(pc, LT), c_gauss.ln10) + __ddiv_rn(log_pv, c_gauss.ln10));
                ++n_valid;
            }
            sum = __dadd_rn(sum, lp);
        }
        if (lane == 0) {
            double w;
            if (rule == MCG_RULE_A)
                // (an empty bin is 0 / 0 here as in aggregate_members_kernel; the reference
                // raises ZeroDivisionError at matching_utils.py:152)
                w = (sum != INFINITY) ? __ddiv_rn(sum, static_cast<double>(hi - lo)) : MCG_MAX_WEIGHT;
            else
                w = (sum != INFINITY && n_valid != 0) ? __ddiv_rn(sum, static_cast<double>(n_valid))
                                                      : MCG_MAX_WEIGHT;
            weights[it] = w;
            if (domain) atomicOr(status, 1);
        }
    }
}

// feature_utils.py:232-233 np.mean(rows, axis=0): rows added in member order, one division.
__global__ void segment_mean_kernel(const double *__restrict__ rows, int width,
                                    const int64_t *__restrict__ members,
                                    const int64_t *__restrict__ seg_offsets, long long nseg,
                                    double *__restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= nseg * width) return;
    const long long b = i / width;
    const int k = static_cast<int>(i - b * width);
    const long long lo = seg_offsets[b], hi = seg_offsets[b + 1];
    double acc = 0.0;
    for (long long j = lo; j < hi; ++j) {
        const double v = rows[members[j] * width + k];
        acc = (j == lo) ? v : __dadd_rn(acc, v);
    }
    out[i] = __ddiv_rn(acc, static_cast<double>(hi - lo));
}

// ---- elementwise primitives ----------------------------------------------------------------
__global__ void normpdf_kernel(const double *__restrict__ x, double mean, double two_var,
                               double denom, double *__restrict__ out, long long n) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = normpdf_dev(x[i], mean, two_var, denom);
}

__global__ void distance_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                long long n, int width, double *__restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int k = 0; k < width; ++k) {
        const double d = u[i * width + k] - v[i * width + k];
        acc = fma(d, d, acc);
    }
    out[i] = sqrt(acc);
}

__global__ void comp_prob_kernel(const double *__restrict__ dist, double *__restrict__ out,
                                 long long n, int32_t *__restrict__ status) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const double p = comp_prob_dev(dist[i]);
    if (p != p) atomicOr(status, 1);
    out[i] = p;
}

// matching_utils.py:42-66 for one pair per thread (no hoisting: inputs are raw coverages).
__global__ void cov_prob_kernel(const double *__restrict__ c1, const double *__restrict__ c2,
                                long long n, int S, double *__restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double p1 = 1.0, p2 = 1.0;
    for (int s = 0; s < S; ++s) {
        const double a = c1[i * S + s], b = c2[i * S + s];
        p1 = __dmul_rn(p1, poisson_term(a, lgamma_cpython(__dadd_rn(a, 1.0)), b, log(b), g_exp_table));
        p2 = __dmul_rn(p2, poisson_term(b, lgamma_cpython(__dadd_rn(b, 1.0)), a, log(a), g_exp_table));
    }
    out[i] = p1 < p2 ? p1 : p2;
}

// ---- FP64 FMA peak -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;   // keeps the loop alive
}

}  // namespace

// ---- launchers -------------------------------------------------------------------------------
typedef void (*pair_fn)(ScoreSide, ScoreSide, int, ScoreOut, int32_t *, const int *, long long);
static pair_fn pair_variant(bool rule_c, bool logsum) {
    if (rule_c) return logsum ? pair_kernel<true, true> : pair_kernel<true, false>;
    return logsum ? pair_kernel<false, true> : pair_kernel<false, false>;
}

static int ensure_gauss(mcg_ctx *ctx) {
    static bool ready[64] = {};
    if (ctx->device < 64 && ready[ctx->device]) return MCG_OK;
    GaussConst g;
    const double sd_i = 0.01037897 / 2, sd_e = 0.03419337;
    const double root = sqrt(2 * 3.141592653589793);   // (2 * math.pi) ** 0.5
    g.two_var_intra = 2 * (sd_i * sd_i);
    g.denom_intra = sd_i * root;
    g.mu_inter = 0.0676654;
    g.two_var_inter = 2 * (sd_e * sd_e);
    g.denom_inter = sd_e * root;
    g.ln10 = log(10.0);
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gauss, &g, sizeof g, 0, cudaMemcpyHostToDevice,
                                          ctx->stream));
    double table[16];
    for (int j = 0; j < 16; ++j) table[j] = exp2(j / 16.0);
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(g_exp_table, table, sizeof table, 0,
                                          cudaMemcpyHostToDevice, ctx->stream));
    double ltab[128];
    for (int j = 0; j < 64; ++j) {
        const double rcp = 1.0 / (1.0 + (j + 0.5) / 64.0);
        ltab[j] = rcp;
        ltab[64 + j] = static_cast<double>(-logl(static_cast<long double>(rcp)));
    }
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(g_log_table, ltab, sizeof ltab, 0,
                                          cudaMemcpyHostToDevice, ctx->stream));
    MathConst m;
    m.inv_ln2_16 = 23.083120654223414;          // 16 / ln 2
    m.magic = 6755399441055744.0;               // 1.5 * 2^52
    m.ln2_16_hi = -0x1.62e42ff000000p-5;        // -(ln2/16), high 32 bits
    m.ln2_16_lo = 0x1.718432a1b0e26p-39;        // minus the low part
    m.e7 = 1.0 / 5040.0; m.e6 = 1.0 / 720.0; m.e5 = 1.0 / 120.0; m.e4 = 1.0 / 24.0;
    m.e3 = 1.0 / 6.0;
    m.ln2_hi = 0x1.62e42fefa3800p-1;            // ln2 with 11 trailing zero bits
    m.ln2_lo = 0x1.ef35793c76730p-45;
    m.l7 = 1.0 / 7.0; m.l6 = -1.0 / 6.0; m.l5 = 1.0 / 5.0; m.l4 = -1.0 / 4.0; m.l3 = 1.0 / 3.0;
    m.log_floor = log(1e-10);
    m.inv_two_var_intra = 1.0 / g.two_var_intra;
    m.inv_denom_intra = 1.0 / g.denom_intra;
    m.inv_two_var_inter = 1.0 / g.two_var_inter;
    m.inv_denom_inter = 1.0 / g.denom_inter;
    m.inv_ln10 = 1.0 / g.ln10;
    m.log_inv_denom_intra = -log(g.denom_intra);
    m.sd_ratio = sd_i / sd_e;
    m.log_sd_ratio = log(sd_i / sd_e);
    {
        const double k1 = m.inv_two_var_intra, k2 = m.inv_two_var_inter, mu = g.mu_inter;
        m.prune_4a = 4.0 * (k1 - k2);
        m.prune_b = 2.0 * mu * k2;
        m.prune_c = m.log_sd_ratio - mu * mu * k2;
        m.prune_inv_2a = 1.0 / (2.0 * (k1 - k2));
    }
    MCG_CUDA(ctx, cudaMemcpyToSymbolAsync(c_math, &m, sizeof m, 0, cudaMemcpyHostToDevice,
                                          ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < 4; ++v) MCG_CUDA(ctx, cudaFuncSetAttribute(
        pair_variant(v & 2, v & 1), cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(kPairSmem)));
    if (ctx->device < 64) ready[ctx->device] = true;
    return MCG_OK;
}

static inline unsigned blocks_for(long long n, int threads) {
    return static_cast<unsigned>((n + threads - 1) / threads);
}

int mcg_k_cov_terms(mcg_ctx *ctx, double *d_cov, double *d_logc, double *d_lgam, int64_t count) {
    if (count == 0) return MCG_OK;
    cov_terms_kernel<<<blocks_for(count, 256), 256, 0, ctx->stream>>>(
        d_cov, d_logc, d_lgam, count, ctx->status.as<int32_t>());
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// argmin / wmin of the re-scored contigs go back to their places.  The redo pass spreads the
// bins over `parts` slices (grid.y of pair_kernel): first index of the minimum over the slices.
__global__ void scatter_redo_kernel(const int *__restrict__ count, const long long *__restrict__ pos,
                                    const int32_t *__restrict__ arg, const double *__restrict__ w,
                                    int parts, long long part_stride,
                                    int32_t *__restrict__ out_arg, double *__restrict__ out_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    int bi = arg[i];
    double bw = w[i];
    for (int p = 1; p < parts; ++p) {
        const int oi = arg[i + p * part_stride];
        const double ow = w[i + p * part_stride];
        if (oi >= 0 && (bi < 0 || ow < bw || (ow == bw && oi < bi))) { bw = ow; bi = oi; }
    }
    out_arg[pos[i]] = bi;
    out_w[pos[i]] = bi < 0 ? MCG_MAX_WEIGHT : bw;
}

static size_t prob_smem_bytes(int S, long long group, int warps = kProbWarps) {
    return sizeof(double) * (16 + 128 + warps * kProbRows * 4 * S + bin_pitch(S) * group);
}

// Can the two-kernel rule-C path take S samples at all (one warp of bins per group)?
static bool split_fits(int S) { return prob_smem_bytes(S, 32) <= 200 * 1024; }

// The two-kernel rule-C path over row chunks of the A side (the d^2 scratch is bounded).
// S <= 30: the log-sum form is exact for every pair.  S > 30: prob_kernel<true> defers the
// pairs whose Poisson product may leave the normal range and lists the contigs whose answer
// could depend on one; pair_kernel re-scores those in the reference's operation order.
static int split_rule_c(mcg_ctx *ctx, const ScoreSide &a, const ScoreSide &b, int S,
                        const ScoreOut &out, int32_t *d_status) {
    static bool attr[64] = {};
    const bool guard = S > 30;
    // the full weight matrix needs every pair; the argmin does not (bound_kernel)
    const bool use_cand = ctx->prune && !out.weights && b.n <= 32LL * kMaxSlabs;
    const long long b_tiles = (b.n + kTile - 1) / kTile;
    const long long ldc = b_tiles * kTile;
    // bins per shared-memory group of prob_kernel, whole warps of columns: S <= 30 keeps the
    // terms under 56 KB (4 CTAs per SM); larger S takes up to ~100 KB (2 CTAs per SM), or one
    // warp of bins when even that does not fit
    const int pitch = bin_pitch(S);
    const long long fixed = static_cast<long long>(prob_smem_bytes(S, 0));
    long long budget = guard ? 104 * 1024 - fixed : 56 * 1024;
    long long group = budget > 0 ? budget / static_cast<long long>(sizeof(double) * pitch) : 0;
    group = (group / 32) * 32;
    if (group < 32) group = 32;
    const long long n_groups = (b.n + group - 1) / group;
    group = ((b.n + n_groups - 1) / n_groups + 31) / 32 * 32;
    // When even the smallest group leaves room for one 8-warp CTA per SM only (S ~ 100), a
    // 16-warp CTA shares the same bin terms between twice the warps, if it fits.
    int warps = kProbWarps;
    if (guard && 2 * prob_smem_bytes(S, group) > 220 * 1024 &&
        prob_smem_bytes(S, group, 16) <= 200 * 1024)
        warps = 16;
    const size_t prob_smem = prob_smem_bytes(S, group, warps);
    if (!(ctx->device < 64 && attr[ctx->device])) {
        MCG_CUDA(ctx, cudaFuncSetAttribute(dist2_kernel<0>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kDistSmem)));
        MCG_CUDA(ctx, cudaFuncSetAttribute(prob_kernel<false, 8>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           96 * 1024));
        MCG_CUDA(ctx, cudaFuncSetAttribute(prob_kernel<true, 8>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
        MCG_CUDA(ctx, cudaFuncSetAttribute(prob_kernel<true, 16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));

        if (ctx->device < 64) attr[ctx->device] = true;
    }
    // rows per chunk: <= 4 GiB of (fallback) float64 scratch, whole tiles.  Every chunk costs a
    // fixed set of launches (at S = 100, B = 1000 that is 31 gated bin-group launches next to the
    // working kernels) and a tail per kernel, so chunks are as large as 180 GB of HBM
    // comfortably allow: 524 288 rows at B = 1000 (d2a 2 GiB, operands 0.47 GiB, candidate list
    // <= 2 GiB).  Measured, config 3 at size / config-5 shard: 512 MiB 13.05 / 17.2 ms, 2 GiB
    // 10.19 / 15.0 ms, 4 GiB 10.00 / 14.6 ms.
    long long rows_per_chunk = (4096LL << 20) / (sizeof(double) * ldc);
    rows_per_chunk = (rows_per_chunk / kTile) * kTile;
    if (rows_per_chunk < kTile) rows_per_chunk = kTile;
    const long long need_rows = ((a.n < rows_per_chunk ? a.n : rows_per_chunk) + kTile - 1) /
                                kTile * kTile;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[5], sizeof(double) * need_rows * ldc));
    double *d2 = ctx->scratch[5].as<double>();
    RedoList redo = {nullptr, nullptr, nullptr};
    int32_t *redo_arg = nullptr;
    double *redo_w = nullptr;
    int redo_parts = 1;   // slices of the bins in the redo pass (few contigs: fill the SMs)
    if (guard) {
        // [count | pad] [rows a.n] [pos a.n] [wmin parts x a.n] [argmin parts x a.n]
        redo_parts = static_cast<int>(b_tiles < 16 ? b_tiles : 16);
        const size_t bytes = 16 + static_cast<size_t>(a.n) * (8 + 8 + (8 + 4) * redo_parts);
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[4], bytes));
        char *base = ctx->scratch[4].as<char>();
        redo.count = reinterpret_cast<int *>(base);
        redo.rows = reinterpret_cast<long long *>(base + 16);
        redo.pos = redo.rows + a.n;
        redo_w = reinterpret_cast<double *>(redo.pos + a.n);
        redo_arg = reinterpret_cast<int32_t *>(redo_w + a.n * redo_parts);
        MCG_CUDA(ctx, cudaMemsetAsync(redo.count, 0, 16, ctx->stream));
    }
    ctx->redo_valid = guard;
    CandList cl = {nullptr, 0, nullptr, nullptr, nullptr, nullptr};
    unsigned long long *tilemin = nullptr;   // per (row, 64-bin tile) smallest d^2, from dist2_kernel
    float2 *headf = nullptr;
    ctx->cand_valid = false;
    if (use_cand) {
        // [total | pad] [start rows] [count rows] [pair cap] [w cap]
        long long cap = need_rows * ldc / 4;
        if (cap > 0x3fffffffLL) cap = 0x3fffffffLL;
        if (cap < 1024) cap = 1024;
        const size_t bytes = 16 + static_cast<size_t>(need_rows) * 8 + static_cast<size_t>(cap) * 16;
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[3], bytes));
        char *base = ctx->scratch[3].as<char>();
        cl.total = reinterpret_cast<int *>(base);
        cl.cap = static_cast<int>(cap);
        cl.start = reinterpret_cast<int *>(base + 16);
        cl.count = cl.start + need_rows;
        cl.pair = reinterpret_cast<int2 *>(cl.count + need_rows);
        cl.w = reinterpret_cast<double *>(cl.pair + cap);
        ctx->cand_valid = true;
        ctx->cand_cap = cap;
        const int nq = S < kHeadSamples ? S : kHeadSamples;
        // [head table 2 nq ldc doubles] [its float2 copy nq ldc] [tile minima need_rows x b_tiles]
        MCG_TRY(mcg_reserve(ctx, ctx->scratch[1], sizeof(double) * (3 * nq * ldc + need_rows * b_tiles)));
        headf = reinterpret_cast<float2 *>(ctx->scratch[1].as<double>() + 2 * nq * ldc);
        tilemin = reinterpret_cast<unsigned long long *>(ctx->scratch[1].as<double>() + 3 * nq * ldc);
        bin_head_kernel<<<blocks_for(b.n * nq, 256), 256, 0, ctx->stream>>>(
            b, S, nq, ldc, ctx->scratch[1].as<double>(), headf);
        MCG_KERNEL_CHECK(ctx);
    }
    // distances for the bound: the tensor-core estimate (default) or dist2_kernel's exact matrix
    const bool use_tc = use_cand && ctx->tc;
    const long long ldf = mcg_tc_ldf(b.n);
    float *d2a = nullptr, *mrow = nullptr;
    BoundPre *pre = nullptr;
    int *todo = nullptr;
    // MCG_BOUND_SINGLE: bound_kernel without the lane-per-contig pre-pass (A/B switch; measured
    // with it: config 2 probability phase 0.280 -> 0.151 ms, config-3 shard 0.98 -> 0.84 ms,
    // config-5 chunk 1.05 -> 1.00 ms)
    const bool bound_single = getenv("MCG_BOUND_SINGLE") != nullptr;
    if (use_tc) {
        const long long rows_pad = (need_rows + 127) / 128 * 128;
        MCG_TRY(mcg_reserve(ctx, ctx->tc_d2a, sizeof(float) * rows_pad * ldf));
        // [margin need_rows floats, padded to 16 bytes] [BoundPre need_rows]
        const size_t mrow_bytes = (sizeof(float) * need_rows + 15) / 16 * 16;
        MCG_TRY(mcg_reserve(ctx, ctx->tc_mrow,
                            mrow_bytes + sizeof(BoundPre) * need_rows + sizeof(int) * (need_rows + 4)));
        d2a = ctx->tc_d2a.as<float>();
        mrow = ctx->tc_mrow.as<float>();
        pre = reinterpret_cast<BoundPre *>(ctx->tc_mrow.as<char>() + mrow_bytes);
        todo = reinterpret_cast<int *>(pre + need_rows);
        MCG_TRY(mcg_k_tc_prep_bins(ctx, b));
    }
    const CandGate gate = {cl.total, cl.cap};
    ScoreOut shifted = out;   // outputs are indexed by (first + r): shift them back
    shifted.argmin = out.argmin - a.first;
    shifted.wmin = out.wmin - a.first;
    if (out.weights) shifted.weights = out.weights - a.first * b.n;
    for (long long r0 = 0; r0 < a.n; r0 += rows_per_chunk) {
        ScoreSide ac = a;
        ac.first = a.first + r0;
        ac.n = a.n - r0 < rows_per_chunk ? a.n - r0 : rows_per_chunk;
        const long long a_tiles = (ac.n + kTile - 1) / kTile;
        // one B tile per CTA (up to 16 slices): short CTAs fill the last wave better than CTAs
        // that walk every B tile (measured: config 2 0.270 -> 0.242 ms, config-3 shard
        // 0.683 -> 0.637 ms); MCG_DIST_GY overrides (experiments)
        long long gy = b_tiles < 16 ? b_tiles : 16;
        if (const char *env = getenv("MCG_DIST_GY")) gy = atoi(env);
        if (gy > b_tiles) gy = b_tiles;
        if (gy < 1) gy = 1;
        dim3 grid(static_cast<unsigned>(a_tiles), static_cast<unsigned>(gy), 1);
        mcg_timer_begin(ctx, T_DIST);
        if (use_tc) {
            MCG_TRY(mcg_k_tc_dist(ctx, ac, b.n, d2a, ldf, mrow, tilemin, b_tiles));
        } else {
            if (tilemin)   // 0x7f7f...: a finite number above every d^2
                MCG_CUDA(ctx, cudaMemsetAsync(tilemin, 0x7f,
                                              sizeof(unsigned long long) * need_rows * b_tiles,
                                              ctx->stream));
            dist2_kernel<0><<<grid, kPairThreads, kDistSmem, ctx->stream>>>(ac, b, d2, ldc, tilemin,
                                                                           b_tiles, nullptr, 0);
            MCG_KERNEL_CHECK(ctx);
        }
        mcg_timer_end(ctx, T_DIST);
        long long blocks = (ac.n + warps * kProbRows - 1) / (warps * kProbRows);
        long long cap = 4LL * ctx->sm_count * 4;
        if (use_cand) {
            // Behind the candidate path these launches leave at once unless the list overflowed
            // (device-side gate), one launch per bin group: 31 groups x 10 chunks of 2048 CTAs
            // with 200 KB of shared memory each were 1.8 ms = 10 % of a config-5 shard step
            // (ncu launch list r02K).  One resident wave is enough: the kernel strides over its
            // rows, so the real fallback runs the same work on fewer, longer CTAs.
            const long long per_sm = prob_smem > 110 * 1024 ? 1 : (prob_smem > 56 * 1024 ? 2 : 4);
            cap = per_sm * ctx->sm_count;
        }
        if (blocks > cap) blocks = cap;
        mcg_timer_begin(ctx, T_PROB);
        if (use_cand) {
            ctx->cand_pairs = ac.n * b.n;
            MCG_CUDA(ctx, cudaMemsetAsync(cl.total, 0, 16, ctx->stream));
            long long bb = (ac.n + 7) / 8;
            if (bb > 8LL * ctx->sm_count) bb = 8LL * ctx->sm_count;
            if (use_tc && !bound_single) {
                MCG_CUDA(ctx, cudaMemsetAsync(todo, 0, sizeof(int), ctx->stream));
                bound_pre_kernel<<<blocks_for(ac.n, 128), 128, 0, ctx->stream>>>(
                    ac, b, S, d2a, ldf, mrow, tilemin, pre, d_status, cl, todo);
                MCG_KERNEL_CHECK(ctx);
                bound_kernel<float, true><<<static_cast<unsigned>(bb), 256, 0, ctx->stream>>>(
                    ac, b, S, d2a, ldf, mrow, ctx->scratch[1].as<double>(), ldc, tilemin, cl,
                    d_status, pre, headf, todo);
            } else if (use_tc)
                bound_kernel<float, false><<<static_cast<unsigned>(bb), 256, 0, ctx->stream>>>(
                    ac, b, S, d2a, ldf, mrow, ctx->scratch[1].as<double>(), ldc, tilemin, cl,
                    d_status, nullptr, headf, nullptr);
            else
                bound_kernel<double, false><<<static_cast<unsigned>(bb), 256, 0, ctx->stream>>>(
                    ac, b, S, d2, ldc, nullptr, ctx->scratch[1].as<double>(), ldc, tilemin, cl,
                    d_status, nullptr, headf, nullptr);
            MCG_KERNEL_CHECK(ctx);
            const unsigned sb = static_cast<unsigned>(8 * ctx->sm_count);
            if (guard) {
                if (use_tc)
                    cand_score_kernel<true, true><<<sb, 256, 0, ctx->stream>>>(ac, b, S, nullptr, 0, cl,
                                                                              d_status);
                else
                    cand_score_kernel<true, false><<<sb, 256, 0, ctx->stream>>>(ac, b, S, d2, ldc, cl,
                                                                               d_status);
                MCG_KERNEL_CHECK(ctx);
                cand_argmin_kernel<true><<<blocks_for(ac.n, 256), 256, 0, ctx->stream>>>(
                    ac, cl, shifted, redo);
            } else {
                if (use_tc)
                    cand_score_kernel<false, true><<<sb, 256, 0, ctx->stream>>>(ac, b, S, nullptr, 0,
                                                                               cl, d_status);
                else
                    cand_score_kernel<false, false><<<sb, 256, 0, ctx->stream>>>(ac, b, S, d2, ldc, cl,
                                                                                d_status);
                MCG_KERNEL_CHECK(ctx);
                cand_argmin_kernel<false><<<blocks_for(ac.n, 256), 256, 0, ctx->stream>>>(
                    ac, cl, shifted, redo);
            }
            MCG_KERNEL_CHECK(ctx);
            if (use_tc) {
                // the candidate list overflowed (a device-side fact): the exact matrix after
                // all, for prob_kernel below; one early-exit CTA wave otherwise
                dist2_kernel<0><<<grid, kPairThreads, kDistSmem, ctx->stream>>>(
                    ac, b, d2, ldc, nullptr, b_tiles, cl.total, cl.cap);
                MCG_KERNEL_CHECK(ctx);
            }
        }
        for (long long g0 = 0; g0 < b.n; g0 += group) {
            const int ncols = static_cast<int>(b.n - g0 < group ? b.n - g0 : group);
            const int ncols_pad = (ncols + 31) / 32 * 32;
            const size_t smem_g = prob_smem - sizeof(double) * pitch * (group - ncols);
            if (guard && warps == 16)
                prob_kernel<true, 16><<<static_cast<unsigned>(blocks), 512, smem_g, ctx->stream>>>(
                    ac, b, S, d2, ldc, g0, ncols, ncols_pad, pitch, shifted, d_status, redo, gate);
            else if (guard)
                prob_kernel<true, 8><<<static_cast<unsigned>(blocks), kProbThreads, smem_g,
                                       ctx->stream>>>(ac, b, S, d2, ldc, g0, ncols, ncols_pad,
                                                      pitch, shifted, d_status, redo, gate);
            else
                prob_kernel<false, 8><<<static_cast<unsigned>(blocks), kProbThreads, smem_g,
                                        ctx->stream>>>(ac, b, S, d2, ldc, g0, ncols, ncols_pad,
                                                       pitch, shifted, d_status, redo, gate);
            MCG_KERNEL_CHECK(ctx);
        }
        // (the timer holds one interval: with S > 30 it closes after the redo pass below)
        if (!(guard && r0 + rows_per_chunk >= a.n)) mcg_timer_end(ctx, T_PROB);
    }
    if (guard) {
        // the listed contigs again, in the reference's operation order; the grid covers the
        // worst case and the CTAs beyond the device-side count leave at once
        ScoreSide ar = a;
        ar.index = reinterpret_cast<const int64_t *>(redo.rows);
        ar.first = 0;
        ScoreOut ro;
        ro.argmin = redo_arg;
        ro.wmin = redo_w;
        const long long a_tiles = (a.n + kTile - 1) / kTile;
        pair_variant(true, false)<<<dim3(static_cast<unsigned>(a_tiles),
                                         static_cast<unsigned>(redo_parts), 1),
                                    kPairThreads, kPairSmem, ctx->stream>>>(
            ar, b, S, ro, d_status, redo.count, redo_parts > 1 ? a.n : 0);
        MCG_KERNEL_CHECK(ctx);
        scatter_redo_kernel<<<blocks_for(a.n, 256), 256, 0, ctx->stream>>>(
            redo.count, redo.pos, redo_arg, redo_w, redo_parts, a.n, shifted.argmin,
            shifted.wmin);
        MCG_KERNEL_CHECK(ctx);
        mcg_timer_end(ctx, T_PROB);
    }
    return MCG_OK;
}

int mcg_k_row_norms(mcg_ctx *ctx, const double *d_prof, int64_t rows, double *d_norm) {
    if (rows == 0) return MCG_OK;
    row_norms_kernel<<<blocks_for(((rows + 7) / 8) * 32, 256), 256, 0, ctx->stream>>>(
        d_prof, nullptr, nullptr, nullptr, rows, d_norm);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

// counts -> normalised profiles and their squared norms in one pass
int mcg_k_normalise_norms(mcg_ctx *ctx, const uint32_t *d_counts, const ContigMeta *d_meta,
                          int64_t rows, double *d_prof, double *d_norm) {
    if (rows == 0) return MCG_OK;
    row_norms_kernel<<<blocks_for(((rows + 7) / 8) * 32, 256), 256, 0, ctx->stream>>>(
        nullptr, d_counts, d_meta, d_prof, rows, d_norm);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_pairs(mcg_ctx *ctx, const ScoreSide &a, const ScoreSide &b, int n_samples,
                const ScoreOut &out, int32_t *d_status) {
    MCG_TRY(ensure_gauss(ctx));
    ctx->redo_valid = false;
    if (a.n == 0) return MCG_OK;
    const bool rule_c = out.argmin != nullptr || out.wmin != nullptr || out.weights != nullptr;
    const long long a_tiles = (a.n + kTile - 1) / kTile;
    const long long b_tiles = (b.n + kTile - 1) / kTile;
    if (a_tiles > 0x7fffffffLL) return mcg_fail(ctx, MCG_ERR_ARG, "too many query rows");
    dim3 grid(static_cast<unsigned>(a_tiles), 1, 1);
    const bool logsum = n_samples <= 30;
    // big rule-C batches: distance GEMM + probability kernel (argmin/wmin are needed to
    // carry the running minimum when the bins do not fit one shared-memory group)
    // (S > 30: only without the weight matrix -- deferred pairs have no exact weight there)
    if (rule_c && (a.n >= 2048 || ctx->force_split || ctx->force_split_calls) && out.argmin &&
        out.wmin &&
        (logsum || (!out.weights && split_fits(n_samples))))
        return split_rule_c(ctx, a, b, n_samples, out, d_status);
    if (out.dist) {
        pair_distance_kernel<<<blocks_for(a.n * b.n, 128), 128, 0, ctx->stream>>>(a, b, out.dist);
        MCG_KERNEL_CHECK(ctx);
    }
    if (!rule_c) {
        // few query tiles: spread the B tiles over grid.y to fill the SMs
        long long gy = (2LL * ctx->sm_count + a_tiles - 1) / a_tiles;
        if (gy > b_tiles) gy = b_tiles;
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        grid.y = static_cast<unsigned>(gy);
    }
    pair_variant(rule_c, logsum)<<<grid, kPairThreads, kPairSmem, ctx->stream>>>(
        a, b, n_samples, out, d_status, nullptr, 0);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_aggregate_members(mcg_ctx *ctx, const double *d_lp, int64_t nq, int64_t nm,
                            const int64_t *d_seg_offsets, int64_t nseg, int rule,
                            double *d_weights) {
    if (nq * nseg == 0) return MCG_OK;
    aggregate_members_kernel<<<blocks_for(nq * nseg, 128), 128, 0, ctx->stream>>>(
        d_lp, nq, nm, d_seg_offsets, nseg, rule, d_weights);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_member_items(mcg_ctx *ctx, const ScoreSide &table, const int64_t *d_item_query,
                       const int32_t *d_item_bin, int64_t n_items, const int64_t *d_members,
                       const int64_t *d_seg_offsets, int n_samples, int rule, double *d_weights,
                       int32_t *d_status) {
    MCG_TRY(ensure_gauss(ctx));
    if (n_items == 0) return MCG_OK;
    long long blocks = (n_items + 7) / 8;
    if (blocks > 16LL * ctx->sm_count) blocks = 16LL * ctx->sm_count;
    member_items_kernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(
        table, d_item_query, d_item_bin, n_items, d_members, d_seg_offsets, n_samples, rule,
        d_weights, d_status);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_segment_mean(mcg_ctx *ctx, const double *d_rows, int width, const int64_t *d_members,
                       const int64_t *d_seg_offsets, int64_t nseg, double *d_out) {
    if (nseg * width == 0) return MCG_OK;
    segment_mean_kernel<<<blocks_for(nseg * width, 128), 128, 0, ctx->stream>>>(
        d_rows, width, d_members, d_seg_offsets, nseg, d_out);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_normpdf(mcg_ctx *ctx, const double *d_x, double mean, double sd, double *d_out,
                  int64_t n) {
    if (n == 0) return MCG_OK;
    const double two_var = 2 * (sd * sd);
    const double denom = sd * sqrt(2 * 3.141592653589793);
    normpdf_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_x, mean, two_var, denom, d_out,
                                                                n);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_distance(mcg_ctx *ctx, const double *d_u, const double *d_v, int64_t n, int width,
                   double *d_out) {
    if (n == 0) return MCG_OK;
    distance_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(d_u, d_v, n, width, d_out);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_comp_prob(mcg_ctx *ctx, const double *d_dist, double *d_out, int64_t n,
                    int32_t *d_status) {
    MCG_TRY(ensure_gauss(ctx));
    if (n == 0) return MCG_OK;
    comp_prob_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_dist, d_out, n, d_status);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_cov_prob(mcg_ctx *ctx, const double *d_c1, const double *d_c2, int64_t n, int S,
                   double *d_out) {
    MCG_TRY(ensure_gauss(ctx));
    if (n == 0) return MCG_OK;
    cov_prob_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(d_c1, d_c2, n, S, d_out);
    MCG_KERNEL_CHECK(ctx);
    return MCG_OK;
}

int mcg_k_fp64_peak(mcg_ctx *ctx, double *tflops) {
    const int iters = 4096;
    const int blocks = ctx->sm_count * 8, threads = 256;
    MCG_TRY(mcg_reserve(ctx, ctx->scratch[5], 64));
    cudaEvent_t e0, e1;
    MCG_CUDA(ctx, cudaEventCreate(&e0));
    MCG_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        MCG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->scratch[5].as<double>(), iters,
                                                              1.0 + rep);
        MCG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        MCG_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        MCG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return MCG_OK;
}

// The candidate list of the last rule-C chunk (bound_kernel): [0] pairs listed, [1] capacity of
// the list, [2] all pairs of the chunk, [3] 1 when the list was used (0: it overflowed and
// prob_kernel scored every pair, or pruning was off).
int mcg_k_prune_stats(mcg_ctx *ctx, int64_t stats[4]) {
    for (int i = 0; i < 4; ++i) stats[i] = 0;
    if (!ctx->cand_valid) return MCG_OK;
    int32_t total = 0;
    MCG_CUDA(ctx, cudaMemcpyAsync(&total, ctx->scratch[3].p, sizeof total, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    stats[0] = total;
    stats[1] = ctx->cand_cap;
    stats[2] = ctx->cand_pairs;
    stats[3] = total <= ctx->cand_cap ? 1 : 0;
    return MCG_OK;
}
