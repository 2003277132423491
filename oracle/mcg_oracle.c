/*
 * oracle/mcg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the arithmetic on MetaCoAG's featurise+score path
 * (reference v1.1.5, pure Python).  It exists so that the CUDA path in
 * metacoag_b200/csrc can be checked against something that runs on any CPU.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against tests/golden/ (npz files), which tests/golden/make_golden.py produced
 * by importing the reference's own modules from /root/reference in the build
 * container (the reference ships no function-level golden vectors of its own,
 * SURVEY.md section 8c).
 *
 * libm note: CPython's math.exp/log/lgamma/pow call the C library, so using
 * glibc here reproduces the reference bit for bit on the same host.
 * Build with -ffp-contract=off (see oracle/Makefile): the reference rounds
 * after every Python-level operation.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_NPROF 136

/* ---- tiny pthread parallel-for (this image has no OpenMP runtime) --------- */

typedef void (*orc_body_fn)(int64_t i, void *ctx);
typedef struct {
    orc_body_fn body;
    void *ctx;
    int64_t n, chunk;
    int64_t next;      /* claimed with __atomic_fetch_add */
} orc_pf_t;

static void *orc_pf_worker(void *arg)
{
    orc_pf_t *pf = (orc_pf_t *)arg;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&pf->next, pf->chunk, __ATOMIC_RELAXED);
        if (lo >= pf->n) break;
        int64_t hi = lo + pf->chunk < pf->n ? lo + pf->chunk : pf->n;
        for (int64_t i = lo; i < hi; ++i) pf->body(i, pf->ctx);
    }
    return NULL;
}

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void orc_parallel_for(int64_t n, int64_t chunk, int nthreads, orc_body_fn body, void *ctx)
{
    if (nthreads <= 0) nthreads = orc_max_threads();
    if (nthreads > 256) nthreads = 256;
    orc_pf_t pf = {body, ctx, n, chunk > 0 ? chunk : 1, 0};
    if (nthreads == 1 || n <= pf.chunk) { orc_pf_worker(&pf); return; }
    pthread_t tid[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t)
        if (pthread_create(&tid[started], NULL, orc_pf_worker, &pf) == 0) ++started;
    orc_pf_worker(&pf);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

/* ---- featurisation ------------------------------------------------------ */

/* feature_utils.py:21,31-35 -- nt_bits.get(c, 0): A0 C1 G2 T3, anything else 0
 * (lower case and N included). */
static inline unsigned orc_base_code(uint8_t c)
{
    switch (c) {
    case 'C': return 1u;
    case 'G': return 2u;
    case 'T': return 3u;
    default:  return 0u;
    }
}

/* feature_utils.py:25-28 -- reverse complement of a 2-bit packed 4-mer. */
static unsigned orc_rc4(unsigned code)
{
    unsigned out = 0;
    for (int i = 0; i < 4; ++i) {
        out = (out << 2) | (3u - (code & 3u));
        code >>= 2;
    }
    return out;
}

/* feature_utils.py:38-57 compute_kmer_inds(4): walk the 4-mers in sorted
 * (= ascending code) order; a code whose reverse complement was already seen
 * shares its slot, otherwise it opens the next slot.  Returns slot count. */
int orc_kmer_table(uint8_t *table /* [256] */)
{
    int16_t seen[256];
    int next = 0;
    for (int i = 0; i < 256; ++i) seen[i] = -1;
    for (unsigned code = 0; code < 256; ++code) {
        unsigned rc = orc_rc4(code);
        if (seen[rc] >= 0) {
            seen[code] = seen[rc];
        } else {
            seen[code] = (int16_t)next++;
        }
        table[code] = (uint8_t)seen[code];
    }
    return next;
}

/* feature_utils.py:60-70 count_kmers with k=4 (the caller strips whitespace,
 * as str.strip() does there).  counts are whole numbers held in float64 like
 * the reference's np.zeros profile; freqs = counts / max(1, sum(counts)). */
void orc_count_kmers(const uint8_t *seq, int64_t len, double *counts, double *freqs)
{
    static uint8_t table[256];
    static int have_table = 0;
    if (!have_table) { orc_kmer_table(table); have_table = 1; }
    uint64_t tally[ORC_NPROF];
    memset(tally, 0, sizeof tally);
    for (int64_t i = 0; i + 4 <= len; ++i) {
        unsigned code = (orc_base_code(seq[i]) << 6) | (orc_base_code(seq[i + 1]) << 4) |
                        (orc_base_code(seq[i + 2]) << 2) | orc_base_code(seq[i + 3]);
        tally[table[code]]++;
    }
    double total = 0.0;
    for (int k = 0; k < ORC_NPROF; ++k) { counts[k] = (double)tally[k]; total += counts[k]; }
    double denom = total > 1.0 ? total : 1.0;   /* max(1, sum(profile)) */
    for (int k = 0; k < ORC_NPROF; ++k) freqs[k] = counts[k] / denom;
}

/* Batch form over a concatenated blob; offsets has n+1 entries. */
typedef struct { const uint8_t *blob; const int64_t *offsets; double *counts, *freqs; } orc_ck_t;
static void orc_ck_body(int64_t i, void *p)
{
    orc_ck_t *a = (orc_ck_t *)p;
    double c[ORC_NPROF];
    orc_count_kmers(a->blob + a->offsets[i], a->offsets[i + 1] - a->offsets[i],
                    a->counts ? a->counts + i * ORC_NPROF : c, a->freqs + i * ORC_NPROF);
}
void orc_count_kmers_batch(const uint8_t *blob, const int64_t *offsets, int64_t n,
                           double *counts /* [n,136] or NULL */, double *freqs /* [n,136] */,
                           int nthreads)
{
    {   /* prime the static table before threads start */
        double c[ORC_NPROF], f[ORC_NPROF];
        orc_count_kmers(blob, 0, c, f);
    }
    orc_ck_t a = {blob, offsets, counts, freqs};
    orc_parallel_for(n, 16, nthreads, orc_ck_body, &a);
}

/* feature_utils.py:150-153 -- coverage clamp (VERY_SMALL_VAL = 1e-4). */
void orc_clamp_coverage(double *cov, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
        if (cov[i] < 0.0001) cov[i] = 0.0001;
}

/* feature_utils.py:217-235 get_bin_profiles: np.mean(rows, axis=0) adds the
 * member rows one after another (outer-axis reduction) and divides once. */
void orc_segment_mean(const double *rows /* [n,width] */, const int64_t *members,
                      const int64_t *seg_offsets /* [nseg+1] */, int64_t nseg, int width,
                      double *out /* [nseg,width] */)
{
    for (int64_t b = 0; b < nseg; ++b) {
        int64_t lo = seg_offsets[b], hi = seg_offsets[b + 1];
        double *dst = out + b * width;
        for (int k = 0; k < width; ++k) {
            double acc = 0.0;
            for (int64_t j = lo; j < hi; ++j) {
                double v = rows[members[j] * width + k];
                acc = (j == lo) ? v : acc + v;
            }
            dst[k] = acc / (double)(hi - lo);
        }
    }
}

/* ---- scoring primitives ------------------------------------------------- */

/* matching_utils.py:12-15 */
#define ORC_MU_INTRA 0.0
#define ORC_SIGMA_INTRA (0.01037897 / 2)
#define ORC_MU_INTER 0.0676654
#define ORC_SIGMA_INTER 0.03419337
#define ORC_FLOOR 1e-10
#define ORC_MAX_WEIGHT DBL_MAX

/* numpy's pairwise summation (what np.add.reduce does on a contiguous
 * float64 vector), restated so that distances match the installed SciPy
 * (1.18: euclidean -> minkowski -> norm(u-v, ord=2) -> sqrt(add.reduce(s*s))). */
static double orc_pairwise(const double *a, int64_t n)
{
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t half = n / 2;
    half -= half % 8;
    return orc_pairwise(a, half) + orc_pairwise(a + half, n - half);
}

/* matching_utils.py:28-29 get_tetramer_distance */
double orc_euclidean(const double *u, const double *v, int n)
{
    double sq[1024];
    if (n > 1024) n = 1024;
    for (int i = 0; i < n; ++i) { double d = u[i] - v[i]; sq[i] = d * d; }
    return sqrt(orc_pairwise(sq, n));
}

/* matching_utils.py:21-25 normpdf; Python's ** is libm pow (built with
 * -fno-builtin-pow: gcc would turn pow(x, 2.0) into x*x, which differs from
 * glibc's pow by an ulp for some x and so from the reference). */
double orc_normpdf(double x, double mean, double sd)
{
    double var = pow(sd, 2.0);
    double denom = sd * pow(2 * M_PI, 0.5);
    double num = exp(-(pow(x - mean, 2.0)) / (2 * var));
    return num / denom;
}

/* matching_utils.py:36-39 get_comp_probability (0/0 when both tails underflow:
 * the reference raises ZeroDivisionError there; here the NaN is returned). */
double orc_comp_prob(double dist)
{
    double gi = orc_normpdf(dist, ORC_MU_INTRA, ORC_SIGMA_INTRA);
    double ge = orc_normpdf(dist, ORC_MU_INTER, ORC_SIGMA_INTER);
    return gi / (gi + ge);
}


/* CPython's math.lgamma is NOT libm's: Modules/mathmodule.c (CPython 3.12,
 * the interpreter the reference runs under; not vendored in /root/reference)
 * evaluates Lanczos' formula with the Boost lanczos13m53 rational, g =
 * 6.0246800407767296.  Restated from the published algorithm for x > 0 (the
 * only range coverage values reach); pinned bit-for-bit against math.lgamma
 * in tests/test_oracle_golden.py. */
static const double orc_lanczos_g = 6.024680040776729583740234375;
static const double orc_lanczos_num[13] = {
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
static const double orc_lanczos_den[13] = {
    0.0, 39916800.0, 120543840.0, 150917976.0, 105258076.0, 45995730.0, 13339535.0,
    2637558.0, 357423.0, 32670.0, 1925.0, 66.0, 1.0};

static double orc_lanczos_sum(double x)
{
    double num = 0.0, den = 0.0;
    if (x < 5.0) {
        for (int i = 13; --i >= 0;) {
            num = num * x + orc_lanczos_num[i];
            den = den * x + orc_lanczos_den[i];
        }
    } else {
        for (int i = 0; i < 13; ++i) {
            num = num / x + orc_lanczos_num[i];
            den = den / x + orc_lanczos_den[i];
        }
    }
    return num / den;
}

double orc_lgamma(double x)
{
    if (!(x > 0.0) || isinf(x)) return lgamma(x);          /* outside the path's domain */
    if (x == floor(x) && x <= 2.0) return 0.0;             /* lgamma(1) = lgamma(2) = 0 */
    if (x < 1e-20) return -log(x);
    double r = log(orc_lanczos_sum(x)) - orc_lanczos_g;
    r += (x - 0.5) * (log(x + orc_lanczos_g - 0.5) - 1);
    return r;
}

/* matching_utils.py:42-66 get_cov_probability (math.lgamma -> orc_lgamma) */
double orc_cov_prob(const double *c1, const double *c2, int S)
{
    double p1 = 1.0, p2 = 1.0;
    for (int i = 0; i < S; ++i) {
        double a = exp((c1[i] * log(c2[i])) - orc_lgamma(c1[i] + 1.0) - c2[i]);
        double b = exp((c2[i] * log(c1[i])) - orc_lgamma(c2[i] + 1.0) - c1[i]);
        if (a < ORC_FLOOR) a = ORC_FLOOR;
        if (b < ORC_FLOOR) b = ORC_FLOOR;
        p1 = p1 * a;
        p2 = p2 * b;
    }
    return p1 < p2 ? p1 : p2;   /* min(p1, p2): p2 only if strictly smaller */
}

/* One (query, target) pair.  math.log(x, 10) is log(x)/log(10).
 * Returns 1 and *lp when prob_comp*prob_cov > 0.0, else 0 (callers apply
 * their own rule for the "invalid" case). */
int orc_pair_lp(const double *prof_q, const double *prof_t, const double *cov_q,
                const double *cov_t, int S, double *lp, double *pc_out, double *pv_out)
{
    double d = orc_euclidean(prof_q, prof_t, ORC_NPROF);
    double pc = orc_comp_prob(d);
    double pv = orc_cov_prob(cov_q, cov_t, S);
    if (pc_out) *pc_out = pc;
    if (pv_out) *pv_out = pv;
    if (pc * pv > 0.0) {
        *lp = -(log(pc) / log(10.0) + log(pv) / log(10.0));
        return 1;
    }
    *lp = ORC_MAX_WEIGHT;
    return 0;
}

/* ---- aggregation rules ------------------ */

/* Rules A and B share the member loop.  rule 0 = A (matching_utils.py:117-155,
 * 368-398: divide by all members); rule 1 = B (label_prop_utils.py:54-86:
 * divide by the members whose product was positive).
 * prof [n,136], cov [n,S] are the dense contig tables; members/seg_offsets is
 * the CSR of bins; out is [nq, nseg]. */
typedef struct {
    const double *prof, *cov; int S; const int64_t *queries, *members, *seg_offsets;
    int64_t nseg; int rule; double *out;
} orc_sm_t;
static void orc_sm_body(int64_t qi, void *p)
{
    orc_sm_t *a = (orc_sm_t *)p;
    int S = a->S;
    int64_t q = a->queries[qi];
    for (int64_t b = 0; b < a->nseg; ++b) {
        double sum = 0.0;
        int64_t valid = 0, total = a->seg_offsets[b + 1] - a->seg_offsets[b];
        for (int64_t j = a->seg_offsets[b]; j < a->seg_offsets[b + 1]; ++j) {
            int64_t t = a->members[j];
            double lp;
            valid += orc_pair_lp(a->prof + q * ORC_NPROF, a->prof + t * ORC_NPROF, a->cov + q * S,
                                 a->cov + t * S, S, &lp, NULL, NULL);
            sum += lp;
        }
        double w;
        if (a->rule == 0)
            w = (sum != INFINITY) ? sum / (double)total : ORC_MAX_WEIGHT;
        else
            w = (sum != INFINITY && valid != 0) ? sum / (double)valid : ORC_MAX_WEIGHT;
        a->out[qi * a->nseg + b] = w;
    }
}
void orc_score_members(const double *prof, const double *cov, int S, const int64_t *queries,
                       int64_t nq, const int64_t *members, const int64_t *seg_offsets,
                       int64_t nseg, int rule, double *out, int nthreads)
{
    orc_sm_t a = {prof, cov, S, queries, members, seg_offsets, nseg, rule, out};
    orc_parallel_for(nq, 1, nthreads, orc_sm_body, &a);
}

/* Rule C (label_prop_utils.py:308-345 assign_long): log_prob stays 0 when the
 * product is not positive, and an exact 0 becomes MAX_WEIGHT; the winner is
 * the first index holding the minimum; "None" (best = -1) when that minimum is
 * MAX_WEIGHT.  weights_out may be NULL. */
typedef struct {
    const double *prof, *cov; int S; const int64_t *queries; const double *cen_prof, *cen_cov;
    int64_t B; double *weights_out; int32_t *best; double *best_w;
} orc_sc_t;
static void orc_sc_body(int64_t qi, void *p)
{
    orc_sc_t *a = (orc_sc_t *)p;
    int S = a->S;
    int64_t q = a->queries ? a->queries[qi] : qi;
    double wmin = 0.0;
    int32_t arg = -1;
    for (int64_t b = 0; b < a->B; ++b) {
        double lp = 0.0, tmp;
        if (orc_pair_lp(a->prof + q * ORC_NPROF, a->cen_prof + b * ORC_NPROF, a->cov + q * S,
                        a->cen_cov + b * S, S, &tmp, NULL, NULL))
            lp = tmp;
        double w = (lp != 0.0) ? lp : ORC_MAX_WEIGHT;
        if (a->weights_out) a->weights_out[qi * a->B + b] = w;
        if (arg < 0 || w < wmin) { wmin = w; arg = (int32_t)b; }
    }
    if (arg >= 0 && wmin == ORC_MAX_WEIGHT) arg = -1;
    a->best[qi] = arg;
    a->best_w[qi] = (arg >= 0) ? wmin : ORC_MAX_WEIGHT;
}
void orc_score_centroids(const double *prof, const double *cov, int S, const int64_t *queries,
                         int64_t nq, const double *cen_prof /* [B,136] */,
                         const double *cen_cov /* [B,S] */, int64_t B, double *weights_out,
                         int32_t *best, double *best_w, int nthreads)
{
    orc_sc_t a = {prof, cov, S, queries, cen_prof, cen_cov, B, weights_out, best, best_w};
    orc_parallel_for(nq, 8, nthreads, orc_sc_body, &a);
}

/* Hoisted per-(contig, sample) terms, for checking the GPU's coverage
 * featurisation kernel: log(c) and lgamma(c + 1). */
void orc_cov_terms(const double *cov, int64_t n, double *logc, double *lgam)
{
    for (int64_t i = 0; i < n; ++i) {
        logc[i] = log(cov[i]);
        lgam[i] = orc_lgamma(cov[i] + 1.0);
    }
}

/* ---- synthetic contigs: the generator of csrc/mcg_scan.cu (synth_kernel) restated ----------
 * SURVEY.md section 8d workload.  bench.py --impl reference draws its bases with this function,
 * so that the reference arm needs neither the CUDA library nor a GPU; tests check that the two
 * generators write the same bytes.  Positions [p_begin, p_end) of the blob, whole 256-base
 * chunks (p_begin a multiple of 256). */
static uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

typedef struct {
    const int64_t *offsets;
    int64_t n_contigs;
    const int32_t *genome_of;
    const float *trans;
    uint64_t seed;
    float n_run_prob, lower_frac;
    uint8_t *bases;
    int64_t n_bases, chunk0;
} orc_synth_t;

static void orc_synth_body(int64_t i, void *p)
{
    orc_synth_t *a = (orc_synth_t *)p;
    const int64_t chunk = a->chunk0 + i;
    const int64_t p0 = chunk * 256;
    if (p0 >= a->n_bases) return;
    const int64_t p1 = p0 + 256 < a->n_bases ? p0 + 256 : a->n_bases;
    int64_t lo = 0, hi = a->n_contigs;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a->offsets[mid] <= p0) lo = mid; else hi = mid;
    }
    int64_t c = lo, cend = a->offsets[c + 1];
    const uint64_t key = orc_splitmix64(a->seed ^ ((uint64_t)chunk * 0xD6E8FEB86659FD93ull));
    int n_lo = 1 << 30, n_hi = 0;
    if ((float)(key & 0xFFFFFF) * (1.0f / 16777216.0f) < a->n_run_prob) {
        n_lo = (int)((key >> 24) & 0xFF);
        n_hi = n_lo + 1 + (int)((key >> 32) % 50);
    }
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    uint32_t prev = (uint32_t)(key >> 40) & 3u;
    uint64_t bits = 0, ctr = 0;
    int have = 0;
    for (int64_t q = p0; q < p1; ++q) {
        while (q >= cend) { ++c; cend = a->offsets[c + 1]; }
        const int g = a->genome_of[c];
        if (have == 0) { bits = orc_splitmix64(key + (++ctr)); have = 4; }
        const float uni = (float)(bits & 0xFFFF) * (1.0f / 65536.0f);
        bits >>= 16;
        --have;
        const float *t = a->trans + (g * 4 + prev) * 4;
        uint32_t b = 3;
        const float c0 = t[0], c1 = c0 + t[1], c2 = c1 + t[2];
        if (uni < c0) b = 0; else if (uni < c1) b = 1; else if (uni < c2) b = 2;
        prev = b;
        uint8_t ch = (uint8_t)alphabet[b];
        const int rel = (int)(q - p0);
        if (rel >= n_lo && rel < n_hi) ch = 'N';
        const uint64_t ck = orc_splitmix64(a->seed * 31ull + (uint64_t)c);
        if ((float)(ck & 0xFFFFFF) * (1.0f / 16777216.0f) < a->lower_frac) ch |= 0x20;
        a->bases[q] = ch;
    }
}

void orc_synth_bases(const int64_t *offsets, int64_t n_contigs, const int32_t *genome_of,
                     const float *trans, uint64_t seed, double n_frac, double lower_frac,
                     uint8_t *bases, int64_t p_begin, int64_t p_end, int nthreads)
{
    orc_synth_t a;
    a.offsets = offsets; a.n_contigs = n_contigs; a.genome_of = genome_of; a.trans = trans;
    a.seed = seed;
    a.n_run_prob = (float)(n_frac * 256.0 / 25.5);
    a.lower_frac = (float)lower_frac;
    a.bases = bases;
    a.n_bases = p_end < offsets[n_contigs] ? p_end : offsets[n_contigs];
    a.chunk0 = p_begin / 256;
    const int64_t chunks = (a.n_bases + 255) / 256 - a.chunk0;
    if (chunks > 0) orc_parallel_for(chunks, 64, nthreads, orc_synth_body, &a);
}

/* ---- bounded walk to labelled contigs: label_prop_utils.py:37-53, 92-100 restated ------------
 * FIFO queue WITH duplicates (membership of the queue is not checked), depth[neighbour]
 * overwritten on every sighting of an unvisited neighbour, labelled nodes (other than the
 * start) reported at every pop and not expanded.  hits: (node, bin, depth) triples in pop
 * order, at most max_hits of them written; returns the number of hits, -1 when out of memory.
 * Pinned against tests/golden/score.npz (the reference's run_bfs_long) by
 * tests/test_oracle_golden.py. */
/* scratch: depth int32[n_vertices] (any content), visited uint8[n_vertices] (all zero on entry
 * and on return: only the touched entries are cleared), so a batch of walks costs its walks */
static int64_t orc_bfs_long_scratch(const int64_t *indptr, const int32_t *indices,
                                    const int32_t *labels, int64_t start, int depth_limit,
                                    int32_t *hits, int64_t max_hits, int32_t *depth, uint8_t *visited)
{
    int64_t cap = 1024, head = 0, tail = 0, n_hits = 0, n_visited = 0;
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * cap);
    if (!queue) return -1;
    queue[tail++] = (int32_t)start;
    depth[start] = 0;
    while (head < tail) {
        const int32_t v = queue[head++];
        if (!visited[v]) { visited[v] = 1; ++n_visited; }
        if (labels[v] >= 0 && n_visited > 1) {
            if (n_hits < max_hits) {
                hits[3 * n_hits] = v;
                hits[3 * n_hits + 1] = labels[v];
                hits[3 * n_hits + 2] = depth[v];
            }
            ++n_hits;
            continue;
        }
        for (int64_t e = indptr[v]; e < indptr[v + 1]; ++e) {
            const int32_t nb = indices[e];
            if (visited[nb]) continue;
            depth[nb] = depth[v] + 1;
            if (depth[nb] > depth_limit) continue;
            if (tail == cap) {
                cap *= 2;
                int32_t *grown = (int32_t *)realloc(queue, sizeof(int32_t) * cap);
                if (!grown) { free(queue); return -1; }
                queue = grown;
            }
            queue[tail++] = nb;
        }
    }
    for (int64_t i = 0; i < tail; ++i) visited[queue[i]] = 0;   /* every visited vertex was queued */
    free(queue);
    return n_hits;
}

int64_t orc_bfs_long(const int64_t *indptr, const int32_t *indices, int64_t n_vertices,
                     const int32_t *labels, int64_t start, int depth_limit, int32_t *hits,
                     int64_t max_hits)
{
    int32_t *depth = (int32_t *)malloc(sizeof(int32_t) * n_vertices);
    uint8_t *visited = (uint8_t *)calloc(n_vertices, 1);
    int64_t n = -1;
    if (depth && visited)
        n = orc_bfs_long_scratch(indptr, indices, labels, start, depth_limit, hits, max_hits, depth,
                                 visited);
    free(depth); free(visited);
    return n;
}

/* A batch of walks, one after the other on one thread with shared scratch (what a C port of the
 * reference's loop over start nodes costs): hit_count[i] hits of start i, the first max_hits of
 * each in hits[i * 3 * max_hits ...].  Returns 0, -1 when out of memory. */
int orc_bfs_long_batch(const int64_t *indptr, const int32_t *indices, int64_t n_vertices,
                       const int32_t *labels, const int64_t *starts, int64_t n_starts,
                       int depth_limit, int32_t *hits, int64_t max_hits, int64_t *hit_count)
{
    int32_t *depth = (int32_t *)malloc(sizeof(int32_t) * n_vertices);
    uint8_t *visited = (uint8_t *)calloc(n_vertices, 1);
    int rc = (depth && visited) ? 0 : -1;
    for (int64_t i = 0; rc == 0 && i < n_starts; ++i) {
        hit_count[i] = orc_bfs_long_scratch(indptr, indices, labels, starts[i], depth_limit,
                                            hits + 3 * max_hits * i, max_hits, depth, visited);
        if (hit_count[i] < 0) rc = -1;
    }
    free(depth); free(visited);
    return rc;
}
