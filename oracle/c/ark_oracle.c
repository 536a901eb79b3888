/*
 * Tier-1 oracle: arkworks-0.3.0-faithful CPU restatement of the distributed_plonk worker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under distributed_plonk_b200/ or include/ links, loads or
 * calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs do, as the checker / CPU baseline.
 *
 * PARITY UNPINNED by reference fixtures: /root/reference holds no golden vectors and is Rust
 * (cannot be compiled in this image: no cargo/rustc; its arithmetic lives in the third-party
 * crates ark-ff / ark-ec / ark-poly / ark-bls12-381 0.3.0, Cargo.toml:31-36, whose sources are
 * not vendored).  This file restates those crates' published algorithms; values are pinned by
 * the independent big-integer oracle oracle/py/bls12_381.py (tests/test_oracle.py) and by the
 * public BLS12-381 constants.
 *
 * What is restated (reference call site -> function here):
 *   src/worker.rs:66-94    fft1_helper ................. orc_fft1_helper
 *   src/worker.rs:96-115   fft2_helper ................. orc_fft2_helper
 *   src/worker.rs:117-123  commit_polynomial ........... orc_commit
 *   src/worker.rs:143-154  Radix2EvaluationDomain::new . domain_gen
 *   src/worker.rs:177-182  VariableBaseMSM ............. orc_msm
 *   src/worker.rs:327-330,432-435 exchange ............. orc_distributed_fft
 *   src/worker.rs:398      ifft_in_place ............... orc_fft_in_place
 *   src/dispatcher2.rs:731-787 Prover::fft ............. orc_distributed_fft
 *   src/playground.rs:67-80 coset_fft / coset_ifft ..... orc_coset_fft_in_place
 *   src/utils.rs:27-43     raw struct layouts .......... the byte formats of every argument
 *
 * ark-poly 0.3.0 radix-2: forward = DIF ("io_helper") with the n/2 roots table, then the
 * bit-reversal permutation ("derange"); inverse = derange, DIT ("oi_helper") with the inverse
 * roots, then * size_inv.  ark-ec 0.3.0 MSM: window c = 3 if n < 32 else ceil(log2 n)*69/100+2;
 * 2^c-1 buckets per window; zero scalars skipped; scalar == 1 added once in window 0;
 * running-sum bucket reduce; Horner combine with c doublings; rayon over windows (OpenMP here).
 * Jacobian formulas: madd-2007-bl, add-2007-bl, dbl-2009-l (a = 0).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ generic Montgomery core */
#define DEF_FIELD(P, N)                                                                         \
    typedef struct { u64 l[N]; } P##_t;                                                         \
    static inline int P##_geq_mod(const u64 *a) {                                               \
        for (int i = N - 1; i >= 0; i--) {                                                      \
            if (a[i] > P##_MOD[i]) return 1;                                                    \
            if (a[i] < P##_MOD[i]) return 0;                                                    \
        }                                                                                       \
        return 1;                                                                               \
    }                                                                                           \
    static inline void P##_sub_mod(u64 *a) {                                                    \
        u64 borrow = 0;                                                                         \
        for (int i = 0; i < N; i++) {                                                           \
            u128 d = (u128)a[i] - P##_MOD[i] - borrow;                                          \
            a[i] = (u64)d;                                                                      \
            borrow = (u64)(d >> 64) & 1;                                                        \
        }                                                                                       \
    }                                                                                           \
    static inline void P##_add(P##_t *z, const P##_t *x, const P##_t *y) {                      \
        u64 carry = 0;                                                                          \
        for (int i = 0; i < N; i++) {                                                           \
            u128 s = (u128)x->l[i] + y->l[i] + carry;                                           \
            z->l[i] = (u64)s;                                                                   \
            carry = (u64)(s >> 64);                                                             \
        }                                                                                       \
        if (carry || P##_geq_mod(z->l)) P##_sub_mod(z->l);                                      \
    }                                                                                           \
    static inline void P##_sub(P##_t *z, const P##_t *x, const P##_t *y) {                      \
        u64 borrow = 0;                                                                         \
        for (int i = 0; i < N; i++) {                                                           \
            u128 d = (u128)x->l[i] - y->l[i] - borrow;                                          \
            z->l[i] = (u64)d;                                                                   \
            borrow = (u64)(d >> 64) & 1;                                                        \
        }                                                                                       \
        if (borrow) {                                                                           \
            u64 carry = 0;                                                                      \
            for (int i = 0; i < N; i++) {                                                       \
                u128 s = (u128)z->l[i] + P##_MOD[i] + carry;                                    \
                z->l[i] = (u64)s;                                                               \
                carry = (u64)(s >> 64);                                                         \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
    static inline void P##_dbl(P##_t *z, const P##_t *x) { P##_add(z, x, x); }                  \
    static inline void P##_neg(P##_t *z, const P##_t *x) {                                      \
        P##_t zero;                                                                             \
        memset(&zero, 0, sizeof zero);                                                          \
        P##_sub(z, &zero, x);                                                                   \
    }                                                                                           \
    static inline int P##_is_zero(const P##_t *x) {                                             \
        u64 o = 0;                                                                              \
        for (int i = 0; i < N; i++) o |= x->l[i];                                               \
        return o == 0;                                                                          \
    }                                                                                           \
    static inline int P##_eq(const P##_t *x, const P##_t *y) {                                  \
        return memcmp(x, y, sizeof(P##_t)) == 0;                                                \
    }                                                                                           \
    /* CIOS Montgomery product x*y*R^-1 mod p */                                                \
    static inline void P##_mul(P##_t *z, const P##_t *x, const P##_t *y) {                      \
        u64 t[N + 2];                                                                           \
        memset(t, 0, sizeof t);                                                                 \
        for (int i = 0; i < N; i++) {                                                           \
            u64 carry = 0;                                                                      \
            for (int j = 0; j < N; j++) {                                                       \
                u128 acc = (u128)x->l[j] * y->l[i] + t[j] + carry;                              \
                t[j] = (u64)acc;                                                                \
                carry = (u64)(acc >> 64);                                                       \
            }                                                                                   \
            u128 acc = (u128)t[N] + carry;                                                      \
            t[N] = (u64)acc;                                                                    \
            t[N + 1] = (u64)(acc >> 64);                                                        \
            u64 mi = t[0] * P##_INV;                                                            \
            acc = (u128)mi * P##_MOD[0] + t[0];                                                 \
            carry = (u64)(acc >> 64);                                                           \
            for (int j = 1; j < N; j++) {                                                       \
                acc = (u128)mi * P##_MOD[j] + t[j] + carry;                                     \
                t[j - 1] = (u64)acc;                                                            \
                carry = (u64)(acc >> 64);                                                       \
            }                                                                                   \
            acc = (u128)t[N] + carry;                                                           \
            t[N - 1] = (u64)acc;                                                                \
            t[N] = t[N + 1] + (u64)(acc >> 64);                                                 \
        }                                                                                       \
        if (t[N] || P##_geq_mod(t)) P##_sub_mod(t);                                             \
        memcpy(z->l, t, sizeof(u64) * N);                                                       \
    }                                                                                           \
    static inline void P##_sqr(P##_t *z, const P##_t *x) { P##_mul(z, x, x); }                  \
    static inline void P##_from_canonical(P##_t *z, const P##_t *x) {                           \
        P##_t r2;                                                                               \
        memcpy(r2.l, P##_R2, sizeof r2);                                                        \
        P##_mul(z, x, &r2);                                                                     \
    }                                                                                           \
    static inline void P##_to_canonical(P##_t *z, const P##_t *x) {                             \
        P##_t one;                                                                              \
        memset(&one, 0, sizeof one);                                                            \
        one.l[0] = 1;                                                                           \
        P##_mul(z, x, &one);                                                                    \
    }                                                                                           \
    static inline void P##_set_one(P##_t *z) { memcpy(z->l, P##_ONE, sizeof(P##_t)); }          \
    /* x^e, e given as little-endian u64 limbs (square-and-multiply, MSB first) */             \
    static void P##_pow_limbs(P##_t *z, const P##_t *x, const u64 *e, int ne) {                 \
        P##_t acc;                                                                              \
        P##_set_one(&acc);                                                                      \
        int started = 0;                                                                        \
        for (int i = ne - 1; i >= 0; i--)                                                       \
            for (int b = 63; b >= 0; b--) {                                                     \
                if (started) P##_sqr(&acc, &acc);                                               \
                if ((e[i] >> b) & 1) {                                                          \
                    started = 1;                                                                \
                    P##_mul(&acc, &acc, x);                                                     \
                }                                                                               \
            }                                                                                   \
        *z = acc;                                                                               \
    }                                                                                           \
    static void P##_inv(P##_t *z, const P##_t *x) {                                             \
        u64 e[N];                                                                               \
        memcpy(e, P##_MOD, sizeof e);                                                           \
        e[0] -= 2; /* low limb of both moduli is >= 2 */                                        \
        P##_pow_limbs(z, x, e, N);                                                              \
    }

/* BLS12-381 scalar field Fr (ark_bls12_381::FrParameters) */
static const u64 fr_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                              0x73eda753299d7d48ULL};
static const u64 fr_ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                              0x1824b159acc5056fULL};
static const u64 fr_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                             0x0748d9d99f59ff11ULL};
static const u64 fr_INV = 0xfffffffeffffffffULL;
/* TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32), Montgomery form */
static const u64 fr_ROOT[4] = {0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL,
                               0x5bf3adda19e9b27bULL};
DEF_FIELD(fr, 4)

/* BLS12-381 base field Fq (ark_bls12_381::FqParameters) */
static const u64 fq_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                              0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 fq_ONE[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                              0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
static const u64 fq_R2[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                             0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};
static const u64 fq_INV = 0x89f3fffcfffcfffdULL;
DEF_FIELD(fq, 6)

/* ------------------------------------------------------------------ Fr helpers */
static void fr_pow_u64(fr_t *z, const fr_t *x, u64 e) { fr_pow_limbs(z, x, &e, 1); }

static void fr_from_u64(fr_t *z, u64 v) {
    fr_t c = {{v, 0, 0, 0}};
    fr_from_canonical(z, &c);
}

static int log2_ceil(u64 n) {
    int l = 0;
    while (((u64)1 << l) < n) l++;
    return l;
}

/* Radix2EvaluationDomain::new(min_size).group_gen (or its inverse) */
static void domain_gen(fr_t *g, int log_size, int inverse) {
    memcpy(g->l, fr_ROOT, sizeof(fr_t));
    for (int i = log_size; i < 32; i++) fr_sqr(g, g);
    if (inverse) fr_inv(g, g);
}

static void derange(fr_t *x, u64 n, int log_n) {
    for (u64 i = 1; i < n; i++) {
        u64 r = 0;
        for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < r) {
            fr_t t = x[i];
            x[i] = x[r];
            x[r] = t;
        }
    }
}

/* ark-poly io_helper: in-order input, bit-reversed output (decimation in frequency) */
static void io_helper(fr_t *x, u64 n, const fr_t *roots /* n/2 powers of root */) {
    u64 gap = n / 2;
    while (gap > 0) {
        u64 chunk = 2 * gap, step = n / chunk; /* = num_chunks */
        for (u64 base = 0; base < n; base += chunk)
            for (u64 i = 0; i < gap; i++) {
                fr_t *lo = &x[base + i], *hi = &x[base + gap + i], neg;
                fr_sub(&neg, lo, hi);
                fr_add(lo, lo, hi);
                fr_mul(hi, &neg, &roots[i * step]);
            }
        gap /= 2;
    }
}

/* ark-poly oi_helper: bit-reversed input, in-order output (decimation in time) */
static void oi_helper(fr_t *x, u64 n, const fr_t *roots) {
    u64 gap = 1;
    while (gap < n) {
        u64 chunk = 2 * gap, step = n / chunk;
        for (u64 base = 0; base < n; base += chunk)
            for (u64 i = 0; i < gap; i++) {
                fr_t *lo = &x[base + i], *hi = &x[base + gap + i], neg;
                fr_mul(hi, hi, &roots[i * step]);
                fr_sub(&neg, lo, hi);
                fr_add(lo, lo, hi);
                *hi = neg;
            }
        gap *= 2;
    }
}

static fr_t *roots_table(const fr_t *root, u64 n) {
    u64 h = n / 2 ? n / 2 : 1;
    fr_t *t = (fr_t *)malloc(sizeof(fr_t) * h);
    fr_set_one(&t[0]);
    for (u64 i = 1; i < h; i++) fr_mul(&t[i], &t[i - 1], root);
    return t;
}

/* Radix2EvaluationDomain::{fft,ifft}_in_place on a vector already of domain size n (pow2) */
void orc_fft_in_place(u64 *data, u64 n, int inverse) {
    fr_t *x = (fr_t *)data;
    if (n <= 1) return;
    int log_n = log2_ceil(n);
    fr_t root;
    domain_gen(&root, log_n, inverse);
    fr_t *roots = roots_table(&root, n);
    if (!inverse) {
        io_helper(x, n, roots);
        derange(x, n, log_n);
    } else {
        derange(x, n, log_n);
        oi_helper(x, n, roots);
        fr_t ninv, nn;
        fr_from_u64(&nn, n);
        fr_inv(&ninv, &nn);
        for (u64 i = 0; i < n; i++) fr_mul(&x[i], &x[i], &ninv);
    }
    free(roots);
}

static void distribute_powers(fr_t *x, u64 n, const fr_t *g) {
    fr_t pw;
    fr_set_one(&pw);
    for (u64 i = 0; i < n; i++) {
        fr_mul(&x[i], &x[i], &pw);
        fr_mul(&pw, &pw, g);
    }
}

/* coset_fft = distribute_powers(g) ; fft   |   coset_ifft = ifft ; distribute_powers(g^-1) */
void orc_coset_fft_in_place(u64 *data, u64 n, int inverse) {
    fr_t *x = (fr_t *)data;
    fr_t g, gi;
    fr_from_u64(&g, 7);
    fr_inv(&gi, &g);
    if (!inverse) {
        distribute_powers(x, n, &g);
        orc_fft_in_place(data, n, 0);
    } else {
        orc_fft_in_place(data, n, 1);
        distribute_powers(x, n, &gi);
    }
}

/* ------------------------------------------------------------------ the worker's 2-D helpers */
/* as_written = 1 mirrors worker.rs:79,93,113 (one Fr::pow per element); 0 uses running products
 * (identical values, the "fair" CPU variant of BASELINE.md §2). */
void orc_fft1_helper(u64 *data, u64 i, int is_coset, int is_inv, u64 domain_size, int as_written) {
    int log_n = log2_ceil(domain_size);
    u64 r = (u64)1 << (log_n >> 1), c = ((u64)1 << log_n) / r;
    fr_t *v = (fr_t *)data;
    if (is_coset && !is_inv) {
        fr_t g;
        fr_from_u64(&g, 7);
        if (as_written) {
            for (u64 j = 0; j < c; j++) {
                fr_t t;
                fr_pow_u64(&t, &g, i + j * r);
                fr_mul(&v[j], &v[j], &t);
            }
        } else {
            fr_t cur, step;
            fr_pow_u64(&cur, &g, i);
            fr_pow_u64(&step, &g, r);
            for (u64 j = 0; j < c; j++) {
                fr_mul(&v[j], &v[j], &cur);
                fr_mul(&cur, &cur, &step);
            }
        }
    }
    orc_fft_in_place(data, c, is_inv);
    fr_t w;
    domain_gen(&w, log_n, is_inv);
    if (as_written) {
        for (u64 j = 0; j < c; j++) {
            fr_t t;
            fr_pow_u64(&t, &w, i * j);
            fr_mul(&v[j], &v[j], &t);
        }
    } else {
        fr_t cur, step;
        fr_set_one(&cur);
        fr_pow_u64(&step, &w, i);
        for (u64 j = 0; j < c; j++) {
            fr_mul(&v[j], &v[j], &cur);
            fr_mul(&cur, &cur, &step);
        }
    }
}

void orc_fft2_helper(u64 *data, u64 i, int is_coset, int is_inv, u64 domain_size, int as_written) {
    int log_n = log2_ceil(domain_size);
    u64 r = (u64)1 << (log_n >> 1), c = ((u64)1 << log_n) / r;
    fr_t *v = (fr_t *)data;
    orc_fft_in_place(data, r, is_inv);
    if (is_coset && is_inv) {
        fr_t g, gi;
        fr_from_u64(&g, 7);
        fr_inv(&gi, &g);
        if (as_written) {
            for (u64 j = 0; j < r; j++) {
                fr_t t;
                fr_pow_u64(&t, &gi, i + j * c);
                fr_mul(&v[j], &v[j], &t);
            }
        } else {
            fr_t cur, step;
            fr_pow_u64(&cur, &gi, i);
            fr_pow_u64(&step, &gi, c);
            for (u64 j = 0; j < r; j++) {
                fr_mul(&v[j], &v[j], &cur);
                fr_mul(&cur, &cur, &step);
            }
        }
    }
}

/* Whole 4-RPC pipeline of dispatcher2.rs:731-787 for n_workers workers in one process.
 * in: domain_size Fr (already zero-padded); out: domain_size Fr.  OpenMP over rows / columns
 * stands in for the workers running concurrently. */
void orc_distributed_fft(const u64 *in, u64 *out, u64 domain_size, int is_inv, int is_coset,
                         u64 n_workers, int as_written) {
    int log_n = log2_ceil(domain_size);
    u64 n = (u64)1 << log_n;
    u64 r = (u64)1 << (log_n >> 1), c = n / r;
    const fr_t *x = (const fr_t *)in;
    fr_t *rows = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_t *cols = (fr_t *)malloc(sizeof(fr_t) * n);
    /* dispatcher2.rs:754 transpose: row b = { x[b + a*r] } */
#pragma omp parallel for schedule(static)
    for (u64 b = 0; b < r; b++) {
        for (u64 a = 0; a < c; a++) rows[b * c + a] = x[b + a * r];
        orc_fft1_helper((u64 *)&rows[b * c], b, is_coset, is_inv, domain_size, as_written);
    }
    /* worker.rs:327-330 + 432-435 for every (p, q) pair */
    u64 W = n_workers;
    for (u64 p = 0; p < W; p++) {
        u64 rs = p * r / W, re = (p + 1) * r / W;
        for (u64 q = 0; q < W; q++) {
            u64 cs = q * c / W, ce = (q + 1) * c / W, ncols = ce - cs, t = 0;
            for (u64 i = rs; i < re; i++)
                for (u64 k = cs; k < ce; k++, t++)
                    cols[(cs + t % ncols) * r + rs + t / ncols] = rows[i * c + k];
        }
    }
#pragma omp parallel for schedule(static)
    for (u64 k = 0; k < c; k++)
        orc_fft2_helper((u64 *)&cols[k * r], k, is_coset, is_inv, domain_size, as_written);
    /* dispatcher2.rs:780-786 */
    fr_t *o = (fr_t *)out;
    for (u64 j = 0; j < r; j++)
        for (u64 i = 0; i < c; i++) o[j * c + i] = cols[i * r + j];
    free(rows);
    free(cols);
}

/* p(z) for p = sum x[j] z^j by Horner, z = omega_n^(+-k) (and optionally the coset shift g): one
 * output element X[k] of an n-point (coset) (i)NTT in O(n) - the size-independent spot check used
 * by the full-size parity tests.  inverse also multiplies by 1/n (and g^-k for coset). */
void orc_ntt_output_at(const u64 *data, u64 n, u64 k, int inverse, int coset, u64 *out) {
    const fr_t *x = (const fr_t *)data;
    int log_n = log2_ceil(n);
    fr_t w, z, g, acc;
    domain_gen(&w, log_n, inverse);
    fr_pow_u64(&z, &w, k);
    fr_from_u64(&g, 7);
    if (coset && !inverse) fr_mul(&z, &z, &g);
    memset(&acc, 0, sizeof acc);
    for (u64 j = n; j-- > 0;) {
        fr_mul(&acc, &acc, &z);
        fr_add(&acc, &acc, &x[j]);
    }
    if (inverse) {
        fr_t nn, ninv;
        fr_from_u64(&nn, n);
        fr_inv(&ninv, &nn);
        fr_mul(&acc, &acc, &ninv);
        if (coset) {
            fr_t gi, gk;
            fr_inv(&gi, &g);
            fr_pow_u64(&gk, &gi, k);
            fr_mul(&acc, &acc, &gk);
        }
    }
    memcpy(out, &acc, 32);
}

/* The same for several positions at once (one thread each) and for a coefficient vector shorter than the
 * domain: X[k] of the domain_size-point transform of x zero-padded from n_coeffs to domain_size - what
 * Prover::fft is fed on the quotient domain (n coefficients on 8n points, dispatcher2.rs:386-388). */
void orc_ntt_outputs_at(const u64 *data, u64 n_coeffs, u64 domain_size, const u64 *ks, u64 n_k, int inverse, int coset, u64 *out) {
    const fr_t *x = (const fr_t *)data;
    int log_n = log2_ceil(domain_size);
    fr_t w, g, ninv, gi;
    domain_gen(&w, log_n, inverse);
    fr_from_u64(&g, 7);
    fr_inv(&gi, &g);
    {
        fr_t nn;
        fr_from_u64(&nn, domain_size);
        fr_inv(&ninv, &nn);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (u64 t = 0; t < n_k; t++) {
        fr_t z, acc;
        fr_pow_u64(&z, &w, ks[t]);
        if (coset && !inverse) fr_mul(&z, &z, &g);
        memset(&acc, 0, sizeof acc);
        for (u64 j = n_coeffs; j-- > 0;) {
            fr_mul(&acc, &acc, &z);
            fr_add(&acc, &acc, &x[j]);
        }
        if (inverse) {
            fr_mul(&acc, &acc, &ninv);
            if (coset) {
                fr_t gk;
                fr_pow_u64(&gk, &gi, ks[t]);
                fr_mul(&acc, &acc, &gk);
            }
        }
        memcpy(out + 4 * t, &acc, 32);
    }
}

/* sum_i s_i * k_i mod r for canonical scalars s (n x 4 words) and 64-bit multipliers k: the discrete
 * logarithm of sum_i s_i * (k_i * G), used to check MSMs over synthetic bases k_i * G whose size makes a
 * second full Pippenger on the CPU too slow.  Result canonical. */
void orc_fr_dot_u64(const u64 *scalars, const u64 *ks, u64 n, u64 *out) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    fr_t *part = (fr_t *)calloc((size_t)nt, sizeof(fr_t));
#pragma omp parallel
    {
        int me = 0;
#ifdef _OPENMP
        me = omp_get_thread_num();
#endif
        fr_t acc;
        memset(&acc, 0, sizeof acc);
#pragma omp for schedule(static)
        for (u64 i = 0; i < n; i++) {
            fr_t s, k, t;
            fr_from_canonical(&s, (const fr_t *)(scalars + 4 * i));
            fr_from_u64(&k, ks[i]);
            fr_mul(&t, &s, &k);
            fr_add(&acc, &acc, &t);
        }
        part[me] = acc;
    }
    fr_t tot;
    memset(&tot, 0, sizeof tot);
    for (int t = 0; t < nt; t++) fr_add(&tot, &tot, &part[t]);
    free(part);
    fr_t c;
    fr_to_canonical(&c, &tot);
    memcpy(out, &c, 32);
}

/* Round-2 permutation grand product exactly as the dispatcher computes it (src/dispatcher2.rs:329-345):
 * product_vec[0] = 1; product_vec[j+1] = product_vec[j] * a / b with one field division per row.
 * wires / id / sigma: [n_types][n] Montgomery Fr (wire value, extended_id_permutation at (i,j), and at
 * the permuted position).  Returns 0, or -1 when a denominator is zero (the reference panics). */
int orc_perm_product(const u64 *wires, const u64 *id, const u64 *sigma, u64 n_types, u64 n, const u64 *beta,
                     const u64 *gamma, u64 *out) {
    const fr_t *w = (const fr_t *)wires, *idp = (const fr_t *)id, *sg = (const fr_t *)sigma;
    const fr_t *be = (const fr_t *)beta, *ga = (const fr_t *)gamma;
    fr_t *z = (fr_t *)out;
    if (n == 0) return 0;
    fr_set_one(&z[0]);
    for (u64 j = 0; j + 1 < n; j++) {
        fr_t a, b;
        fr_set_one(&a);
        fr_set_one(&b);
        for (u64 i = 0; i < n_types; i++) {
            fr_t tmp, t;
            fr_add(&tmp, &w[i * n + j], ga);
            fr_mul(&t, be, &idp[i * n + j]);
            fr_add(&t, &t, &tmp);
            fr_mul(&a, &a, &t);
            fr_mul(&t, be, &sg[i * n + j]);
            fr_add(&t, &t, &tmp);
            fr_mul(&b, &b, &t);
        }
        if (fr_is_zero(&b)) return -1;
        fr_t bi;
        fr_inv(&bi, &b);
        fr_mul(&a, &a, &bi);
        fr_mul(&z[j + 1], &z[j], &a);
    }
    return 0;
}

/* ------------------------------------------------------------------ rounds 3-5 of Prover::prove
 * ("next" row 1 of SURVEY.md 8f: the arithmetic the dispatcher does between the transforms)        */

/* Round 3: coset evaluations of the quotient polynomial over the quotient domain, before the final
 * coset iFFT (src/dispatcher2.rs:363-504).  Same expression, same order of terms:
 *   z_h_inv[i % (m/n)] * (gate(i) + alpha * (acc1 - acc2)) + alpha^2/n * (z[i] - 1) / (x_i - 1)
 * with x_i = g * omega_m^i (lines 366-369), z_h_inv[i] = 1/(x_i^n - 1) (372-379), GATE_WIDTH = 4,
 * selectors in the order q_lc[4], q_mul[2], q_hash[4], q_o, q_c, q_ecc (437-450), five wire types.
 * selectors [13][m], sigmas [5][m], wires [5][m], perm [m], pub_input [m], k [5]; out [m].        */
void orc_quotient_evals(const u64 *selectors, const u64 *sigmas, const u64 *wires, const u64 *perm,
                        const u64 *pub_input, const u64 *k, const u64 *alpha_, const u64 *beta_,
                        const u64 *gamma_, u64 n, u64 m, u64 *out) {
    const fr_t *sel = (const fr_t *)selectors, *sg = (const fr_t *)sigmas, *w = (const fr_t *)wires;
    const fr_t *z = (const fr_t *)perm, *pi = (const fr_t *)pub_input, *vk_k = (const fr_t *)k;
    const fr_t alpha = *(const fr_t *)alpha_, beta = *(const fr_t *)beta_, gamma = *(const fr_t *)gamma_;
    fr_t *q = (fr_t *)out;
    const u64 ratio = m / n;
    fr_t one, nf, ninv, alpha_sq_div_n, gen, omega;
    fr_set_one(&one);
    fr_from_u64(&nf, n);
    fr_inv(&ninv, &nf);
    fr_sqr(&alpha_sq_div_n, &alpha);
    fr_mul(&alpha_sq_div_n, &alpha_sq_div_n, &ninv); /* alpha.square() / Fr::from(n), line 363 */
    fr_from_u64(&gen, 7);
    domain_gen(&omega, log2_ceil(m), 0);
    fr_t *x = (fr_t *)malloc(m * sizeof(fr_t));
    x[0] = gen;
    for (u64 i = 1; i < m; i++) fr_mul(&x[i], &x[i - 1], &omega);
    fr_t *z_h_inv = (fr_t *)malloc(ratio * sizeof(fr_t));
    for (u64 i = 0; i < ratio; i++) {
        fr_t t;
        fr_pow_u64(&t, &x[i], n);
        fr_sub(&t, &t, &one);
        fr_inv(&z_h_inv[i], &t);
    }
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < m; i++) {
        const fr_t a = w[0 * m + i], b = w[1 * m + i], c = w[2 * m + i], d = w[3 * m + i], e = w[4 * m + i];
        fr_t ab, cd, t, u, gate;
        fr_mul(&ab, &a, &b);
        fr_mul(&cd, &c, &d);
        fr_add(&gate, &sel[11 * m + i], &pi[i]); /* q_c + pub_input */
        const fr_t lc_in[4] = {a, b, c, d};
        for (int j = 0; j < 4; j++) { /* q_lc[j] * wire_j */
            fr_mul(&t, &sel[j * m + i], &lc_in[j]);
            fr_add(&gate, &gate, &t);
        }
        fr_mul(&t, &sel[4 * m + i], &ab);
        fr_add(&gate, &gate, &t);
        fr_mul(&t, &sel[5 * m + i], &cd);
        fr_add(&gate, &gate, &t);
        fr_mul(&t, &sel[12 * m + i], &ab); /* q_ecc * ab * cd * e */
        fr_mul(&t, &t, &cd);
        fr_mul(&t, &t, &e);
        fr_add(&gate, &gate, &t);
        for (int j = 0; j < 4; j++) { /* q_hash[j] * wire_j^5 */
            fr_sqr(&u, &lc_in[j]);
            fr_sqr(&u, &u);
            fr_mul(&u, &u, &lc_in[j]);
            fr_mul(&t, &sel[(6 + j) * m + i], &u);
            fr_add(&gate, &gate, &t);
        }
        fr_mul(&t, &sel[10 * m + i], &e); /* - q_o * e */
        fr_sub(&gate, &gate, &t);
        fr_t acc1 = z[i], acc2 = z[(i + ratio) % m];
        for (int j = 0; j < 5; j++) {
            fr_t wg;
            fr_add(&wg, &w[j * m + i], &gamma);
            fr_mul(&t, &vk_k[j], &x[i]);
            fr_mul(&t, &t, &beta);
            fr_add(&t, &t, &wg);
            fr_mul(&acc1, &acc1, &t);
            fr_mul(&t, &sg[j * m + i], &beta);
            fr_add(&t, &t, &wg);
            fr_mul(&acc2, &acc2, &t);
        }
        fr_sub(&t, &acc1, &acc2);
        fr_mul(&t, &t, &alpha);
        fr_add(&gate, &gate, &t);
        fr_mul(&gate, &gate, &z_h_inv[i % ratio]);
        fr_sub(&t, &z[i], &one); /* alpha^2/n * (z - 1) / (x - 1) */
        fr_mul(&t, &t, &alpha_sq_div_n);
        fr_sub(&u, &x[i], &one);
        fr_inv(&u, &u);
        fr_mul(&t, &t, &u);
        fr_add(&q[i], &gate, &t);
    }
    free(x);
    free(z_h_inv);
}

/* Round 4: DensePolynomial::evaluate (Horner), src/dispatcher2.rs:535-548 */
void orc_poly_eval(const u64 *coeffs, u64 n, const u64 *point, u64 *out) {
    const fr_t *c = (const fr_t *)coeffs, *zt = (const fr_t *)point;
    fr_t acc;
    memset(&acc, 0, sizeof acc);
    for (u64 j = n; j-- > 0;) {
        fr_mul(&acc, &acc, zt);
        fr_add(&acc, &acc, &c[j]);
    }
    memcpy(out, &acc, sizeof acc);
}

/* Round 5: sum_k coeff_k * poly_k (the folds of src/dispatcher2.rs:566-633 and 646-649); polys of
 * different lengths are zero-extended to out_len.  polys = k pointers to raw Fr arrays.          */
void orc_poly_lincomb(const u64 *const *polys, const u64 *lens, const u64 *coeffs, u64 k, u64 *out, u64 out_len) {
    const fr_t *cf = (const fr_t *)coeffs;
    fr_t *o = (fr_t *)out;
#pragma omp parallel for schedule(static)
    for (u64 j = 0; j < out_len; j++) {
        fr_t acc, t;
        memset(&acc, 0, sizeof acc);
        for (u64 i = 0; i < k; i++)
            if (j < lens[i]) {
                fr_mul(&t, &((const fr_t *)polys[i])[j], &cf[i]);
                fr_add(&acc, &acc, &t);
            }
        o[j] = acc;
    }
}

/* Round 5: witness polynomial = quotient of p(X) by (X - point), the long division of
 * src/dispatcher2.rs:651-666 (and 672-688 with point = omega * zeta): top coefficient down,
 * quotient[d-1] = remainder[d]; remainder[d-1] += remainder[d] * point.  out has n-1 coefficients. */
void orc_poly_div_linear(const u64 *coeffs, u64 n, const u64 *point, u64 *out) {
    const fr_t *c = (const fr_t *)coeffs, *zt = (const fr_t *)point;
    fr_t *q = (fr_t *)out;
    if (n < 2) return;
    fr_t carry = c[n - 1];
    for (u64 d = n - 1; d >= 1; d--) {
        q[d - 1] = carry;
        fr_t t;
        fr_mul(&t, &carry, zt);
        fr_add(&carry, &c[d - 1], &t);
    }
}

/* ------------------------------------------------------------------ raw Fr utilities for tests */
void orc_fr_mul(const u64 *a, const u64 *b, u64 *out) { fr_mul((fr_t *)out, (const fr_t *)a, (const fr_t *)b); }
void orc_fr_add(const u64 *a, const u64 *b, u64 *out) { fr_add((fr_t *)out, (const fr_t *)a, (const fr_t *)b); }
void orc_fr_sub(const u64 *a, const u64 *b, u64 *out) { fr_sub((fr_t *)out, (const fr_t *)a, (const fr_t *)b); }
void orc_fq_mul(const u64 *a, const u64 *b, u64 *out) { fq_mul((fq_t *)out, (const fq_t *)a, (const fq_t *)b); }
void orc_fq_add(const u64 *a, const u64 *b, u64 *out) { fq_add((fq_t *)out, (const fq_t *)a, (const fq_t *)b); }
void orc_fq_sub(const u64 *a, const u64 *b, u64 *out) { fq_sub((fq_t *)out, (const fq_t *)a, (const fq_t *)b); }
/* elementwise vector ops for building test instances: op 0 a+b, 1 a-b, 2 a*b, 3 1/a (0 -> 0) */
void orc_fr_vec_op(const u64 *a, const u64 *b, u64 *out, u64 n, int op) {
    const fr_t *x = (const fr_t *)a, *y = (const fr_t *)b;
    fr_t *o = (fr_t *)out;
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) {
        if (op == 0) fr_add(&o[i], &x[i], &y[i]);
        else if (op == 1) fr_sub(&o[i], &x[i], &y[i]);
        else if (op == 2) fr_mul(&o[i], &x[i], &y[i]);
        else if (fr_is_zero(&x[i])) memset(&o[i], 0, sizeof(fr_t));
        else fr_inv(&o[i], &x[i]);
    }
}
/* Fr::into_repr over a vector (worker.rs:118) */
void orc_fr_into_repr(const u64 *in, u64 *out, u64 n) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) fr_to_canonical((fr_t *)(out + 4 * i), (const fr_t *)(in + 4 * i));
}
void orc_fr_from_repr(const u64 *in, u64 *out, u64 n) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) fr_from_canonical((fr_t *)(out + 4 * i), (const fr_t *)(in + 4 * i));
}

/* ------------------------------------------------------------------ G1 (ark-ec 0.3.0 Jacobian) */
typedef struct { fq_t x, y; uint8_t infinity; uint8_t pad[7]; } g1a_t; /* 104 B raw GroupAffine */
typedef struct { fq_t x, y, z; } g1j_t;                                /* 144 B raw GroupProjective */

static void g1j_set_zero(g1j_t *p) {
    memset(p, 0, sizeof *p);
    fq_set_one(&p->y); /* ark: (0, 1, 0) */
}
static int g1j_is_zero(const g1j_t *p) { return fq_is_zero(&p->z); }

/* dbl-2009-l */
static void g1j_double(g1j_t *p) {
    if (g1j_is_zero(p)) return;
    fq_t a, b, c, d, e, f, t;
    fq_sqr(&a, &p->x);
    fq_sqr(&b, &p->y);
    fq_sqr(&c, &b);
    fq_add(&t, &p->x, &b);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &a);
    fq_sub(&t, &t, &c);
    fq_dbl(&d, &t);
    fq_dbl(&e, &a);
    fq_add(&e, &e, &a);
    fq_sqr(&f, &e);
    fq_mul(&p->z, &p->z, &p->y);
    fq_dbl(&p->z, &p->z);
    fq_sub(&p->x, &f, &d);
    fq_sub(&p->x, &p->x, &d);
    fq_sub(&t, &d, &p->x);
    fq_mul(&t, &t, &e);
    fq_dbl(&c, &c);
    fq_dbl(&c, &c);
    fq_dbl(&c, &c);
    fq_sub(&p->y, &t, &c);
}

/* madd-2007-bl */
static void g1j_add_mixed(g1j_t *p, const g1a_t *q) {
    if (q->infinity) return;
    if (g1j_is_zero(p)) {
        p->x = q->x;
        p->y = q->y;
        fq_set_one(&p->z);
        return;
    }
    fq_t z1z1, u2, s2;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&p->x, &u2) && fq_eq(&p->y, &s2)) {
        g1j_double(p);
        return;
    }
    fq_t h, hh, i, j, r, v;
    fq_sub(&h, &u2, &p->x);
    fq_sqr(&hh, &h);
    fq_dbl(&i, &hh);
    fq_dbl(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &p->y);
    fq_dbl(&r, &r);
    fq_mul(&v, &p->x, &i);
    fq_sqr(&p->x, &r);
    fq_sub(&p->x, &p->x, &j);
    fq_sub(&p->x, &p->x, &v);
    fq_sub(&p->x, &p->x, &v);
    fq_mul(&j, &j, &p->y);
    fq_dbl(&j, &j);
    fq_sub(&p->y, &v, &p->x);
    fq_mul(&p->y, &p->y, &r);
    fq_sub(&p->y, &p->y, &j);
    fq_add(&p->z, &p->z, &h);
    fq_sqr(&p->z, &p->z);
    fq_sub(&p->z, &p->z, &z1z1);
    fq_sub(&p->z, &p->z, &hh);
}

/* add-2007-bl */
static void g1j_add(g1j_t *p, const g1j_t *q) {
    if (g1j_is_zero(p)) {
        *p = *q;
        return;
    }
    if (g1j_is_zero(q)) return;
    fq_t z1z1, z2z2, u1, u2, s1, s2;
    fq_sqr(&z1z1, &p->z);
    fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z);
    fq_mul(&s1, &s1, &z2z2);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&u1, &u2) && fq_eq(&s1, &s2)) {
        g1j_double(p);
        return;
    }
    fq_t h, i, j, r, v, t;
    fq_sub(&h, &u2, &u1);
    fq_dbl(&i, &h);
    fq_sqr(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &s1);
    fq_dbl(&r, &r);
    fq_mul(&v, &u1, &i);
    fq_sqr(&p->x, &r);
    fq_sub(&p->x, &p->x, &j);
    fq_sub(&p->x, &p->x, &v);
    fq_sub(&p->x, &p->x, &v);
    fq_sub(&t, &v, &p->x);
    fq_mul(&t, &t, &r);
    fq_mul(&s1, &s1, &j);
    fq_dbl(&s1, &s1);
    fq_sub(&p->y, &t, &s1);
    fq_add(&t, &p->z, &q->z);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &z1z1);
    fq_sub(&t, &t, &z2z2);
    fq_mul(&p->z, &t, &h);
}

static void g1j_to_affine(g1a_t *a, const g1j_t *p) {
    memset(a, 0, sizeof *a);
    if (g1j_is_zero(p)) {
        fq_set_one(&a->y);
        a->infinity = 1;
        return;
    }
    fq_t zi, zi2, zi3;
    fq_inv(&zi, &p->z);
    fq_sqr(&zi2, &zi);
    fq_mul(&zi3, &zi2, &zi);
    fq_mul(&a->x, &p->x, &zi2);
    fq_mul(&a->y, &p->y, &zi3);
}

/* GroupProjective -> GroupAffine (the `.into()` of dispatcher2.rs:892) */
void orc_g1_normalize(const uint8_t *jac144, uint8_t *aff104) {
    g1j_t p;
    g1a_t a;
    memcpy(&p, jac144, 144);
    g1j_to_affine(&a, &p);
    memcpy(aff104, &a, 104);
}

/* sum of two raw Jacobian points (the dispatcher's `.reduce(|a, b| a + b)`, dispatcher2.rs:887-890) */
void orc_g1_add(const uint8_t *a144, const uint8_t *b144, uint8_t *out144) {
    g1j_t p, q;
    memcpy(&p, a144, 144);
    memcpy(&q, b144, 144);
    g1j_add(&p, &q);
    memcpy(out144, &p, 144);
}

static const u64 G1_GEN_X[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
static const u64 G1_GEN_Y[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};

static void g1_generator(g1a_t *g) {
    memset(g, 0, sizeof *g);
    fq_t x, y;
    memcpy(x.l, G1_GEN_X, 48);
    memcpy(y.l, G1_GEN_Y, 48);
    fq_from_canonical(&g->x, &x);
    fq_from_canonical(&g->y, &y);
}

/* k*P by double-and-add, k a canonical 256-bit integer */
static void g1_scalar_mul(g1j_t *out, const g1a_t *p, const u64 k[4]) {
    g1j_t acc;
    g1j_set_zero(&acc);
    for (int i = 255; i >= 0; i--) {
        g1j_double(&acc);
        if ((k[i / 64] >> (i % 64)) & 1) g1j_add_mixed(&acc, p);
    }
    *out = acc;
}

void orc_g1_mul(const uint8_t *aff104, const u64 *k, uint8_t *out104) {
    g1a_t p, a;
    g1j_t q;
    memcpy(&p, aff104, 104);
    g1_scalar_mul(&q, &p, k);
    g1j_to_affine(&a, &q);
    memcpy(out104, &a, 104);
}

void orc_g1_generator(uint8_t *out104) {
    g1a_t g;
    g1_generator(&g);
    memcpy(out104, &g, 104);
}

/* KZG SRS with a KNOWN trapdoor, for tests only: out[i] = tau^i * G (tau canonical, 4 limbs).
 * With tau known, commit(p) = p(tau) * G, so opening proofs can be checked in the group without a
 * pairing: (tau - z) * commit(q) + p(z) * G == commit(p)  for q = (p - p(z)) / (X - z).            */
void orc_gen_srs(const u64 *tau, u64 n, uint8_t *out104) {
    fr_t t, cur;
    fr_from_canonical(&t, (const fr_t *)tau);
    fr_t *pw = (fr_t *)malloc((n ? n : 1) * sizeof(fr_t));
    fr_set_one(&cur);
    for (u64 i = 0; i < n; i++) {
        fr_to_canonical(&pw[i], &cur);
        fr_mul(&cur, &cur, &t);
    }
    g1a_t g;
    g1_generator(&g);
#pragma omp parallel for schedule(dynamic, 16)
    for (u64 i = 0; i < n; i++) {
        g1j_t q;
        g1a_t a;
        g1_scalar_mul(&q, &g, pw[i].l);
        g1j_to_affine(&a, &q);
        memcpy(out104 + 104 * i, &a, 104);
    }
    free(pw);
}

/* ------------------------------------------------------------------ canonical point encoding
 * ark-serialize 0.3.0, `CanonicalSerialize for GroupAffine<P>` (short Weierstrass) [3P-recall: the
 * crate is not vendored; "next" row 4 of SURVEY.md 8f, the format jellyfish SRS files use]:
 *   compressed, 48 B: canonical x, little-endian; the two top bits of the last byte are SWFlags:
 *   bit 7 = PositiveY (y > -y as canonical integers), bit 6 = Infinity (x written as 0).
 * deserialize: x >= p, both flags set, x^3 + 4 not a square -> error; y = sqrt via (p+1)/4 (p = 3 mod 4),
 * the root with (y > -y) == flag; the checked variant also requires r * P = 0.                      */
static int fq_canon_gt(const fq_t *a, const fq_t *b) { /* canonical integers */
    for (int i = 5; i >= 0; i--)
        if (a->l[i] != b->l[i]) return a->l[i] > b->l[i];
    return 0;
}

void orc_g1_compress(const uint8_t *aff104, uint8_t *out48) {
    g1a_t p;
    memcpy(&p, aff104, 104);
    if (p.infinity) {
        memset(out48, 0, 48);
        out48[47] |= 1 << 6;
        return;
    }
    fq_t xc, yc, ny, nyc;
    fq_to_canonical(&xc, &p.x);
    fq_to_canonical(&yc, &p.y);
    fq_neg(&ny, &p.y);
    fq_to_canonical(&nyc, &ny);
    memcpy(out48, xc.l, 48);
    if (fq_canon_gt(&yc, &nyc)) out48[47] |= 1 << 7;
}

/* 0 ok; -1 x not canonical; -2 bad flags; -3 not on the curve; -4 not in the r-torsion subgroup */
int orc_g1_decompress(const uint8_t *in48, uint8_t *out104, int check_subgroup) {
    uint8_t buf[48];
    memcpy(buf, in48, 48);
    const int positive = (buf[47] >> 7) & 1, infinity = (buf[47] >> 6) & 1;
    buf[47] &= 0x3f;
    g1a_t p;
    memset(&p, 0, sizeof p);
    if (positive && infinity) return -2;
    if (infinity) { /* GroupAffine::zero() = (0, 1, true) */
        fq_set_one(&p.y);
        p.infinity = 1;
        memcpy(out104, &p, 104);
        return 0;
    }
    fq_t xc, mod;
    memcpy(xc.l, buf, 48);
    memcpy(mod.l, fq_MOD, 48);
    if (!fq_canon_gt(&mod, &xc)) return -1;
    fq_from_canonical(&p.x, &xc);
    fq_t rhs, four, y, y2;
    fq_t c4 = {{4, 0, 0, 0, 0, 0}};
    fq_from_canonical(&four, &c4);
    fq_sqr(&rhs, &p.x);
    fq_mul(&rhs, &rhs, &p.x);
    fq_add(&rhs, &rhs, &four);
    u64 e[6]; /* (p + 1) / 4 */
    memcpy(e, fq_MOD, 48);
    e[0] += 1; /* low limb ...aaab + 1 does not carry */
    for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
    fq_pow_limbs(&y, &rhs, e, 6);
    fq_sqr(&y2, &y);
    if (memcmp(&y2, &rhs, sizeof y2) != 0) return -3;
    fq_t ny, yc, nyc;
    fq_neg(&ny, &y);
    fq_to_canonical(&yc, &y);
    fq_to_canonical(&nyc, &ny);
    const int y_is_larger = fq_canon_gt(&yc, &nyc);
    p.y = (y_is_larger == positive) ? y : ny;
    if (check_subgroup) {
        g1j_t q;
        g1_scalar_mul(&q, &p, fr_MOD);
        if (!g1j_is_zero(&q)) return -4;
    }
    memcpy(out104, &p, 104);
    return 0;
}

/* a point of the curve outside the r-torsion subgroup (for the negative test): the first x = 1, 2, ...
 * with x^3 + 4 a square, NOT multiplied by the cofactor */
int orc_g1_point_outside_subgroup(uint8_t *out48) {
    for (u64 x = 1; x < 1000; x++) {
        uint8_t buf[48] = {0};
        uint8_t tmp[104];
        memcpy(buf, &x, 8);
        if (orc_g1_decompress(buf, tmp, 0) == 0 && orc_g1_decompress(buf, tmp, 1) == -4) {
            memcpy(out48, buf, 48);
            return 0;
        }
    }
    return -1;
}

/* ------------------------------------------------------------------ seeded inputs */
typedef struct { u64 s; } splitmix_t;
static u64 splitmix_next(splitmix_t *g) {
    u64 z = (g->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void splitmix_fr(splitmix_t *g, fr_t *canon) {
    for (;;) {
        for (int i = 0; i < 4; i++) canon->l[i] = splitmix_next(g);
        canon->l[3] &= 0x7fffffffffffffffULL;
        if (!fr_geq_mod(canon->l)) return;
    }
}

/* n uniform Fr; montgomery=1 -> raw ark Fr (Montgomery), 0 -> canonical BigInteger256.
 * Element i depends only on (seed, i) through a per-1024-block stream so generation can be
 * parallel and any sub-range reproducible. */
void orc_gen_fr(u64 seed, u64 n, u64 *out, int montgomery) {
#pragma omp parallel for schedule(static)
    for (u64 blk = 0; blk < (n + 1023) / 1024; blk++) {
        splitmix_t g = {seed ^ (0xD15791B07E5EEDULL + blk * 0x632BE59BD9B4E019ULL)};
        u64 end = (blk + 1) * 1024 < n ? (blk + 1) * 1024 : n;
        for (u64 i = blk * 1024; i < end; i++) {
            fr_t c;
            splitmix_fr(&g, &c);
            if (montgomery)
                fr_from_canonical((fr_t *)(out + 4 * i), &c);
            else
                memcpy(out + 4 * i, &c, 32);
        }
    }
}

/* MSM bases in the style of dispatcher.rs:190-196 / dispatcher2.rs:1097-1104:
 * `distinct` points k_i*G (k_i seeded), index 3 = infinity when with_infinity, tiled by doubling
 * up to n, raw 104-byte GroupAffine each. */
void orc_gen_bases(u64 seed, u64 n, u64 distinct, int with_infinity, uint8_t *out104) {
    if (distinct > n) distinct = n;
    g1a_t gen;
    g1_generator(&gen);
    g1a_t *o = (g1a_t *)out104;
#pragma omp parallel for schedule(dynamic, 8)
    for (u64 i = 0; i < distinct; i++) {
        splitmix_t g = {seed ^ (0xBA5E5ULL + i * 0x9E3779B97F4A7C15ULL)};
        fr_t k;
        splitmix_fr(&g, &k);
        g1j_t p;
        g1_scalar_mul(&p, &gen, k.l);
        g1j_to_affine(&o[i], &p);
    }
    if (with_infinity && distinct > 3) {
        memset(&o[3], 0, sizeof(g1a_t));
        fq_set_one(&o[3].y);
        o[3].infinity = 1;
    }
    u64 have = distinct;
    while (have < n) {
        u64 cp = have < n - have ? have : n - have;
        memcpy(&o[have], &o[0], cp * sizeof(g1a_t));
        have += cp;
    }
}

/* ------------------------------------------------------------------ MSM (ark-ec 0.3.0) */
u64 orc_msm_window_c(u64 n) {
    if (n < 32) return 3;
    return (u64)log2_ceil(n) * 69 / 100 + 2;
}

/* G1-adds of one MSM(n) in the shared numerator of BASELINE.md §3 */
double orc_msm_work_adds(u64 n_nonzero, u64 n) {
    u64 c = orc_msm_window_c(n), w = (255 + c - 1) / c;
    return (double)n_nonzero * (double)w + 2.0 * (double)(((u64)1 << c) - 1) * (double)w;
}

/* VariableBaseMSM::multi_scalar_mul(bases, scalars): scalars canonical BigInteger256 */
void orc_msm(const uint8_t *bases104, const u64 *scalars, u64 n, uint8_t *out144) {
    const g1a_t *bases = (const g1a_t *)bases104;
    u64 c = orc_msm_window_c(n);
    int num_bits = 255;
    int n_windows = (num_bits + (int)c - 1) / (int)c;
    g1j_t *window_sums = (g1j_t *)malloc(sizeof(g1j_t) * n_windows);
    static const u64 one[4] = {1, 0, 0, 0};
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < n_windows; w++) {
        u64 w_start = (u64)w * c;
        g1j_t res;
        g1j_set_zero(&res);
        u64 nb = ((u64)1 << c) - 1;
        g1j_t *buckets = (g1j_t *)malloc(sizeof(g1j_t) * nb);
        for (u64 b = 0; b < nb; b++) g1j_set_zero(&buckets[b]);
        for (u64 i = 0; i < n; i++) {
            const u64 *s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
            if (memcmp(s, one, 32) == 0) {
                if (w_start == 0) g1j_add_mixed(&res, &bases[i]);
                continue;
            }
            /* scalar.divn(w_start); scalar.as_ref()[0] % (1 << c) */
            u64 limb = w_start / 64, sh = w_start % 64;
            u64 lo = s[limb] >> sh;
            if (sh && limb + 1 < 4) lo |= s[limb + 1] << (64 - sh);
            u64 d = lo & (((u64)1 << c) - 1);
            if (d) g1j_add_mixed(&buckets[d - 1], &bases[i]);
        }
        g1j_t running;
        g1j_set_zero(&running);
        for (u64 b = nb; b-- > 0;) {
            g1j_add(&running, &buckets[b]);
            g1j_add(&res, &running);
        }
        free(buckets);
        window_sums[w] = res;
    }
    g1j_t total;
    g1j_set_zero(&total);
    for (int w = n_windows - 1; w >= 1; w--) {
        g1j_add(&total, &window_sums[w]);
        for (u64 k = 0; k < c; k++) g1j_double(&total);
    }
    g1j_t lowest = window_sums[0];
    g1j_add(&lowest, &total);
    memcpy(out144, &lowest, 144);
    free(window_sums);
}

/* commit_polynomial (worker.rs:117-123): into_repr, zero-pad to bases.len(), MSM */
void orc_commit(const uint8_t *bases104, u64 n_bases, const u64 *fr_mont, u64 n, uint8_t *out144) {
    u64 *sc = (u64 *)calloc(n_bases, 32);
    orc_fr_into_repr(fr_mont, sc, n < n_bases ? n : n_bases);
    orc_msm(bases104, sc, n_bases, out144);
    free(sc);
}

/* the checker's own thread count: torchrun exports OMP_NUM_THREADS=1 to its children, which would make the
 * O(N) checks of a multi-GPU bench run crawl on one core */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
