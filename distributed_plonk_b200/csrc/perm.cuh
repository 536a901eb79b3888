// Round-2 grand product of the TurboPlonk permutation argument on the GPU.
//
// "Next" row 3 of SURVEY.md §8(f): the reference computes it serially on the dispatcher with one
// field division per row (src/dispatcher2.rs:329-345):
//     z[0] = 1,   z[j+1] = z[j] * a_j / b_j,      j < n-1
//     a_j = prod_i (w_i[j] + gamma + beta * id_i[j]),   b_j = prod_i (w_i[j] + gamma + beta * sigma_i[j])
// Here: one elementwise kernel for (a_j, b_j), a multiplicative exclusive-prefix scan of a, a
// multiplicative inclusive-suffix scan of b, ONE field inversion (of T = prod b), and
//     z[j] = (prod_{k<j} a_k) * (prod_{k>=j} b_k) / T
// so no per-row division at all.  Every value is a canonical Montgomery Fr, hence byte-identical to
// the reference's sequential result.
#pragma once
#include "ntt.cuh"

namespace dp {

constexpr int PERM_TPB = 256;
constexpr int PERM_ITEMS = 4;
constexpr int PERM_BLOCK = PERM_TPB * PERM_ITEMS;

// a[j], b[j] for j < n-1; a[n-1] = b[n-1] = 1
__global__ void perm_terms_kernel(const Fr *wires, const Fr *id, const Fr *sigma, uint32_t n_types, uint64_t n, Fr beta, Fr gamma,
                                  Fr *a, Fr *b) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fr pa = Fr::one(), pb = Fr::one();
    if (j + 1 < n) {
        for (uint32_t i = 0; i < n_types; i++) {
            const Fr t = gmem_ld(wires + i * n + j) + gamma;
            pa = pa * (t + beta * gmem_ld(id + i * n + j));
            pb = pb * (t + beta * gmem_ld(sigma + i * n + j));
        }
    }
    gmem_st(a + j, pa);
    gmem_st(b + j, pb);
}

// logical index -> memory index (reverse = 1 scans from the end: suffix products)
DP_D uint64_t perm_idx(uint64_t i, uint64_t n, uint32_t reverse) { return reverse ? n - 1 - i : i; }

// block-wide exclusive multiplicative scan of one value per thread; returns this thread's exclusive
// prefix and the block total (Hillis-Steele over shared memory)
DP_D Fr perm_block_scan(const Fr &v, Fr *sh, Fr &total) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    sh[tid] = v;
    __syncthreads();
    for (uint32_t off = 1; off < nt; off <<= 1) {
        Fr t = sh[tid];
        if (tid >= off) t = sh[tid - off] * t;
        __syncthreads();
        sh[tid] = t;
        __syncthreads();
    }
    total = sh[nt - 1];
    const Fr excl = tid ? sh[tid - 1] : Fr::one();
    __syncthreads();
    return excl;
}

// phase 1: product of each block of PERM_BLOCK logical elements
__global__ void __launch_bounds__(PERM_TPB) perm_block_products_kernel(const Fr *x, uint64_t n, uint32_t reverse, Fr *block_tot) {
    __shared__ Fr sh[PERM_TPB];
    const uint64_t base = (uint64_t)blockIdx.x * PERM_BLOCK + (uint64_t)threadIdx.x * PERM_ITEMS;
    Fr p = Fr::one();
    for (int k = 0; k < PERM_ITEMS; k++)
        if (base + k < n) p = p * gmem_ld(x + perm_idx(base + k, n, reverse));
    Fr total;
    perm_block_scan(p, sh, total);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

// phase 2 (single block): exclusive scan of the block products in place; grand total -> *total_out
__global__ void __launch_bounds__(PERM_TPB) perm_block_offsets_kernel(Fr *block_tot, uint32_t n_blocks, Fr *total_out) {
    __shared__ Fr sh[PERM_TPB];
    Fr run = Fr::one();
    for (uint32_t base = 0; base < n_blocks; base += PERM_TPB) {
        const uint32_t i = base + threadIdx.x;
        const Fr v = i < n_blocks ? block_tot[i] : Fr::one();
        Fr total;
        const Fr e = perm_block_scan(v, sh, total);
        if (i < n_blocks) block_tot[i] = run * e;
        run = run * total;
    }
    if (threadIdx.x == 0) *total_out = run;
}

// phase 3: out[i] = prod of logical elements before i (inclusive = 0) or up to and including i
__global__ void __launch_bounds__(PERM_TPB) perm_scan_write_kernel(const Fr *x, uint64_t n, uint32_t reverse, uint32_t inclusive,
                                                                    const Fr *block_tot, Fr *out) {
    __shared__ Fr sh[PERM_TPB];
    const uint64_t base = (uint64_t)blockIdx.x * PERM_BLOCK + (uint64_t)threadIdx.x * PERM_ITEMS;
    Fr v[PERM_ITEMS];
    Fr p = Fr::one();
    for (int k = 0; k < PERM_ITEMS; k++) {
        v[k] = base + k < n ? gmem_ld(x + perm_idx(base + k, n, reverse)) : Fr::one();
        p = p * v[k];
    }
    Fr total;
    Fr run = block_tot[blockIdx.x] * perm_block_scan(p, sh, total);
    for (int k = 0; k < PERM_ITEMS; k++)
        if (base + k < n) {
            if (inclusive) run = run * v[k];
            gmem_st(out + perm_idx(base + k, n, reverse), run);
            if (!inclusive) run = run * v[k];
        }
}

__global__ void perm_invert_kernel(Fr *t) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *t = t->inverse();
}

// z[j] = prefA_excl[j] * sufB_incl[j] * t_inv
__global__ void perm_finish_kernel(const Fr *pref_a, const Fr *suf_b, const Fr *t_inv, uint64_t n, Fr *z) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    gmem_st(z + j, gmem_ld(pref_a + j) * (gmem_ld(suf_b + j) * gmem_ld(t_inv)));
}

}  // namespace dp
