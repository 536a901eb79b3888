// Throughput of the unsaturated-limb Montgomery products (ufield.cuh) against the carry-chain ones.
#include <cstdio>
#include <cuda_runtime.h>
#include "../distributed_plonk_b200/csrc/field.cuh"
#include "experiments/ufield.cuh"
using namespace dp;

template <class F>
__global__ void mb_sat(F *out, int iters) {
    F x = F::one(), y = F::r2();
    x.l[0] += threadIdx.x;
    F u = y, v = x;
    for (int it = 0; it < iters; it++) { x = x * y; u = u * v; y = y * x; v = v * u; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + u + v;
}
template <class F>
__global__ void mb_unsat(F *out, int iters) {
    F x = F::modulus(), y = F::modulus();
    x.l[0] -= threadIdx.x + 1;
    y.l[1] -= 77;
    F u = y, v = x;
    for (int it = 0; it < iters; it++) { x = x * y; u = u * v; y = y * x; v = v * u; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + u + v;
}
// butterfly-like mix: lazy add/sub feeding the product
template <class F>
__global__ void mb_unsat_bfly(F *out, const F *bias, int iters) {
    F x = F::modulus(), y = F::modulus(), w = F::modulus();
    x.l[0] -= threadIdx.x + 1;
    y.l[1] -= 77;
    w.l[2] -= 5;
    const F b = bias[0];
    for (int it = 0; it < iters; it++) {
        F s = (x + y).normalized();
        F d = F::sub(x, y, b) * w;
        x = s;
        y = d;
        x = x * w;  // keep the value bounded in this synthetic loop
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}
template <class K, class... A>
float tk(K k, int blocks, int tpb, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, tpb>>>(args...);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<blocks, tpb>>>(args...);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    void *buf; cudaMalloc(&buf, (size_t)sms * 8 * 256 * 64);
    FrU hb; make_sub_bias<FrUParams>(3, 28, hb);
    FrU *db; cudaMalloc(&db, sizeof(FrU)); cudaMemcpy(db, &hb, sizeof hb, cudaMemcpyHostToDevice);
    for (int occ = 1; occ <= 4; occ *= 2) {
        int blocks = sms * occ, tpb = 256, iters = 512;
        double n = (double)blocks * tpb * iters * 4;
        float a = tk(mb_sat<Fr>, blocks, tpb, (Fr *)buf, iters);
        float b = tk(mb_unsat<FrU>, blocks, tpb, (FrU *)buf, iters);
        float c = tk(mb_sat<Fq>, blocks, tpb, (Fq *)buf, iters);
        float d = tk(mb_unsat<FqU>, blocks, tpb, (FqU *)buf, iters);
        float e = tk(mb_unsat_bfly<FrU>, blocks, tpb, (FrU *)buf, (const FrU *)db, iters);
        printf("blocks/SM=%d  Fr mul: carry-chain %.1f G/s, unsaturated %.1f G/s | Fq mul: carry-chain %.1f G/s, unsaturated %.1f G/s | FrU butterfly(2 mul+add+sub) %.1f G/s\n",
               occ, n / a / 1e6, n / b / 1e6, n / c / 1e6, n / d / 1e6, (double)blocks * tpb * iters / e / 1e6);
    }
    return 0;
}
