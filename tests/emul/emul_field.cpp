// Host-emulation shim: compiles the device arithmetic headers with g++ (carry flag emulated in
// ptx_arith.cuh) and exports them for ctypes so tests can check the exact limb algorithms that
// run on the GPU against the oracle on a CPU-only box.  TEST INFRASTRUCTURE ONLY.
#include "../../distributed_plonk_b200/csrc/field.cuh"
#include <cstring>
using namespace dp;
extern "C" {
#define BINOP(name, T, op)                                              \
    void name(const void *a, const void *b, void *o, uint64_t n) {      \
        const T *x = (const T *)a, *y = (const T *)b;                   \
        T *z = (T *)o;                                                  \
        for (uint64_t i = 0; i < n; i++) z[i] = x[i] op y[i];           \
    }
BINOP(emu_fr_mul, Fr, *)
BINOP(emu_fr_add, Fr, +)
BINOP(emu_fr_sub, Fr, -)
BINOP(emu_fq_mul, Fq, *)
BINOP(emu_fq_add, Fq, +)
BINOP(emu_fq_sub, Fq, -)
void emu_fr_inverse(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->inverse(); }
void emu_fq_inverse(const void *a, void *o) { *(Fq *)o = ((const Fq *)a)->inverse(); }
void emu_fr_sqr(const void *a, void *o, uint64_t n) { for (uint64_t i = 0; i < n; i++) ((Fr *)o)[i] = Fr::sqr_inline(((const Fr *)a)[i]); }
void emu_fq_sqr(const void *a, void *o, uint64_t n) { for (uint64_t i = 0; i < n; i++) ((Fq *)o)[i] = Fq::sqr_inline(((const Fq *)a)[i]); }
void emu_fr_inverse_vartime(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->inverse_vartime(); }
void emu_fq_inverse_vartime(const void *a, void *o) { *(Fq *)o = ((const Fq *)a)->inverse_vartime(); }
void emu_fr_pow(const void *a, uint64_t e, void *o) { *(Fr *)o = ((const Fr *)a)->pow(e); }
void emu_fr_to_mont(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->to_mont(); }
void emu_fr_from_mont(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->from_mont(); }
}

#include "../../distributed_plonk_b200/csrc/g1.cuh"
extern "C" {
// k*P with XYZZ double-and-add (exercises dbl + add_mixed incl. the special cases)
void emu_g1_mul(const void *aff96, const uint64_t *k, void *out96) {
    G1Affine p = *(const G1Affine *)aff96;
    G1XYZZ acc = G1XYZZ::inf();
    for (int i = 255; i >= 0; i--) {
        acc = acc.dbl();
        if ((k[i / 64] >> (i % 64)) & 1) acc = acc.add_mixed(p);
    }
    *(G1Affine *)out96 = acc.to_affine();
}
// sum of n affine points: two halves accumulated with add_mixed, merged with the full add
void emu_g1_sum(const void *pts96, uint64_t n, void *out96) {
    const G1Affine *p = (const G1Affine *)pts96;
    G1XYZZ a = G1XYZZ::inf(), b = G1XYZZ::inf();
    for (uint64_t i = 0; i < n; i++) (i & 1 ? b : a) = (i & 1 ? b : a).add_mixed(p[i]);
    *(G1Affine *)out96 = a.add(b).to_affine();
}
void emu_g1_add_xyzz_self(const void *aff96, void *out96) {  // P + P through the full add
    G1XYZZ a = G1XYZZ::from_affine(*(const G1Affine *)aff96);
    *(G1Affine *)out96 = a.add(a).to_affine();
}
}

#include "../../tools/experiments/ufield.cuh"
extern "C" {
// unsaturated-limb fields: operands arrive as raw limb arrays (lazy forms allowed), result limbs out
void emu_fru_mul(const uint32_t *a, const uint32_t *b, uint32_t *o) { *(FrU *)o = FrU::mul(*(const FrU *)a, *(const FrU *)b); }
void emu_fqu_mul(const uint32_t *a, const uint32_t *b, uint32_t *o) { *(FqU *)o = FqU::mul(*(const FqU *)a, *(const FqU *)b); }
void emu_fru_unpack(const uint32_t *w, uint32_t *o) { *(FrU *)o = FrU::unpack(w); }
void emu_fru_pack(const uint32_t *l, uint32_t *w) { ((const FrU *)l)->pack(w); }
void emu_fqu_unpack(const uint32_t *w, uint32_t *o) { *(FqU *)o = FqU::unpack(w); }
void emu_fqu_pack(const uint32_t *l, uint32_t *w) { ((const FqU *)l)->pack(w); }
void emu_fru_normalize(const uint32_t *l, uint32_t *o) { *(FrU *)o = ((const FrU *)l)->normalized(); }
void emu_fru_canonical(const uint32_t *l, uint32_t *o) { *(FrU *)o = ((const FrU *)l)->canonical_from_2p(); }
void emu_fqu_canonical(const uint32_t *l, uint32_t *o) { *(FqU *)o = ((const FqU *)l)->canonical_from_2p(); }
int emu_fru_bias(uint32_t k_log2, uint32_t floor_log2, uint32_t *o) { return make_sub_bias<FrUParams>(k_log2, floor_log2, *(FrU *)o) ? 1 : 0; }
int emu_fqu_bias(uint32_t k_log2, uint32_t floor_log2, uint32_t *o) { return make_sub_bias<FqUParams>(k_log2, floor_log2, *(FqU *)o) ? 1 : 0; }
}
