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
void emu_fr_pow(const void *a, uint64_t e, void *o) { *(Fr *)o = ((const Fr *)a)->pow(e); }
void emu_fr_to_mont(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->to_mont(); }
void emu_fr_from_mont(const void *a, void *o) { *(Fr *)o = ((const Fr *)a)->from_mont(); }
}
