// Drives the C++ host mirror (distributed_plonk_b200/host/plonk_worker.hpp) the way the
// reference's tests drive the workers: reads a request file, runs Prover::fft and
// Prover::commit_polynomial against one in-process PlonkImpl, writes the replies.  The Python test
// (tests/test_host_mirror.py) compares them with the oracle.
//   request : u64 n_bases | u64 log_n | u64 log_q | u64 n_coeffs | u64 flags(bit0 quot, bit1 inv, bit2 coset)
//             | bases (n_bases*104) | scalars (n_bases*32 canonical) | coeffs (n_coeffs*32 Fr)
//   reply   : msm partial (144) | fft output (2^L * 32) | p(z) (32) | p / (X - z) ((n_coeffs-1) * 32)
//             | z * p + z * p[..n/2] (n_coeffs * 32)        with p = coeffs, z = coeffs[0]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../distributed_plonk_b200/host/plonk_worker.hpp"

int main(int argc, char **argv) {
    if (argc != 3 && argc != 4) {
        std::fprintf(stderr, "usage: %s request.bin reply.bin [rounds]\n", argv[0]);
        return 2;
    }
    const bool rounds = argc == 4;  // also exercise Promise-style varMsm and the round 4-5 bodies
    std::ifstream in(argv[1], std::ios::binary);
    uint64_t hdr[5];
    in.read(reinterpret_cast<char *>(hdr), sizeof hdr);
    const uint64_t n_bases = hdr[0], log_n = hdr[1], log_q = hdr[2], n_coeffs = hdr[3], flags = hdr[4];
    dplonk::Bytes bases(n_bases * DP_G1_AFFINE_BYTES), scalars(n_bases * DP_FR_BYTES), coeffs(n_coeffs * DP_FR_BYTES);
    in.read(reinterpret_cast<char *>(bases.data()), bases.size());
    in.read(reinterpret_cast<char *>(scalars.data()), scalars.size());
    in.read(reinterpret_cast<char *>(coeffs.data()), coeffs.size());
    if (!in) {
        std::fprintf(stderr, "short request file\n");
        return 2;
    }
    try {
        dplonk::PlonkImpl worker(0, 0, 1);
        worker.init(dplonk::chunks(bases.data(), bases.size()), uint64_t(1) << log_n, uint64_t(1) << log_q);
        std::vector<dplonk::PlonkImpl *> conns{&worker};
        auto parts = dplonk::Prover::commit_polynomial(conns, n_bases, scalars);
        const bool is_quot = flags & 1, is_inv = flags & 2, is_coset = flags & 4;
        auto none = [](void *, void *, uint64_t) {};
        dplonk::Bytes out = dplonk::Prover::fft(conns, is_quot ? log_q : log_n, coeffs, is_quot, is_inv, is_coset, 0xC0FFEE, none);
        // error behaviour: an unknown task id must surface as an exception carrying DP_E_ARG
        bool threw = false;
        try {
            worker.fft2(12345);
        } catch (const dplonk::Error &e) {
            threw = e.code == DP_E_ARG;
        }
        if (!threw) {
            std::fprintf(stderr, "fft2 on an unknown task did not raise DP_E_ARG\n");
            return 1;
        }
        std::ofstream o(argv[2], std::ios::binary);
        o.write(reinterpret_cast<const char *>(parts[0].data()), parts[0].size());
        o.write(reinterpret_cast<const char *>(out.data()), out.size());
        if (!rounds) return 0;
        // varMsm from a Promise: two jobs pending around another request, same partial as the blocking call
        worker.var_msm_begin(1, {0, n_bases}, dplonk::chunks(scalars.data(), scalars.size()));
        worker.var_msm_begin(2, {0, n_bases}, dplonk::chunks(scalars.data(), scalars.size()));
        dplonk::Bytes again = dplonk::Prover::fft(conns, is_quot ? log_q : log_n, coeffs, is_quot, is_inv, is_coset, 0xC0FFEF, none);
        if (worker.var_msm_end(2) != parts[0] || worker.var_msm_end(1) != parts[0] || again != out) {
            std::fprintf(stderr, "asynchronous varMsm / repeated fft differ from the first answers\n");
            return 1;
        }
        // rounds 4-5 on p = coeffs at z = coeffs[0]
        dplonk::Bytes z(coeffs.begin(), coeffs.begin() + DP_FR_BYTES);
        dplonk::Bytes pz = worker.evaluate(coeffs, z), q = worker.witness_poly(coeffs, z);
        dplonk::Bytes one_and_z(2 * DP_FR_BYTES);  // coefficients (z, z): z*p + z*p would need Fr one; use (z, z) on (p, p)
        std::memcpy(one_and_z.data(), z.data(), DP_FR_BYTES);
        std::memcpy(one_and_z.data() + DP_FR_BYTES, z.data(), DP_FR_BYTES);
        dplonk::Bytes lc = worker.lin_comb({coeffs, dplonk::Bytes(coeffs.begin(), coeffs.begin() + coeffs.size() / 2)}, one_and_z);
        o.write(reinterpret_cast<const char *>(pz.data()), pz.size());
        o.write(reinterpret_cast<const char *>(q.data()), q.size());
        o.write(reinterpret_cast<const char *>(lc.data()), lc.size());
    } catch (const dplonk::Error &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
