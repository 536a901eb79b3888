// C++ host-side mirror of the reference worker / dispatcher for the hot path, over the C ABI.
//
// The reference's host code is Rust (src/worker.rs, src/dispatcher2.rs); no Rust toolchain exists in
// this image, so the layer that sits between the Cap'n Proto surface and include/dplonk.h is
// written in C++ with the reference's names, argument meaning and error behaviour:
//
//   dplonk::PlonkImpl            <->  impl plonk_slave::Server / plonk_peer::Server for PlonkImpl
//                                     (worker.rs:125-439; methods of hello_world.capnp:15-52)
//   dplonk::Prover::fft          <->  Prover::fft                  (dispatcher2.rs:731-787)
//   dplonk::Prover::commit_polynomial <-> Prover::commit_polynomial (dispatcher2.rs:834-893)
//   PlonkImpl::quotient_evals / evaluate / lin_comb / witness_poly  <->  the round 3-5 arithmetic of
//                                     Prover::prove (dispatcher2.rs:363-690), i.e. the bodies of the
//                                     round3*/round4*/round5* RPCs the schema declares
//
// `ListData` is the in-memory form of a capnp `List(Data)`: byte chunks cut at 2^28 bytes
// (dispatcher.rs:61-63).  Where the reference `unwrap()`s (panics -> capnp error to the caller)
// this throws dplonk::Error carrying the DP_E_* code and dp_last_error() text.
// Header-only; link against libdplonk.so.  The Python twin is distributed_plonk_b200/worker.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dplonk.h"

namespace dplonk {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("dplonk " + std::to_string(c) + ": " + m), code(c) {}
};

using Bytes = std::vector<uint8_t>;
using ListData = std::vector<Bytes>;
constexpr size_t kChunk = size_t(1) << 28;

// v.chunks(1 << 28)   (dispatcher.rs:61-63)
inline ListData chunks(const void *p, size_t n) {
    ListData out;
    const uint8_t *b = static_cast<const uint8_t *>(p);
    for (size_t off = 0; off < n; off += kChunk) out.emplace_back(b + off, b + off + (n - off < kChunk ? n - off : kChunk));
    if (out.empty()) out.emplace_back();
    return out;
}
// extend_from_slice over the chunks (worker.rs:136-141, 172-175, 244-247)
inline Bytes concat(const ListData &l) {
    if (l.size() == 1) return l[0];
    Bytes out;
    for (const Bytes &c : l) out.insert(out.end(), c.begin(), c.end());
    return out;
}

struct MsmWorkload { uint64_t start, end; };                       // utils.rs:21-25
using FftWorkload = dp_fft_workload;                               // utils.rs:3-19

class PlonkImpl {
   public:
    PlonkImpl(int cuda_device, uint64_t me, uint64_t n_workers) : me_(me), n_workers_(n_workers) {
        int rc = dp_create(cuda_device, me, n_workers, &ctx_);
        if (rc != DP_OK) throw Error(rc, dp_last_error(nullptr));
    }
    ~PlonkImpl() { dp_destroy(ctx_); }
    PlonkImpl(const PlonkImpl &) = delete;
    PlonkImpl &operator=(const PlonkImpl &) = delete;

    // init @0 (bases :List(Data), domainSize :UInt64, quotDomainSize :UInt64)      worker.rs:126-157
    void init(const ListData &bases, uint64_t domain_size, uint64_t quot_domain_size) {
        Bytes b = concat(bases);
        check(dp_init(ctx_, b.data(), b.size() / DP_G1_AFFINE_BYTES, domain_size, quot_domain_size));
        log_[0] = log2_ceil(domain_size);
        log_[1] = log2_ceil(quot_domain_size);
    }
    // varMsm @1 (workload, scalars :List(Data)) -> (result :Data)                   worker.rs:159-185
    Bytes var_msm(const MsmWorkload &w, const ListData &scalars) {
        Bytes s = concat(scalars), out(DP_G1_PROJECTIVE_BYTES);
        check(dp_msm(ctx_, w.start, w.end, s.data(), s.size() / DP_FR_BYTES, out.data()));
        return out;
    }
    // fftInit @2                                                                    worker.rs:187-233
    void fft_init(uint64_t id, const std::vector<FftWorkload> &workloads, bool is_quot, bool is_inv, bool is_coset) {
        check(dp_fft_init(ctx_, id, workloads.data(), workloads.size(), is_quot, is_inv, is_coset));
        const uint32_t L = log_[is_quot ? 1 : 0];
        Dims d;
        d.r = uint64_t(1) << (L >> 1);
        d.n_cols = workloads[me_].col_end - workloads[me_].col_start;
        dims_[id] = d;
    }
    // fft1 @3 (id, i, v :List(Data))                                                worker.rs:235-278
    void fft1(uint64_t id, uint64_t i, const ListData &v) {
        Bytes row = concat(v);
        check(dp_fft1(ctx_, id, i, row.data(), row.size() / DP_FR_BYTES));
    }
    // fft2Prepare @4 (id)  [+ PlonkPeer.fftExchange for several workers]            worker.rs:280-345, 412-438
    // `exchange(send_dev, recv_dev, block_elems)` performs the all-to-all for n_workers > 1.
    template <class Exchange>
    void fft2_prepare(uint64_t id, Exchange &&exchange) {
        if (n_workers_ == 1) {
            check(dp_fft2_prepare(ctx_, id));
            return;
        }
        void *s = nullptr, *r = nullptr;
        uint64_t blk = 0;
        check(dp_fft_exchange_begin(ctx_, id, &s, &r, &blk));
        exchange(s, r, blk);
        check(dp_fft_exchange_end(ctx_, id));
    }
    void fft2_prepare(uint64_t id) {
        fft2_prepare(id, [](void *, void *, uint64_t) { throw Error(DP_E_COMM, "no exchange for a multi-worker task"); });
    }
    // fft2 @5 (id) -> (v :List(Data)), one Data per local column                    worker.rs:347-381
    ListData fft2(uint64_t id) {
        auto it = dims_.find(id);
        if (it == dims_.end()) throw Error(DP_E_ARG, "fft2: unknown task");
        const Dims d = it->second;
        dims_.erase(it);
        Bytes all(d.n_cols * d.r * DP_FR_BYTES);
        check(dp_fft2(ctx_, id, all.data(), all.size()));
        ListData out;
        const size_t col = d.r * DP_FR_BYTES;
        for (uint64_t k = 0; k < d.n_cols; k++) out.emplace_back(all.begin() + k * col, all.begin() + (k + 1) * col);
        return out;
    }
    // round1 @6 (w :List(Data)) -> (c :Data)                                        worker.rs:383-408
    // blind_2fr: two uniformly random secret Fr (the reference draws them from ThreadRng); nullptr = the
    // library draws them from the operating system's entropy pool (getrandom), never from a fixed seed
    Bytes round1(const ListData &w, const uint8_t *blind_2fr = nullptr) {
        Bytes e = concat(w), out(DP_G1_PROJECTIVE_BYTES);
        check(dp_round1(ctx_, e.data(), e.size() / DP_FR_BYTES, blind_2fr, out.data()));
        return out;
    }
    // varMsm answered from a Promise (the way fft2Prepare is, worker.rs:293): begin in the RPC body,
    // end when the reply is built; other requests are served - and their kernels queued - in between
    void var_msm_begin(uint64_t id, const MsmWorkload &w, const ListData &scalars) {
        Bytes &s = pending_scalars_[id];   // the copy-in is asynchronous: keep the bytes until the job is collected
        s = concat(scalars);
        check(dp_msm_submit(ctx_, id, w.start, w.end, s.data(), s.size() / DP_FR_BYTES));
    }
    Bytes var_msm_end(uint64_t id) {
        Bytes out(DP_G1_PROJECTIVE_BYTES);
        int rc = dp_msm_collect(ctx_, id, out.data());
        pending_scalars_.erase(id);
        check(rc);
        return out;
    }
    // init from the canonical encoding of the SRS (ark-serialize compressed G1, 48 B per point)
    void init_compressed(const ListData &bases48, uint64_t domain_size, uint64_t quot_domain_size, bool check_subgroup = true) {
        Bytes b = concat(bases48);
        check(dp_init_compressed(ctx_, b.data(), b.size() / DP_G1_COMPRESSED_BYTES, domain_size, quot_domain_size, check_subgroup));
        log_[0] = log2_ceil(domain_size);
        log_[1] = log2_ceil(quot_domain_size);
    }

    // ---- the bodies of the declared-but-unimplemented round3* / round4* / round5* RPCs
    // (hello_world.capnp:26-44); the arithmetic is the dispatcher's, src/dispatcher2.rs:363-690.
    // Polynomials are raw Fr vectors (32 B per coefficient / evaluation).
    // round 3: quotient evaluations over the quotient coset (434-504); every array quot_domain_size long
    Bytes quotient_evals(const std::vector<Bytes> &selectors /*13*/, const std::vector<Bytes> &sigmas /*5*/,
                         const std::vector<Bytes> &wires /*5*/, const Bytes &perm, const Bytes &pub_input, const Bytes &k /*5 Fr*/,
                         const Bytes &alpha, const Bytes &beta, const Bytes &gamma) {
        if (selectors.size() != 13 || sigmas.size() != 5 || wires.size() != 5) throw Error(DP_E_ARG, "quotient_evals: 13 selectors, 5 sigmas, 5 wires");
        dp_quotient_args a;
        for (int i = 0; i < 13; i++) a.selectors[i] = selectors[i].data();
        for (int i = 0; i < 5; i++) {
            a.sigmas[i] = sigmas[i].data();
            a.wires[i] = wires[i].data();
        }
        a.perm = perm.data();
        a.pub_input = pub_input.data();
        a.k = k.data();
        a.alpha = alpha.data();
        a.beta = beta.data();
        a.gamma = gamma.data();
        Bytes out(perm.size());
        check(dp_quotient_evals(ctx_, &a, out.data()));
        return out;
    }
    // round 4: poly.evaluate(&point) (535-548)
    Bytes evaluate(const Bytes &coeffs, const Bytes &point) {
        Bytes out(DP_FR_BYTES);
        check(dp_poly_eval(ctx_, coeffs.data(), coeffs.size() / DP_FR_BYTES, point.data(), out.data()));
        return out;
    }
    // round 5: sum_i coeffs[i] * polys[i] (566-649)
    Bytes lin_comb(const std::vector<Bytes> &polys, const Bytes &coeffs) {
        std::vector<const void *> ptr;
        std::vector<size_t> len;
        size_t longest = 0;
        for (const Bytes &p : polys) {
            ptr.push_back(p.data());
            len.push_back(p.size() / DP_FR_BYTES);
            longest = len.back() > longest ? len.back() : longest;
        }
        Bytes out(longest * DP_FR_BYTES);
        check(dp_poly_lincomb(ctx_, ptr.data(), len.data(), coeffs.data(), polys.size(), out.data(), longest));
        return out;
    }
    // round 5: witness polynomial p(X) / (X - point) (651-666, 672-688)
    Bytes witness_poly(const Bytes &coeffs, const Bytes &point) {
        const size_t n = coeffs.size() / DP_FR_BYTES;
        Bytes out(n ? (n - 1) * DP_FR_BYTES : 0);
        check(dp_poly_div_linear(ctx_, coeffs.data(), n, point.data(), out.data(), nullptr));
        return out;
    }
    dp_ctx *raw() { return ctx_; }
    uint64_t me() const { return me_; }

   private:
    struct Dims { uint64_t r, n_cols; };
    void check(int rc) {
        if (rc != DP_OK) throw Error(rc, dp_last_error(ctx_));
    }
    static uint32_t log2_ceil(uint64_t n) {
        uint32_t l = 0;
        while ((uint64_t(1) << l) < n) l++;
        return l;
    }
    dp_ctx *ctx_ = nullptr;
    uint64_t me_, n_workers_;
    uint32_t log_[2] = {0, 0};
    std::map<uint64_t, Dims> dims_;
    std::map<uint64_t, Bytes> pending_scalars_;
};

// The dispatcher's side of the path, against in-process workers instead of capnp connections.
struct Prover {
    // FftWorkload per worker: equal row / column blocks (dispatcher2.rs:1143-1156)
    static std::vector<FftWorkload> workloads(uint32_t domain_log, uint64_t n_slaves) {
        const uint64_t r = uint64_t(1) << (domain_log >> 1), c = (uint64_t(1) << domain_log) / r;
        std::vector<FftWorkload> wl;
        for (uint64_t p = 0; p < n_slaves; p++) wl.push_back({p * r / n_slaves, (p + 1) * r / n_slaves, p * c / n_slaves, (p + 1) * c / n_slaves});
        return wl;
    }
    // Prover::fft (dispatcher2.rs:731-787).  coeffs: raw Fr (32 B each), resized to the domain.
    template <class Exchange>
    static Bytes fft(std::vector<PlonkImpl *> &connections, uint32_t domain_log, Bytes coeffs, bool is_quot, bool is_inv,
                     bool is_coset, uint64_t id, Exchange &&exchange) {
        const uint64_t n_slaves = connections.size();
        const uint64_t N = uint64_t(1) << domain_log, r = uint64_t(1) << (domain_log >> 1), c = N / r;
        coeffs.resize(N * DP_FR_BYTES, 0);                                        // :746
        auto wl = workloads(domain_log, n_slaves);
        for (auto *conn : connections) conn->fft_init(id, wl, is_quot, is_inv, is_coset);    // :748-753
        // t = transpose(coeffs.chunks(r)): row b = { x[b + a*r] }                  :754
        Bytes row(c * DP_FR_BYTES);
        for (uint64_t p = 0; p < n_slaves; p++)
            for (uint64_t j = 0; j < r / n_slaves; j++) {                             // one fft1 per row :756-766
                const uint64_t b = p * r / n_slaves + j;
                for (uint64_t a = 0; a < c; a++) std::memcpy(&row[a * DP_FR_BYTES], &coeffs[(b + a * r) * DP_FR_BYTES], DP_FR_BYTES);
                connections[p]->fft1(id, j, chunks(row.data(), row.size()));
            }
        for (auto *conn : connections) conn->fft2_prepare(id, exchange);          // :767-772
        Bytes out(N * DP_FR_BYTES);
        for (uint64_t p = 0; p < n_slaves; p++) {                                  // :774-786
            ListData cols = connections[p]->fft2(id);
            for (uint64_t k = 0; k < cols.size(); k++) {
                const uint64_t i = p * c / n_slaves + k;                           // u[i] = column i
                for (uint64_t j = 0; j < r; j++) std::memcpy(&out[(j * c + i) * DP_FR_BYTES], &cols[k][j * DP_FR_BYTES], DP_FR_BYTES);
            }
        }
        return out;
    }
    // Prover::commit_polynomial (dispatcher2.rs:834-893) with the tested contract of
    // dispatcher.rs:213-229 (global index ranges over full bases).  Returns the workers' partials.
    static std::vector<Bytes> commit_polynomial(std::vector<PlonkImpl *> &connections, uint64_t n_bases, Bytes plain_coeffs) {
        const uint64_t n_slaves = connections.size();
        plain_coeffs.resize(n_bases * DP_FR_BYTES, 0);                            // :843
        std::vector<Bytes> parts;
        for (uint64_t i = 0; i < n_slaves; i++) {
            const uint64_t lo = i * n_bases / n_slaves, hi = (i + 1) * n_bases / n_slaves;
            parts.push_back(connections[i]->var_msm({lo, hi}, chunks(&plain_coeffs[lo * DP_FR_BYTES], (hi - lo) * DP_FR_BYTES)));
        }
        return parts;
    }
};

}  // namespace dplonk
