// BLS12-381 G1 group law (y^2 = x^3 + 4 over Fq) for the MSM kernels.
//
// Replaces ark-ec 0.3.0 GroupProjective::{add_assign_mixed, add_assign, double_in_place}
// (Jacobian) that VariableBaseMSM::multi_scalar_mul spends its time in (src/worker.rs:122,179).
// Accumulators here use extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// the mixed addition costs 8M + 2S against 7M + 4S for Jacobian madd-2007-bl and needs no field
// inversion; the representation is free because parity is defined on the group element
//  and the result is normalised before it leaves the device.
//   mixed add : EFD madd-2008-s      full add : EFD add-2008-s      doubling : EFD dbl-2008-s-1
#pragma once
#include "field.cuh"

namespace dp {

DP_HD Fq fq_from_u32(uint32_t v) {
    Fq x = Fq::zero();
    x.l[0] = v;
    return x.to_mont();
}

// affine point as stored on the device: 96 B, (0,0) encodes the point at infinity
// ((0,0) is not on the curve since b = 4 != 0)
struct alignas(16) G1Affine {
    Fq x, y;
    DP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    DP_HD static G1Affine inf() { return G1Affine{Fq::zero(), Fq::zero()}; }
    DP_HD G1Affine neg() const { return G1Affine{x, is_inf() ? y : y.neg()}; }
};

struct alignas(16) G1XYZZ {
    Fq x, y, zz, zzz;
    DP_HD bool is_inf() const { return zz.is_zero(); }
    DP_HD static G1XYZZ inf() { return G1XYZZ{Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
    DP_HD static G1XYZZ from_affine(const G1Affine &p) {
        if (p.is_inf()) return inf();
        return G1XYZZ{p.x, p.y, Fq::one(), Fq::one()};
    }

    // dbl-2008-s-1 (a = 0)
    DP_HD G1XYZZ dbl() const {
        if (is_inf()) return *this;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq xx = x.sqr();
        Fq m = xx.dbl() + xx;
        G1XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this + q, q affine (madd-2008-s); handles infinity, doubling and P + (-P)
    DP_HD G1XYZZ add_mixed(const G1Affine &q) const {
        if (q.is_inf()) return *this;
        if (is_inf()) return from_affine(q);
        Fq u2 = q.x * zz;
        Fq s2 = q.y * zzz;
        Fq p = u2 - x;
        Fq r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) return from_affine(q).dbl();
            return inf();
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq qq = x * pp;
        G1XYZZ o;
        o.x = r.sqr() - ppp - qq.dbl();
        o.y = r * (qq - o.x) - y * ppp;
        o.zz = zz * pp;
        o.zzz = zzz * ppp;
        return o;
    }

    // this + q (add-2008-s)
    DP_HD G1XYZZ add(const G1XYZZ &q) const {
        if (q.is_inf()) return *this;
        if (is_inf()) return q;
        Fq u1 = x * q.zz;
        Fq u2 = q.x * zz;
        Fq s1 = y * q.zzz;
        Fq s2 = q.y * zzz;
        Fq p = u2 - u1;
        Fq r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) return dbl();
            return inf();
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq qq = u1 * pp;
        G1XYZZ o;
        o.x = r.sqr() - ppp - qq.dbl();
        o.y = r * (qq - o.x) - s1 * ppp;
        o.zz = zz * q.zz * pp;
        o.zzz = zzz * q.zzz * ppp;
        return o;
    }

    // -> affine (one field inversion); lone_thread: the caller is a single active lane (see Field::inverse_vartime)
    DP_HD G1Affine to_affine(bool lone_thread = false) const {
        if (is_inf()) return G1Affine::inf();
        const Fq d = zz * zzz;
        const Fq t = lone_thread ? d.inverse_vartime() : d.inverse();
        return G1Affine{x * (t * zzz), y * (t * zz)};
    }
};

// raw ark-ec GroupProjective (Jacobian X,Y,Z; 144 B) as returned over the wire
// (worker.rs:177-179): we emit the normalised representative (x, y, 1) or ark's identity (0, 1, 0).
struct alignas(16) G1JacobianOut {
    Fq x, y, z;
    DP_HD static G1JacobianOut from_affine(const G1Affine &a) {
        if (a.is_inf()) return G1JacobianOut{Fq::zero(), Fq::one(), Fq::zero()};
        return G1JacobianOut{a.x, a.y, Fq::one()};
    }
};

}  // namespace dp
