// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the product path
// (moleengine_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// CPU restatement of the arithmetic primitives the Starframe physics tick relies on.
// The reference uses the third-party crate `ultraviolet` ^0.9 (feature f64), which is NOT
// vendored under /root/reference (Cargo.toml:15; Cargo.lock is git-ignored), so its published
// algorithm is restated here and anchored on the reference's own call sites:
//   * src/math.rs:9-12 (type aliases), :199-203 and src/physics/solver.rs:95
//     (angle = -2*atan2(bv.xy, s)  =>  from_angle(t) = {cos(t/2), -sin(t/2)})
//   * src/physics/collision/query.rs:343-348 (pose*ray and pose.inversed()*ray are inverses)
//   * src/physics/solver.rs:435-452 (+w*l*n translation pairs with +I^-1*l*(r^n) rotation
//     => positive angle = counter-clockwise)
// Parity status of these primitives: pinned only through the reference's collision unit
// tests (tests/test_oracle_kat.py ports all 11); ulp-level details (mul_add in dot products)
// are unverifiable offline and are far below every parity tolerance used in this repo.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace sfo {

template <class R>
struct Vec2 {
    R x{0}, y{0};
    constexpr Vec2() = default;
    constexpr Vec2(R x_, R y_) : x(x_), y(y_) {}
    Vec2 operator+(Vec2 o) const { return {x + o.x, y + o.y}; }
    Vec2 operator-(Vec2 o) const { return {x - o.x, y - o.y}; }
    Vec2 operator-() const { return {-x, -y}; }
    Vec2 operator*(R s) const { return {x * s, y * s}; }
    Vec2 operator/(R s) const { return {x / s, y / s}; }
    Vec2& operator+=(Vec2 o) { x += o.x; y += o.y; return *this; }
    Vec2& operator-=(Vec2 o) { x -= o.x; y -= o.y; return *this; }
    R dot(Vec2 o) const { return x * o.x + y * o.y; }
    R mag_sq() const { return x * x + y * y; }
    R mag() const { return std::sqrt(mag_sq()); }
    // ultraviolet: normalize() multiplies by the reciprocal of the magnitude.
    Vec2 normalized() const { R r = R(1) / mag(); return {x * r, y * r}; }
    Vec2 abs() const { return {std::fabs(x), std::fabs(y)}; }
    // a.wedge(b).xy
    R wedge(Vec2 o) const { return x * o.y - o.x * y; }
};
template <class R> inline Vec2<R> operator*(R s, Vec2<R> v) { return {s * v.x, s * v.y}; }

// math.rs:393-396
template <class R> inline Vec2<R> left_normal(Vec2<R> v) { return {-v.y, v.x}; }

template <class R>
struct Rotor2 {
    R s{1}, b{0};  // b = bv.xy
    static Rotor2 from_angle(R angle) {
        R h = angle / R(2);
        return {std::cos(h), -std::sin(h)};
    }
    Rotor2 reversed() const { return {s, -b}; }
    Rotor2 operator*(Rotor2 q) const { return {s * q.s - b * q.b, s * q.b + q.s * b}; }
    Vec2<R> operator*(Vec2<R> v) const {
        R fx = s * v.x + b * v.y;
        R fy = s * v.y - b * v.x;
        return {s * fx + b * fy, s * fy - b * fx};
    }
};

template <class R>
struct Iso2 {
    Vec2<R> t;
    Rotor2<R> r;
    Vec2<R> operator*(Vec2<R> v) const { return r * v + t; }
    Iso2 operator*(Iso2 o) const { return {r * o.t + t, r * o.r}; }
    Iso2 inversed() const {
        Rotor2<R> rr = r.reversed();
        return {rr * (-t), rr};
    }
    void append_translation(Vec2<R> d) { t += d; }
    void prepend_rotation(Rotor2<R> q) { r = q * r; }
};

// Rust float semantics the reference relies on.
template <class R> inline R rs_signum(R v) {
    if (std::isnan(v)) return v;
    return std::signbit(v) ? R(-1) : R(1);
}
template <class R> inline R rs_max(R a, R b) {  // f64::max: ignores NaN
    if (std::isnan(a)) return b;
    if (std::isnan(b)) return a;
    return a > b ? a : b;
}
template <class R> inline R rs_min(R a, R b) {
    if (std::isnan(a)) return b;
    if (std::isnan(b)) return a;
    return a < b ? a : b;
}

template <class R> constexpr R real_max() { return std::numeric_limits<R>::max(); }

}  // namespace sfo
