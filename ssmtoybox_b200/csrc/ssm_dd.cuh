// Double-double arithmetic (~32 significant digits) for the BQ-weights kernel.
// Why: the weights are  wm = q K^-1,  Wc = K^-1 Q K^-1  with K the RBF kernel matrix + 1e-8 I.  For the
// length-scales the reference's own research scripts use (e.g. [25, 25, 1e4, 1e4, 1e4] on the radar model)
// cond(K) ~ 1e9, so a float64 evaluation -- the reference's included -- returns Wc with O(1) relative
// rounding noise.  Evaluating the same formulas in double-double and rounding once at the
// end gives the correctly rounded weights: deterministic, and equal to the float64 evaluation wherever that
// one is meaningful.  Algorithms: Dekker / Knuth error-free transformations (QD library, Hida-Li-Bailey).
#pragma once
#include "ssm_common.cuh"

namespace ssm {

struct dd {
    double hi, lo;
    SSM_DEV dd() : hi(0.0), lo(0.0) {}
    SSM_DEV dd(double h) : hi(h), lo(0.0) {}
    SSM_DEV dd(double h, double l) : hi(h), lo(l) {}
};

SSM_DEV dd quick_two_sum(double a, double b) {
    const double s = __dadd_rn(a, b);
    return dd(s, __dsub_rn(b, __dsub_rn(s, a)));
}
SSM_DEV dd two_sum(double a, double b) {
    const double s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    return dd(s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb)));
}
SSM_DEV dd two_prod(double a, double b) {
    const double p = __dmul_rn(a, b);
    return dd(p, __fma_rn(a, b, -p));
}
SSM_DEV dd operator+(const dd &a, const dd &b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
SSM_DEV dd operator-(const dd &a) { return dd(-a.hi, -a.lo); }
SSM_DEV dd operator-(const dd &a, const dd &b) { return a + (-b); }
SSM_DEV dd operator*(const dd &a, const dd &b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = __dadd_rn(p.lo, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}
SSM_DEV dd operator/(const dd &a, const dd &b) {
    const double q1 = a.hi / b.hi;
    dd r = a - dd(q1) * b;
    const double q2 = r.hi / b.hi;
    r = r - dd(q2) * b;
    const double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return q + dd(q3);
}
SSM_DEV bool operator>(const dd &a, double b) { return a.hi > b || (a.hi == b && a.lo > 0.0); }
SSM_DEV dd tsqrt(const dd &a) {
    if (!(a.hi > 0.0)) return dd(sqrt(a.hi));
    const double x = 1.0 / sqrt(a.hi);
    const double ax = __dmul_rn(a.hi, x);
    const dd e = a - two_prod(ax, ax);
    return two_sum(ax, __dmul_rn(__dmul_rn(e.hi, x), 0.5));
}
SSM_DEV double tsqrt(double a) { return sqrt(a); }
SSM_DEV dd tabs(const dd &a) { return a.hi < 0.0 ? -a : a; }
SSM_DEV double tabs(double a) { return fabs(a); }
SSM_DEV double to_double(const dd &a) { return a.hi; }
SSM_DEV double to_double(double a) { return a; }

// exp in double-double: x = k ln2 + r, exp(r) = (exp(r / 512))^512 with a 12-term Taylor series
SSM_DEV dd texp(const dd &x) {
    if (x.hi <= -709.0) return dd(0.0);
    if (x.hi >= 709.0) return dd(exp(x.hi));
    const dd ln2(6.931471805599452862e-01, 2.319046813846299558e-17);
    const double k = rint(x.hi / ln2.hi);
    dd r = x - dd(k) * ln2;
    r.hi = ldexp(r.hi, -9);
    r.lo = ldexp(r.lo, -9);
    // Taylor: sum_{n>=1} r^n / n!
    dd term = r, sum = r;
    for (int n = 2; n <= 12; ++n) {
        term = term * r / dd((double)n);
        sum = sum + term;
    }
    // (1 + s)^2 - 1 = 2 s + s^2, nine times
    for (int i = 0; i < 9; ++i) sum = sum * dd(2.0) + sum * sum;
    sum = sum + dd(1.0);
    return dd(ldexp(sum.hi, (int)k), ldexp(sum.lo, (int)k));
}
SSM_DEV double texp(double x) { return exp(x); }
// integer power by repeated multiplication
template <class T>
SSM_DEV T tpowi(T x, int n) {
    T r(1.0);
    for (int i = 0; i < n; ++i) r = r * x;
    return r;
}

}  // namespace ssm
