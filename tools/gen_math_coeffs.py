"""Developer tool: coefficient tables of the lean fp64 exp / atan2 of csrc/ssm_math.cuh (mpmath, 80 digits).

exp(r),  |r| <= ln2/2:  1 + r + r^2 g(r),      g = degree-9 interpolant of (exp(r) - 1 - r) / r^2 at Chebyshev nodes
atan(t), |t| <= tan(pi/8):  t + t s h(s), s = t^2,  h = degree-12 interpolant of (atan(sqrt s)/sqrt s - 1) / s at Chebyshev nodes
Prints the tables as C initialisers and the maximum error of the double-rounded polynomials (exact evaluation).
"""
import mpmath as mp
mp.mp.dps = 80


def cheb_fit(f, a, b, deg):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[j] for j in range(n)]


def to_double(c):
    return [float(x) for x in c]


def horner(c, x):
    p = mp.mpf(c[-1])
    for v in reversed(c[:-1]):
        p = p * x + mp.mpf(v)
    return p


def g_exp(r):
    if abs(r) < mp.mpf(10) ** -30:
        return mp.mpf(1) / 2 + r / 6
    return (mp.exp(r) - 1 - r) / (r * r)


def h_atan(s):
    if s < mp.mpf(10) ** -40:
        return -mp.mpf(1) / 3 + s / 5
    q = mp.sqrt(s)
    return (mp.atan(q) / q - 1) / s


def main():
    a = mp.log(2) / 2 * (1 + mp.mpf(2) ** -40)
    ce = to_double(cheb_fit(g_exp, -a, a, 9))
    worst = 0
    for i in range(4001):
        r = -a + 2 * a * i / 4000
        approx = 1 + r + r * r * horner(ce, r)
        worst = max(worst, abs(approx / mp.exp(r) - 1))
    print('// exp: g coefficients c2..c11, max relative error of the polynomial %.3e (%.3f ulp of 2^-53)' % (float(worst), float(worst * 2 ** 53)))
    print('{' + ', '.join('%.17e' % v for v in ce) + '}')
    T = mp.tan(mp.pi / 8)
    smax = (T * (1 + mp.mpf(2) ** -40)) ** 2
    ca = to_double(cheb_fit(h_atan, 0, smax, 12))
    worst = 0
    for i in range(1, 4001):
        t = T * i / 4000
        s = t * t
        approx = t + t * s * horner(ca, s)
        worst = max(worst, abs(approx / mp.atan(t) - 1))
    print('// atan: h coefficients, max relative error %.3e (%.3f ulp)' % (float(worst), float(worst * 2 ** 53)))
    print('{' + ', '.join('%.17e' % v for v in ca) + '}')
    for name, v in (('L2E', 1 / mp.log(2)), ('LN2_HI', None), ('PIO4', mp.pi / 4), ('PIO2', mp.pi / 2), ('PI', mp.pi), ('TAN_PIO8', T)):
        if v is None:
            continue
        hi = float(v)
        lo = float(v - mp.mpf(hi))
        print('// %s hi %.17e lo %.17e' % (name, hi, lo))
    ln2 = mp.log(2)
    # ln2 split: hi with 32 trailing zero bits would make t * hi exact for |t| < 2^21; with fma a full-precision hi is fine
    hi = float(ln2)
    print('// LN2 hi %.17e lo %.17e' % (hi, float(ln2 - mp.mpf(hi))))


if __name__ == '__main__':
    main()
