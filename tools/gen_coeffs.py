"""Generate the polynomial coefficients of ggdmc_b200/csrc/gg_fastmath.cuh (run once, offline).

  g(q)  = Q(a) / exp(-a^2/2) * (a + c),  a = c (1 + q) / (1 - q),  q in [-1, 1]
          (Q = upper normal tail; so  Q(a) = exp(-a^2/2) * g(q) / (a + c))
  e(r)  = exp(r) on [-ln2/2, ln2/2]
Near-minimax polynomials by Chebyshev interpolation in 60-digit arithmetic, converted to the
monomial basis, rounded to double; the script then measures the error of double Horner evaluation.
"""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_interp_monomial(f, lo, hi, deg):
    n = deg + 1
    xs = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fv = [f((hi - lo) / 2 * x + (hi + lo) / 2) for x in xs]
    c = [sum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) * 2 / n for j in range(n)]
    c[0] /= 2
    # Chebyshev -> monomial in t in [-1, 1]
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        a = [mp.mpf(0)] + [2 * v for v in T[j - 1]]
        b = T[j - 2] + [mp.mpf(0)] * (len(a) - len(T[j - 2]))
        T.append([x - y for x, y in zip(a, b)])
    mono_t = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono_t[i] += c[j] * v
    # t = (2x - (hi+lo)) / (hi - lo): substitute to get monomial in x
    al, be = 2 / (hi - lo), -(hi + lo) / (hi - lo)
    mono_x = [mp.mpf(0)] * n
    # (al x + be)^i expansion
    for i, ci in enumerate(mono_t):
        for k in range(i + 1):
            mono_x[k] += ci * mp.binomial(i, k) * al ** k * be ** (i - k)
    return mono_x


def horner_double(coef, x):
    r = np.full_like(x, coef[-1])
    for c in coef[-2::-1]:
        r = r * x + c  # numpy has no fma; good enough for an error estimate
    return r


def main():
    out = {}
    C = mp.mpf(sys.argv[1]) if len(sys.argv) > 1 else mp.mpf(4)

    def g(q):
        if q >= 1:
            return 1 / mp.sqrt(2 * mp.pi)
        a = C * (1 + q) / (1 - q)
        return mp.mpf(0.5) * mp.erfc(a / mp.sqrt(2)) * mp.exp(a * a / 2) * (a + C)

    for deg in (22, 23, 24, 25, 26):
        co = cheb_interp_monomial(g, mp.mpf(-1), mp.mpf(1), deg)
        cd = [float(v) for v in co]
        qs = np.linspace(-1, 0.9999, 4001)
        got = horner_double(cd, qs)
        ref = np.array([float(g(mp.mpf(float(q)))) for q in qs])
        # also error of the exact (unrounded) polynomial
        ex = max(abs(sum(co[i] * mp.mpf(float(q)) ** i for i in range(len(co))) - g(mp.mpf(float(q)))) / g(mp.mpf(float(q))) for q in qs[::40])
        print(f"g c={float(C)} deg {deg}: max rel err double-horner {np.max(np.abs(got - ref) / ref):.2e}  approx err {float(ex):.2e}  sum|c|={sum(abs(v) for v in cd):.2f}")
        out[("g", deg)] = cd
    ln2h = mp.log(2) / 2
    for deg in (10, 11, 12, 13):
        co = cheb_interp_monomial(mp.exp, -ln2h, ln2h, deg)
        cd = [float(v) for v in co]
        rs = np.linspace(float(-ln2h), float(ln2h), 2001)
        got = horner_double(cd, rs)
        ref = np.array([float(mp.exp(mp.mpf(float(r)))) for r in rs])
        print(f"exp deg {deg}: max rel err {np.max(np.abs(got - ref) / ref):.2e}")
        out[("e", deg)] = cd
    import json
    json.dump({f"{k[0]}{k[1]}": v for k, v in out.items()}, open("/tmp/coeffs.json", "w"))


if __name__ == "__main__":
    main()
