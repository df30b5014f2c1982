"""Coefficients of the device exp polynomial (physics.cuh: fexp) and an accuracy check of the
operation sequences of flog / fexp, emulated in NumPy doubles against mpmath.

exp(r) on |r| <= ln2/2:  1 + r + r^2 q(r), q of degree D-2 from Chebyshev interpolation of
(exp(r) - 1 - r) / r^2 at 60 digits, rounded to double.  log uses fdlibm's Lg1..Lg7 (e_log.c)."""
import sys
import numpy as np
import mpmath as mp

mp.mp.dps = 60
D = int(sys.argv[1]) if len(sys.argv) > 1 else 12
a = mp.log(2) / 2 * mp.mpf("1.0001")


def q(r):
    r = mp.mpf(r)
    if abs(r) < mp.mpf("1e-25"):
        return mp.mpf(1) / 2 + r / 6
    return (mp.exp(r) - 1 - r) / (r * r)


n = D - 1   # coefficients of q
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = q(x)
c = mp.lu_solve(A, b)
coef = [float(ci) for ci in c]
print("// exp: q(r) coefficients, degree", D)
print(", ".join("%.17e" % v for v in coef))

LOG2E = float(mp.mpf(1) / mp.log(2))
LN2_HI = 6.93147180369123816490e-01
LN2_LO = 1.90821492927058770002e-10
SHIFT = 6755399441055744.0


def fexp(x):
    x = np.asarray(x, dtype=np.float64)
    t = x * LOG2E + SHIFT          # (an fma on the device; the difference cannot change n off a tie)
    fn = t - SHIFT
    n_ = fn.astype(np.int64)
    # r = fma(fn, -ln2_hi, x): exact in doubles here because ln2_hi has 32 significant bits and |fn| < 2^11
    r = x - fn * LN2_HI
    r = r - fn * LN2_LO
    p = np.full_like(r, coef[-1])
    for cf in coef[-2::-1]:
        p = p * r + cf
    r2 = r * r
    res = 1.0 + (r + r2 * p)
    return np.ldexp(res, n_.astype(np.int32))


Lg = [6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
      2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
      1.479819860511658591e-01]


def flog(x):
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    hi = (bits >> 32).astype(np.int64)
    lo = bits & 0xffffffff
    hi = hi + (0x3ff00000 - 0x3fe6a09e)
    k = (hi >> 20) - 0x3ff
    hi = (hi & 0x000fffff) + 0x3fe6a09e
    m = ((hi << 32) | lo).view(np.float64)
    f = m - 1.0
    s = f / (2.0 + f)              # the device uses f * frcp(2 + f), ~1 ulp from this
    z = s * s
    w = z * z
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    R = t2 + t1
    hfsq = 0.5 * f * f
    dk = k.astype(np.float64)
    return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f)


rng = np.random.default_rng(0)
xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-1, 1, 20000), rng.uniform(-40, 40, 20000)])
got = fexp(xs)
err = max(abs((mp.mpf(float(g)) - mp.exp(mp.mpf(float(x)))) / mp.exp(mp.mpf(float(x)))) for g, x in zip(got[::7], xs[::7]))
print("exp: max rel err = %.3e (%.2f ulp)" % (float(err), float(err) / 2.2e-16))
xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000), rng.uniform(1e-3, 1e3, 20000),
                     1.0 + rng.uniform(-1e-6, 1e-6, 2000)])
got = flog(xs.copy())
e_rel = 0; e_abs = 0
for g, x in zip(got[::5], xs[::5]):
    ref = mp.log(mp.mpf(float(x)))
    d = abs(mp.mpf(float(g)) - ref)
    e_abs = max(e_abs, d / max(1, abs(ref)))
    if ref != 0: e_rel = max(e_rel, d / abs(ref))
print("log: max rel err = %.3e, max err / max(1,|log|) = %.3e" % (float(e_rel), float(e_abs)))
