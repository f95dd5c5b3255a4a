"""CPU emulation (numpy, fp32 op by op) of where K1d's residual gradient error near the posterior mode comes from.

Measured on B200 (profiles/r02_strict_k1d.txt): K1d with MUFU functions AND K1d with libm exp2f / log2f / IEEE division
both sit at 6.3e-6 of |g|inf near the mode, the libm engine (exp only at negative arguments) at 1.4e-6.  Hypothesis:
MUFU.EX2 -- which libm's exp2f also ends in -- has a SIGN-DEPENDENT mean relative error (profiles/r02_mufu_error.txt:
-5e-8 for negative, +3e-8 for positive arguments).  A uniform relative bias of every e_j scales the gradient by
(1 + O(bias)) and vanishes against |g|; a bias that differs between "expected" visits (x2 < 0) and "surprising" ones
(x2 > 0) does not cancel.  This script evaluates the K1d formulation and the engine formulation in fp32 with exactly
rounded functions, with and without that bias injected, against an fp64 reference.

A second finding of the same emulation (not shipped -- found after the round's GPU budget was spent): the remaining
1e-6 of K1d's formulation with exact functions is the LAST fma of the argument chain, x2 = fma(sgn, A_hi, t).  A_hi =
fl(-log2(e) alpha_0) carries 24 significant bits; whenever |t| is in a coarser binade than A_hi, its low bits are
rounded away the same way for every such visit (mean error -1.2e-8 sgn, 150 standard errors from zero).  Putting A_hi
on a 2^-17 grid (and the rest, < 4e-6, into A_lo, which is then added exactly enough at the start of the chain) makes
the rounding of that fma depend on t's random low bits only: the emulated error drops to the engine's 1.4e-7."""
import sys
import numpy as np

S = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
J, KS, KO = 8, 5, 3
rng = np.random.default_rng(0)
f32 = np.float32
beta = rng.normal(size=KS + 1); alpha = rng.normal(size=KO + 1)
X = rng.normal(size=(S, KS)); W = rng.normal(size=(S, J, KO))
psi = 1 / (1 + np.exp(-(beta[0] + X @ beta[1:])))
z = rng.uniform(size=S) < psi
p = 1 / (1 + np.exp(-(alpha[0] + W @ alpha[1:])))
y = (rng.uniform(size=(S, J)) < p) & z[:, None]
theta = np.concatenate([beta, alpha]) + 2e-3 * rng.standard_normal(KS + KO + 2)
theta = theta.astype(f32).astype(np.float64)
X = X.astype(f32); W = W.astype(f32)
b, a = theta[:KS + 1], theta[KS + 1:]
sgn = np.where(y, 1.0, -1.0)
LOG_TINY = np.log(np.finfo(np.float32).tiny)


def site_level(eta, L1, n1, ga, Xd, dt):
    """logaddexp over z and the gradient, in dtype dt with exactly rounded functions"""
    eta = eta.astype(dt); L1 = L1.astype(dt)
    d = (eta + L1 - n1.astype(dt) * dt(LOG_TINY)).astype(dt)
    rr = (1 / (1 + np.exp(-d.astype(np.float64)))).astype(dt)
    ps = (1 / (1 + np.exp(-eta.astype(np.float64)))).astype(dt)
    geta = (rr - ps).astype(dt)
    gb = np.concatenate([[geta.astype(np.float64).sum()], (geta[:, None] * Xd.astype(dt)).astype(np.float64).sum(0)])
    gal = (rr[:, None] * ga.astype(dt)).astype(np.float64).sum(0)
    return np.concatenate([gb, gal])


def reference():
    nu = a[0] + W.astype(np.float64) @ a[1:]
    xs = sgn * nu
    L1 = -np.log1p(np.exp(-xs)).sum(1)
    q = 1 / (1 + np.exp(xs))
    ga = np.concatenate([(q * sgn).sum(1)[:, None], np.einsum("sj,sjk->sk", q * sgn, W.astype(np.float64))], 1)
    eta = b[0] + X.astype(np.float64) @ b[1:]
    return site_level(eta, L1, y.sum(1), ga, X, np.float64) - theta


def k1d(bias_neg, bias_pos, grid_bits=None, delta=0.0):
    """grid_bits: put the high part of the two-float intercept on a 2^-grid_bits grid (None: A_hi = fl(A), as shipped);
    delta: offset of the exponent's argument (the shipped kEx2Shift = 5e-8 / ln 2)."""
    L2E = 1.4426950408889634
    v = (-(L2E) * (sgn[:, :, None] * W.astype(np.float64))).astype(f32)          # records, rounded per element
    A = -L2E * a[0]
    A_hi = f32(A) if grid_bits is None else f32(np.rint(A * 2.0 ** grid_bits) / 2.0 ** grid_bits)
    A_lo = f32(A - np.float64(A_hi))
    a2 = a[1:].astype(f32)
    s32 = sgn.astype(f32)
    x2 = (s32.astype(np.float64) * np.float64(A_lo) + delta).astype(f32)
    for k in range(KO):
        x2 = (v[:, :, k].astype(np.float64) * np.float64(a2[k]) + x2.astype(np.float64)).astype(f32)   # fma: one rounding
    x2 = (s32.astype(np.float64) * np.float64(A_hi) + x2.astype(np.float64)).astype(f32)
    e = np.exp2(x2.astype(np.float64)) * np.where(x2 < 0, 1 + bias_neg, 1 + bias_pos)
    e = e.astype(f32)
    u = (f32(1) + e).astype(f32)
    pr = (u[:, 0::2] * u[:, 1::2]).astype(f32)
    pa = (pr[:, 0] * pr[:, 1]).astype(f32); pb = (pr[:, 2] * pr[:, 3]).astype(f32); pp = (pa * pb).astype(f32)
    rinv = (1 / pp.astype(np.float64)).astype(f32)
    lg = np.log2(pp.astype(np.float64)).astype(f32)
    ra = (rinv * pb).astype(f32); rb = (rinv * pa).astype(f32)
    rp = np.stack([(ra * pr[:, 1]).astype(f32), (ra * pr[:, 0]).astype(f32), (rb * pr[:, 3]).astype(f32),
                   (rb * pr[:, 2]).astype(f32)], 1)
    uo = u.reshape(S, 4, 2)[:, :, ::-1].reshape(S, J)                             # u[j ^ 1]
    q = (e * (np.repeat(rp, 2, axis=1) * uo).astype(f32)).astype(f32)
    ga = np.zeros((S, KO + 1), f32)
    vv = np.concatenate([s32[:, :, None], v], 2)
    for j in range(J):
        ga = (q[:, j, None].astype(np.float64) * vv[:, j].astype(np.float64) + ga.astype(np.float64)).astype(f32)
    ga = ga.astype(np.float64); ga[:, 1:] *= -np.log(2.0)                          # record units -> natural
    L1 = (-np.float64(f32(np.log(2.0))) * lg.astype(np.float64)).astype(f32)
    eta = np.full(S, f32(b[0]))
    for k in range(KS):
        eta = (X[:, k].astype(np.float64) * np.float64(f32(b[1 + k])) + eta.astype(np.float64)).astype(f32)
    return site_level(eta, L1, y.sum(1), ga.astype(f32), X, f32) - theta


def engine(bias_neg):
    nu = np.full((S, J), f32(a[0]))
    for k in range(KO):
        nu = (W[:, :, k].astype(np.float64) * np.float64(f32(a[1 + k])) + nu.astype(np.float64)).astype(f32)
    t = (np.exp(-np.abs(nu.astype(np.float64))) * (1 + bias_neg)).astype(f32)     # always a negative argument
    l = np.log1p(t.astype(np.float64)).astype(f32)
    inv = (1 / (1 + t.astype(np.float64))).astype(f32)
    pj = np.where(nu >= 0, inv, (t * inv).astype(f32)).astype(f32)
    term = np.where(y, np.minimum(nu, 0) - l, -np.maximum(nu, 0) - l).astype(f32)
    g = np.where(y, f32(1) - pj, -pj).astype(f32)
    L1 = term.astype(np.float64).sum(1).astype(f32)
    ga = np.concatenate([g.astype(np.float64).sum(1)[:, None], np.einsum("sj,sjk->sk", g.astype(np.float64), W.astype(np.float64))], 1)
    eta = np.full(S, f32(b[0]))
    for k in range(KS):
        eta = (X[:, k].astype(np.float64) * np.float64(f32(b[1 + k])) + eta.astype(np.float64)).astype(f32)
    return site_level(eta, L1, y.sum(1), ga.astype(f32), X, f32) - theta


ref = reference()
ginf = np.abs(ref).max()
print(f"S = {S}, |g|inf = {ginf:.1f} (sum |terms| ~ {S * 0.3:.0f})")
for name, g in (("K1d formulation, exact functions", k1d(0.0, 0.0)),
                ("K1d formulation, ex2 bias -5e-8 (x2 < 0) / +3e-8 (x2 > 0)", k1d(-5e-8, 3e-8)),
                ("K1d formulation, uniform ex2 bias -5e-8", k1d(-5e-8, -5e-8)),
                ("K1d, sign-dependent bias + the shipped argument offset 5e-8 / ln 2", k1d(-5e-8, 3e-8, None, 5e-8 / np.log(2.0))),
                ("K1d, exact functions, A_hi on a 2^-17 grid (NOT shipped: next step)", k1d(0.0, 0.0, 17)),
                ("K1d, bias + offset, A_hi on a 2^-17 grid (NOT shipped: next step)", k1d(-5e-8, 3e-8, 17, 5e-8 / np.log(2.0))),
                ("engine formulation, exact functions", engine(0.0)),
                ("engine formulation, ex2 bias -5e-8 (always negative argument)", engine(-5e-8))):
    print(f"{name:70s} gradient error / |g|inf = {np.abs(g - ref).max() / ginf:.2e}")
