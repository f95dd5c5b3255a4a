"""CPU restatement of numpyro's NUTS (one chain, numpy) -- TEST INFRASTRUCTURE ONLY.

Follows numpyro/infer/hmc_util.py (build_tree, _double_tree, _iterative_build_subtree,
_combine_tree, _is_turning, _leaf_idx_to_ckpt_idxs, warmup_adapter, dual_averaging,
welford_covariance, build_adaptation_schedule) and hmc.py's sample_kernel for a diagonal mass
matrix, which is what biolith/utils/fit.py:93 instantiates (NUTS(model, init_to_uniform)).
numpyro is not vendored in /root/reference and not installable here, so parity of the sampler is
*statistical* (posterior moments within Monte Carlo error), not bitwise: jax's threefry key
splitting is replaced by numpy's Generator.  Used by tests/ to check the device sampler.
"""

from __future__ import annotations

import numpy as np


def build_adaptation_schedule(num_steps):
    if num_steps <= 0:
        return []
    if num_steps < 20:
        return [(0, num_steps - 1)]
    start_buffer, end_buffer, init_window = 75, 50, 25
    if start_buffer + end_buffer + init_window > num_steps:
        start_buffer = int(0.15 * num_steps)
        end_buffer = int(0.1 * num_steps)
        init_window = num_steps - start_buffer - end_buffer
    sched = [(0, start_buffer - 1)]
    end_window_start = num_steps - end_buffer
    next_size, next_start = init_window, start_buffer
    while next_start < end_window_start:
        cur_start, cur_size = next_start, next_size
        if 3 * cur_size <= end_window_start - cur_start:
            next_size = 2 * cur_size
        else:
            cur_size = end_window_start - cur_start
        next_start = cur_start + cur_size
        sched.append((cur_start, next_start - 1))
    sched.append((end_window_start, num_steps - 1))
    return sched


def _is_turning(imm, r_left, r_right, r_sum):
    r_sum = r_sum - (r_left + r_right) / 2
    return (np.dot(imm * r_left, r_sum) <= 0) or (np.dot(imm * r_right, r_sum) <= 0)


def _leaf_idx_to_ckpt_idxs(n):
    idx_max = bin(n >> 1).count("1")
    num_subtrees = 0
    m = n
    while m & 1:
        m >>= 1
        num_subtrees += 1
    return idx_max - num_subtrees + 1, idx_max


class _Tree:
    __slots__ = ("zl", "rl", "gl", "zr", "rr", "gr", "zp", "up", "gp", "depth", "weight", "r_sum", "turning",
                 "diverging", "sum_acc", "num")


def nuts_chain(logp_grad, theta0, num_warmup, num_samples, rng, max_tree_depth=10, target_accept=0.8,
               step_size=1.0, max_delta_energy=1000.0):
    """logp_grad(theta) -> (logp, grad).  Returns dict(samples, accept_prob, num_steps, diverging, step_size, imm)."""
    D = theta0.size
    z = np.array(theta0, dtype=np.float64)
    lp, g = logp_grad(z)
    U, g = -lp, np.asarray(g, np.float64)
    imm = np.ones(D)
    eps = step_size
    sched = build_adaptation_schedule(num_warmup)
    win = 0
    # dual averaging state
    prox, x_t, x_avg, g_avg, tt = np.log(10 * eps), 0.0, 0.0, 0.0, 0
    wn, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
    out = dict(samples=[], accept_prob=[], num_steps=[], diverging=[])
    n_leap = 0
    for t in range(num_warmup + num_samples):
        r = rng.standard_normal(D) / np.sqrt(imm)
        E0 = U + 0.5 * np.dot(imm * r, r)
        T = _Tree()
        T.zl = T.zr = T.zp = z
        T.rl = T.rr = r
        T.gl = T.gr = T.gp = g
        T.up, T.depth, T.weight, T.r_sum = U, 0, 0.0, r.copy()
        T.turning = T.diverging = False
        T.sum_acc, T.num = 0.0, 0
        r_ck = np.zeros((max_tree_depth, D))
        rs_ck = np.zeros((max_tree_depth, D))
        while T.depth < max_tree_depth and not T.turning and not T.diverging:
            right = rng.uniform() < 0.5
            # ---- _iterative_build_subtree
            S = None
            s_turning = False
            max_num = 2 ** T.depth
            num = 0
            while num < max_num and not s_turning and not (S is not None and S.diverging):
                if S is None:
                    z0, r0, g0 = (T.zr, T.rr, T.gr) if right else (T.zl, T.rl, T.gl)
                else:
                    z0, r0, g0 = (S.zr, S.rr, S.gr) if right else (S.zl, S.rl, S.gl)
                de = eps if right else -eps
                rh = r0 + 0.5 * de * g0
                zn = z0 + de * imm * rh
                lpn, gn = logp_grad(zn)
                n_leap += 1
                gn = np.asarray(gn, np.float64)
                rn = rh + 0.5 * de * gn
                delta = (-lpn + 0.5 * np.dot(imm * rn, rn)) - E0
                if np.isnan(delta):
                    delta = np.inf
                w_leaf, div, acc = -delta, delta > max_delta_energy, min(1.0, np.exp(-delta))
                if S is None:
                    S = _Tree()
                    S.zl = S.zr = S.zp = zn
                    S.rl = S.rr = rn
                    S.gl = S.gr = S.gp = gn
                    S.up, S.weight, S.r_sum = -lpn, w_leaf, rn.copy()
                    S.diverging, S.sum_acc = div, acc
                else:
                    if right:
                        S.zr, S.rr, S.gr = zn, rn, gn
                    else:
                        S.zl, S.rl, S.gl = zn, rn, gn
                    S.r_sum = S.r_sum + rn
                    tp = 1.0 / (1.0 + np.exp(-(w_leaf - S.weight)))
                    if rng.uniform() < tp:
                        S.zp, S.gp, S.up = zn, gn, -lpn
                    S.weight = np.logaddexp(S.weight, w_leaf)
                    S.diverging = div
                    S.sum_acc += acc
                leaf_idx = num
                num += 1
                imin, imax = _leaf_idx_to_ckpt_idxs(leaf_idx)
                if leaf_idx % 2 == 0:
                    r_ck[imax], rs_ck[imax] = rn, S.r_sum
                else:
                    i = imax
                    while i >= imin and not s_turning:
                        sub = S.r_sum - rs_ck[i] + r_ck[i]
                        s_turning = _is_turning(imm, r_ck[i], rn, sub)
                        i -= 1
            # ---- _combine_tree (biased)
            if right:
                T.zr, T.rr, T.gr = S.zr, S.rr, S.gr
            else:
                T.zl, T.rl, T.gl = S.zl, S.rl, S.gl
            T.r_sum = T.r_sum + S.r_sum
            tp = 0.0 if (s_turning or S.diverging) else min(1.0, np.exp(S.weight - T.weight))
            turning = True if s_turning else _is_turning(imm, T.rl, T.rr, T.r_sum)
            if rng.uniform() < tp:
                T.zp, T.gp, T.up = S.zp, S.gp, S.up
            T.depth += 1
            T.weight = np.logaddexp(T.weight, S.weight)
            T.diverging, T.turning = S.diverging, turning
            T.sum_acc += S.sum_acc
            T.num += num
        accept_prob = T.sum_acc / T.num
        z, g, U = T.zp, T.gp, T.up
        if t >= num_warmup:
            out["samples"].append(z.copy())
            out["accept_prob"].append(accept_prob)
            out["num_steps"].append(T.num)
            out["diverging"].append(T.diverging)
        else:
            tt += 1
            g_avg = (1 - 1 / (tt + 10.0)) * g_avg + (target_accept - accept_prob) / (tt + 10.0)
            x_t = prox - np.sqrt(tt) / 0.05 * g_avg
            wgt = tt ** (-0.75)
            x_avg = (1 - wgt) * x_avg + wgt * x_t
            eps = float(np.exp(x_avg if t == num_warmup - 1 else x_t))
            middle = 0 < win < len(sched) - 1
            if middle:
                wn += 1
                pre = z - wmean
                wmean = wmean + pre / wn
                wm2 = wm2 + pre * (z - wmean)
            at_end = t == sched[win][1]
            if at_end:
                win += 1
            if at_end and middle:
                var = wm2 / (wn - 1)
                imm = (wn / (wn + 5.0)) * var + 1e-3 * (5.0 / (wn + 5.0))
                wn, wmean, wm2 = 0, np.zeros(D), np.zeros(D)
                prox, x_t, x_avg, g_avg, tt = np.log(10 * eps), 0.0, 0.0, 0.0, 0
    out = {k: np.array(v) for k, v in out.items()}
    out.update(step_size=eps, inverse_mass_matrix=imm, leapfrogs=n_leap)
    return out
