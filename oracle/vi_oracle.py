"""Value-iteration oracle (numpy, float64).  TEST INFRASTRUCTURE ONLY.

Restates pomdpy/solvers/value_iteration.py of the reference:
  * value_iteration (:24-73): the reference's projection formula
        v_new[g,u,z,i] = sum_j v_g[i] * O[u,i,z] * T[u,j,i]        (:51-58; indexed by i, summed over j)
        candidate(g,u,z) = discount * (R[u] + v_new[g,u,z])           (:60-67; no cross-sum)
    the candidate list the reference builds (|A| * |G|^|Z| * |Z| objects) has exactly the |A|*|Z|*|G| distinct
    members enumerated here;
  * prune (:88-142): the pairwise LP  min d  s.t.  b >= 0, sum b = 1, d >= 0, (a_i - a_j).b + d >= delta  has
    d* = max(0, delta - max_s (a_i[s] - a_j[s])), so "res.x[n] > 0" is the pointwise test
    max_s (a_i[s] - a_j[s]) < delta with delta = 1e-10 (test/vi_test.py:27-29 pins exactly this).  The reference's
    set iteration order is by object address (A.9), so parity is defined on the upper envelope and on the
    canonically pruned set, both order-independent.
  * select_action (:161-181): arg-max of alpha.b.
Also the textbook projection Gamma^{a,o} = T_a diag(O_a[:,o]) Gamma used by the synthetic |S| = 4096 configuration.
"""
import numpy as np

DELTA = 1e-10


def reference_candidates(gamma, T, O, R, discount):
    """gamma [G][S] -> candidates [(u, z, g) order][S] and their actions.  Summation over j in index order, each
    term (v*o)*t as the reference evaluates it."""
    G, S = gamma.shape
    A, Z = T.shape[0], O.shape[2]
    cand = np.zeros((A, Z, G, S), dtype=np.float64)
    for u in range(A):
        for z in range(Z):
            vo = gamma * O[u, :, z][None, :]                 # v_g[i] * o[u][i][z]
            acc = np.zeros((G, S), dtype=np.float64)
            for j in range(S):
                acc = acc + vo * T[u, j, :][None, :]          # += (v*o) * t[u][j][i]
            cand[u, z] = discount * (R[u][None, :] + acc)
    actions = np.repeat(np.arange(A), Z * G)
    return cand.reshape(A * Z * G, S), actions


def pointwise_prune(vecs, actions, delta=DELTA):
    """Keep vector i unless some j has max_s(v_i - v_j) < delta and (j does not lose to i the same way, or j comes
    first).  Returns the kept indices in order."""
    n = len(vecs)
    keep = []
    for i in range(n):
        d_ij = (vecs[i][None, :] - vecs).max(axis=1)          # max_s (v_i - v_j) for every j
        d_ji = (vecs - vecs[i][None, :]).max(axis=1)
        dom = d_ij < delta
        dom[i] = False
        mutual = dom & (d_ji < delta)
        beaten = dom & (~mutual | (np.arange(n) < i))
        if not beaten.any():
            keep.append(i)
    return np.array(keep, dtype=np.int64)


def value_iteration(T, O, R, discount, horizon):
    """The reference's loop with the canonical prune.  Returns (alpha [n][S], action [n])."""
    T, O, R = (np.asarray(x, dtype=np.float64) for x in (T, O, R))
    S = T.shape[1]
    gamma = np.zeros((1, S))
    acts = np.array([-1])
    for k in range(horizon):
        cand, cact = reference_candidates(gamma, T, O, R, discount)
        if k == 0:                                           # the dummy zero vector is dropped after the first pass
            allv, alla = cand, cact
        else:
            allv, alla = np.concatenate([gamma, cand]), np.concatenate([acts, cact])
        keep = pointwise_prune(allv, alla)
        gamma, acts = allv[keep], alla[keep]
    return gamma, acts


def envelope(gamma, beliefs):
    return (beliefs @ gamma.T).max(axis=1)


def select_action(belief, gamma, acts):
    v = gamma @ np.asarray(belief, dtype=np.float64)
    i = int(np.argmax(v))                                     # first maximiser, like the reference's strict '>'
    return int(acts[i]), i


def canonical_set(vecs, actions, delta=DELTA):
    """Order-independent form of a vector set: pointwise-pruned, then sorted lexicographically."""
    vecs = np.asarray(vecs, dtype=np.float64)
    actions = np.asarray(actions)
    order = np.lexsort(tuple(vecs[:, s] for s in reversed(range(vecs.shape[1]))) + (actions,))
    order = np.lexsort((actions,) + tuple(vecs[:, s] for s in reversed(range(vecs.shape[1]))))
    vecs, actions = vecs[order], actions[order]
    keep = pointwise_prune(vecs, actions, delta)
    return vecs[keep], actions[keep]


def textbook_projection(gamma, T, O):
    """proj[a, o, g, s] = sum_s' T[a, s, s'] * O[a, s', o] * gamma[g, s']   (float64 einsum)."""
    return np.einsum("ast,ato,gt->aogs", T, O, gamma, optimize=True)


def point_based_backup(T, O, R, gamma, beliefs, discount):
    """float64 restatement of vi_point_based_backup_dev (include/vi_b200.h): returns alpha [P][S], action [P],
    best [A][P][Z], value [A][P], score [A][P][Z][G]."""
    proj = textbook_projection(gamma, T, O)                       # [a][o][g][s]
    score = np.einsum("ps,aogs->apog", beliefs, proj, optimize=True)
    best = score.argmax(axis=3)                                   # first maximiser
    value = beliefs @ R.T                                         # [p][a]
    value = value.T + discount * score.max(axis=3).sum(axis=2)    # [a][p]
    action = value.argmax(axis=0)
    P, S = beliefs.shape
    Z = O.shape[2]
    alpha = np.empty((P, S))
    for p in range(P):
        a = action[p]
        alpha[p] = R[a] + discount * proj[a, np.arange(Z), best[a, p], :].sum(axis=0)
    return alpha, action, best, value, score


# ---------------------------------------------------------------------------------------------------------------
# Policy execution and belief expansion (checkers of vi_policy_rollout_dev / vi_expand_beliefs_dev): the same
# arithmetic in the same order -- sequential sums where the kernel sums sequentially, lane-strided partial sums and a
# butterfly / pairwise-halving reduction where a warp / the block reduces -- so that the comparison is exact.
# ---------------------------------------------------------------------------------------------------------------
VI_RT = 256


def _u53(a, b):
    return float(((int(a) >> 5) << 26) | (int(b) >> 6)) * (1.0 / 9007199254740992.0)


def _sample_index(p, u):
    acc = 0.0
    for j in range(len(p)):
        acc = acc + float(p[j])
        if u < acc:
            return j
    return len(p) - 1


def _block_sum(parts):
    red = np.array(parts, dtype=np.float64)
    off = VI_RT // 2
    while off >= 1:
        red[:off] = red[:off] + red[off:2 * off]
        off //= 2
    return float(red[0])


def _strided_parts(vals):
    """per-thread partial sums: thread t adds the elements t, t + 256, ... in that order"""
    parts = np.zeros(VI_RT)
    for j0 in range(0, len(vals), VI_RT):
        chunk = vals[j0:j0 + VI_RT]
        parts[:len(chunk)] = parts[:len(chunk)] + chunk
    return parts


def _pred(Ta, b):
    """pred[j] = sum_i T[a][i][j] b[i], i in index order"""
    acc = np.zeros(Ta.shape[1])
    for i in range(Ta.shape[0]):
        acc = acc + Ta[i, :] * b[i]
    return acc


def _warp_dot(v, b):
    lanes = np.zeros(32)
    for j0 in range(0, len(v), 32):
        c = v[j0:j0 + 32] * b[j0:j0 + 32]
        lanes[:len(c)] = lanes[:len(c)] + c
    x = lanes
    for off in (16, 8, 4, 2, 1):
        x = x + x[np.arange(32) ^ off]
    return float(x[0])


def belief_update(b, Ta, Oa, z):
    nb = _pred(Ta, b) * Oa[:, z]
    tot = _block_sum(_strided_parts(nb))
    return (nb / tot if tot > 0.0 else b), tot


def policy_rollout(alpha, act, T, O, R, b0, discount, max_steps, n_epochs, seed=0, term_action_mask=0):
    """agent.py:215-264 on a tabular model; returns [n_epochs][4] = discounted, undiscounted, steps, last action."""
    import oracle
    key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    S = T.shape[1]
    out = np.zeros((n_epochs, 4))
    for e in range(n_epochs):
        b = np.array(b0, dtype=np.float64)
        r = oracle.philox([0xFFFFFFFF, e, 0, 5], key)
        s = _sample_index(b, _u53(r[0], r[1]))
        ret = dret = 0.0
        disc, steps, last = 1.0, 0, -1
        for t in range(max_steps):
            vals = [_warp_dot(alpha[g], b) for g in range(alpha.shape[0])]
            g = int(np.argmax(vals))                                  # first maximum = lowest g on ties
            a = int(act[g])
            r = oracle.philox([t, e, 0, 5], key)
            s2 = _sample_index(T[a, s], _u53(r[0], r[1]))
            z = _sample_index(O[a, s2], _u53(r[2], r[3]))
            rew = float(R[a, s])
            ret = ret + rew
            dret = dret + disc * rew
            disc = disc * discount
            s, steps, last = s2, steps + 1, a
            if (term_action_mask >> a) & 1:
                break
            b, _ = belief_update(b, T[a], O[a], z)
        out[e] = (dret, ret, steps, last)
    return out


def expand_beliefs(B, P, T, O, round_index, seed=0):
    """rows P..2P-1 = one successor of each of the first P points (random action, sampled observation)."""
    import oracle
    key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    A, Z = T.shape[0], O.shape[2]
    B = np.array(B, dtype=np.float64)
    for p in range(P):
        r = oracle.philox([round_index, p, 0, 6], key)
        a = (int(r[0]) * A) >> 32
        pred = _pred(T[a], B[p])
        pz = [_block_sum(_strided_parts(pred * O[a][:, z])) for z in range(Z)]
        tot = 0.0
        for v in pz:
            tot = tot + v
        z = _sample_index(pz, _u53(r[1], r[2]) * tot)
        nb = pred * O[a][:, z]
        t2 = _block_sum(_strided_parts(nb))
        B[P + p] = nb / t2 if t2 > 0.0 else B[p]
    return B
