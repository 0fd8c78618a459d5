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
