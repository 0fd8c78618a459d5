"""Pin the C oracle to the EXECUTED reference and (re)generate tests/golden/.

TEST INFRASTRUCTURE ONLY; runs in the build container where /root/reference exists:

    python oracle/refharness/validate.py            # validate + write tests/golden/*.npz|json

What is compared (reference file:line in parentheses):
  * per-cell legal-action sets (examples/rock_sample/rock_model.py:218-247) and cell codes (:110-138);
  * sensor thresholds: np.random.binomial(1.0, p) (:395) == [U <= exp(log(1-(1-p)))] against the REAL
    numpy legacy generator, incl. draws within 1e-9 of the threshold;
  * generate_step (:451-465) over a grid of (cell, action, rocks, world, uniform at threshold +-1 ulp);
  * the complete sequential POMCP (pomdpy/solvers/pomcp.py:69-137, belief_tree_solver.py:42-212,
    action_selectors.py:6-37, discrete_action_mapping.py:137-172, pomdpy/pomdp/model.py:221-252): with
    random.choice / random.shuffle / np.random.binomial patched to the sequential Philox stream, root
    (visit_count, total_q, mean_q, legal), the chosen action and the real env step must be identical to
    the oracle's "ref" mode, move after move, including the belief updates in between.
"""
import contextlib
import io
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
import ref_env  # noqa: E402
from philox_py import SeqStream  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden")
MAPS = ["map-7-8.txt", "map-11-11.txt", "map-15-15.txt"]
TWO53 = float(1 << 53)


def thr_of_p(p):
    """numpy legacy binomial(1, p), p > 0.5: returns 1 iff U <= exp(1*log(1-q)), q = 1-p."""
    p = float(p)
    q = 1.0 - p
    return math.exp(math.log(1.0 - q))


def tables_from_reference(ref, m):
    rows, cols, K = m.n_rows, m.n_cols, m.n_rocks
    cell = np.array(m.env_map, dtype=np.int8).reshape(-1)
    GP = ref["GridPosition"]
    thr53 = np.zeros((rows * cols, K), dtype=np.uint64)
    prob = np.zeros((rows * cols, K), dtype=np.float64)
    legal = np.zeros(rows * cols, dtype=np.uint32)
    for i in range(rows):
        for j in range(cols):
            c = i * cols + j
            for k in range(K):
                d = GP(i, j).euclidean_distance(m.rock_positions[k])
                p = m.get_sensor_correctness_probability(d)
                prob[c, k] = p
                thr53[c, k] = int(math.floor(thr_of_p(p) * TWO53))
            if cell[c] != -3:
                st = ref["RockState"](GP(i, j), [0] * K)
                for a in m.get_legal_actions(st):
                    legal[c] |= np.uint32(1 << a)
    start = m.start_position.i * cols + m.start_position.j
    return dict(rows=rows, cols=cols, n_rocks=K, cell=cell, start_cell=start, thr53=thr53, prob=prob, legal=legal,
                rewards=np.array([-m.illegal_move_penalty, m.exit_reward, m.good_rock_reward, -m.bad_rock_penalty]))


def oracle_model(t):
    return oracle.Model(t["rows"], t["cols"], t["cell"], t["start_cell"], t["thr53"], t["rewards"])


def check_binomial_semantics(t, n=20000):
    """Real np.random.binomial(1.0, p) vs the threshold rule, on the real legacy generator."""
    ps = np.unique(t["prob"])
    rng = np.random.RandomState(12345)
    bad = 0
    for it in range(n):
        p = float(ps[it % len(ps)])
        st = rng.get_state()
        u = rng.random_sample()
        rng.set_state(st)
        got = rng.binomial(1.0, p)
        st2 = rng.get_state()
        rng.set_state(st)
        rng.random_sample()
        same_consumption = all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b
                               for a, b in zip(st2, rng.get_state()))
        want = 1 if int(u * TWO53) <= int(math.floor(thr_of_p(p) * TWO53)) else 0
        if got != want or not same_consumption:
            bad += 1
    return bad


class Injector:
    """Patch the reference's random sources to the sequential Philox stream."""

    def __init__(self, stream):
        self.s = stream

    def __enter__(self):
        self._choice, self._shuffle, self._binom = random.choice, random.shuffle, np.random.binomial
        s = self.s

        def choice(seq):
            return seq[s.randbelow(len(seq))]

        def shuffle(x):
            for i in reversed(range(1, len(x))):
                j = s.randbelow(i + 1)
                x[i], x[j] = x[j], x[i]

        def binomial(n, p):
            assert n == 1.0
            return 1 if s.u53() <= int(math.floor(thr_of_p(p) * TWO53)) else 0

        random.choice, random.shuffle, np.random.binomial = choice, shuffle, binomial
        return self

    def __exit__(self, *a):
        random.choice, random.shuffle, np.random.binomial = self._choice, self._shuffle, self._binom


class FixedU:
    def __init__(self, u53):
        self.u = u53

    def randbelow(self, n):
        raise AssertionError

    def u53(self):
        return self.u


def check_generate_step(ref, m, t, om, rng, n_random=6):
    """Exhaustive over (cell, action) x sampled rock masks x worlds x threshold-straddling uniforms."""
    K, A = t["n_rocks"], 5 + t["n_rocks"]
    GP, RS = ref["GridPosition"], ref["RockState"]
    rows, cols = t["rows"], t["cols"]
    vectors, mism = [], 0
    for c in range(rows * cols):
        if t["cell"][c] == -3:
            continue
        i, j = divmod(c, cols)
        for a in range(A):
            for rep in range(n_random):
                rocks = int(rng.integers(0, 1 << K))
                actual = int(rng.integers(0, 1 << K))
                sampled = int(rng.integers(0, 1 << K)) if rep % 2 else 0
                if a >= 5:
                    th = int(t["thr53"][c, a - 5])
                    u = [th, max(th - 1, 0), min(th + 1, (1 << 53) - 1), int(rng.integers(0, 1 << 53))][rep % 4]
                else:
                    u = int(rng.integers(0, 1 << 53))
                m.actual_rock_states = [actual & (1 << k) for k in range(K)]
                m.unique_rocks_sampled = [k for k in range(K) if (sampled >> k) & 1]
                st = RS(GP(i, j), [rocks & (1 << k) for k in range(K)])
                with Injector(FixedU(u)):
                    res, legal = m.generate_step(st, a)
                nr = sum((1 << k) for k in range(K) if res.next_state.rock_states[k])
                ncell = res.next_state.position.i * cols + res.next_state.position.j
                mutated = sum((1 << k) for k in range(K) if st.rock_states[k]) != rocks
                o = om.step(c, rocks, a, u, actual, sampled)
                want = (ncell, nr, int(bool(res.observation.is_good)), int(bool(res.observation.is_empty)),
                        float(res.reward), int(bool(res.is_terminal)), int(bool(legal)), int(mutated))
                got = (o.next_cell, o.next_rocks, o.obs_good, o.obs_empty, o.reward, o.terminal, o.legal,
                       o.rewarded_sample)
                if want != got:
                    mism += 1
                    if mism < 5:
                        print("MISMATCH", (c, rocks, a, u, actual, sampled), want, got)
                vectors.append((c, rocks, a, u, actual, sampled) + want)
    return mism, vectors


def world_and_bag(m, seed, n_start):
    """agent.py:135 reset_for_epoch then belief_tree_solver.py:36-38, with the real numpy generator."""
    np.random.seed(seed)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.reset_for_epoch()
        actual = sum((1 << k) for k in range(m.n_rocks) if m.actual_rock_states[k])
    return actual


def run_reference_episode(ref, map_file, seed, n_sims, n_moves, stream_seed, **kw):
    """Drive the executed reference exactly as Agent.run_pomcp does (agent.py:150-195), draws injected."""
    import warnings
    m = ref_env.make_rock_model(ref, map_file, n_sims=n_sims, **kw)
    K = m.n_rocks
    trace = []
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        np.random.seed(seed)
        m.reset_for_epoch()
        actual = sum((1 << k) for k in range(K) if m.actual_rock_states[k])
        agent = ref["Agent"](m, ref["POMCP"])
        old_N, old_n = ref["POMCP"].UCB_N, ref["POMCP"].UCB_n
        ref["POMCP"].UCB_N, ref["POMCP"].UCB_n = 1, 1          # table == closed form; saves 3 s
        solver = agent.solver_factory(agent)                    # (stays 1x1 for the episode: find_fast_ucb
        bag = [sum((1 << k) for k in range(K) if p.rock_states[k]) for p in solver.belief_tree_index.state_particles]
        stream = SeqStream(stream_seed)
        with Injector(stream):
            state = solver.belief_tree_index.sample_particle()
            for mv in range(n_moves):
                import time
                action = solver.select_eps_greedy_action(0.5, time.time())
                amap = solver.belief_tree_index.action_map
                stats = dict(
                    n=[amap.entries[a].visit_count for a in range(5 + K)],
                    total=[float(amap.entries[a].total_q_value) for a in range(5 + K)],
                    mean=[float(amap.entries[a].mean_q_value) for a in range(5 + K)],
                    legal=sum((1 << a) for a in range(5 + K) if amap.entries[a].is_legal),
                    N=amap.total_visit_count)
                res, legal = m.generate_step(state, action)
                state = res.next_state
                rec = dict(action=action.bin_number, stats=stats, obs_good=int(bool(res.observation.is_good)),
                           reward=float(res.reward), terminal=int(bool(res.is_terminal)), legal=int(bool(legal)),
                           next_cell=state.position.i * m.n_cols + state.position.j,
                           next_rocks=sum((1 << k) for k in range(K) if state.rock_states[k]),
                           draws=stream.i)
                if not res.is_terminal or not legal:
                    solver.update(res)
                rec["draws_after_update"] = stream.i
                rec["disable_tree"] = bool(solver.disable_tree)
                trace.append(rec)
                if res.is_terminal or not legal:
                    break
        ref["POMCP"].UCB_N, ref["POMCP"].UCB_n = old_N, old_n     # reads the class attributes, pomcp.py:61)
    return dict(actual=actual, bag=bag, trace=trace)


def run_oracle_episode(t, actual, bag, n_sims, n_moves, stream_seed, max_depth=100, max_particle_count=2000,
                       ucb_c=3.0, discount=0.95):
    om = oracle_model(t)
    o = oracle.RefPOMCP(om, actual, 0, bag, max_depth=max_depth, max_particle_count=max_particle_count,
                        ucb_c=ucb_c, discount=discount, seed=stream_seed, stream_id=0)
    trace = []
    p = o.sample_root_particle()
    sampled = 0
    for mv in range(n_moves):
        action = o.search(n_sims)
        s = o.root_stats()
        p, so = o.env_step(p, action)
        cell, rocks = o.particle(p)
        rec = dict(action=action, stats=dict(n=s["n"].tolist(), total=s["total"].tolist(), mean=s["mean"].tolist(),
                                             legal=s["legal"], N=s["N"]),
                   obs_good=so.obs_good, reward=so.reward, terminal=so.terminal, legal=so.legal,
                   next_cell=cell, next_rocks=rocks, draws=o.draws())
        if not so.terminal or not so.legal:
            if action == 4:
                sampled |= 1 << int(t["cell"][cell])       # rock_model.py:267-272
            o.update(action, so.obs_good, sampled)
        rec["draws_after_update"] = o.draws()
        rec["disable_tree"] = o.disabled()
        trace.append(rec)
        if so.terminal or not so.legal:
            break
    return trace, o.counters()


def compare_traces(a, b):
    if len(a) != len(b):
        return "length %d vs %d" % (len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        for k in x:
            if x[k] != y[k]:
                return "move %d key %s: %r vs %r" % (i, k, x[k], y[k])
    return None


EPISODES = [  # (map, numpy seed, n_sims, moves, stream seed, overrides)
    ("map-7-8.txt", 123, 500, 8, 1001, {}),
    ("map-7-8.txt", 124, 200, 30, 1002, {}),
    ("map-11-11.txt", 2024, 300, 5, 1003, {}),
    ("map-15-15.txt", 7, 400, 5, 1004, {}),
    ("map-7-8.txt", 9, 150, 12, 1005, dict(ucb_coefficient=5.0, max_depth=20, discount=0.9,
                                            n_start_states=300, max_particle_count=400)),
]


def main(write=True):
    ref = ref_env.import_reference()
    ref["model"].pp = lambda *_a, **_k: None
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(20261019)
    ok = True
    for mp in MAPS:
        m = ref_env.make_rock_model(ref, mp)
        t = tables_from_reference(ref, m)
        om = oracle_model(t)
        lg = np.array([om.legal_mask(c) if t["cell"][c] != -3 else 0 for c in range(t["rows"] * t["cols"])],
                      dtype=np.uint32)
        legal_bad = int((lg != t["legal"]).sum())
        binom_bad = check_binomial_semantics(t)
        mism, vec = check_generate_step(ref, m, t, om, rng)
        print("%s: legal mismatches %d, binomial-rule mismatches %d, generate_step mismatches %d / %d"
              % (mp, legal_bad, binom_bad, mism, len(vec)))
        ok &= legal_bad == 0 and binom_bad == 0 and mism == 0
        if write:
            tag = mp.replace(".txt", "")
            v = np.array([[float(x) for x in row] for row in vec])
            np.savez_compressed(os.path.join(GOLDEN, "rocksample_%s.npz" % tag), rows=t["rows"], cols=t["cols"],
                                cell=t["cell"], start_cell=t["start_cell"], thr53=t["thr53"], prob=t["prob"],
                                legal=t["legal"], rewards=t["rewards"],
                                step_in=np.array([r[:6] for r in vec], dtype=np.uint64),
                                step_out=np.array([r[6:] for r in vec], dtype=np.float64))
            del v
    eps = []
    for (mp, seed, n_sims, moves, sseed, kw) in EPISODES:
        m = ref_env.make_rock_model(ref, mp)
        t = tables_from_reference(ref, m)
        r = run_reference_episode(ref, mp, seed, n_sims, moves, sseed, **kw)
        tr, counters = run_oracle_episode(t, r["actual"], r["bag"], n_sims, moves, sseed,
                                          max_depth=kw.get("max_depth", 100),
                                          max_particle_count=kw.get("max_particle_count", 2000),
                                          ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95))
        diff = compare_traces(r["trace"], tr)
        print("%s seed %d n_sims %d: %d moves, %d draws -> %s  %s" % (
            mp, seed, n_sims, len(tr), tr[-1]["draws_after_update"], "IDENTICAL" if diff is None else "DIFF " + diff,
            counters))
        ok &= diff is None
        eps.append(dict(map=mp, seed=seed, n_sims=n_sims, moves=moves, stream_seed=sseed, overrides=kw,
                        actual=r["actual"], bag=r["bag"], trace=r["trace"]))
    if write:
        with open(os.path.join(GOLDEN, "pomcp_ref_episodes.json"), "w") as f:
            json.dump(dict(generator="oracle/refharness/validate.py", note="executed reference, injected draws",
                           episodes=eps), f)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
