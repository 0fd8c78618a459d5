"""GPU parity of the tabular-model planner: the CUDA kernels, through the C ABI, against the
CPU oracle -- bit-exact root / inner-node statistics and beliefs -- on Tiger and on random POMDPs that exercise both
the shared-memory staging of depth-1 edges and the global-atomic path (|A|*|Z| too large to stage)."""
import numpy as np
import pytest

import oracle
from pomdpy_b200 import native
from pomdpy_b200.tabular import random_pomdp, tiger

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

MODELS = {
    "tiger": lambda: tiger(),
    "rand-32-6-8": lambda: random_pomdp(32, 6, 8, seed=1),          # depth-1 edges staged in shared memory
    "rand-64-16-32": lambda: random_pomdp(64, 16, 32, seed=2),      # (1 + 32*16)*16 edges: not staged
    "rand-300-32-3": lambda: random_pomdp(300, 32, 3, seed=3),      # all 32 action lanes
}


def _pair(name, seed=0, n_bag=2000, **kw):
    m = MODELS[name]()
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(n_bag, np.random.RandomState(seed))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=kw.get("max_depth", 100), dcap=kw.get("tree_depth_cap", 64),
                       ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95), seed=kw.get("seed", 0))
    pl = native.Planner(**kw)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    return m, om, g, pl


def _same(a, b):
    assert a["legal"] == b["legal"]
    assert np.array_equal(a["n"], b["n"]), (a["n"], b["n"])
    assert np.array_equal(a["qfix"], b["qfix"]), (a["qfix"], b["qfix"])


@pytest.mark.parametrize("name", list(MODELS))
def test_tabular_generate_step_matches_oracle(name):
    m = MODELS[name]()
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl = native.Planner()
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    rng = np.random.RandomState(5)
    n = 4000
    s = rng.randint(0, m.num_states, size=n).astype(np.uint32)
    a = rng.randint(0, m.num_actions, size=n).astype(np.uint8)
    w = rng.randint(0, 1 << 32, size=(2, n), dtype=np.uint64).astype(np.uint32)
    w[0, :8] = [0, 0xFFFFFFFF, 0xFFFFFFFE, 1 << 31, (1 << 31) - 1, 1, 2, 3]
    w[1, 8:16] = [0, 0xFFFFFFFF, 0xFFFFFFFE, 1 << 31, (1 << 31) - 1, 1, 2, 3]
    ns, ob, rw, tm = pl.tabular_steps(s, a, w[0], w[1])
    for i in range(n):
        assert (int(ns[i]), int(ob[i]), float(rw[i]), int(tm[i])) == om.step(int(s[i]), int(a[i]), int(w[0, i]), int(w[1, i]))
    pl.close()


CASES = [
    # model, n_sims, (w0, growth, wmax), planner kwargs
    ("tiger", 300, (1, 1, 1), {}),
    ("tiger", 20000, (32, 2, 2048), dict(ucb_coefficient=30.0)),
    ("tiger", 50000, (64, 4, 16384), dict(ucb_coefficient=30.0, seed=7)),
    ("rand-32-6-8", 500, (1, 1, 1), dict(max_depth=30)),
    ("rand-32-6-8", 30000, (32, 2, 4096), dict(max_depth=30, ucb_coefficient=10.0)),
    ("rand-64-16-32", 30000, (32, 2, 4096), dict(max_depth=20, ucb_coefficient=10.0)),
    ("rand-64-16-32", 3000, (8, 2, 64), dict(max_depth=20, tree_depth_cap=4)),
    ("rand-300-32-3", 40000, (64, 4, 8192), dict(max_depth=25, ucb_coefficient=8.0, discount=0.9)),
]


@pytest.mark.parametrize("name,n_sims,sched,kw", CASES)
def test_tabular_search_bit_exact(name, n_sims, sched, kw):
    w0, growth, wmax = sched
    m, om, g, pl = _pair(name, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 17, **kw)
    g.search(n_sims, w0=w0, growth=growth, wmax=wmax)
    pl.search(n_sims)
    a, b = g.root_stats(), pl.root_stats()
    _same(a, b)
    assert int(b["n"].sum()) == n_sims
    gc, pc = g.counters(), pl.counters()
    for k in ("levels", "rollout_steps", "created", "nodes", "generations"):
        assert gc[k] == pc[k], (k, gc[k], pc[k])
    # inner nodes: every (action, observation) child of the root and one grandchild layer
    checked = 0
    for act in range(m.num_actions):
        for z in range(m.num_observations):
            x, y = g.node_stats([(act, z)]), pl.node_stats([(act, z)])
            assert (x is None) == (y is None)
            if x is None:
                continue
            _same(x, y)
            checked += 1
            for a2 in range(min(m.num_actions, 3)):
                x2, y2 = g.node_stats([(act, z), (a2, 0)]), pl.node_stats([(act, z), (a2, 0)])
                assert (x2 is None) == (y2 is None)
                if x2 is not None:
                    _same(x2, y2)
    assert checked > 0 or n_sims < 10
    pl.close()


@pytest.mark.parametrize("name,kw", [("tiger", dict(ucb_coefficient=30.0)), ("rand-32-6-8", dict(max_depth=30)),
                                     ("rand-64-16-32", dict(max_depth=20))])
def test_tabular_search_update_loop_bit_exact(name, kw):
    """Several real steps: search, greedy action, belief update with a sampled observation, re-root."""
    m, om, g, pl = _pair(name, gen_w0=32, gen_growth=2, gen_wmax=2048, node_capacity=1 << 17, max_particle_count=1500,
                         **kw)
    rng = np.random.RandomState(11)
    state = int(rng.randint(m.num_states))
    for move in range(5):
        g.search(6000, move=move, w0=32, growth=2, wmax=2048)
        pl.search(6000, move=move)
        a, b = g.root_stats(), pl.root_stats()
        _same(a, b)
        act = g.greedy(move, a)
        if name == "tiger":
            act = 0                                              # keep listening: opening ends the episode
        ns, z, _, _ = om.step(state, act, int(rng.randint(1 << 32)), int(rng.randint(1 << 32)))
        state = ns
        u = g.update(act, z, 0, move=move, need=1500)
        v = pl.advance(act, z, 0, move=move)
        assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"])
        if u["status"] == 2:
            break
        bag, _ = pl.root_particles()
        assert np.array_equal(bag, g.bag())
        _same(g.root_stats(), pl.root_stats())
    pl.close()


def test_tabular_argument_errors():
    m = tiger()
    pl = native.Planner()
    bad = m.Tc.copy()
    bad[0, 0, -1] = 5
    with pytest.raises(native.PomcpError):
        pl.set_model_tabular(bad, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    with pytest.raises(native.PomcpError):
        pl.set_root_particles(np.array([0, 1, 2], dtype=np.uint32))      # state 2 does not exist
    pl.set_root_particles(np.array([0, 1], dtype=np.uint32))
    with pytest.raises(native.PomcpError):
        pl.advance(0, 2)                                                 # observation 2 does not exist
    with pytest.raises(native.PomcpError):
        pl.set_preferred_actions(True, np.zeros(4))
    pl.close()


def test_tiger_agent_episodes():
    """The reference's Agent loop (agent.py:150-213) on Tiger through the CUDA solver: listens first, then opens; the
    mean discounted return beats both opening at once (-5) and never opening."""
    from pomdpy_b200 import Agent
    from pomdpy_b200.solvers import POMCP
    args = dict(max_steps=30, n_epochs=60, n_sims=4000, seed=5, ucb_coefficient=30.0, discount=0.95, max_depth=30,
                max_particle_count=1000, n_start_states=1000, epsilon_start=0.5, epsilon_minimum=0.1,
                epsilon_decay=0.95, action_selection_timeout=60, timeout=3600, preferred_actions=False,
                save=False, test=10, env="Tiger", solver="POMCP")
    np.random.seed(5)
    agent = Agent(tiger(args), POMCP)
    agent.discounted_return()
    er = agent.experiment_results
    mean = er.discounted_return.mean
    print("tiger: discounted return %.2f +- %.2f over %d epochs" % (mean, er.discounted_return.std_err(), args["n_epochs"]))
    assert mean > 0.0


def test_pomcp_agrees_with_value_iteration_on_tiger():
    """The two device solvers on the same matrices (continuing Tiger, horizon 6): the POMCP kernel's greedy action at
    clear-cut beliefs (0.5, 0.03, 0.97) equals the action of the value-iteration kernel's alpha vectors, and its root value at the
    uniform belief stays below the true horizon-6 Bellman value (no planner whose actions depend on the history alone
    can beat that; rollouts beyond the tree only cost) and above the value of listening for ever (-5.3).  (The root value
    is an average that includes the exploratory simulations: over seeds 1..8 the oracle gives 2.2 ... 6.6 at this budget,
    so the reference-formula value 3.85 -- value_iteration.py:51-67 discounts the immediate reward too, SURVEY A.8 -- is
    inside that band, not a bound.)"""
    from pomdpy_b200 import vi
    from pomdpy_b200.tabular import TabularModel
    from pomdpy_b200.tiger.tiger_model import TigerModel as TM
    T, O, R = TM.get_transition_matrix(), TM.get_observation_matrix(), TM.get_reward_matrix()
    H = 6
    alpha, acts = vi.solve_reference(T, O, R, 0.95, H)
    m = TabularModel({}, T, O, R, [0.5, 0.5], name="TigerContinuing")
    pl = native.Planner(max_depth=H, ucb_coefficient=30.0, seed=1, gen_w0=64, gen_growth=2, gen_wmax=4096,
                        node_capacity=1 << 18)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, 0, None)
    for p0 in (0.5, 0.03, 0.97, 0.3, 0.7):
        b = np.array([p0, 1.0 - p0])
        best = int(np.argmax(alpha @ b))
        n0 = int(round(2000 * p0))
        pl.set_root_particles(np.array([0] * n0 + [1] * (2000 - n0), dtype=np.uint32))
        pl.search(100000)
        st = pl.root_stats()
        q = st["qfix"] / np.maximum(st["n"], 1) / native.QFIX_SCALE
        if p0 in (0.5, 0.03, 0.97):
            assert int(np.argmax(q)) == int(acts[best]), (p0, q, acts[best])
        else:
            # at 0.3 / 0.7 listening (7.5 by value iteration) and opening the likelier door (4.4) are close for a search
            # whose value of listening depends on how well the subtree below it is explored: over seeds the planner
            # takes either (the oracle does the same, with the 10-round stream too); it must not open the other door
            per_action = [max([float(v) for v, a in zip(alpha @ b, acts) if a == k] or [-1e30]) for k in range(3)]
            assert int(np.argmax(q)) != int(np.argmin(per_action)), (p0, q, per_action)
        if p0 == 0.5:
            exact = float((alpha @ b).max())
            Tm, Om, Rm = np.asarray(T, float), np.asarray(O, float), np.asarray(R, float)

            def bellman(bel, h):
                if h == 0:
                    return 0.0
                best = -1e30
                for a in range(Tm.shape[0]):
                    v, nxt = float(bel @ Rm[a]), bel @ Tm[a]
                    for z in range(Om.shape[2]):
                        pz = float(nxt @ Om[a][:, z])
                        if pz > 1e-12:
                            v += 0.95 * pz * bellman(nxt * Om[a][:, z] / pz, h - 1)
                    best = max(best, v)
                return best
            listen_for_ever = -sum(0.95 ** t for t in range(H))
            assert listen_for_ever < q.max() <= bellman(b, H) + 0.05, (q.max(), exact, bellman(b, H))
    pl.close()


def _custom_pair(T, O, R, b0, term_actions=(), term_states=(), n_bag=500, **kw):
    from pomdpy_b200.tabular import TabularModel
    m = TabularModel({}, T, O, R, b0, terminal_actions=term_actions, terminal_states=term_states)
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(n_bag, np.random.RandomState(0))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=kw.get("max_depth", 100), dcap=kw.get("tree_depth_cap", 64),
                       ucb_c=kw.get("ucb_coefficient", 3.0), discount=kw.get("discount", 0.95), seed=kw.get("seed", 0))
    pl = native.Planner(**kw)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    return m, om, g, pl


def test_tabular_edge_models_bit_exact():
    """Degenerate shapes: absorbing terminal states with deterministic (zero-probability-laden) transitions, a single
    observation, a single action, a single state."""
    rng = np.random.RandomState(3)
    # a 6-state chain: action 0 moves right, action 1 stays; state 5 is absorbing and terminal; reward on arrival
    S = 6
    T = np.zeros((2, S, S))
    for s in range(S):
        T[0, s, min(s + 1, S - 1)] = 1.0
        T[1, s, s] = 0.7
        T[1, s, max(s - 1, 0)] += 0.3
    O = np.zeros((2, S, 3))
    O[:, :, 0], O[:, :, 1], O[:, :, 2] = 0.5, 0.5, 0.0                           # observation 2 never occurs
    R = np.zeros((2, S))
    R[0, 4] = 10.0
    R[1, :] = -0.5
    cases = [dict(T=T, O=O, R=R, b0=np.array([0.6, 0.4, 0, 0, 0, 0.0]), term_states=(5,)),
             dict(T=rng.dirichlet(np.ones(9), size=(4, 9)), O=np.ones((4, 9, 1)), R=rng.randn(4, 9), b0=np.ones(9) / 9),
             dict(T=rng.dirichlet(np.ones(5), size=(1, 5)), O=rng.dirichlet(np.ones(4), size=(1, 5)), R=rng.randn(1, 5),
                  b0=np.ones(5) / 5),
             dict(T=np.ones((3, 1, 1)), O=rng.dirichlet(np.ones(2), size=(3, 1)), R=np.array([[1.0], [2.0], [-1.0]]),
                  b0=np.ones(1))]
    for c in cases:
        m, om, g, pl = _custom_pair(c["T"], c["O"], c["R"], c["b0"], term_states=c.get("term_states", ()),
                                    gen_w0=16, gen_growth=2, gen_wmax=1024, node_capacity=1 << 16, max_depth=40)
        for move in range(3):
            g.search(5000, move=move, w0=16, growth=2, wmax=1024)
            pl.search(5000, move=move)
            a, b = g.root_stats(), pl.root_stats()
            _same(a, b)
            act = g.greedy(move, a)
            u, v = g.update(act, 0, 0, move=move, need=2000), pl.advance(act, 0, 0, move=move)
            assert (u["status"], u["tries"], u["accepted"]) == (v["status"], v["tries"], v["accepted"])
            if u["status"] == 2:
                break
            assert np.array_equal(pl.root_particles()[0], g.bag())
        pl.close()
    # an observation that cannot occur empties the belief: status 2 on both sides, no exception
    c = cases[0]
    m, om, g, pl = _custom_pair(c["T"], c["O"], c["R"], c["b0"], term_states=(5,), gen_w0=16, gen_growth=2,
                                gen_wmax=1024, node_capacity=1 << 16, max_depth=40)
    g.search(1000, w0=16, growth=2, wmax=1024)
    pl.search(1000)
    u, v = g.update(1, 2, 0, need=2000, max_tries=20000), pl.advance(1, 2, 0, max_tries=20000)
    assert u["status"] == 2 and v["status"] == 2 and v["accepted"] == 0 and v["tries"] == u["tries"] == 20000
    pl.close()


def test_tabular_large_state_space_and_capacity_reset():
    """|S| = 2048 (binary search over 8 KB rows in global memory) and a search that does not fit the arena (the tree
    is dropped, the belief kept: status 1, as for RockSample)."""
    m = random_pomdp(2048, 2, 4, seed=8, sparsity=0.99)
    om = oracle.TabularModel(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    bag = m.sample_init_states_packed(1000, np.random.RandomState(1))
    g = oracle.GsPOMCP(om, 0, 0, bag, max_depth=30, seed=4)
    pl = native.Planner(max_depth=30, seed=4, gen_w0=32, gen_growth=2, gen_wmax=2048, node_capacity=30000)
    pl.set_model_tabular(m.Tc, m.Oc, m.R, m.term_action_mask, m.term_state)
    pl.set_root_particles(bag)
    g.search(20000, w0=32, growth=2, wmax=2048)
    assert pl.search(20000) == 0
    _same(g.root_stats(), pl.root_stats())
    assert pl.search(20000, move=1) == 1                      # 20000 + 20000 nodes could exceed 30000: fresh tree
    g2 = oracle.GsPOMCP(om, 0, 0, bag, max_depth=30, seed=4)
    g2.search(20000, move=1, w0=32, growth=2, wmax=2048)
    _same(g2.root_stats(), pl.root_stats())
    pl.close()


def test_tiger_batched_epochs():
    """Episode-level batching (--concurrent_epochs) on a tabular model: same level of return as the serial Agent."""
    from pomdpy_b200.batch import BatchedAgent
    args = dict(max_steps=30, n_epochs=96, n_sims=4000, seed=5, ucb_coefficient=30.0, discount=0.95, max_depth=30,
                max_particle_count=1000, n_start_states=1000, preferred_actions=False)
    np.random.seed(6)
    agent = BatchedAgent(tiger(args), lanes=16)
    mean, err = agent.run(args["n_epochs"])
    print("tiger batched: %.2f +- %.2f over %d epochs, %.1f ms/epoch" % (mean, err, args["n_epochs"],
                                                                      1e3 * agent.wall_seconds / args["n_epochs"]))
    assert agent.discounted_return.count == 96 and mean > 0.0
