"""Action agreement with the reference algorithm where its own choice is decisive (north star: "the chosen action
matches on fixed seeds where reference Q-gaps exceed Monte-Carlo noise").

The reference side is the oracle's "ref" mode (the reference's sequential POMCP with all its quirks, pinned bit for
bit to the executed reference).  Situations (belief, world, position) are harvested along reference-algorithm
episodes on RockSample(7,8) at the reference's budget (n_sims = 500).  A situation is *decisive* when 16
independent reference searches (different random streams) all pick the same action; there the CUDA planner, at the
same budget and on 4 seeds, must agree: >= 85 % of the searches, and the majority vote in >= 90 % of the situations
(the planner is the canonical variant, DESIGN.md 3.1, so its own Q estimates differ slightly from the reference's)."""
import numpy as np
import pytest

import oracle
from helpers import golden_tables, make_planner, oracle_model

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _harvest(om, t, n_episodes=12, n_sims=500):
    K = om.n_rocks
    out = []
    for ep in range(n_episodes):
        rng = np.random.default_rng(4000 + ep)
        actual = int(rng.integers(0, 1 << K))
        bag_rocks = rng.integers(0, 1 << K, size=2000).astype(np.uint32)
        ref = oracle.RefPOMCP(om, actual, 0, bag_rocks, seed=4000 + ep)
        cell, rocks, sampled = om.start_cell, int(bag_rocks[0]), 0
        bag = (bag_rocks << np.uint32(8)) | np.uint32(cell)
        for mv in range(25):
            out.append(dict(actual=actual, sampled=sampled, cell=cell, bag=bag.copy()))
            a = ref.search(n_sims)
            o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
            cell, rocks = o.next_cell, o.next_rocks
            if o.terminal or not o.legal:
                break
            if a == 4:
                sampled |= 1 << int(t["cell"][cell])
            if ref.update(a, o.obs_good, sampled) != 0:
                break
            # the belief the reference holds now = the particles of its new root: regenerate an equivalent bag
            g = oracle.GsPOMCP(om, actual, sampled, bag, root_cell=out[-1]["cell"], seed=ep * 100 + mv)
            g.set_world(actual, 0 if a != 4 else sampled)
            if g.update(a, o.obs_good, sampled, move=mv)["status"] == 2:
                break
            bag = g.bag()
    return out


def test_decisive_reference_actions_are_matched():
    t = golden_tables("map-7-8")
    om = oracle_model(t)
    situations = _harvest(om, t)
    decisive, pairs, pair_agree, majority, checked = 0, 0, 0, 0, []
    pl = None
    for si, s in enumerate(situations):
        picks = []
        for r in range(16):
            picks.append(_ref_search_at(om, t, s, 9000 + 16 * si + r))
        if len(set(picks)) != 1:
            continue
        decisive += 1
        a_ref = picks[0]
        got = []
        for seed in range(4):
            pl = make_planner(t, seed=700 + seed, gen_w0=8, gen_growth=2, gen_wmax=64, node_capacity=1 << 14)
            pl.set_world(s["actual"], s["sampled"])
            pl.set_root_particles(s["bag"], s["cell"])
            pl.search(500, move=si)
            st = pl.root_stats()
            mean = np.where(st["n"] > 0, st["qfix"] / 2.0 ** 24 / np.maximum(st["n"], 1), 0.0)
            legal = np.array([(st["legal"] >> a) & 1 for a in range(om.n_actions)], dtype=bool)
            mean[~legal] = -np.inf
            got.append(int(np.argmax(mean)))
            pl.close()
        pairs += 4
        pair_agree += sum(1 for g in got if g == a_ref)
        majority += sum(1 for g in got if g == a_ref) >= 3
        checked.append((si, a_ref, got))
    print("situations %d, decisive for the reference %d; CUDA planner: %d/%d searches agree, majority agrees in %d"
          % (len(situations), decisive, pair_agree, pairs, majority))
    print([c for c in checked if c[2].count(c[1]) < 4])
    assert decisive >= 15, "too few decisive situations to say anything"
    assert pair_agree >= 0.85 * pairs and majority >= 0.9 * decisive, checked


def _ref_search_at(om, t, s, seed):
    """One reference-algorithm search (n_sims = 500) from the situation's belief at the situation's cell."""
    m = oracle.Model(int(t["rows"]), int(t["cols"]), t["cell"], s["cell"], t["thr53"], t["rewards"])
    ref = oracle.RefPOMCP(m, s["actual"], s["sampled"], (s["bag"] >> np.uint32(8)).astype(np.uint32), seed=seed)
    return ref.search(500)


def test_decisive_sequential_actions_are_matched_on_rs15_at_10k_sims():
    """The same question on RockSample(15,15) at 10^4 simulations per move.  The reference's own
    algorithm is no yardstick there -- with its quirks its return collapses at this budget (0.1 +- 0.04 over 160
    episodes, profiles/r02_quality.md) -- so the sequential side is the canonical sequential POMCP (oracle "gs" with
    generation width 1).  Situations are harvested along its episodes; one is decisive when 6 independent sequential
    searches pick the same action; there the batched CUDA planner on the solver's own schedule for this budget
    (auto_schedule: generations of 39 ... 1250 simulations) must agree.  (Four wide generations, 312 x4 <= 5000, agree in
    only 82 % of the searches; sequential searches with other seeds in 92 %; this schedule in 97 %.)"""
    from pomdpy_b200.solvers.pomcp import auto_schedule
    w0, growth, wmax = auto_schedule(10000)
    t = golden_tables("map-15-15")
    om = oracle_model(t)
    K, n_sims = om.n_rocks, 10000
    situations = []
    for ep in range(3):
        rng = np.random.default_rng(7000 + ep)
        actual = int(rng.integers(0, 1 << K))
        bag = (rng.integers(0, 1 << K, size=2000).astype(np.uint32) << np.uint32(8)) | np.uint32(om.start_cell)
        g = oracle.GsPOMCP(om, actual, 0, bag, seed=7000 + ep)
        cell, rocks, sampled = om.start_cell, int(bag[0] >> 8), 0
        for mv in range(10):
            situations.append(dict(actual=actual, sampled=sampled, cell=cell, bag=g.bag()))
            g.search(n_sims, move=mv, w0=1, growth=1, wmax=1)
            a = g.greedy(move=mv)
            o = om.step(cell, rocks, a, int(rng.integers(0, 1 << 53)), actual, sampled)
            cell, rocks = o.next_cell, o.next_rocks
            if o.terminal or not o.legal:
                break
            if a == 4:
                sampled |= 1 << int(t["cell"][cell])
            if g.update(a, o.obs_good, sampled, move=mv)["status"] == 2:
                break
    decisive, pairs, agree, majority, misses = 0, 0, 0, 0, []
    for si, s in enumerate(situations):
        picks = []
        for r in range(6):
            g = oracle.GsPOMCP(om, s["actual"], s["sampled"], s["bag"], root_cell=s["cell"], seed=555 + 17 * si + r)
            g.search(n_sims, move=si, w0=1, growth=1, wmax=1)
            picks.append(g.greedy(move=si))
        if len(set(picks)) != 1:
            continue
        decisive += 1
        got = []
        for seed in range(4):
            pl = make_planner(t, seed=900 + seed, gen_w0=w0, gen_growth=growth, gen_wmax=wmax, node_capacity=1 << 16)
            pl.set_world(s["actual"], s["sampled"])
            pl.set_root_particles(s["bag"], s["cell"])
            pl.search(n_sims, move=si)
            st = pl.root_stats()
            mean = np.where(st["n"] > 0, st["qfix"] / 2.0 ** 24 / np.maximum(st["n"], 1), 0.0)
            legal = np.array([(st["legal"] >> a) & 1 for a in range(om.n_actions)], dtype=bool)
            mean[~legal] = -np.inf
            got.append(int(np.argmax(mean)))
            pl.close()
        pairs += 4
        agree += sum(1 for x in got if x == picks[0])
        majority += sum(1 for x in got if x == picks[0]) >= 3
        if got.count(picks[0]) < 4:
            misses.append((si, picks[0], got))
    print("RS(15,15) @ 1e4: situations %d, decisive for the sequential planner %d; CUDA planner: %d/%d searches agree, "
          "majority agrees in %d; misses %s" % (len(situations), decisive, agree, pairs, majority, misses))
    assert decisive >= 8, "too few decisive situations to say anything"
    assert agree >= 0.85 * pairs and majority >= 0.9 * decisive, misses
