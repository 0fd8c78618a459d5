"""POMCP on the B200: the drop-in for the reference's solver path.

Same surface as the reference's POMCP / BeliefTreeSolver (pomdpy/solvers/pomcp.py:14-137,
pomdpy/solvers/belief_tree_solver.py:14-212) as far as Agent.run_pomcp touches it (agent.py:150-213):

    POMCP.reset(agent)                        per-epoch factory; hyper-parameters come from agent.model.<flag>
    solver.belief_tree_index.sample_particle()
    solver.select_eps_greedy_action(eps, start_time) -> DiscreteAction
    solver.update(step_result)                belief update + re-root; sets solver.disable_tree on failure
    solver.history                            HistorySequence

What happens inside is different: select_eps_greedy_action runs ``n_sims`` simulations as ONE cooperative CUDA
kernel over a device-resident structure-of-arrays belief tree (pomcp_search), update() runs the batched
rejection-sampling belief update (pomcp_advance).  With torch.distributed initialised the search is
root-parallel: every rank grows its own tree with its share of the budget and the root statistics are summed
with one all-reduce before the greedy choice (solvers/pomcp.py:78).  There is no CPU fallback."""
import os
import time

import numpy as np

from .. import adapters
from .. import dist as dist_mod
from .. import native, philox
from ..rock_sample.history_data import PositionAndRockData, RockData
from ..rock_sample.types import GridPosition
from ..util import console
from .belief_view import BeliefNodeView
from .solver import Solver

module = "pomcp"


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def greedy_action(n, qfix, legal_mask, seed, move):
    """solvers/pomcp.py:78 ucb_action(greedy=True): arg-max of the mean over legal actions (an unvisited edge
    has mean 0, discrete_action_mapping.py:126), uniform tie-break."""
    best, cand = None, []
    for a in range(len(n)):
        if not (legal_mask >> a) & 1:
            continue
        q = 0.0 if n[a] == 0 else (float(qfix[a]) * (1.0 / native.QFIX_SCALE)) / float(n[a])
        if best is None or q > best:
            best, cand = q, [a]
        elif q == best:
            cand.append(a)
    return cand[philox.randbelow(seed, len(cand), 0, 0, move, philox.DOM_FINAL)]


def auto_schedule(n_sims):
    """Generation schedule (w0, growth, wmax) for a per-rank budget of n_sims per move.  A generation is
    clamp(growth * visits(root), w0, wmax) simulations wide (pomcp_kernels.cu), so w0 is the width against an empty root
    and wmax the width once the root has statistics.  Measured (profiles/r02_quality.md):
      n <= 4096    narrow waves, w0 = n/64, x2, <= n/4: at n_sims = 500 widths up to 64 match the sequential planner's
                   return, 128 starts to cost
      n <  50000   w0 = n/256, x2, <= n/8 (about a dozen generations): at 1e4 simulations this agrees with the
                   sequential planner's decisive choices in 97 % of the searches; four wide generations only in 82 %
      n >= 50000   the throughput schedule, w0 = n/8, x32, <= n: two generations from a fresh tree (n/8, then the rest),
                   ONE inside an episode whenever the re-rooted tree arrives with at least n/32 visits.  At 1e6
                   simulations per move on RockSample(15,15), 192 paired episodes each: 14.66 +- 0.42 at 0.77 ms per move
                   (growth 8: 14.64 at 0.96 ms; growth 128 ... 1e5: 14.54 ... 14.49, RockSample(11,11) loses 0.3 ... 0.8),
                   against 13.15 +- 0.35 at 1.32 ms for n/32, x4, <= n/2 (round 1's rule; 2 ... 62 generations per move
                   all return 12.8 ... 13.4); 100 % agreement with the sequential planner's decisive choices at 1e5"""
    n = int(n_sims)
    if n <= 4096:
        return max(1, min(1024, (n // 4 or 1) // 16)), 2, max(32, n // 4 or 1)
    if n < 50000:
        return max(1, n // 256), 2, max(64, n // 8)
    wm = min(1 << 20, n)
    return max(1, wm // 8), 32, wm


class POMCP(Solver):
    # planners are expensive to create (the tree arena); one is kept per (agent, configuration)
    _EPOCH_STRIDE = 1 << 12

    def __init__(self, agent):
        super(POMCP, self).__init__(agent)
        model = self.model
        # the model's device hooks: its own (this package's models) or extracted (an unmodified reference RockModel)
        self.dev = adapters.hooks_for(model)
        self.agent = agent
        self.action_pool = agent.action_pool
        self.observation_pool = agent.observation_pool
        self.history = agent.histories.create_sequence()
        self.disable_tree = False

        self.n_sims = int(getattr(model, "n_sims", 500))
        self.seed = int(getattr(model, "seed", 1993))
        self.timeout = float(getattr(model, "action_selection_timeout", 60))
        self.max_particle_count = int(getattr(model, "max_particle_count", 2000))
        self.preferred = bool(getattr(model, "preferred_actions", False))
        d = _dist()
        self.world = d.get_world_size() if d else 1
        self.rank = d.get_rank() if d else 0
        self.sims_per_rank, self.sim_offset = dist_mod.shard_budget(self.n_sims, self.rank, self.world)

        self.planner = self._acquire_planner(agent)
        self.epoch_index = agent._pomcp_epochs = getattr(agent, "_pomcp_epochs", -1) + 1
        self.move_in_epoch = 0

        spec = self.dev.device_spec()
        self.spec = spec
        # belief_tree_solver.py:36-38: n_start_states draws of sample_an_init_state (numpy generator)
        n_start = int(getattr(model, "n_start_states", 2000))
        if hasattr(model, "sample_init_states_packed"):
            bag = np.ascontiguousarray(model.sample_init_states_packed(n_start), dtype=np.uint32)
        else:
            bag = np.fromiter((self.dev.pack_state(model.sample_an_init_state()) for _ in range(n_start)),
                              dtype=np.uint32, count=n_start)
        self.tabular = spec.get("kind", "rocksample") == "tabular"
        if self.tabular:
            self.preferred = False                                            # a RockSample heuristic
        else:
            self.planner.set_world(*self.dev.world_masks())
            mode = (2 if getattr(model, "preferred_rollouts", False) else 1) if self.preferred else 0
            if mode != getattr(self.planner, "_preferred", 0):                # rock_position_history.py:52-55
                self.planner.set_preferred_actions(self.preferred, spec["prob"], rollouts=mode == 2)
                self.planner._preferred = mode
        self.planner.set_root_particles(bag, spec.get("start_cell", 0))
        if self.world > 1 and getattr(self.planner, "_fused_merge", None) is not None:
            # the in-kernel merge waits a bounded time (POMCP_MERGE_TIMEOUT_S, default 2 s) for the peers' root
            # statistics: line the ranks up once per epoch so that per-epoch setup skew cannot eat into it
            d.barrier()
        self._root_cell = spec.get("start_cell", 0)
        self._stats = None
        self.belief_tree_index = BeliefNodeView(self)
        self.sims_done = 0
        self.search_seconds = 0.0

    # ---- plumbing -------------------------------------------------------------------------------------
    def _acquire_planner(self, agent):
        model = self.model
        # Generation schedule (simulations in flight): small budgets keep the waves narrow (quality first: at
        # n_sims = 500 widths up to 64 match the sequential planner's return, 128 starts to cost); large budgets use
        # the throughput schedule of bench.py (first wave = 1/32 of the budget, x4 growth up to half of it: four
        # generations; return at 1e5 and 1e6 simulations per move does not depend on the first width,
        # profiles/r01_schedule_sweep.txt).
        auto = auto_schedule(self.sims_per_rank)
        w0 = int(getattr(model, "gen_w0", 0)) or auto[0]
        growth = int(getattr(model, "gen_growth", 0)) or auto[1]
        wmax = max(int(getattr(model, "gen_wmax", 0)) or auto[2], w0)
        max_steps = int(getattr(model, "max_steps", 200))
        cap = int(getattr(model, "node_capacity", 0)) or \
            int(min(1 << 24, max(1 << 14, self.sims_per_rank * min(max_steps, 8) + 64)))
        key = (int(model.max_depth), self.max_particle_count, float(model.ucb_coefficient), float(model.discount),
               self.seed, cap, w0, growth, wmax, id(model))
        cached = getattr(agent, "_pomcp_planner", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        if cached is not None:
            cached[1].close()
        device = 0
        try:
            import torch
            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except Exception:
            pass
        pl = native.Planner(max_depth=int(model.max_depth), tree_depth_cap=min(64, max(2, int(model.max_depth))),
                            max_particle_count=self.max_particle_count, ucb_coefficient=float(model.ucb_coefficient),
                            discount=float(model.discount), seed=self.seed, node_capacity=cap, gen_w0=w0,
                            gen_growth=growth, gen_wmax=wmax, device=device)
        spec = self.dev.device_spec()
        if spec.get("kind", "rocksample") == "tabular":
            pl.set_model_tabular(spec["Tc"], spec["Oc"], spec["R"], spec["term_action_mask"], spec["term_state"])
        else:
            pl.set_model_rocksample(spec["rows"], spec["cols"], spec["cell"], spec["start_cell"], spec["thr53"],
                                    spec["rewards"])
        if self.world > 1:                                   # root-parallel: fuse the merge into the search kernel
            d = _dist()
            try:
                import torch
                fused = torch.cuda.is_available() and d.get_backend() == "nccl" and \
                    os.environ.get("POMCP_FUSED_MERGE", "1") != "0"
            except Exception:
                fused = False
            pl._fused_merge = dist_mod.try_fused_merge(pl) if fused else None
        agent._pomcp_planner = (key, pl)
        return pl

    @property
    def move(self):
        """32-bit move id of the random streams: unique per (epoch, real step)."""
        return (self.epoch_index * self._EPOCH_STRIDE + self.move_in_epoch) & 0xFFFFFFFF

    @staticmethod
    def reset(agent):
        return POMCP(agent)

    def root_statistics(self):
        if self._stats is None:
            self._stats = self.planner.root_stats()
        return self._stats

    def root_data(self):
        if self.tabular:
            return None                                  # create_root_historical_data() of a tabular model
        cell = self._root_cell
        pos = GridPosition(cell // self.model.n_cols, cell % self.model.n_cols)
        rocks = [RockData() for _ in range(self.model.n_rocks)]
        stats = self.planner.root_rock_data() if self.preferred else None
        if stats is not None:
            for rd, ch, gd in zip(rocks, stats[0], stats[1]):
                rd.chance_good, rd.goodness_number = float(ch), int(gd)
        return PositionAndRockData(self.model, pos, rocks, self)

    def _merged_root_stats(self):
        """Root statistics summed over the ranks: one all-reduce of 2*32 int64 per move."""
        if self.world > 1 and getattr(self.planner, "_fused_merge", None) is not None:
            return self.planner.merged_root_stats()          # summed inside the search kernel over NVLink peer memory
        local = self.planner.root_stats()
        if self.world == 1:
            return local
        d = _dist()
        import torch
        if torch.cuda.is_available() and d.get_backend() == "nccl":
            n, q = dist_mod.merge_device(self.planner, len(local["n"]))
        else:
            n, q = dist_mod.merge_host(local["n"], local["qfix"])
        return dict(n=n, qfix=q, legal=local["legal"], sims_done=local["sims_done"])

    # ---- the solver API -------------------------------------------------------------------------------
    def select_eps_greedy_action(self, eps, start_time):
        """Run the Monte-Carlo tree search from the current belief and return the greedy action
        (solvers/pomcp.py:69-78; `eps` is ignored there as well)."""
        remaining = self.timeout - (time.time() - start_time)
        stream = None
        try:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        except Exception:
            pass
        t0 = time.perf_counter()
        self.planner.search(self.sims_per_rank, move=self.move, sim_offset=self.sim_offset,
                            timeout_s=max(remaining, 1e-3), stream=stream)
        st = self._merged_root_stats()
        self.search_seconds += time.perf_counter() - t0
        self.sims_done += int(st["sims_done"])
        self._stats = st
        a = greedy_action(st["n"], st["qfix"], st["legal"], self.seed, self.move)
        return self.action_pool.sample_an_action(a)

    def update(self, step_result, prune=True):
        """belief_tree_solver.py:154-212: fold the real step into the model, re-root at (action, observation),
        refill the belief by rejection sampling (on the device).  On failure: disable_tree, no exception."""
        self.model.update(step_result)                                           # :165
        _, sampled = self.dev.world_masks()
        obs = step_result.observation
        res = self.planner.advance(step_result.action.bin_number,
                                   int(obs.bin_number) if self.tabular else bool(obs.is_good), sampled, move=self.move)
        self._stats = None
        self.move_in_epoch += 1
        if res["status"] == 2:
            self._recover(step_result)
            return
        if res["status"] == 1:
            console(4, module, "observed child was never expanded: fresh root")
        _, self._root_cell = self.planner.root_particles(fetch=False)

    def _recover(self, step_result):
        """No particle of the old belief reproduces the real observation (belief_tree_solver.py:196-208).  As the
        reference does, first try a new particle set at the new position with uniformly drawn rock states that match
        the observation (generate_particles_uninformed, rock_model.py:252-261).  If that fails too the reference sets
        disable_tree and finishes the episode with rollout_search from the STALE node (belief_tree_solver.py:90-121:
        its actions are those of the old position); here disable_tree is set as well, but the old belief is pushed
        through the real action without conditioning on the observation and planted as a fresh root at the true
        position, so that select_eps_greedy_action keeps returning actions that are legal there."""
        m = self.model
        a = step_result.action.bin_number
        bag = None
        if not self.tabular and hasattr(m, "generate_particles_uninformed_packed"):
            bag = m.generate_particles_uninformed_packed(self._root_cell, a, step_result.observation,
                                                         int(getattr(m, "min_particle_count", 1000)))
        if bag is not None and len(bag):
            cell = int(bag[0]) & 0xFF
            self.planner.set_world(*self.dev.world_masks())
            self.planner.set_root_particles(np.ascontiguousarray(bag, dtype=np.uint32), cell)
            self._root_cell = cell
            console(2, module, "belief re-created from uninformed particles: fresh tree")
            return
        console(1, module, "Couldn't refill particles, must use random rollout to finish episode")
        self.disable_tree = True
        _, sampled = self.dev.world_masks()
        res = self.planner.advance(a, None, sampled, move=self.move - 1)
        if res["status"] != 2:
            _, self._root_cell = self.planner.root_particles(fetch=False)

    def prune(self, belief_node=None):
        """Sibling subtrees are simply left behind in the arena (belief_tree.py:87-109 frees them)."""
