"""Episode-level batching: many epochs of Agent.run_pomcp at once on one GPU.

At the reference's budget (n_sims = 500) a move is a 0.8 ms search that keeps ONE thread block busy; the epoch loop
of pomdpy/agent.py:132-213 is embarrassingly parallel across epochs.  `BatchedAgent` runs `lanes` epochs concurrently:
every lane owns a planner (its own tree arena and belief) and a CUDA stream; per round all lanes launch their search
asynchronously, then each lane reads its root statistics, takes the greedy action, steps its environment on the host
and updates its belief.  The searches of different lanes overlap on the device.

Semantics per epoch are those of Agent.run_pomcp; the worlds (true rock states, initial beliefs) are drawn from the
numpy generator in epoch order exactly as in the serial loop, the environment noise of concurrently running epochs
interleaves (a different, equally valid, stream)."""
import copy
import time

import numpy as np

from . import native, philox
from .solvers.pomcp import greedy_action
from .statistic import Statistic


class _Lane(object):
    def __init__(self, model, planner, stream):
        self.model, self.planner, self.stream = model, planner, stream
        self.active = False


class BatchedAgent(object):
    def __init__(self, model, lanes=32):
        self.model = model
        self.tabular = model.device_spec().get("kind", "rocksample") == "tabular"
        self.lanes = int(lanes)
        self.discounted_return = Statistic("discounted return")
        self.undiscounted_return = Statistic("undiscounted return")
        self.steps = Statistic("steps")
        self.total_sims = 0
        self.wall_seconds = 0.0

    def _make_lanes(self):
        import torch
        m = self.model
        spec = m.device_spec()
        n_sims = int(m.n_sims)
        wmax = max(32, n_sims // 4 or 1)
        lanes = []
        for _ in range(self.lanes):
            pl = native.Planner(max_depth=int(m.max_depth), tree_depth_cap=min(64, max(2, int(m.max_depth))),
                                max_particle_count=int(m.max_particle_count), ucb_coefficient=float(m.ucb_coefficient),
                                discount=float(m.discount), seed=int(getattr(m, "seed", 1993)),
                                node_capacity=max(1 << 12, n_sims * 8 + 64), gen_w0=max(1, min(1024, wmax // 16)),
                                gen_growth=2, gen_wmax=wmax)
            if self.tabular:
                pl.set_model_tabular(spec["Tc"], spec["Oc"], spec["R"], spec["term_action_mask"], spec["term_state"])
            else:
                pl.set_model_rocksample(spec["rows"], spec["cols"], spec["cell"], spec["start_cell"], spec["thr53"],
                                        spec["rewards"])
            if not self.tabular and getattr(m, "preferred_actions", False):
                pl.set_preferred_actions(True, spec["prob"], rollouts=bool(getattr(m, "preferred_rollouts", False)))
            lane_model = copy.copy(m)                      # shares the tables, owns the per-epoch world
            lanes.append(_Lane(lane_model, pl, torch.cuda.Stream()))
        return lanes, spec

    def _start_epoch(self, lane, epoch, spec):
        m = lane.model
        if not self.tabular:
            m.unique_rocks_sampled = []
        m.reset_for_epoch()                                # agent.py:135,143: the true world (numpy generator)
        bag = np.ascontiguousarray(m.sample_init_states_packed(int(getattr(m, "n_start_states", 2000))), dtype=np.uint32)
        if not self.tabular:
            lane.planner.set_world(*m.world_masks())
        lane.planner.set_root_particles(bag, spec.get("start_cell", 0))
        idx = philox.randbelow(int(getattr(m, "seed", 1993)), len(bag), 0, 0, epoch << 12, philox.DOM_AGENT)
        lane.state = m.unpack_state(bag[idx])              # agent.py:157
        lane.epoch, lane.step, lane.ret, lane.dret, lane.disc, lane.active = epoch, 0, 0.0, 0.0, 1.0, True

    def run(self, n_epochs):
        """Run n_epochs epochs; returns (mean discounted return, standard error)."""
        import torch
        m = self.model
        lanes, spec = self._make_lanes()
        n_sims, max_steps, seed = int(m.n_sims), int(m.max_steps), int(getattr(m, "seed", 1993))
        next_epoch, finished = 0, 0
        t0 = time.time()
        for lane in lanes:
            if next_epoch < n_epochs:
                self._start_epoch(lane, next_epoch, spec)
                next_epoch += 1
        while finished < n_epochs:
            live = [ln for ln in lanes if ln.active]
            for ln in live:                                # all searches in flight together
                ln.move = ((ln.epoch << 12) + ln.step) & 0xFFFFFFFF
                ln.planner.search(n_sims, move=ln.move, stream=ln.stream.cuda_stream)
            for ln in live:
                st = ln.planner.root_stats()               # waits for this lane's stream only
                self.total_sims += int(st["sims_done"])
                a = greedy_action(st["n"], st["qfix"], st["legal"], seed, ln.move)
                res, legal = ln.model.generate_step(ln.state, a)          # agent.py:174
                ln.ret += res.reward
                ln.dret += ln.disc * res.reward
                ln.disc *= m.discount
                ln.state = res.next_state
                ln.step += 1
                done = res.is_terminal or not legal or ln.step >= max_steps
                if not res.is_terminal or not legal:       # agent.py:186
                    ln.model.update(res)
                    _, sampled = ln.model.world_masks()
                    obs = int(res.observation.bin_number) if self.tabular else bool(res.observation.is_good)
                    if ln.planner.advance(a, obs, sampled, move=ln.move)["status"] == 2:
                        # no particle reproduces the observation: as the serial solver does (POMCP._recover), push the
                        # old belief through the action unconditioned and go on from a fresh root at the true position
                        if ln.planner.advance(a, None, sampled, move=ln.move)["status"] == 2:
                            done = True
                if done:
                    self.discounted_return.add(ln.dret)
                    self.undiscounted_return.add(ln.ret)
                    self.steps.add(ln.step)
                    finished += 1
                    ln.active = False
                    if next_epoch < n_epochs:
                        self._start_epoch(ln, next_epoch, spec)
                        next_epoch += 1
        torch.cuda.synchronize()
        self.wall_seconds = time.time() - t0
        for ln in lanes:
            ln.planner.close()
        return self.discounted_return.mean, self.discounted_return.std_err()
