"""Read-only views of the device-resident belief tree with the attribute names of the reference's
BeliefNode (pomdpy/pomdp/belief_node.py:5-124).  The tree itself lives in the CUDA planner's arena."""
from ..discrete import ActionMappingView
from ..native import QFIX_SCALE
from .. import philox


class BeliefNodeView(object):
    """What Agent and user code touch on ``solver.belief_tree_index``: sample_particle(), state_particles,
    action_map (a snapshot of the root edge statistics), data (position + legal-action generator)."""

    def __init__(self, solver):
        self.solver = solver
        self._draws = 0

    @property
    def depth(self):
        return self.solver.move_in_epoch

    @property
    def state_particles(self):
        packed, _ = self.solver.planner.root_particles()
        return [self.solver.dev.unpack_state(p) for p in packed]

    def sample_particle(self):
        """belief_node.py:47-48 (random.choice over the bag), drawn from the host Philox stream."""
        packed, _ = self.solver.planner.root_particles()
        idx = philox.randbelow(self.solver.seed, len(packed), self._draws, 0, self.solver.move, philox.DOM_AGENT)
        self._draws += 1
        return self.solver.dev.unpack_state(packed[idx])

    @property
    def action_map(self):
        st = self.solver.root_statistics()
        total_q = [float(q) / QFIX_SCALE for q in st["qfix"]]
        return ActionMappingView(self.solver.action_pool, [int(x) for x in st["n"]], total_q, st["legal"])

    @property
    def data(self):
        return self.solver.root_data()
