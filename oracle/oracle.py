"""ctypes front-end of the CPU oracle (``oracle/liboracle.so``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product (``pomdpy_b200/``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DOM_SIM, DOM_UPDATE, DOM_FINAL, DOM_AGENT, DOM_SEQ = 0, 1, 2, 3, 7
QFIX_SCALE = float(1 << 24)


class StepOut(C.Structure):
    _fields_ = [("next_cell", C.c_int32), ("next_rocks", C.c_uint32), ("obs_good", C.c_int32),
                ("obs_empty", C.c_int32), ("terminal", C.c_int32), ("legal", C.c_int32),
                ("used_uniform", C.c_int32), ("rewarded_sample", C.c_int32), ("reward", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pomcp_oracle.c",)]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, u32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_double
        P = C.POINTER
        L.orc_philox4x32_10.argtypes = [P(u32), P(u32), P(u32)]
        L.orc_philox4x32.argtypes = [P(u32), P(u32), P(u32), C.c_int]
        L.orc_model_create.restype = vp
        L.orc_model_create.argtypes = [i32, i32, P(C.c_int8), i32, P(u64), P(dbl)]
        L.orc_model_destroy.argtypes = [vp]
        L.orc_legal_mask.restype = u32
        L.orc_legal_mask.argtypes = [vp, i32]
        L.orc_model_n_actions.argtypes = [vp]
        L.orc_step.argtypes = [vp, P(u32), i32, u32, i32, u64, P(StepOut)]
        L.orc_ref_create.restype = vp
        L.orc_ref_create.argtypes = [vp, u32, u32, P(u32), i32, i32, i32, dbl, dbl, u64, u32]
        L.orc_ref_destroy.argtypes = [vp]
        L.orc_ref_set_quirks.argtypes = [vp, i32, i32, i32]
        L.orc_ref_set_world.argtypes = [vp, u32, u32]
        L.orc_ref_draws.restype = u64
        L.orc_ref_draws.argtypes = [vp]
        L.orc_ref_disabled.argtypes = [vp]
        L.orc_ref_search.argtypes = [vp, i32]
        L.orc_ref_root_stats.argtypes = [vp, P(i64), P(dbl), P(dbl), P(u32), P(i64)]
        L.orc_ref_counters.argtypes = [vp, P(i64)]
        L.orc_ref_sample_root_particle.argtypes = [vp]
        L.orc_ref_particle.argtypes = [vp, i32, P(i32), P(u32)]
        L.orc_ref_env_step.argtypes = [vp, i32, i32, P(StepOut)]
        L.orc_ref_update.argtypes = [vp, i32, i32, u32]
        L.orc_det_log_u64.restype = dbl
        L.orc_det_log_u64.argtypes = [u64]
        L.orc_gs_create.restype = vp
        L.orc_gs_create.argtypes = [vp, u32, u32, P(u32), i32, i32, i32, i32, dbl, dbl, u64]
        L.orc_gs_destroy.argtypes = [vp]
        L.orc_gs_set_world.argtypes = [vp, u32, u32]
        L.orc_gs_search.argtypes = [vp, i64, u32, u32, i32, i32, i32]
        L.orc_gs_root_stats.argtypes = [vp, P(i64), P(i64), P(u32)]
        L.orc_gs_counters.argtypes = [vp, P(i64)]
        L.orc_gs_node_stats.argtypes = [vp, P(i32), i32, P(i64), P(i64), P(u32), P(i32)]
        L.orc_gs_greedy.argtypes = [P(i64), P(i64), u32, i32, u64, u32]
        L.orc_gs_update.restype = i64
        L.orc_gs_update.argtypes = [vp, i32, i32, u32, u32, i32, i64, P(i32), P(i64)]
        L.orc_gs_bag.argtypes = [vp, P(u32)]
        L.orc_gs_bag_size.argtypes = [vp]
        L.orc_gs_root_cell.argtypes = [vp]
        L.orc_gs_set_preferred.argtypes = [vp, i32]
        L.orc_model_set_prob.argtypes = [vp, P(dbl)]
        L.orc_tab_create.restype = vp
        L.orc_tab_create.argtypes = [i32, i32, i32, P(u32), P(u32), P(dbl), u32, P(C.c_uint8)]
        L.orc_tab_destroy.argtypes = [vp]
        L.orc_tab_step.argtypes = [vp, i32, i32, u32, u32, i32, P(i32), P(i32), P(dbl), P(i32)]
        L.orc_gs_create_tab.restype = vp
        L.orc_gs_create_tab.argtypes = [vp, P(u32), i32, i32, i32, dbl, dbl, u64]
        L.orc_rockstats_init.argtypes = [vp, vp]
        L.orc_rockstats_child.argtypes = [vp, vp, i32, i32, i32, vp]
        L.orc_smart_mask.restype = u32
        L.orc_smart_mask.argtypes = [vp, vp, i32]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def philox(ctr, key, rounds=10):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32), int(rounds))
    return out


def det_log(x):
    return lib().orc_det_log_u64(int(x))


class Model:
    """RockSample tables: cell codes, start cell, 53-bit sensor thresholds [cell][rock], signed rewards
    (illegal, exit, good, bad)."""

    def __init__(self, rows, cols, cell, start_cell, thr53, rewards=(-100.0, 10.0, 10.0, -10.0)):
        self.rows, self.cols = int(rows), int(cols)
        self.cell = np.ascontiguousarray(cell, dtype=np.int8).reshape(-1)
        self.n_rocks = int(self.cell.max()) + 1
        self.n_actions = 5 + self.n_rocks
        self.start_cell = int(start_cell)
        self.thr53 = np.ascontiguousarray(thr53, dtype=np.uint64).reshape(self.rows * self.cols, self.n_rocks)
        self.rewards = np.asarray(rewards, dtype=np.float64)
        self.h = lib().orc_model_create(self.rows, self.cols, _p(self.cell, C.c_int8), self.start_cell,
                                        _p(self.thr53, C.c_uint64), _p(self.rewards, C.c_double))
        if not self.h:
            raise ValueError("model too large for the oracle")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_destroy(self.h)
            self.h = None

    def legal_mask(self, cell):
        return int(lib().orc_legal_mask(self.h, int(cell)))

    def set_prob(self, prob):
        """Sensor correctness probabilities [cell][rock] (needed by the preferred-action heuristic)."""
        p = np.ascontiguousarray(prob, dtype=np.float64).reshape(-1)
        lib().orc_model_set_prob(self.h, _p(p, C.c_double))

    def step(self, cell, rocks, action, u53, actual_mask, sampled_mask):
        w = np.array([actual_mask, sampled_mask], dtype=np.uint32)
        o = StepOut()
        lib().orc_step(self.h, _p(w, C.c_uint32), int(cell), int(rocks), int(action), int(u53), C.byref(o))
        return o


class TabularModel:
    """Tabular POMDP tables: cumulative 32-bit thresholds Tc[a][s][s'], Oc[a][s'][z], rewards R[a][s], the mask of
    actions that end the episode and per-state terminal flags."""

    def __init__(self, Tc, Oc, R, term_action_mask=0, term_state=None):
        self.Tc = np.ascontiguousarray(Tc, dtype=np.uint32)
        self.Oc = np.ascontiguousarray(Oc, dtype=np.uint32)
        self.R = np.ascontiguousarray(R, dtype=np.float64)
        self.n_actions, self.n_states, _ = self.Tc.shape
        self.n_obs = self.Oc.shape[2]
        self.start_cell = 0
        ts = np.zeros(self.n_states, dtype=np.uint8) if term_state is None else \
            np.ascontiguousarray(term_state, dtype=np.uint8)
        self.h = lib().orc_tab_create(self.n_states, self.n_actions, self.n_obs, _p(self.Tc, C.c_uint32),
                                      _p(self.Oc, C.c_uint32), _p(self.R, C.c_double), int(term_action_mask),
                                      _p(ts, C.c_uint8))
        if not self.h:
            raise ValueError("tabular model outside the oracle's limits")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tab_destroy(self.h)
            self.h = None

    def step(self, s, a, w1, w2, with_obs=True):
        nxt, obs, term, rew = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
        lib().orc_tab_step(self.h, int(s), int(a), int(w1), int(w2), int(bool(with_obs)), C.byref(nxt),
                           C.byref(obs), C.byref(rew), C.byref(term))
        return nxt.value, obs.value, rew.value, term.value


class RockStats(C.Structure):
    """Per-node rock statistics of the preferred-action heuristic (rock_position_history.py:14-26)."""
    _fields_ = [("chance", C.c_double * 24), ("good", C.c_int32 * 24)]


def rockstats_init(model):
    s = RockStats()
    lib().orc_rockstats_init(model.h, C.byref(s))
    return s


def rockstats_child(model, parent, cell, action, obs_good):
    c = RockStats()
    lib().orc_rockstats_child(model.h, C.byref(parent), int(cell), int(action), int(bool(obs_good)), C.byref(c))
    return c


def smart_mask(model, stats, cell):
    return int(lib().orc_smart_mask(model.h, C.byref(stats), int(cell)))


class RefPOMCP:
    """The reference's sequential POMCP (quirks on by default), one sequential Philox draw stream."""

    def __init__(self, model, actual_mask, sampled_mask, bag_rocks, max_depth=100, max_particle_count=2000,
                 ucb_c=3.0, discount=0.95, seed=0, stream_id=0):
        self.model = model
        b = np.ascontiguousarray(bag_rocks, dtype=np.uint32)
        self.h = lib().orc_ref_create(model.h, int(actual_mask), int(sampled_mask), _p(b, C.c_uint32), len(b),
                                      int(max_depth), int(max_particle_count), float(ucb_c), float(discount),
                                      int(seed), int(stream_id))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ref_destroy(self.h)
            self.h = None

    def set_quirks(self, stale_mean=1, mutate_input=1, libm_log=1):
        lib().orc_ref_set_quirks(self.h, int(stale_mean), int(mutate_input), int(libm_log))

    def search(self, n_sims):
        return int(lib().orc_ref_search(self.h, int(n_sims)))

    def root_stats(self):
        A = self.model.n_actions
        n = np.zeros(A, dtype=np.int64)
        tot = np.zeros(A, dtype=np.float64)
        mean = np.zeros(A, dtype=np.float64)
        legal = C.c_uint32()
        N = C.c_int64()
        lib().orc_ref_root_stats(self.h, _p(n, C.c_int64), _p(tot, C.c_double), _p(mean, C.c_double),
                                 C.byref(legal), C.byref(N))
        return dict(n=n, total=tot, mean=mean, legal=legal.value, N=N.value)

    def counters(self):
        c = np.zeros(6, dtype=np.int64)
        lib().orc_ref_counters(self.h, _p(c, C.c_int64))
        return dict(zip(("steps", "levels", "rollout_steps", "backups", "zero_backups", "nodes"), c.tolist()))

    def draws(self):
        return int(lib().orc_ref_draws(self.h))

    def disabled(self):
        return bool(lib().orc_ref_disabled(self.h))

    def sample_root_particle(self):
        return int(lib().orc_ref_sample_root_particle(self.h))

    def particle(self, p):
        cell, rocks = C.c_int32(), C.c_uint32()
        lib().orc_ref_particle(self.h, int(p), C.byref(cell), C.byref(rocks))
        return cell.value, rocks.value

    def env_step(self, p, action):
        o = StepOut()
        q = lib().orc_ref_env_step(self.h, int(p), int(action), C.byref(o))
        return int(q), o

    def update(self, action, obs_good, sampled_mask_after):
        return int(lib().orc_ref_update(self.h, int(action), int(bool(obs_good)), int(sampled_mask_after)))


class GsPOMCP:
    """Generation-synchronous POMCP (the algorithm the CUDA kernels implement), on the CPU."""

    def __init__(self, model, actual_mask, sampled_mask, bag_packed, root_cell=None, max_depth=100, dcap=64,
                 ucb_c=3.0, discount=0.95, seed=0):
        self.model = model
        self.seed = int(seed)
        b = np.ascontiguousarray(bag_packed, dtype=np.uint32)
        rc = model.start_cell if root_cell is None else int(root_cell)
        self.tabular = isinstance(model, TabularModel)
        if self.tabular:
            self.h = lib().orc_gs_create_tab(model.h, _p(b, C.c_uint32), len(b), int(max_depth), int(dcap),
                                             float(ucb_c), float(discount), self.seed)
        else:
            self.h = lib().orc_gs_create(model.h, int(actual_mask), int(sampled_mask), _p(b, C.c_uint32), len(b), rc,
                                         int(max_depth), int(dcap), float(ucb_c), float(discount), self.seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gs_destroy(self.h)
            self.h = None

    def set_world(self, actual, sampled):
        lib().orc_gs_set_world(self.h, int(actual), int(sampled))

    def set_preferred(self, enabled=True, rollouts=False):
        """rollouts=True: the rollout policy draws from generate_smart_actions of the rollout's own history too."""
        lib().orc_gs_set_preferred(self.h, (2 if rollouts else 1) if enabled else 0)

    def search(self, n_sims, move=0, sim_offset=0, w0=1, growth=1, wmax=1):
        lib().orc_gs_search(self.h, int(n_sims), int(sim_offset), int(move), int(w0), int(growth), int(wmax))

    def root_stats(self):
        A = self.model.n_actions
        n = np.zeros(A, dtype=np.int64)
        q = np.zeros(A, dtype=np.int64)
        legal = C.c_uint32()
        lib().orc_gs_root_stats(self.h, _p(n, C.c_int64), _p(q, C.c_int64), C.byref(legal))
        return dict(n=n, qfix=q, legal=legal.value)

    def node_stats(self, path):
        """path: list of (action, obs_good) from the root."""
        A = self.model.n_actions
        p = np.asarray(path, dtype=np.int32).reshape(-1)
        n = np.zeros(A, dtype=np.int64)
        q = np.zeros(A, dtype=np.int64)
        legal, cell = C.c_uint32(), C.c_int32()
        rc = lib().orc_gs_node_stats(self.h, _p(p, C.c_int32), len(p) // 2, _p(n, C.c_int64), _p(q, C.c_int64),
                                     C.byref(legal), C.byref(cell))
        if rc:
            return None
        return dict(n=n, qfix=q, legal=legal.value, cell=cell.value)

    def counters(self):
        c = np.zeros(8, dtype=np.int64)
        lib().orc_gs_counters(self.h, _p(c, C.c_int64))
        return dict(zip(("levels", "rollout_steps", "created", "nodes", "alloc_nodes", "generations",
                         "max_tree_depth", "bag"), c.tolist()))

    def greedy(self, move=0, stats=None):
        s = stats or self.root_stats()
        n = np.ascontiguousarray(s["n"], dtype=np.int64)
        q = np.ascontiguousarray(s["qfix"], dtype=np.int64)
        return int(lib().orc_gs_greedy(_p(n, C.c_int64), _p(q, C.c_int64), int(s["legal"]), self.model.n_actions,
                                       self.seed, int(move)))

    def update(self, action, obs_good, sampled_mask_after, move=0, need=2000, max_tries=None):
        status, tries = C.c_int32(), C.c_int64()
        mt = int(max_tries if max_tries is not None else 256 * need)
        got = lib().orc_gs_update(self.h, int(action), int(obs_good), int(sampled_mask_after), int(move),
                                  int(need), mt, C.byref(status), C.byref(tries))
        return dict(accepted=int(got), status=status.value, tries=tries.value)

    def bag(self):
        out = np.zeros(lib().orc_gs_bag_size(self.h), dtype=np.uint32)
        lib().orc_gs_bag(self.h, _p(out, C.c_uint32))
        return out

    def root_cell(self):
        return int(lib().orc_gs_root_cell(self.h))
