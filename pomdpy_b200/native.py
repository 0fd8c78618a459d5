"""ctypes binding of the C ABI in include/pomcp_b200.h (libpomcp_b200.so).

There is no CPU fallback: if the library is missing or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

QFIX_SCALE = float(1 << 24)
MERGE_BYTES = (2 * 8 * 64 + 2 * 8) * 8       # POMCP_MERGE_BYTES of include/pomcp_b200.h
MAX_ACTIONS = 32
_LIB = None


class PomcpError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("tree_depth_cap", C.c_int32), ("max_particle_count", C.c_int32),
                ("reserved0", C.c_int32), ("ucb_coefficient", C.c_double), ("discount", C.c_double),
                ("seed", C.c_uint64), ("node_capacity", C.c_int64), ("gen_w0", C.c_int32),
                ("gen_growth", C.c_int32), ("gen_wmax", C.c_int32), ("child_table", C.c_int32)]


EXPORTS = ["pomcp_create", "pomcp_destroy", "pomcp_last_error", "pomcp_set_model_rocksample", "pomcp_set_world",
           "pomcp_set_root_particles", "pomcp_search", "pomcp_root_stats", "pomcp_node_stats", "pomcp_advance",
           "pomcp_root_particles", "pomcp_generate_steps", "pomcp_counters", "pomcp_root_stats_device",
           "pomcp_reset_tree", "pomcp_debug_steplog", "pomcp_set_preferred_actions",
           "pomcp_root_rock_data", "pomcp_refexact_create", "pomcp_refexact_destroy", "pomcp_refexact_last_error",
           "pomcp_refexact_search", "pomcp_refexact_sample_root_particle", "pomcp_refexact_env_step",
           "pomcp_refexact_update", "pomcp_set_model_tabular", "pomcp_tabular_steps",
           "pomcp_set_peer_merge", "pomcp_merged_root_stats"]


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libpomcp_b200.so (building it first if the sources are newer and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("POMCP_B200_LIB") or library_path()      # override: experiment builds
    if build_if_missing and not os.environ.get("POMCP_B200_LIB") and _build.stale():
        try:
            _build.build()
        except Exception as e:            # no nvcc on this box: a prebuilt library must already be there
            if not os.path.exists(path):
                raise PomcpError("libpomcp_b200.so is missing and cannot be built: %s" % e)
    if not os.path.exists(path):
        raise PomcpError("libpomcp_b200.so not found; run `python -m pomdpy_b200.build`")
    L = C.CDLL(path)
    vp, i32, u32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    L.pomcp_create.argtypes = [P(Config), C.c_int, P(vp)]
    L.pomcp_destroy.argtypes = [vp]
    L.pomcp_destroy.restype = None
    L.pomcp_last_error.argtypes = [vp]
    L.pomcp_last_error.restype = C.c_char_p
    L.pomcp_set_model_rocksample.argtypes = [vp, C.c_int, C.c_int, P(C.c_int8), C.c_int, P(u64), P(dbl)]
    L.pomcp_set_model_tabular.argtypes = [vp, C.c_int, C.c_int, C.c_int, P(u32), P(u32), P(dbl), u32, P(C.c_uint8)]
    L.pomcp_tabular_steps.argtypes = [vp, i64, P(u32), P(C.c_uint8), P(u32), P(u32), P(u32), P(i32), P(dbl),
                                      P(C.c_uint8)]
    L.pomcp_set_peer_merge.argtypes = [vp, C.c_int, C.c_int, P(u64)]
    L.pomcp_merged_root_stats.argtypes = [vp, P(i64), P(i64), P(u32), P(i64)]
    L.pomcp_set_world.argtypes = [vp, u32, u32]
    L.pomcp_set_root_particles.argtypes = [vp, P(u32), C.c_int, C.c_int]
    L.pomcp_search.argtypes = [vp, i64, u32, u32, dbl, vp]
    L.pomcp_root_stats.argtypes = [vp, P(i64), P(i64), P(u32), P(i64)]
    L.pomcp_node_stats.argtypes = [vp, P(i32), C.c_int, P(i64), P(i64), P(u32), P(i32)]
    L.pomcp_advance.argtypes = [vp, C.c_int, C.c_int, u32, u32, i64, P(i32), P(i64), P(i32)]
    L.pomcp_root_particles.argtypes = [vp, P(u32), C.c_int, P(i32), P(i32)]
    L.pomcp_generate_steps.argtypes = [vp, i64, P(u32), P(C.c_uint8), P(u64), P(u32), P(u32), P(u32), P(C.c_uint8),
                                       P(dbl)]
    L.pomcp_counters.argtypes = [vp, P(i64)]
    L.pomcp_root_stats_device.argtypes = [vp, vp, vp]
    L.pomcp_reset_tree.argtypes = [vp, vp]
    L.pomcp_debug_steplog.argtypes = [vp, P(i64)]
    L.pomcp_set_preferred_actions.argtypes = [vp, C.c_int, P(dbl)]
    L.pomcp_root_rock_data.argtypes = [vp, P(dbl), P(i32)]
    L.pomcp_refexact_create.argtypes = [C.c_int, C.c_int, C.c_int, P(C.c_int8), C.c_int, P(u64), P(dbl), u32, P(u32),
                                        C.c_int, C.c_int, C.c_int, dbl, dbl, u64, u32, P(dbl), C.c_int, i64, i64, P(vp)]
    L.pomcp_refexact_destroy.argtypes = [vp]
    L.pomcp_refexact_destroy.restype = None
    L.pomcp_refexact_last_error.argtypes = [vp]
    L.pomcp_refexact_last_error.restype = C.c_char_p
    L.pomcp_refexact_search.argtypes = [vp, C.c_int, P(i32), P(i64), P(dbl), P(dbl), P(u32), P(i64), P(u64)]
    L.pomcp_refexact_sample_root_particle.argtypes = [vp, P(i32)]
    L.pomcp_refexact_env_step.argtypes = [vp, C.c_int, C.c_int, P(i64), P(dbl), P(u64)]
    L.pomcp_refexact_update.argtypes = [vp, C.c_int, C.c_int, u32, P(i32), P(u64)]
    _LIB = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Planner:
    """One device-resident POMCP planner (tree arena + belief) behind the C ABI."""

    def __init__(self, max_depth=100, tree_depth_cap=64, max_particle_count=2000, ucb_coefficient=3.0,
                 discount=0.95, seed=0, node_capacity=1 << 20, gen_w0=32, gen_growth=2, gen_wmax=1 << 17, device=0, hashed_children=False):
        self.lib = load()
        self.cfg = Config(int(max_depth), int(tree_depth_cap), int(max_particle_count), 0, float(ucb_coefficient),
                          float(discount), int(seed) & 0xFFFFFFFFFFFFFFFF, int(node_capacity), int(gen_w0),
                          int(gen_growth), int(gen_wmax), 1 if hashed_children else 0)
        self.h = C.c_void_p()
        self.n_actions = 0
        self.tabular = False
        self._bag_hint = None
        rc = self.lib.pomcp_create(C.byref(self.cfg), int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.pomcp_last_error(self.h).decode() if self.h else "pomcp_create failed"
            if self.h:
                self.lib.pomcp_destroy(self.h)
                self.h = C.c_void_p()
            raise PomcpError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pomcp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise PomcpError(self.lib.pomcp_last_error(self.h).decode())
        return rc

    def set_model_rocksample(self, rows, cols, cell, start_cell, thr53, rewards):
        cell = np.ascontiguousarray(cell, dtype=np.int8).reshape(-1)
        n_rocks = int(cell.max()) + 1
        thr53 = np.ascontiguousarray(thr53, dtype=np.uint64).reshape(-1)
        if thr53.size != rows * cols * n_rocks:
            raise ValueError("thr53 must be [rows*cols][n_rocks]")
        rewards = np.ascontiguousarray(rewards, dtype=np.float64)
        self._ck(self.lib.pomcp_set_model_rocksample(self.h, int(rows), int(cols), _p(cell, C.c_int8),
                                                     int(start_cell), _p(thr53, C.c_uint64),
                                                     _p(rewards, C.c_double)))
        self.n_actions = 5 + n_rocks
        self.tabular = False

    def set_model_tabular(self, Tc, Oc, R, term_action_mask=0, term_state=None):
        """Tabular POMDP: cumulative thresholds Tc[a][s][s'], Oc[a][s'][z] (uint32, rows end at 0xFFFFFFFF), rewards
        R[a][s]; see pomdpy_b200.tabular.cumulative_thresholds."""
        Tc = np.ascontiguousarray(Tc, dtype=np.uint32)
        Oc = np.ascontiguousarray(Oc, dtype=np.uint32)
        R = np.ascontiguousarray(R, dtype=np.float64)
        A, S, S2 = Tc.shape
        if S2 != S or Oc.shape[:2] != (A, S) or R.shape != (A, S):
            raise ValueError("expected Tc[A][S][S], Oc[A][S][Z], R[A][S]")
        ts = None if term_state is None else np.ascontiguousarray(term_state, dtype=np.uint8)
        if ts is not None and ts.shape != (S,):
            raise ValueError("term_state must be [S]")
        self._ck(self.lib.pomcp_set_model_tabular(self.h, S, A, Oc.shape[2], _p(Tc, C.c_uint32), _p(Oc, C.c_uint32),
                                                  _p(R, C.c_double), int(term_action_mask),
                                                  None if ts is None else _p(ts, C.c_uint8)))
        self.n_actions = A
        self.tabular = True

    def tabular_steps(self, state, action, w1, w2):
        state = np.ascontiguousarray(state, dtype=np.uint32)
        action = np.ascontiguousarray(action, dtype=np.uint8)
        w1 = np.ascontiguousarray(w1, dtype=np.uint32)
        w2 = np.ascontiguousarray(w2, dtype=np.uint32)
        n = state.size
        ns, ob = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.int32)
        rw, tm = np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.pomcp_tabular_steps(self.h, n, _p(state, C.c_uint32), _p(action, C.c_uint8),
                                              _p(w1, C.c_uint32), _p(w2, C.c_uint32), _p(ns, C.c_uint32),
                                              _p(ob, C.c_int32), _p(rw, C.c_double), _p(tm, C.c_uint8)))
        return ns, ob, rw, tm

    def set_preferred_actions(self, enabled, prob=None, rollouts=False):
        """--preferred_actions: in-tree action sets from per-node rock statistics; prob [cells][n_rocks] float64.
        rollouts=True: the rollout policy draws from the smart actions of the rollout's own history as well."""
        if enabled:
            prob = np.ascontiguousarray(prob, dtype=np.float64).reshape(-1)
            self._ck(self.lib.pomcp_set_preferred_actions(self.h, 2 if rollouts else 1, _p(prob, C.c_double)))
        else:
            self._ck(self.lib.pomcp_set_preferred_actions(self.h, 0, None))

    def root_rock_data(self):
        ch = np.zeros(24, dtype=np.float64)
        gd = np.zeros(24, dtype=np.int32)
        rc = self._ck(self.lib.pomcp_root_rock_data(self.h, _p(ch, C.c_double), _p(gd, C.c_int32)))
        n = self.n_actions - 5
        return None if rc else (ch[:n].copy(), gd[:n].copy())

    def set_world(self, actual_mask, sampled_mask):
        self._ck(self.lib.pomcp_set_world(self.h, int(actual_mask), int(sampled_mask)))

    def set_root_particles(self, packed, root_cell=0):
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        self._bag_hint = int(packed.size)
        self._ck(self.lib.pomcp_set_root_particles(self.h, _p(packed, C.c_uint32), int(packed.size), int(root_cell)))

    def search(self, n_sims, move=0, sim_offset=0, timeout_s=0.0, stream=None):
        return self._ck(self.lib.pomcp_search(self.h, int(n_sims), int(sim_offset), int(move), float(timeout_s),
                                              C.c_void_p(stream or 0)))

    def reset_tree(self, stream=None):
        self._ck(self.lib.pomcp_reset_tree(self.h, C.c_void_p(stream or 0)))

    def root_stats_device(self, dev_ptr, stream=None):
        """Write n[0..32), qfix[0..32) (int64) to the device buffer at `dev_ptr` (e.g. a torch tensor's data_ptr)."""
        self._ck(self.lib.pomcp_root_stats_device(self.h, C.c_void_p(int(dev_ptr)), C.c_void_p(stream or 0)))

    def root_stats(self):
        n = np.zeros(self.n_actions, dtype=np.int64)
        q = np.zeros(self.n_actions, dtype=np.int64)
        legal, done = C.c_uint32(), C.c_int64()
        self._ck(self.lib.pomcp_root_stats(self.h, _p(n, C.c_int64), _p(q, C.c_int64), C.byref(legal),
                                           C.byref(done)))
        return dict(n=n, qfix=q, legal=legal.value, sims_done=done.value)

    def set_peer_merge(self, world, rank, peer_buffers):
        """Fuse the root-parallel merge into the search kernel: peer_buffers[r] = device address of rank r's exchange
        buffer (MERGE_BYTES, zeroed) as mapped into this process.  world <= 1 switches it off."""
        ptrs = np.ascontiguousarray(list(peer_buffers) + [0] * (8 - len(peer_buffers)), dtype=np.uint64)
        self._ck(self.lib.pomcp_set_peer_merge(self.h, int(world), int(rank), _p(ptrs, C.c_uint64)))

    def merged_root_stats(self):
        """Root statistics summed over the ranks by the last search (set_peer_merge); raises if a peer was missing."""
        n = np.zeros(self.n_actions, dtype=np.int64)
        q = np.zeros(self.n_actions, dtype=np.int64)
        legal, done = C.c_uint32(), C.c_int64()
        self._ck(self.lib.pomcp_merged_root_stats(self.h, _p(n, C.c_int64), _p(q, C.c_int64), C.byref(legal),
                                                  C.byref(done)))
        return dict(n=n, qfix=q, legal=legal.value, sims_done=done.value)

    def node_stats(self, path):
        p = np.asarray(path, dtype=np.int32).reshape(-1)
        n = np.zeros(self.n_actions, dtype=np.int64)
        q = np.zeros(self.n_actions, dtype=np.int64)
        legal, cell = C.c_uint32(), C.c_int32()
        rc = self._ck(self.lib.pomcp_node_stats(self.h, _p(p, C.c_int32), p.size // 2, _p(n, C.c_int64),
                                                _p(q, C.c_int64), C.byref(legal), C.byref(cell)))
        if rc:
            return None
        return dict(n=n, qfix=q, legal=legal.value, cell=cell.value)

    def advance(self, action, obs_good, sampled_mask_after=0, move=0, max_tries=0):
        """Belief update + re-root.  obs_good=None: unconditioned on the observation (every try is accepted, the root is
        re-planted as a fresh node at the next position) -- the degraded mode after a failed update."""
        status, tries, acc = C.c_int32(), C.c_int64(), C.c_int32()
        obs = -1 if obs_good is None else int(obs_good) if self.tabular else int(bool(obs_good))   # tabular: the index
        self._ck(self.lib.pomcp_advance(self.h, int(action), obs, int(sampled_mask_after), int(move),
                                        int(max_tries), C.byref(status), C.byref(tries), C.byref(acc)))
        return dict(status=status.value, tries=tries.value, accepted=acc.value)

    def root_particles(self, fetch=True):
        """(packed states, root cell) of the current root belief; fetch=False returns only (count, root cell)."""
        n, cell = C.c_int32(), C.c_int32()
        if not fetch:
            self._ck(self.lib.pomcp_root_particles(self.h, None, 0, C.byref(n), C.byref(cell)))
            return n.value, cell.value
        cap = int(self.cfg.max_particle_count) if self._bag_hint is None else max(self._bag_hint, 1)
        out = np.empty(max(cap, int(self.cfg.max_particle_count)), dtype=np.uint32)
        self._ck(self.lib.pomcp_root_particles(self.h, _p(out, C.c_uint32), out.size, C.byref(n), C.byref(cell)))
        if n.value > out.size:                           # an initial belief larger than max_particle_count
            out = np.empty(n.value, dtype=np.uint32)
            self._ck(self.lib.pomcp_root_particles(self.h, _p(out, C.c_uint32), out.size, C.byref(n), C.byref(cell)))
        return out[:n.value].copy(), cell.value

    def generate_steps(self, state, action, u53, actual_mask, sampled_mask):
        state = np.ascontiguousarray(state, dtype=np.uint32)
        actual_mask = np.ascontiguousarray(np.broadcast_to(actual_mask, state.shape), dtype=np.uint32)
        sampled_mask = np.ascontiguousarray(np.broadcast_to(sampled_mask, state.shape), dtype=np.uint32)
        action = np.ascontiguousarray(action, dtype=np.uint8)
        u53 = np.ascontiguousarray(u53, dtype=np.uint64)
        n = state.size
        ns = np.zeros(n, dtype=np.uint32)
        fl = np.zeros(n, dtype=np.uint8)
        rw = np.zeros(n, dtype=np.float64)
        self._ck(self.lib.pomcp_generate_steps(self.h, n, _p(state, C.c_uint32), _p(action, C.c_uint8),
                                               _p(u53, C.c_uint64), _p(actual_mask, C.c_uint32),
                                               _p(sampled_mask, C.c_uint32), _p(ns, C.c_uint32), _p(fl, C.c_uint8),
                                               _p(rw, C.c_double)))
        return ns, fl, rw

    def steplog(self):
        out = np.zeros((128, 8), dtype=np.int64)
        self._ck(self.lib.pomcp_debug_steplog(self.h, _p(out, C.c_int64)))
        return out

    def counters(self):
        c = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.pomcp_counters(self.h, _p(c, C.c_int64)))
        names = ("levels", "rollout_steps", "created", "nodes", "alloc_nodes", "generations", "level_steps",
                 "last_search_ns", "last_search_sims", "launches", "coop_blocks", "threads_per_block",
                 "root_stats_d2h_bytes", "phase_descend_ns", "merge_wait_ns", "check_failed")
        return dict(zip(names, c.tolist()))


class ReferenceExact:
    """The reference's sequential POMCP, quirks included, as a single-thread CUDA kernel (parity path; see
    csrc/pomcp_refexact.cu).  Draws come from one sequential Philox stream keyed by `seed`."""

    def __init__(self, rows, cols, cell, start_cell, thr53, rewards, actual_mask, bag_rocks, max_depth=100,
                 max_particle_count=2000, ucb_c=3.0, discount=0.95, seed=0, stream_id=0, n_log=1 << 16,
                 cap_nodes=1 << 16, cap_particles=1 << 22, device=0):
        self.lib = load()
        cell = np.ascontiguousarray(cell, dtype=np.int8).reshape(-1)
        self.n_actions = 5 + int(cell.max()) + 1
        thr53 = np.ascontiguousarray(thr53, dtype=np.uint64).reshape(-1)
        rewards = np.ascontiguousarray(rewards, dtype=np.float64)
        bag = np.ascontiguousarray(bag_rocks, dtype=np.uint32)
        # ln(N + 1) exactly as the reference evaluates it: np.log on a Python int, one scalar at a time
        log_table = np.array([np.log(N + 1) for N in range(n_log)], dtype=np.float64)
        self.h = C.c_void_p()
        rc = self.lib.pomcp_refexact_create(int(device), int(rows), int(cols), _p(cell, C.c_int8), int(start_cell),
                                            _p(thr53, C.c_uint64), _p(rewards, C.c_double), int(actual_mask),
                                            _p(bag, C.c_uint32), int(bag.size), int(max_depth), int(max_particle_count),
                                            float(ucb_c), float(discount), int(seed), int(stream_id),
                                            _p(log_table, C.c_double), int(n_log), int(cap_nodes), int(cap_particles),
                                            C.byref(self.h))
        if rc != 0:
            msg = self.lib.pomcp_refexact_last_error(self.h).decode()
            self.close()
            raise PomcpError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pomcp_refexact_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise PomcpError(self.lib.pomcp_refexact_last_error(self.h).decode())
        return rc

    def search(self, n_sims):
        A = self.n_actions
        n = np.zeros(A, dtype=np.int64)
        tot = np.zeros(A, dtype=np.float64)
        mean = np.zeros(A, dtype=np.float64)
        act, legal, N, draws = C.c_int32(), C.c_uint32(), C.c_int64(), C.c_uint64()
        self._ck(self.lib.pomcp_refexact_search(self.h, int(n_sims), C.byref(act), _p(n, C.c_int64), _p(tot, C.c_double),
                                                _p(mean, C.c_double), C.byref(legal), C.byref(N), C.byref(draws)))
        return dict(action=act.value, n=n, total=tot, mean=mean, legal=legal.value, N=N.value, draws=draws.value)

    def sample_root_particle(self):
        p = C.c_int32()
        self._ck(self.lib.pomcp_refexact_sample_root_particle(self.h, C.byref(p)))
        return p.value

    def env_step(self, particle, action):
        out = np.zeros(6, dtype=np.int64)
        rew, draws = C.c_double(), C.c_uint64()
        self._ck(self.lib.pomcp_refexact_env_step(self.h, int(particle), int(action), _p(out, C.c_int64), C.byref(rew),
                                                  C.byref(draws)))
        return dict(obs_good=int(out[0]), terminal=int(out[1]), legal=int(out[2]), next_cell=int(out[3]),
                    next_rocks=int(out[4]), particle=int(out[5]), reward=rew.value, draws=draws.value)

    def update(self, action, obs_good, sampled_after):
        dis, draws = C.c_int32(), C.c_uint64()
        rc = self._ck(self.lib.pomcp_refexact_update(self.h, int(action), int(bool(obs_good)), int(sampled_after),
                                                     C.byref(dis), C.byref(draws)))
        return dict(status=rc, disable_tree=bool(dis.value), draws=draws.value)
