/*
 * pomcp_b200.h -- C ABI of the B200-native POMCP planner (libpomcp_b200.so).
 *
 * The reference (pemami4911/POMDPy) is pure Python and has no FFI; its solver boundary is the
 * duck-typed object that pomdpy/agent.py drives.  Each entry point below names the reference
 * call it replaces (file:line in the reference tree).  Plain pointers and sizes only; all host
 * arrays are borrowed for the duration of the call; one host thread per handle.
 *
 * Return value: 0 ok, >0 degraded (documented per call), <0 error (text via pomcp_last_error).
 */
#ifndef POMCP_B200_H
#define POMCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POMCP_MAX_ACTIONS 32   /* |A| = 5 + n_rocks <= 32: one warp lane per action */
#define POMCP_MAX_CELLS 256
#define POMCP_MAX_ROCKS 24
#define POMCP_MAX_PEERS 8      /* root-parallel merge inside the search kernel: the GPUs of one NVSwitch box */
#define POMCP_MAX_OBS 2048     /* tabular models: |Z| <= 2048 observations (11 bits of a path entry) */
#define POMCP_MAX_OBS_DIRECT 32 /* ... up to here the child table is direct-indexed [node][action][Z]; above, hashed */
#define POMCP_QFIX_SCALE 16777216.0 /* root statistics carry sum(q) as a 2^-24 fixed-point integer */

typedef struct pomcp_handle pomcp_handle;

/* Hyper-parameters the reference solver reads from agent.model.<flag>
 * (pomdpy/pomdp/model.py:30-31; flags pomcp.py:13-43). */
typedef struct {
    int32_t max_depth;          /* --max_depth: search horizon and rollout length */
    int32_t tree_depth_cap;     /* path-buffer depth of the device tree (<= 64) */
    int32_t max_particle_count; /* --max_particle_count: size of a re-rooted belief */
    int32_t reserved0;
    double ucb_coefficient;     /* --ucb_coefficient */
    double discount;            /* --discount */
    uint64_t seed;              /* Philox key */
    int64_t node_capacity;      /* device tree arena, in belief nodes */
    int32_t gen_w0;             /* generation schedule: w0, w0*growth, ... <= wmax simulations in flight */
    int32_t gen_growth;
    int32_t gen_wmax;
    int32_t child_table;        /* tabular models: 0 = direct-indexed child table up to POMCP_MAX_OBS_DIRECT observations and
                                 * hashed above; 1 = hashed always (same results; the reference's observation map is a dict
                                 * per action node, discrete_pomdp/discrete_observation_mapping.py:12-47) */
} pomcp_config;

/* POMCP.__init__ / POMCP.reset(agent)  (pomdpy/solvers/pomcp.py:23-50, belief_tree_solver.py:20-40) */
int pomcp_create(const pomcp_config *cfg, int device, pomcp_handle **out);
void pomcp_destroy(pomcp_handle *h);
const char *pomcp_last_error(const pomcp_handle *h);

/* RockModel tables (examples/rock_sample/rock_model.py:40-138): cell codes (rock k >= 0, EMPTY -1, GOAL -2,
 * OBSTACLE -3), 53-bit sensor thresholds [cell][n_rocks] (floor(thr * 2^53), :148-150 + :395), signed rewards
 * {illegal, exit, good sample, bad sample}. */
int pomcp_set_model_rocksample(pomcp_handle *h, int rows, int cols, const int8_t *cell, int start_cell,
                               const uint64_t *thr53, const double rewards[4]);

/* Any model given by explicit matrices -- what the reference's Model ABC exposes through get_transition_matrix /
 * get_observation_matrix / get_reward_matrix (pomdpy/pomdp/model.py:133-160, examples/tiger/tiger_model.py:102-139) --
 * as 32-bit cumulative thresholds: Tc[a][s][s'] = min(floor(2^32 * sum_{j<=s'} T[a][s][j]), 2^32 - 1) and
 * Oc[a][s'][z] likewise (every row must end at 0xFFFFFFFF); R[a][s]; term_action_mask marks the actions that end an
 * episode (Tiger: opening a door), term_state[s'] the absorbing states (may be NULL).  A packed state is the state
 * index; observations are indices 0..n_obs-1 (the obs_good argument of pomcp_advance / pomcp_node_stats).  Rows of
 * 64 or more states get a guide table on the device (same draws, fewer loads).  The reference keeps the children of an
 * action node in a dict keyed by observation (discrete_pomdp/discrete_observation_mapping.py:12-47): up to
 * POMCP_MAX_OBS_DIRECT observations the device table is direct-indexed, above it is an open-addressing hash table over
 * (node, action, observation) -- same tree, same statistics (pomcp_config.child_table forces the hashed table). */
int pomcp_set_model_tabular(pomcp_handle *h, int n_states, int n_actions, int n_obs, const uint32_t *Tc,
                            const uint32_t *Oc, const double *R, uint32_t term_action_mask,
                            const uint8_t *term_state);

/* --preferred_actions (pomcp.py:30-31; examples/rock_sample/rock_position_history.py:52-55, 96-137, 170-234): nodes
 * carry per-rock statistics and their in-tree action set is generate_smart_actions(); prob is the sensor
 * correctness probability [cell][n_rocks] (rock_model.py:148-150).  Call after the model, before the root belief.
 * enabled = 2 (additive; the reference has no such policy): the rollouts draw from generate_smart_actions of their own
 * history as well -- the leaf's statistics, updated with every (action, observation) the rollout generates. */
int pomcp_set_preferred_actions(pomcp_handle *h, int enabled, const double *prob);

/* RockData of the current root (rock_position_history.py:14-26): chance_good[n_rocks], goodness_number[n_rocks];
 * returns 1 when preferred actions are off. */
int pomcp_root_rock_data(pomcp_handle *h, double *chance, int32_t *good);

/* model.actual_rock_states / model.unique_rocks_sampled as bit masks
 * (rock_model.py:263-265 reset_for_epoch, :267-272 update). */
int pomcp_set_world(pomcp_handle *h, uint32_t actual_rock_mask, uint32_t sampled_rock_mask);

/* Root belief of a fresh tree: n packed states (cell | rocks << 8)
 * (belief_tree_solver.py:30-40: tree reset + n_start_states x sample_an_init_state). */
int pomcp_set_root_particles(pomcp_handle *h, const uint32_t *packed_states, int n, int root_cell);

/* BeliefTreeSolver.monte_carlo_approx (belief_tree_solver.py:42-56): n_sims simulations with stream ids
 * [sim_offset, sim_offset + n_sims) from the current root; asynchronous on `stream` (a cudaStream_t, may be 0).
 * timeout_s <= 0 disables the wall-clock guard (solvers/pomcp.py:93-95).  Needs node_capacity - used >= n_sims;
 * returns 1 (degraded) after dropping the subtree to make room. */
int pomcp_search(pomcp_handle *h, int64_t n_sims, uint32_t sim_offset, uint32_t move, double timeout_s,
                 void *stream);

/* Root DiscreteActionMapping entries (discrete_action_mapping.py:113-184): visit_count[a], sum of q as a
 * 2^-24 fixed-point integer, legal mask.  Synchronises with the search. */
int pomcp_root_stats(pomcp_handle *h, int64_t *n, int64_t *qfix, uint32_t *legal_mask, int64_t *sims_done);

/* Root statistics into a caller-provided DEVICE buffer of 64 int64 (n[0..32), qfix[0..32)), asynchronous on
 * `stream`: the operand of the root-parallel merge (all-reduce(sum) across ranks before solvers/pomcp.py:78). */
int pomcp_root_stats_device(pomcp_handle *h, int64_t *dev_out, void *stream);

/* Root-parallel search with the merge fused into the search kernel.  peer_buffers[r] is rank r's exchange
 * buffer (POMCP_MERGE_BYTES, zero-initialised, 16-byte aligned) as mapped into THIS process -- e.g. the buffer_ptrs of
 * a torch symmetric-memory rendezvous, or cudaIpcOpenMemHandle pointers.  From then on every pomcp_search pushes the
 * rank's 64 root words to all peers over NVLink, waits for theirs and sums; all ranks must search in lockstep.
 * pomcp_merged_root_stats returns the sums (the operand of the greedy choice, solvers/pomcp.py:78); pomcp_root_stats
 * keeps returning the local tree's.  world <= 1 switches the exchange off. */
#define POMCP_MERGE_BYTES ((2 * POMCP_MAX_PEERS * 64 + 2 * POMCP_MAX_PEERS) * 8)
int pomcp_set_peer_merge(pomcp_handle *h, int world, int rank, const uint64_t *peer_buffers);
int pomcp_merged_root_stats(pomcp_handle *h, int64_t *n, int64_t *qfix, uint32_t *legal_mask, int64_t *sims_done);

/* Fresh tree over the current root belief (BeliefTree.reset + initialize, belief_tree.py:19-49), asynchronous. */
int pomcp_reset_tree(pomcp_handle *h, void *stream);

/* Same, for the node reached from the root through (action, obs_good) pairs; 1 if it does not exist. */
int pomcp_node_stats(pomcp_handle *h, const int32_t *path_ao, int len, int64_t *n, int64_t *qfix,
                     uint32_t *legal_mask, int32_t *cell);

/* BeliefTreeSolver.update (belief_tree_solver.py:154-212) + Model.generate_particles (model.py:221-252):
 * re-root at (action, obs_good), refill the belief with max_particle_count rejection samples drawn from the
 * old root belief.  *status: 0 re-rooted at an existing child, 1 fresh node (child never expanded),
 * 2 no particle accepted (the reference's disable_tree).  obs_good = -1: unconditioned -- every try is accepted
 * and the root becomes a fresh node at the next position (what the host does after status 2, so that the next
 * search returns an action that is legal at the true position; the reference goes on from the stale node with
 * rollout_search, belief_tree_solver.py:90-121). */
int pomcp_advance(pomcp_handle *h, int action, int obs_good, uint32_t sampled_mask_after, uint32_t move,
                  int64_t max_tries, int32_t *status, int64_t *tries_used, int32_t *accepted);

/* The current root belief (BeliefNode.state_particles) */
int pomcp_root_particles(pomcp_handle *h, uint32_t *out, int capacity, int32_t *n, int32_t *root_cell);

/* generate_step (rock_model.py:451-465) for a batch: state, action, 53-bit uniform and world masks per item ->
 * next state, flags (bit0 obs_good, bit1 obs_empty, bit2 terminal, bit3 legal), reward.  Host arrays. */
int pomcp_generate_steps(pomcp_handle *h, int64_t n, const uint32_t *state, const uint8_t *action,
                         const uint64_t *u53, const uint32_t *actual_mask, const uint32_t *sampled_mask,
                         uint32_t *next_state, uint8_t *flags, double *reward);

/* generate_step of the tabular model for a batch: state, action and two 32-bit draws per item (w1 -> next state,
 * w2 -> observation) -> next state, observation, reward, terminal flag.  Host arrays. */
int pomcp_tabular_steps(pomcp_handle *h, int64_t n, const uint32_t *state, const uint8_t *action, const uint32_t *w1,
                        const uint32_t *w2, uint32_t *next_state, int32_t *obs, double *reward, uint8_t *terminal);

/* counters: [0] tree levels, [1] rollout steps, [2] nodes created, [3] nodes in arena, [4] allocation nodes,
 * [5] generations, [6] level steps, [7] last search ns (device globaltimer), [8] sims of last search,
 * [9] kernel launches so far, [10] cooperative grid blocks, [11] threads per block,
 * [12] device-to-host bytes of one pomcp_root_stats call, [13] / [14] ns of the last search spent waiting at the
 * level-step barrier / allocating, [15] first failed device-side bounds check (-DPOMCP_CHECKS builds; 0 = none) */
int pomcp_counters(pomcp_handle *h, int64_t out[16]);

/* profiling aid: per level step of the last search, ns of {barrier wait, allocate (thread 0), barrier wait,
 * expand (thread 0)}, worklist nodes, 3 spare: [128][8] */
int pomcp_debug_steplog(pomcp_handle *h, int64_t *out);

/* ------------------------------------------------------------------------------------------------------------
 * Reference-exact mode: the reference's SEQUENTIAL POMCP with all its quirks as a
 * single-thread kernel.  Parity, not speed: with the sequential Philox draw stream that oracle/refharness/ injects
 * into the executed reference, root statistics, chosen action, real step and belief update equal the reference's
 * bit for bit (pomdpy/solvers/pomcp.py:69-137, belief_tree_solver.py:42-212, action_selectors.py:6-37).
 * log_table[N] = ln(N + 1) as numpy computes it, N < n_log.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pomcp_refexact pomcp_refexact;
int pomcp_refexact_create(int device, int rows, int cols, const int8_t *cell, int start_cell, const uint64_t *thr53,
                          const double rewards[4], uint32_t actual_mask, const uint32_t *bag_rocks, int n_bag,
                          int max_depth, int max_particle_count, double ucb_c, double discount, uint64_t seed,
                          uint32_t stream_id, const double *log_table, int n_log, int64_t cap_nodes,
                          int64_t cap_particles, pomcp_refexact **out);
void pomcp_refexact_destroy(pomcp_refexact *h);
const char *pomcp_refexact_last_error(const pomcp_refexact *h);
/* POMCP.select_eps_greedy_action (solvers/pomcp.py:69-78) */
int pomcp_refexact_search(pomcp_refexact *h, int n_sims, int32_t *action, int64_t *n, double *total, double *mean,
                          uint32_t *legal, int64_t *N, uint64_t *draws);
/* solver.belief_tree_index.sample_particle() (agent.py:157) */
int pomcp_refexact_sample_root_particle(pomcp_refexact *h, int32_t *particle);
/* model.generate_step(state, action) on the true state (agent.py:174): out = obs_good, terminal, legal, next cell,
 * next rocks, next particle */
int pomcp_refexact_env_step(pomcp_refexact *h, int particle, int action, int64_t out[6], double *reward, uint64_t *draws);
/* BeliefTreeSolver.update (belief_tree_solver.py:154-212) */
int pomcp_refexact_update(pomcp_refexact *h, int action, int obs_good, uint32_t sampled_after, int32_t *disable_tree,
                          uint64_t *draws);

#ifdef __cplusplus
}
#endif
#endif
