// pomcp_state.cuh -- device-resident state of the B200 POMCP planner: the state block, the structure-of-arrays tree
// arena, the per-simulation buffers and the kernel parameter blocks, plus the tuning constants.
#pragma once
#include <cstddef>
#include <stdint.h>

#include "../../include/pomcp_b200.h"
#include "pomcp_device.cuh"

#define AS POMCP_AS

// Bounds checks of the device code (compute-sanitizer is not available on the GPU pool): a -DPOMCP_CHECKS build
// records the largest failing check code in DevState.check_failed; tests/test_gpu_debug_checks.py runs it.
#ifdef POMCP_CHECKS
#define DCHECK(S_, cond, code) do { if (!(cond)) atomicMax(&(S_)->check_failed, (code)); } while (0)
#else
#define DCHECK(S_, cond, code) do { } while (0)
#endif
// Section timers of the expand phase (-DPOMCP_TSEC diagnostic build only): per-thread clock64 deltas, summed over the
// grid into steplog[127][0..7] at the end of the search.
#ifdef POMCP_TSEC
#define TSEC_DECL long long tsec_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tsec_prev_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tsec_t_ = 0, tsec_bk_ = 0
#define TSEC_START() (tsec_t_ = clock64())
#define TSEC(k) do { const long long n_ = clock64(); tsec_[k] += n_ - tsec_t_; tsec_t_ = n_; } while (0)
#else
#define TSEC_DECL
#define TSEC_START() do { } while (0)
#define TSEC(k) do { } while (0)
#endif
#define CHILD_LOCKED 0xFFFFFFFFu
#define CHILD_VISITED 0x80000000u   // set on a child slot once a backup has passed through the child (N > 0)
#define CHILD_ID(c) ((int)((c) & 0x7FFFFFFFu) - 1)
#define SEARCH_THREADS 256
#ifndef SEARCH_MIN_BLOCKS
#define SEARCH_MIN_BLOCKS 4
#endif
#define ADV_THREADS 1024
#define ADV_TRIES 4
#define ROLL_SLICE 8            // rollout steps per slice while waiting at a barrier (even: Philox block pairs)
#ifndef ARR_HT
#define ARR_HT 1024            // per-block arrival table (shared memory): slots of {node id, arrivals in this level step}
#endif
#define ARR_PROBES 8
#define PLAN_MASK 0            // kinds of a node's action plan (warp_plan)
#define PLAN_PULLS 1
#define PLAN_THRESHOLDS 2
#define SEQ_ALLOC_MAX 16u      // arrivals <= this: exact virtual-pull allocation; more: water-filling
#ifndef BISECT_ITERS
#define BISECT_ITERS 14
#endif

// ---------------------------------------------------------------------------------------------------------
// Device-resident state (one per handle).  Kernels read and write it without host round trips.
// ---------------------------------------------------------------------------------------------------------
struct DevState {
    int32_t root;
    int32_t node_count;
    uint32_t actual, sampled;
    int32_t n_bag, bag_sel;
    int32_t wl_count[4];
    int32_t stop;
    uint32_t barrier;                  // arrival counter of grid_barrier (zeroed before each search)
    int32_t check_failed, check_pad;
    int32_t ro_cursor[2];              // work counters of the final rollout + backup phase (alternating per generation)
    // root edge statistics as of the end of the last search / advance / plant (one copy serves pomcp_root_stats)
    long long res_n[AS], res_q[AS];
    uint32_t res_legal; int32_t res_valid;
    // root edge statistics summed over the ranks of a root-parallel search (fused merge in the search kernel's tail)
    long long mrg_n[AS], mrg_q[AS];
    unsigned long long mrg_seq;        // sequence number of the merge that filled mrg_*; 0 = the exchange timed out
    unsigned long long gstep;          // global level-step stamp, never reused
    long long sims_done;               // simulations completed by the last search
    long long counters[16];
    long long phase_ns[8];            // device time per phase of the last search (thread 0's globaltimer)
    // results of the last advance
    int32_t adv_status, adv_accepted, adv_root_cell, adv_pad;
    long long adv_tries;
    // ---- profiling aid, not copied by the hot calls ----
    long long steplog[128][8];        // last search, per level step: ns of [wait, allocate, wait, expand], worklist
};
#define DEVSTATE_HEAD (offsetof(DevState, steplog))

struct Tree {                          // structure-of-arrays belief-tree arena
    int32_t *n;                        // [cap][AS]   visit_count of edge (node, action)
    long long *qfix;                   // [cap][AS]   sum of q in 2^-24 fixed point
    uint32_t *child;                   // [cap][AS][ZS] child id + 1 keyed (action, observation), bit 31 = has statistics;
                                       //              0 = none.  ZS = 2 for RockSample (obs_good)
    uint32_t *legal;                   // [cap]       legal-action mask (fixed at creation)
    int32_t *cell;                     // [cap]       tag of the node: robot cell (RockSample), 0 (tabular)
    unsigned long long *arrive;        // [cap]       (stamp << 32) | arrivals in the current level step
    uint4 *plan;                       // [cap][2]    action plan of the node for the current level step:
                                       //             {kind << 16 | arrivals << 8 | last action, worklist slot, -, -},
                                       //             {action mask (kind 0) | <= 16 pull bytes (kind 1)}; kind 2: thresholds in wl_thr[slot]
    double *chance;                    // [cap][24]   --preferred_actions: chance_good per rock (null otherwise)
    int16_t *good;                     // [cap][24]   --preferred_actions: goodness number per rock
    const double *prob;                // [cells][n_rocks] sensor correctness probability
};

struct SimBuf {                        // per simulation-in-flight scratch, SoA over the generation
    uint32_t *state;                   // packed state cell | rocks << 8
    int32_t *node;                     // current node, or child-slot index when `pending`
    uint8_t *flag;                     // status: 0 descending, 1 leaf (rollout pending / in progress), 2 done without rollout,
                                       //         3 rollout finished
    uint8_t *depth;                    // path entries recorded
    int32_t *path_node;                // [dcap][wmax]
    uint16_t *path_ar;                 // [dcap][wmax] action | obs_good << 5 | reward id << 8   (tabular: action | obs << 5)
    double *path_rew;                  // [dcap][wmax] tabular models: the step reward (null for RockSample)
    int32_t *wl_node;                  // [2][wmax] worklist of nodes with arrivals in this level step (set = step parity)
    uint32_t *wl_thr;                  // [wmax][32] cumulative action thresholds of worklist nodes with a kind-2 plan
    double *ro_sum;                    // [wmax] discounted reward sum of a rollout in progress
    uint8_t *ro_t;                     // [wmax] rollout steps done so far (rollouts advance in slices)
};

struct SearchParams {
    DevState *S;
    const ModelHeader *model;
    const uint64_t *thr53;
    TabHeader tab;                     // tabular models (KIND 1); unused by RockSample
    TabTables tt;
    Tree t;
    SimBuf b;
    const uint32_t *bag[2];
    long long n_sims;
    uint32_t sim_offset, move, key0, key1;
    uint32_t ks[20];                   // Philox key schedule: ks[2r], ks[2r+1] = (key0 + r*0x9E3779B9, key1 + r*0xBB67AE85)
    int32_t w0, growth, wmax, max_depth, dcap, n_thr;
    long long node_cap;
    double ucb_c, discount;
    unsigned long long timeout_ns;     // 0 = no wall-clock guard; measured from kernel start
    // root-parallel merge over NVLink peer memory (world <= 8 GPUs of one box): every rank's exchange buffer, mapped
    // into this process; layout MERGE_* below
    unsigned long long peer_buf[POMCP_MAX_PEERS];
    int32_t world, rank;
    unsigned long long merge_seq;
    int32_t pref_rollout, pad_pr;      // --preferred_actions with preferred rollouts (RockSample)
};

// Exchange buffer of one rank (int64 words): slots[2][POMCP_MAX_PEERS][64] then flags[2][POMCP_MAX_PEERS].  Set
// `seq & 1` is used by merge number seq, so a peer that is already one search ahead writes the other set.
#define MERGE_SLOT_WORDS (2 * POMCP_MAX_PEERS * 64)
#define MERGE_WORDS (MERGE_SLOT_WORDS + 2 * POMCP_MAX_PEERS)
