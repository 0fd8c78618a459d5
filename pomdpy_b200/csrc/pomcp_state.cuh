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
#define CHILD_ID(c) ((int)((c) & 0x7FFFFFFFu) - 1)
#define SEARCH_THREADS 256
#ifndef SEARCH_MIN_BLOCKS
#define SEARCH_MIN_BLOCKS 4
#endif
#define ADV_THREADS 1024
#define ADV_TRIES 4
#ifndef ARR_HT
#define ARR_HT 1024            // per-block backup table (shared memory): slots of {edge, visits, return sum} of shared edges
#endif
#define ARR_PROBES 8
#define BACKUP_TABLE_DEPTH 3   // backup: unshared-in-the-warp edges down to this path depth still go through the block table
#define BACKUP_TABLE_PROBES 2
#define PLAN_MASK 0            // kinds of a node's action plan (warp_plan)
#define PLAN_PULLS 1
#define PLAN_THRESHOLDS 2
#ifndef SMALL_COOP_MAX
#define SMALL_COOP_MAX 6        // at most this many fringe plans per warp and level: computed by the warp, else per thread
#endif
#define PLAN_LOCKED 0xFFFFFFFFu // stamp of a plan record while one warp computes it
#define SEQ_ALLOC_MAX 16u      // expected arrivals < 16.5: virtual-pull sequence / arg-max set; more: water-filling
// Per-block cache of the hot plans (the upper tree levels, where thousands of simulations of a generation pass) in
// shared memory: PC_SLOTS direct-mapped entries of PC_WORDS 32-bit words
//   [0] key = node id + 1 (0: free)  [1] ready  [2] expected arrivals (float bits)  [4..7] head  [8..11] data
//   [12..43] the 32 cumulative thresholds
#ifndef PC_BITS
#define PC_BITS 6
#endif
#define PC_SLOTS (1 << PC_BITS)
#define PC_WORDS 44
#define SQ_CAP 192                // per-block queue of the hot nodes whose plans are computed ahead of the simulations
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
    int32_t cursor[4];                 // work counters of a generation: [2 * parity] descent + rollout, [2 * parity + 1] backup
    int32_t stop;
    uint32_t barrier;                  // arrival counter of grid_barrier (zeroed before each search)
    int32_t check_failed, check_pad;
    int32_t pad_c[2];
    // root edge statistics as of the end of the last search / advance / plant (one copy serves pomcp_root_stats)
    long long res_n[AS], res_q[AS];
    uint32_t res_legal; int32_t res_valid;
    // root edge statistics summed over the ranks of a root-parallel search (fused merge in the search kernel's tail)
    long long mrg_n[AS], mrg_q[AS];
    unsigned long long mrg_seq;        // sequence number of the merge that filled mrg_*; 0 = the exchange timed out
    unsigned long long gstep;          // generation stamp of the plan records, never reused
    long long sims_done;               // simulations completed by the last search
    long long counters[16];
    long long phase_ns[8];            // device time per phase of the last search (thread 0's globaltimer)
    // results of the last advance
    int32_t adv_status, adv_accepted, adv_root_cell, adv_pad;
    long long adv_tries;
    // ---- profiling aid, not copied by the hot calls ----
    long long steplog[128][8];        // last search: row g < 120 = ns of generation g's [start, descent + rollout, backup, flush]; row 126 = totals
};
#define DEVSTATE_HEAD (offsetof(DevState, steplog))

struct Tree {                          // structure-of-arrays belief-tree arena
    int32_t *n;                        // [cap][AS]   visit_count of edge (node, action)
    long long *qfix;                   // [cap][AS]   sum of q in 2^-24 fixed point
    uint32_t *child;                   // [cap][AS][ZS] child id + 1 keyed (action, observation); 0 = none, ~0 = being
                                       //              created.  ZS = 2 for RockSample (obs_good)
                                       // hashed table (ckey != null): [cmask + 1] values, slot found through ckey
    unsigned long long *ckey;          // hashed child table (tabular models with |Z| > 32): [cmask + 1] keys
                                       //              1 + ((node * AS + action) << 11 | observation); 0 = free slot
    uint32_t cmask, cpad;              //              slots - 1 (a power of two >= 2 * cap: load factor <= 1/2)
    uint32_t *legal;                   // [cap]       legal-action mask (fixed at creation)
    int32_t *cell;                     // [cap]       tag of the node: robot cell (RockSample), 0 (tabular)
    uint4 *plan;                       // [cap][2]    action plan of the node for the generation `stamp`:
                                       //             {kind << 16 | pulls << 8 | last action, stamp, coarse thresholds 4, 5},
                                       //             {action mask (kind 0) | <= 16 pull bytes (kind 1) | coarse thresholds 0..3 (kind 2)}
    uint32_t *thr;                     // [cap][32]   kind 2: the 32 cumulative action thresholds
    double *chance;                    // [cap][24]   --preferred_actions: chance_good per rock (null otherwise)
    int16_t *good;                     // [cap][24]   --preferred_actions: goodness number per rock
    const double *prob;                // [cells][n_rocks] sensor correctness probability
};

// Hashed child table.  The reference's observation map is a dict per action node that grows with the observations
// seen (discrete_pomdp/discrete_observation_mapping.py:12-47), so |Z| is unbounded; a direct-indexed [node][action][Z]
// table is 128 * Z bytes per node, while a node has exactly one parent edge.  For |Z| > 32 the table is open
// addressing over (node, action, observation) with linear probing: the KEY array fixes which slot an edge owns
// (insert-only, CAS 0 -> key: every thread that asks for the same edge gets the same slot, whoever came first), the
// VALUE array keeps the protocol of the direct table (0 -> CHILD_LOCKED -> id + 1).  Slot numbers depend on insertion
// order and never leave the table code; ids, statistics and results do not.
__device__ __forceinline__ unsigned long long child_key(int X, int a, int obs)
{ return (((((unsigned long long)(unsigned)X * AS) + (unsigned)a) << 11) | (unsigned)obs) + 1ull; }

template <bool CREATE>
__device__ __forceinline__ long long child_slot_hashed(const Tree &t, int X, int a, int obs)
{
    const unsigned long long key = child_key(X, a, obs);
    uint32_t s = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & t.cmask;
    for (uint32_t probe = 0; probe <= t.cmask; ++probe, s = (s + 1u) & t.cmask) {
        unsigned long long k = __ldcg(&t.ckey[s]);                  // (keys land in this very phase: never from L1)
        if (k == key) return (long long)s;
        if (k == 0ull) {
            if (!CREATE) return -1;
            k = atomicCAS(&t.ckey[s], 0ull, key);
            if (k == 0ull || k == key) return (long long)s;
        }
    }
    return -1;                                                       // table full (cannot happen below node_capacity)
}

struct SimBuf {                        // what a simulation hands from the descent to the backup, SoA over the generation
    uint8_t *flag;                     // 2 ended in the tree (terminal / horizon), 3 ended with a rollout
    uint8_t *depth;                    // path entries recorded
    int32_t *path_node;                // [dcap][wmax]
    uint16_t *path_ar;                 // [dcap][wmax] action | obs_good << 5 | reward id << 8   (tabular: action | obs << 5)
    double *path_rew;                  // [dcap][wmax] tabular models: the step reward (null for RockSample)
    double *ro_sum;                    // [wmax] discounted reward sum of the rollout
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
    unsigned long long merge_timeout_ns;
    int32_t pref_rollout, pad_pr;      // --preferred_actions with preferred rollouts (RockSample)
    int32_t pc_mult, pad_pm;           // tuning (no effect on results): a node is hot from k >= pc_mult * blocks
};

// Exchange buffer of one rank (int64 words): slots[2][POMCP_MAX_PEERS][64] then flags[2][POMCP_MAX_PEERS].  Set
// `seq & 1` is used by merge number seq, so a peer that is already one search ahead writes the other set.
#define MERGE_SLOT_WORDS (2 * POMCP_MAX_PEERS * 64)
#define MERGE_WORDS (MERGE_SLOT_WORDS + 2 * POMCP_MAX_PEERS)
