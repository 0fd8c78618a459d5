// pomcp_kernels.cu -- B200 (sm_100a) POMCP planner: device tree, search / belief-update kernels and the C ABI.
//
// Replaces what happens inside POMCP.select_eps_greedy_action() and BeliefTreeSolver.update() of the reference
// (pomdpy/solvers/pomcp.py:69-137, pomdpy/solvers/belief_tree_solver.py:42-56,123-212,
// pomdpy/action_selection/action_selectors.py:6-37, pomdpy/pomdp/model.py:221-252).  See DESIGN.md.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pomcp_b200.h"
#include "pomcp_device.cuh"

#include "pomcp_state.cuh"
#include "pomcp_models.cuh"
#include "pomcp_plan.cuh"

// ---------------------------------------------------------------------------------------------------------
// K1: the search.  One persistent cooperative kernel per move; simulations advance in generations.  Within a
// generation the tree statistics are read-only, and the action plan of a node depends only on its statistics and on
// the number of simulations EXPECTED there (pomcp_plan.cuh) -- so a generation is two phases and two grid barriers:
//     [descend + rollout]  warps claim 32 simulations at a time; every simulation draws a root particle, walks down
//                          the tree on its own (plan lookup: block cache in shared memory -> node record in L2 ->
//                          computed by this warp), creates / finds the child and, at the first node without
//                          statistics, takes a uniform action and rolls out
//     [backup]             the recorded paths update the statistics (commutative integer atomics, folded per warp and
//                          staged per block in shared memory)
// ---------------------------------------------------------------------------------------------------------

// 64-bit add to a shared-memory word as two native 32-bit atomic adds (low word, then high word plus the carry): a
// 64-bit atomicAdd on shared memory is a compare-and-swap loop, which spins when the warps of a block update the same
// hot edge.  The sums are only read after a block barrier, when every add has landed.
__device__ __forceinline__ void smem_add64(long long *dst, long long v)
{

    uint32_t *w = reinterpret_cast<uint32_t *>(dst);
    const uint32_t lo = (uint32_t)(unsigned long long)v, hi = (uint32_t)((unsigned long long)v >> 32);
    const uint32_t old = atomicAdd(&w[0], lo);
    atomicAdd(&w[1], hi + (uint32_t)(old + lo < old));
}

__device__ __forceinline__ int row_sum(const int32_t *nrow)
{
    int N = 0;
#pragma unroll
    for (int q = 0; q < AS / 4; ++q) { const int4 v = reinterpret_cast<const int4 *>(nrow)[q]; N += v.x + v.y + v.z + v.w; }
    return N;
}

template <int KIND>
__device__ __forceinline__ void search_body(const SearchParams &p, unsigned char *smem)
{
    using MD = Mdl<KIND>;
    const unsigned FULL = 0xFFFFFFFFu;
    unsigned int bar_epoch = 0;
#define GRID_SYNC() grid_barrier(&p.S->barrier, gridDim.x, bar_epoch)
    typename MD::Ctx mc;
    size_t off = MD::stage(smem, p.model, p.thr53, p.n_thr, p.tab, p.tt, mc);
    double *disc = reinterpret_cast<double *>(smem + off);          // discount^t, t < max_depth
    off += ((size_t)p.max_depth * 8 + 15) & ~size_t(15);
    const int A = MD::n_actions(mc, p.model);
    const int ZS = MD::zs(mc);                                       // observation stride of the child table
    const bool stage_d1 = MD::stage_d1(mc);                          // depth-1 edges staged per block too
    const int n_acc = stage_d1 ? (1 + ZS * A) * A : A;               // root (+ depth-1) edges staged per block
    long long *acc_q = reinterpret_cast<long long *>(smem + off);
    off += (size_t)n_acc * 8;
    int *acc_n = reinterpret_cast<int *>(smem + off);
    off += (size_t)n_acc * 4;
    off = (off + 15) & ~size_t(15);                                  // the tables below hold 64-bit words
    uint32_t *ht_key = reinterpret_cast<uint32_t *>(smem + off);    // backup: visits and return sums of shared deeper edges
    int *ht_cnt = reinterpret_cast<int *>(smem + off + ARR_HT * 4);
    long long *ht_q = reinterpret_cast<long long *>(smem + off + ARR_HT * 8);
    off += (size_t)ARR_HT * 16;
    uint32_t *pc = reinterpret_cast<uint32_t *>(smem + off);        // plan cache of the hot nodes (PC_SLOTS x PC_WORDS)
    off += (size_t)PC_SLOTS * PC_WORDS * 4;
    int *sq = reinterpret_cast<int *>(smem + off);                   // hot-node queue: head, tail, pending, pad, nodes[], k[]
    int *sq_node = sq + 4;
    float *sq_k = reinterpret_cast<float *>(sq + 4 + SQ_CAP);
    for (int i = threadIdx.x; i < ARR_HT; i += blockDim.x) { ht_key[i] = 0u; ht_cnt[i] = 0; ht_q[i] = 0; }
    for (int i = threadIdx.x; i < PC_SLOTS; i += blockDim.x) { pc[i * PC_WORDS] = 0u; pc[i * PC_WORDS + 1] = 0u; }
    if (threadIdx.x == 0) { double d = 1.0; for (int t = 0; t < p.max_depth; ++t) { disc[t] = d; d *= p.discount; } }
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) { acc_q[i] = 0; acc_n[i] = 0; }
    __syncthreads();

    DevState *S = p.S;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int root = S->root;
    if constexpr (KIND == 0) { mc.actual = S->actual; mc.sampled = S->sampled; }
    const uint32_t *bag = p.bag[S->bag_sel];
    const int n_bag = S->n_bag;
    unsigned long long gstep = S->gstep;
    // bookkeeping that lives for the whole kernel is kept out of the registers of the hot loops: the phase clocks of
    // thread 0 in shared memory, the per-thread work counters 32-bit (a thread runs far fewer than 2^31 steps a search)
    __shared__ unsigned long long tclk[6];                           // [0..3] phase totals, [4] last mark, [5] kernel start
    if (threadIdx.x == 0) { tclk[0] = tclk[1] = tclk[2] = tclk[3] = 0ull; tclk[4] = tclk[5] = globaltimer_ns(); }
    int c_levels = 0, c_rollout = 0, c_created = 0, c_alloc = 0;
    long long done = 0; int gens = 0;
#define PHASE_MARK(k) do { if (gtid == 0) { const unsigned long long now_ = globaltimer_ns(), d_ = now_ - tclk[4]; tclk[k] += d_; \
                                             if (gens <= 120) S->steplog[gens - 1][k] = (long long)d_; tclk[4] = now_; } } while (0)
    TSEC_DECL;
    // expected arrivals that earn a node a slot in the block caches (batch-UCB allocations only: >= 16.5)
    const float pc_min_k = fmaxf((float)p.pc_mult * (float)gridDim.x, 16.5f);

    if (gtid == 0) S->phase_ns[4] = 0;
    // (no grid barrier here: what a block reads from S above -- root, world, belief, gstep -- is written again only in
    // the kernel's tail, behind the barriers of the generations; a block that starts late misses nothing, the
    // statistics stay read-only until the first generation's barrier, which waits for it)

    while (done < p.n_sims) {
        // Width of the generation: `growth` times the visits the root has seen, within [w0, wmax] -- from a fresh tree
        // w0, growth*w0, ... as a geometric schedule would, and inside an episode, where the re-rooted tree arrives
        // with statistics, wide generations right away (statistics are as stale relative to their mass either way)
        const int Nroot = row_sum(p.t.n + (size_t)root * AS);
        long long W = (long long)p.growth * Nroot;
        W = W < p.w0 ? p.w0 : W > p.wmax ? p.wmax : W;
        const int w = (int)((p.n_sims - done) < W ? (p.n_sims - done) : W);
        ++gens; ++gstep;
        const uint32_t stamp = (uint32_t)gstep;                      // of the plans computed in this generation
        const uint32_t root_mask = p.t.legal[root];
        if (gtid == 0) { S->cursor[2 * ((gens + 1) & 1)] = 0; S->cursor[2 * ((gens + 1) & 1) + 1] = 0; }   // the next generation's
        // ---- the hot plans, ahead of the simulations.  The expected arrivals of a child follow from its parent's plan
        // alone (k * P_plan(a) * N(child) / n(X, a)): no simulation has to get there first.  So every block walks the
        // hot part of the tree (k >= pc_min_k: the upper levels, and the whole chain a re-rooted tree starts with) top
        // down, its warps taking nodes from a queue, and fills its plan cache; the plan of a level then costs one
        // allocation, not an allocation plus the step, child lookup and polling of the warps that wait for it.  What
        // does not fit (queue, cache set) is left to the simulations (computed on arrival, as every cooler node is).
        if (threadIdx.x == 0) {
            sq[0] = 0; sq[1] = Nroot > 0 ? 1 : 0; sq[2] = sq[1]; sq[3] = sq[1];     // head, tail, pending, reserved
            sq_node[0] = root; sq_k[0] = (float)w;
        }
        __syncthreads();
        if ((float)w >= pc_min_k) for (;;) {
            int idx = -1;
            if (lane == 0) {
                for (;;) {
                    const int hd = *((volatile int *)&sq[0]);
                    if (hd >= *((volatile int *)&sq[1])) { idx = *((volatile int *)&sq[2]) == 0 ? -2 : -1; break; }
                    if (atomicCAS(&sq[0], hd, hd + 1) == hd) { idx = hd; break; }
                }
            }
            idx = __shfl_sync(FULL, idx, 0);
            if (idx == -2) break;
            if (idx < 0) { __nanosleep(64); continue; }
            const int X = *((volatile int *)&sq_node[idx]);
            const float kx = *((volatile float *)&sq_k[idx]);
#ifdef POMCP_TSEC
            const long long sc_t0 = clock64();
#endif
            // a slot of the block cache (2-way set): nobody else is filling X, every node is queued once
            uint32_t *e = nullptr;
            if (lane == 0) {
                uint32_t *pset = pc + (((uint32_t)X * 2654435761u) >> (33 - PC_BITS)) * (2 * PC_WORDS);
                if (atomicCAS(pset, 0u, (uint32_t)X + 1u) == 0u) e = pset;
                else if (atomicCAS(pset + PC_WORDS, 0u, (uint32_t)X + 1u) == 0u) e = pset + PC_WORDS;
            }
            const int eoff = __shfl_sync(FULL, e ? (int)(e - pc) : -1, 0);
            const PlanRec r = warp_allocation(p, X, kx, lane, A, stamp);
#ifdef POMCP_TSEC
            const long long sc_t1 = clock64();
#endif
            if (lane == 0) ++c_alloc;
            if (eoff >= 0) {
                e = pc + eoff;
                e[12 + lane] = r.th;
                if (lane == 0) {
                    e[2] = __float_as_uint(kx);
                    *reinterpret_cast<uint4 *>(e + 4) = r.head; *reinterpret_cast<uint4 *>(e + 8) = r.data;
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *((volatile uint32_t *)(e + 1)) = 1u;
            }
            // the hot children: lane = action, same arithmetic as a simulation that descends (kpre = k * p, then * N / n)
            const uint32_t th_prev = __shfl_up_sync(FULL, r.th, 1);
            const float pa = (float)(r.th - (lane == 0 ? 0u : th_prev)) * (1.0f / 4294967296.0f);
            const float kpre = kx * pa;
            const int na = lane < A ? p.t.n[(size_t)X * AS + lane] : 0;
            // queue the hot children a round of lanes found: reserve, write, publish in reservation order
            const auto enqueue = [&](int child, float kc) {
                const unsigned hot = __ballot_sync(FULL, child >= 0);
                if (hot) {
                    const int m = __popc(hot);
                    int base_slot = 0;
                    if (lane == 0) { atomicAdd(&sq[2], m); base_slot = atomicAdd(&sq[3], m); }
                    base_slot = __shfl_sync(FULL, base_slot, 0);
                    const int slot = base_slot + __popc(hot & ((1u << lane) - 1u));
                    if (child >= 0 && slot < SQ_CAP) { sq_node[slot] = child; sq_k[slot] = kc; }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) {
                        const int n_ok = max(0, min(m, SQ_CAP - base_slot));      // the rest does not fit: left to the simulations
                        if (n_ok < m) atomicSub(&sq[2], m - n_ok);
                        if (n_ok > 0) {
                            while (*((volatile int *)&sq[1]) != base_slot) { }     // earlier reservations publish first
                            *((volatile int *)&sq[1]) = base_slot + n_ok;
                        }
                    }
                }
            };
            bool hashed = false;
            if constexpr (KIND == 1) hashed = p.t.ckey != nullptr;
            if (!hashed) {
                for (int o = 0; o < ZS; ++o) {
                    int child = -1; float kc = 0.0f;
                    if (lane < A && na > 0 && kpre >= pc_min_k) {
                        const uint32_t c = p.t.child[((size_t)X * AS + lane) * ZS + o];
                        if (c != 0u && c != CHILD_LOCKED) {
                            const int id = CHILD_ID(c);
                            const int N = row_sum(p.t.n + (size_t)id * AS);
                            kc = kpre * ((float)N / (float)na);
                            if (N > 0 && kc >= pc_min_k) child = id;
                        }
                    }
                    enqueue(child, kc);
                }
            } else {
                // hashed child table: only a few actions carry enough simulations; for each of them the lanes probe
                // 32 observations at a time (which children end up in the cache changes no result, see above)
                unsigned hot_a = __ballot_sync(FULL, lane < A && na > 0 && kpre >= pc_min_k);
                while (hot_a) {
                    const int a = __ffs((int)hot_a) - 1;
                    hot_a &= hot_a - 1;
                    const float kpre_a = __shfl_sync(FULL, kpre, a);
                    const int na_a = __shfl_sync(FULL, na, a);
                    for (int o0 = 0; o0 < MD::n_obs(mc); o0 += 32) {
                        int child = -1; float kc = 0.0f;
                        const long long cs = o0 + lane < MD::n_obs(mc) ? child_slot_hashed<false>(p.t, X, a, o0 + lane) : -1;
                        if (cs >= 0) {
                            const uint32_t c = __ldcg(&p.t.child[cs]);
                            if (c != 0u && c != CHILD_LOCKED) {
                                const int id = CHILD_ID(c);
                                const int N = row_sum(p.t.n + (size_t)id * AS);
                                kc = kpre_a * ((float)N / (float)na_a);
                                if (N > 0 && kc >= pc_min_k) child = id;
                            }
                        }
                        enqueue(child, kc);
                    }
                }
            }
            __threadfence_block();
            if (lane == 0) atomicSub(&sq[2], 1);
#ifdef POMCP_TSEC
            if (blockIdx.x == 0 && lane == 0 && gens == 1 && idx < 100) {   // scout log of block 0, first generation
                S->steplog[idx][4] = sc_t1 - sc_t0; S->steplog[idx][5] = clock64() - sc_t1;
                S->steplog[idx][6] = 100 + (threadIdx.x >> 5); S->steplog[idx][7] = (long long)kx;
            }
#endif
        }
        PHASE_MARK(0);
        // ---- descent (solvers/pomcp.py:87-137) + rollout (belief_tree_solver.py:123-152, from the post-action state) ----
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&S->cursor[2 * (gens & 1)], 32);
            base = __shfl_sync(FULL, base, 0);
            if (base >= w) break;
            const int i = base + lane;
            const uint32_t sim = p.sim_offset + (uint32_t)(done + i);
            int status = i < w ? 0 : 4;                              // 0 descending, 1 leaf: rollout pending, 2 ended in the tree, 4 no simulation
            uint32_t st = 0u; int X = root, d = 0, na = Nroot;       // X: node id, or -1 - child slot of a node being created
            float kpre = (float)w;                                   // expected arrivals at X = kpre * N(X) / na
            float kx = -1.0f;                                        // ... once computed (< 0: not yet)
            // A simulation at a node WITHOUT statistics (N = 0; "virgin"): every legal action is untried, so UCB1 is a
            // uniform draw (action_selectors.py:6-37 with +inf bonuses), the node cannot be expanded
            // (solvers/pomcp.py:107) and the simulation ends in a rollout.
            bool virgin = Nroot == 0; uint32_t vmask = root_mask;
            bool filler = false;                                     // this lane computes X's plan for the whole block
            int last_ref = root; uint32_t last_ar = 0u;              // the last tree step (preferred-action rollouts)
            TSEC_START();
            if (status == 0) {                                       // root particle (belief_node.py:47-48)
                const Philox4 r = philox4x32_10_ks(0u, sim, p.move, DOM_SIM, p.ks);
                st = bag[randbelow32(r.x, (uint32_t)n_bag)];
            }
#ifdef POMCP_TSEC
            int it_log = 0; long long it_t0 = clock64();
            const bool it_on = false;                               // (iteration log of a chunk: off, the scout log uses the columns)
#endif
            while (__any_sync(FULL, status == 0)) {                  // one tree level per iteration, for the lanes whose
                bool act = status == 0;                              // plan is at hand; the others retry (nobody blocks)
                if (act && d >= p.max_depth) { status = 2; act = false; }                  // :100
                // -- the plan of node X: block cache (shared memory), node record (L2), or computed here
                // the block cache is 2-way set associative: entry = way 0 or 1 of the set the node id hashes to
                uint32_t *pset = pc + (((uint32_t)X * 2654435761u) >> (33 - PC_BITS)) * (2 * PC_WORDS);
                uint32_t *pce = nullptr;
                if (status == 0 && !virgin) {
                    if (*((volatile uint32_t *)pset) == (uint32_t)X + 1u) pce = pset;
                    else if (*((volatile uint32_t *)(pset + PC_WORDS)) == (uint32_t)X + 1u) pce = pset + PC_WORDS;
                }
                bool have = false, cand = false, self = false, small = false, won = false, progressed = false;
                bool wait_block = false, wait_grid = false;          // the plan is being computed by another warp
                uint4 head = make_uint4(0u, 0u, 0u, 0u), data = head;
                const uint32_t *sthr = nullptr;
                if (act && !virgin) {
                    if (!filler && pce != nullptr) {
                        if (*((volatile uint32_t *)(pce + 1)) != 0u) {   // cached by this block
                            __threadfence_block();
                            kx = __uint_as_float(pce[2]);
                            head = *reinterpret_cast<const uint4 *>(pce + 4); data = *reinterpret_cast<const uint4 *>(pce + 8);
                            sthr = pce + 12;
                            have = true;
                        } else wait_block = true;                    // being computed by another warp of the block
                    } else {
                        if (kx < 0.0f) {
                            // a node met for the first time by this simulation: besides its visit counts it will need
                            // the return sums, the action set and maybe the plan record -- in a big tree those are
                            // DRAM misses; ask for them now instead of one after the other
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.t.qfix + (size_t)X * AS));
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.t.qfix + (size_t)X * AS + 16));
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.t.legal + X));
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.t.plan + (size_t)X * 2));
                            const int N = row_sum(p.t.n + (size_t)X * AS);
                            if (N == 0) {
                                virgin = true;
                                vmask = MD::tag_legal(mc, MD::state_tag(st));
                                // published in this very phase, maybe by another SM: with preferred actions its action
                                // set is read from L2, never from a possibly stale L1 line
                                if (KIND == 0 && p.t.chance) vmask = __ldcg(&p.t.legal[X]);
                            } else {
                                kx = kpre * ((float)N / (float)na);
                                if (kx >= pc_min_k) {                // hot: every block keeps its own copy of the plan
                                    pce = pset;
                                    uint32_t prev = atomicCAS(pce, 0u, (uint32_t)X + 1u);
                                    if (prev != 0u && prev != (uint32_t)X + 1u) {        // way 0 holds another node
                                        pce = pset + PC_WORDS;
                                        prev = atomicCAS(pce, 0u, (uint32_t)X + 1u);
                                    }
                                    filler = prev == 0u;             // this warp computes it for the block (no waiting for
                                    wait_block = prev == (uint32_t)X + 1u;   // other SMs); another warp of the block is at it
                                    if (!filler && !wait_block) pce = nullptr;          // set full: use the node's record
                                }
                            }
                        }
                        if (filler) self = true;
                        else if (!virgin && !wait_block) {
                            if (kx >= 1.5f) {
                                head = __ldcg(&p.t.plan[(size_t)X * 2]); data = __ldcg(&p.t.plan[(size_t)X * 2 + 1]);
                                have = head.y == stamp;
                            }
                            if (!have && kx < 16.5f) {               // few expected arrivals: computed below, by this thread
                                small = true;                        // or by the warp (also if another thread is at it: no waiting)
                                won = kx >= 1.5f && head.y != PLAN_LOCKED &&
                                    atomicCAS(reinterpret_cast<uint32_t *>(&p.t.plan[(size_t)X * 2]) + 1, head.y, PLAN_LOCKED) == head.y;
                            } else if (!have) {
                                cand = head.y != PLAN_LOCKED;        // a batch-UCB allocation nobody has claimed
                                wait_grid = !cand;
                            }
                        }
                    }
                }
                // plans for few expected arrivals (the fringe of the tree).  One thread computes one in ~10^3 dependent
                // instructions; the warp (lane = action) needs a few hundred cycles but takes them one at a time: so
                // the warp does them if only a few lanes ask (a chain level with a stray sibling must not stall its warp
                // for microseconds), and every lane its own if many do (a bushy level).  Same arithmetic, same plans.
                {
                    unsigned sm_few = __ballot_sync(FULL, small);
                    const bool coop = __popc(sm_few) <= SMALL_COOP_MAX;
                    const int kq = (int)(kx + 0.5f);
                    PlanRec r;
                    if (coop) {
                        while (sm_few) {
                            const int src = __ffs((int)sm_few) - 1;
                            sm_few &= sm_few - 1;
                            const int kqs = __shfl_sync(FULL, kq, src);
                            const PlanRec rs = warp_plan_few(p, __shfl_sync(FULL, X, src), (uint32_t)(kqs < 1 ? 1 : kqs), lane, A, stamp);
                            if (lane == src) r = rs;
                        }
                    } else if (small) r = thread_plan_few(p, X, (uint32_t)(kq < 1 ? 1 : kq), A, stamp);
                    if (small) {
                        head = r.head; data = r.data; have = true;
                        if (won) {                                   // publish: data, then the head with the stamp
                            __stcg(&p.t.plan[(size_t)X * 2 + 1], data);
                            __threadfence();
                            __stcg(&p.t.plan[(size_t)X * 2], head);
                        }
                    }
                }
                // batch-UCB allocations (lane = action): the hot ones of this block go straight into its cache; of the
                // others the warp takes the first unclaimed one it can get per iteration -- the rest is left to the
                // warps that arrive meanwhile, or to the next iteration -- and publishes it in the node's record
                {
                    unsigned todo = __ballot_sync(FULL, self);
                    unsigned cm = __ballot_sync(FULL, cand);
                    while (cm) {
                        const int src = __ffs((int)cm) - 1;
                        cm &= cm - 1;
                        int got = 0;
                        if (lane == src)
                            got = atomicCAS(reinterpret_cast<uint32_t *>(&p.t.plan[(size_t)X * 2]) + 1, head.y, PLAN_LOCKED) == head.y;
                        if (__shfl_sync(FULL, got, src)) { todo |= 1u << src; break; }
                    }
                    while (todo) {
                        const int src = __ffs((int)todo) - 1;
                        todo &= todo - 1;
                        const int Xs = __shfl_sync(FULL, X, src);
                        const bool hot = __shfl_sync(FULL, (int)self, src) != 0;
                        const int eoff = __shfl_sync(FULL, pce ? (int)(pce - pc) : 0, src);
                        const PlanRec r = warp_allocation(p, Xs, __shfl_sync(FULL, kx, src), lane, A, stamp);
                        if (lane == 0) ++c_alloc;                    // a diagnostic: hot plans are computed once per block
                        if (hot) {
                            uint32_t *e = pc + eoff;
                            e[12 + lane] = r.th;
                            if (lane == src) {
                                e[2] = __float_as_uint(kx);
                                *reinterpret_cast<uint4 *>(e + 4) = r.head; *reinterpret_cast<uint4 *>(e + 8) = r.data;
                            }
                            __threadfence_block();
                            __syncwarp();
                            if (lane == src) { *((volatile uint32_t *)(e + 1)) = 1u; filler = false; sthr = e + 12; }
                        } else {                                     // thresholds and data first, then the head with the stamp
                            __threadfence();
                            __syncwarp();
                            if (lane == 0) {
                                __stcg(&p.t.plan[(size_t)Xs * 2 + 1], r.data);
                                __threadfence();
                                __stcg(&p.t.plan[(size_t)Xs * 2], r.head);
                            }
                        }
                        if (lane == src) { head = r.head; data = r.data; have = true; }
                    }
                }
#ifdef POMCP_TSEC
                const long long it_t1 = clock64();
                const int it_kind = have ? (sthr ? 1 : small ? 3 : 4) : virgin ? 2 : (wait_block ? 5 : wait_grid ? 6 : 7);
#endif
                TSEC(0);                                             // plan acquisition
                // -- choose, step the model, look up / create the child (solvers/pomcp.py:97-117)
                if (act && (have || virgin)) {
                    DCHECK(S, d < p.dcap && (virgin ? vmask != 0u : (X >= 0 && X < S->node_count)), 21);
                    const Philox4 r = philox4x32_10_ks((uint32_t)(d + 1), sim, p.move, DOM_SIM, p.ks);
                    float pa = 0.0f;
                    int a;
                    if (virgin) a = nth_bit(vmask, randbelow32(r.x, (uint32_t)__popc(vmask)));
                    else a = plan_action(head, data, sthr, p.t.thr + (size_t)X * 32, r.x, A, pa);
                    DCHECK(S, a >= 0 && a < A && (virgin || ((p.t.legal[X] >> a) & 1u)), 22);
                    DCHECK(S, MD::tag_ok(mc, MD::state_tag(st)) && (X < 0 || MD::state_tag(st) == (uint32_t)__ldcg(&p.t.cell[X])), 23);
                    const StepOut so = MD::step(mc, st, a, r.y, r.z);                       // :104
                    ++c_levels;
                    const uint16_t code = MD::path_code(a, so);
                    p.b.path_node[(size_t)d * p.wmax + i] = X;
                    p.b.path_ar[(size_t)d * p.wmax + i] = code;
                    if (KIND == 1) p.b.path_rew[(size_t)d * p.wmax + i] = so.reward;
                    last_ref = X; last_ar = code;
                    const uint32_t pcell = st & 0xFFu;
                    progressed = true;
                    d = d + 1;
                    st = so.next;
                    TSEC(1);                                         // choice + step + path stores
                    if (so.terminal) status = 2;
                    else if (virgin) status = 1;                     // leaf
                    else {
                        size_t cslot = ((size_t)X * AS + a) * ZS + so.obs;
                        bool have_slot = true;
                        if constexpr (KIND == 1) if (p.t.ckey) {                            // hashed table: the edge's slot
                            const long long cs = d < p.dcap ? child_slot_hashed<true>(p.t, X, a, so.obs)
                                                            : child_slot_hashed<false>(p.t, X, a, so.obs);
                            have_slot = cs >= 0; cslot = have_slot ? (size_t)cs : 0;
                        }
                        uint32_t cmask = MD::tag_legal(mc, so.tag); bool have_mask = false;   // the child's action set
                        uint32_t c = have_slot ? p.t.child[cslot] : 0u;                     // :106
                        if (c == 0 && d < p.dcap && have_slot) {                            // :107-108 (N(X) > 0 here)
                            // one CAS per distinct child slot among the lanes that got here together
                            const unsigned cm = __activemask();
                            const unsigned same = __match_any_sync(cm, (unsigned long long)cslot);
                            if (lane == __ffs((int)same) - 1) c = atomicCAS(&p.t.child[cslot], 0u, CHILD_LOCKED);
                            c = __shfl_sync(same, c, __ffs((int)same) - 1);
                            if (lane != __ffs((int)same) - 1 && c == 0) c = CHILD_LOCKED;   // the group leader creates it
                            if (c == 0) {                                                   // this thread creates it
                                const unsigned am = __activemask();                         // one arena bump per warp
                                const int rk = __popc(am & ((1u << lane) - 1u));
                                int id0 = 0;
                                if (rk == 0) id0 = atomicAdd(&S->node_count, __popc(am));
                                const int id = __shfl_sync(am, id0, __ffs((int)am) - 1) + rk;
                                DCHECK(S, id > 0 && (long long)id < p.node_cap, 24);
                                p.t.cell[id] = (int32_t)so.tag;
                                if constexpr (KIND == 0) if (p.t.chance)                    // create_child + smart action set
                                    cmask = child_smart_mask(mc.M.hdr, p.t.prob, p.t.chance + (size_t)X * POMCP_RS,
                                                             p.t.good + (size_t)X * POMCP_RS,
                                                             p.t.chance + (size_t)id * POMCP_RS,
                                                             p.t.good + (size_t)id * POMCP_RS, (int)pcell, a,
                                                             so.obs, (int)so.tag);
                                p.t.legal[id] = cmask;
                                // The id may be picked up by another SM in this very phase.  Without preferred actions
                                // such a reader derives the action set from its own state and reads only the (zero)
                                // statistics of the node; with them it reads legal[id] and the rock statistics: node
                                // fields before the id
                                if (KIND == 0 && p.t.chance) __threadfence();
                                __stcg(&p.t.child[cslot], (uint32_t)id + 1u);
                                ++c_created;
                                c = CHILD_LOCKED; have_mask = true;
                            }
                        }
                        if (c == 0) status = 1;                                             // depth cap: leaf (:119)
                        else if (c == CHILD_LOCKED) {                                       // created in this generation: no
                            if constexpr (KIND == 0) if (p.t.chance && !have_mask)          // statistics; id maybe not published yet
                                cmask = child_smart_mask(mc.M.hdr, p.t.prob, p.t.chance + (size_t)X * POMCP_RS,
                                                         p.t.good + (size_t)X * POMCP_RS, nullptr, nullptr,
                                                         (int)pcell, a, so.obs, (int)so.tag);
                            virgin = true; vmask = cmask; X = -1 - (int)cslot;              // next iteration: uniform action, leaf
                        } else {                                                            // descend
                            na = p.t.n[(size_t)X * AS + a];
                            kpre = kx * pa;
                            X = CHILD_ID(c);
                            kx = -1.0f;
                        }
                    }
                }
                TSEC(2);                                             // child lookup / creation / leaf
#ifdef POMCP_TSEC
                if (it_on && it_log < 100) {
                    const long long it_t2 = clock64();
                    S->steplog[it_log][4] = it_t1 - it_t0; S->steplog[it_log][5] = it_t2 - it_t1;
                    S->steplog[it_log][6] = it_kind; S->steplog[it_log][7] = d;
                    ++it_log; it_t0 = it_t2;
                }
#endif
                if (!__any_sync(FULL, progressed)) {                 // every lane waits for a plan another warp computes:
                    for (int spin = 0; spin < 4096; ++spin) {        // sleep until one of them is there
                        __nanosleep(spin < 4 ? 64 : 256);
                        bool there = false;
                        if (wait_block) there = *((volatile uint32_t *)(pce + 1)) != 0u;
                        else if (wait_grid) there = __ldcg(reinterpret_cast<const uint32_t *>(&p.t.plan[(size_t)X * 2]) + 1) != PLAN_LOCKED;
                        else there = status == 0;                    // (lost every claim: look again)
                        if (__any_sync(FULL, there)) break;
                    }
                }
            }
            // -- rollout
            double sum = 0.0;
            if (status == 1) {
                int t = 0;
                if constexpr (KIND == 0) {
                    if (p.pref_rollout) {
                        int Xl = last_ref;
                        if (Xl < 0) {                                // the leaf's id is being published by its creator
                            uint32_t c;
                            while ((c = __ldcg(&p.t.child[(size_t)(-1 - Xl)])) == CHILD_LOCKED) __nanosleep(64);
                            Xl = CHILD_ID(c);
                        }
                        const typename Mdl<0>::PrefOut po =
                            Mdl<0>::rollout_preferred(mc.M, mc.actual, mc.sampled, p.t.prob, p.t.chance + (size_t)Xl * POMCP_RS,
                                                      p.t.good + (size_t)Xl * POMCP_RS, __ldcg(&p.t.cell[Xl]), (int)(last_ar & 31u),
                                                      MD::path_obs(last_ar), sim, p.move, p.key0, p.key1, d, disc, st, 0,
                                                      p.max_depth, 0.0);
                        t = po.t; sum = po.sum;
                    } else {
                        const typename Mdl<0>::RollOut ro = Mdl<0>::rollout(mc.M.hdr, mc.M.thr32, mc.actual, mc.sampled, sim, p.move,
                                                                            p.ks, d, disc, st, p.max_depth);
                        t = ro.t; sum = ro.sum;
                    }
                } else {
                    const typename Mdl<1>::RollOut ro = Mdl<1>::rollout(mc, sim, p.move, p.ks, d, disc, st, p.max_depth);
                    t = ro.t; sum = ro.sum;
                }
                c_rollout += t;
            }
            if (status != 4) {
                p.b.flag[i] = status == 1 ? 3 : 2;
                p.b.depth[i] = (uint8_t)d;
                p.b.ro_sum[i] = sum;
            }
            TSEC(3);                                                 // rollout
        }
        GRID_SYNC();              // the statistics were read-only up to here; every path and return is stored
        PHASE_MARK(1);
        // ---- backup (solvers/pomcp.py:126-137) ----
        // (chunks of 32 simulations are dealt out statically here: the work per chunk is uniform, and a claim counter
        // that ~5000 warps hit at the same moment serialises in L2)
        for (int base = ((threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x) * 32; base < w; base += (int)gridDim.x * (SEARCH_THREADS / 32) * 32) {
            const int i = base + lane;
            const bool have = i < w;
            TSEC_START();
            // Level by level from the deepest path entry of the warp: at level d the active lanes usually update the
            // SAME edge (inside an episode the re-rooted tree is a long chain), so the warp first checks for a common
            // edge and then issues one update for all of them.  Updates go to the per-block tables in shared memory
            // -- root and depth-1 edges by index, deeper edges into the hash table keyed by edge -- which are flushed
            // once per block after the loop: a hot edge sees one global atomic per block instead of one per
            // simulation.  Integer sums: any order gives the same result.
            const int depth = have ? (int)p.b.depth[i] : 0;
            double delayed = (have && p.b.flag[i] == 3) ? p.b.ro_sum[i] : 0.0;
            int dmax = depth;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) dmax = max(dmax, __shfl_xor_sync(FULL, dmax, o));
            // the path entries are read two levels ahead of their use (and the id of a node that was still being
            // published when the simulation passed one level ahead): the loads overlap the updates of the levels above
            const auto load_entry = [&](int d, uint32_t &ar, int &X, double &rw) {
                const bool on = d >= 0 && d < depth;
                ar = on ? (uint32_t)p.b.path_ar[(size_t)d * p.wmax + i] : 0u;
                X = (on && d > 0 && !(d == 1 && stage_d1)) ? p.b.path_node[(size_t)d * p.wmax + i] : 0;
                if (KIND == 1) rw = on ? p.b.path_rew[(size_t)d * p.wmax + i] : 0.0;
            };
            // ... and before that the whole path is prefetched into L1 (one L2 round trip for all levels instead of one
            // per level: a warp has only a few chunks, nothing else hides the latency of these loads)
            for (int d = 0; d < depth; ++d) {
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.b.path_ar + (size_t)d * p.wmax + i));
                if (d > 0 && !(d == 1 && stage_d1)) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.b.path_node + (size_t)d * p.wmax + i));
                if (KIND == 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(p.b.path_rew + (size_t)d * p.wmax + i));
            }
            uint32_t ar0 = 0u, ar1 = 0u, ar2 = 0u; int x0 = 0, x1 = 0, x2 = 0; double rw0 = 0.0, rw1 = 0.0, rw2 = 0.0;
            load_entry(dmax - 1, ar0, x0, rw0);
            load_entry(dmax - 2, ar1, x1, rw1);
            if (x0 < 0) x0 = CHILD_ID(p.t.child[(size_t)(-1 - x0)]);
            const uint32_t ar_root = (have && depth > 1 && stage_d1) ? (uint32_t)p.b.path_ar[i] : 0u;   // keys of depth-1 edges
            for (int d = dmax - 1; d >= 0; --d) {
                load_entry(d - 2, ar2, x2, rw2);
                if (x1 < 0) x1 = CHILD_ID(p.t.child[(size_t)(-1 - x1)]);   // id published after this simulation passed
                const bool act = d < depth;
                uint32_t key = 0u;                                   // 1 + staged index, or 0x80000000 | edge index
                long long qi = 0;
                if (act) {
                    const uint32_t ar = ar0;
                    const int a = ar & 31;
                    const double q = MD::path_reward(mc, ar, KIND == 1 ? &rw0 : nullptr) + p.discount * delayed;
                    qi = __double2ll_rn(q * POMCP_QFIX_SCALE);
                    delayed = q;
                    if (d == 0) key = 1u + (uint32_t)a;
                    else if (d == 1 && stage_d1)
                        key = 1u + (uint32_t)((1 + (int)(ar_root & 31) * ZS + MD::path_obs(ar_root)) * A + a);
                    else {
                        DCHECK(S, x0 > 0 && x0 < S->node_count && a < A, 31);
                        key = 0x80000000u | (uint32_t)(x0 * AS + a);
                    }
                }
                // up to three rounds of "the first pending lane's edge, for every lane that shares it" (a chain level has
                // one or two nodes), aggregated into the block tables; what is left updates on its own, deeper edges
                // directly in the arena (a diverse warp gains nothing from hashing)
                bool pending = act;
                unsigned alone = 0u;                                     // lanes found to share their edge with nobody
                // cheap test first: a warp whose neighbouring lanes rarely share an edge (a bushy tree) skips the rounds
                const uint32_t nkey = __shfl_down_sync(FULL, key, 1);
                const bool shared = __popc(__ballot_sync(FULL, act && key == nkey) & 0x7FFFFFFFu) >= 4;
#pragma unroll
                for (int round = 0; round < 3 && shared; ++round) {
                    const unsigned todo = __ballot_sync(FULL, pending) & ~alone;
                    if (todo == 0u) break;
                    const int leader = __ffs((int)todo) - 1;
                    const uint32_t lkey = __shfl_sync(FULL, key, leader);
                    const bool in_grp = pending && key == lkey;
                    const unsigned grp = __ballot_sync(FULL, in_grp);
                    if (__popc(grp) < 2) { alone |= grp; continue; }         // nothing to share: leave it to the tail
                    long long sum = in_grp ? qi : 0;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
                    if (lane == leader) {
                        const int cnt = __popc(grp);
                        if (!(lkey & 0x80000000u)) {
                            atomicAdd(&acc_n[lkey - 1u], cnt);
                            smem_add64(&acc_q[lkey - 1u], sum);
                        } else {
                            uint32_t hq = ((lkey * 2654435761u) >> 16) & (ARR_HT - 1);
                            int probe = 0;
                            for (; probe < ARR_PROBES; ++probe, hq = (hq + 1u) & (ARR_HT - 1)) {
                                const uint32_t prev = atomicCAS(&ht_key[hq], 0u, lkey);
                                if (prev == 0u || prev == lkey) {
                                    atomicAdd(&ht_cnt[hq], cnt);
                                    smem_add64(&ht_q[hq], sum);
                                    break;
                                }
                            }
                            if (probe == ARR_PROBES) {                       // table crowded: update the arena directly
                                const size_t e = lkey & 0x7FFFFFFFu;
                                atomicAdd(&p.t.n[e], cnt);
                                atomicAdd(reinterpret_cast<unsigned long long *>(&p.t.qfix[e]), (unsigned long long)sum);
                            }
                        }
                    }
                    if (in_grp) pending = false;
                }
                if (pending) {
                    if (!(key & 0x80000000u)) {
                        atomicAdd(&acc_n[key - 1u], 1);
                        smem_add64(&acc_q[key - 1u], qi);
                    } else {
                        // an edge a few levels below the root is shared by many simulations of the BLOCK even when no
                        // other lane of this warp holds it: those go through the block table too (one global update
                        // per block and edge instead of one per simulation); deeper down edges rarely repeat within a
                        // block, a table slot would only add the flush to the update
                        bool staged = false;
                        if (d <= BACKUP_TABLE_DEPTH) {
                            uint32_t hq = ((key * 2654435761u) >> 16) & (ARR_HT - 1);
                            for (int probe = 0; probe < BACKUP_TABLE_PROBES && !staged; ++probe, hq = (hq + 1u) & (ARR_HT - 1)) {
                                const uint32_t prev = atomicCAS(&ht_key[hq], 0u, key);
                                if (prev == 0u || prev == key) { atomicAdd(&ht_cnt[hq], 1); smem_add64(&ht_q[hq], qi); staged = true; }
                            }
                        }
                        if (!staged) {
                            const size_t e = key & 0x7FFFFFFFu;
                            atomicAdd(&p.t.n[e], 1);
                            atomicAdd(reinterpret_cast<unsigned long long *>(&p.t.qfix[e]), (unsigned long long)qi);
                        }
                    }
                }
                // rotate the read-ahead registers only now: the loads issued at the top of the iteration had the whole
                // iteration to complete
                ar0 = ar1; x0 = x1; rw0 = rw1; ar1 = ar2; x1 = x2; rw1 = rw2;
            }
            TSEC(4);                                                 // backup
        }
        PHASE_MARK(2);
        __syncthreads();
        for (int e = threadIdx.x; e < n_acc; e += blockDim.x) {      // flush the staged root / depth-1 edges
            const int dn = acc_n[e];
            if (dn == 0) continue;
            const int slot = e / A, a = e - slot * A;
            int X = root;
            if (slot > 0) X = CHILD_ID(p.t.child[(size_t)root * AS * ZS + (slot - 1)]);   // a depth-1 node
            atomicAdd(&p.t.n[(size_t)X * AS + a], dn);
            atomicAdd(reinterpret_cast<unsigned long long *>(&p.t.qfix[(size_t)X * AS + a]), (unsigned long long)acc_q[e]);
            acc_n[e] = 0; acc_q[e] = 0;
        }
        for (int t = threadIdx.x; t < ARR_HT; t += blockDim.x) {    // ... and the deeper edges of the hash table
            const uint32_t key = ht_key[t];
            if (key == 0u) continue;
            const size_t e = key & 0x7FFFFFFFu;
            atomicAdd(&p.t.n[e], ht_cnt[t]);
            atomicAdd(reinterpret_cast<unsigned long long *>(&p.t.qfix[e]), (unsigned long long)ht_q[t]);
            ht_key[t] = 0u; ht_cnt[t] = 0; ht_q[t] = 0;
        }
        for (int t = threadIdx.x; t < PC_SLOTS; t += blockDim.x) { pc[t * PC_WORDS] = 0u; pc[t * PC_WORDS + 1] = 0u; }   // next generation: new plans
        done += w;
        if (p.timeout_ns && gtid == 0 && globaltimer_ns() - tclk[5] > p.timeout_ns) S->stop = 1;
        GRID_SYNC();
        PHASE_MARK(3);
        if (p.timeout_ns && *((volatile int32_t *)&S->stop)) break;   // action_selection_timeout
    }

    if (blockIdx.x == 0) publish_root(S, p.t, root, threadIdx.x);     // after the last generation's barrier
    // ---- fused root merge: instead of a separate all-reduce, block 0 pushes this rank's 64 root words
    // into every peer's exchange buffer with peer stores over NVLink, raises its flag there, waits for the peers'
    // flags in its own buffer and sums.  Integer operands: every rank gets the same sums in any order.
    if (p.world > 1 && blockIdx.x == 0) {
        __syncthreads();                                             // res_n / res_q are in place
        const int par = (int)(p.merge_seq & 1ull);
        for (int idx = threadIdx.x; idx < 64 * p.world; idx += blockDim.x) {
            const int peer = idx >> 6, wq = idx & 63;
            const long long v = wq < AS ? S->res_n[wq] : S->res_q[wq - AS];
            long long *dst = reinterpret_cast<long long *>(p.peer_buf[peer]);
            dst[(size_t)(par * POMCP_MAX_PEERS + p.rank) * 64 + wq] = v;
        }
        __threadfence_system();                                      // data before the flags, system-wide
        __syncthreads();
        long long *mine = reinterpret_cast<long long *>(p.peer_buf[p.rank]);
        int ok = 1;
        if (threadIdx.x < p.world) {
            volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(p.peer_buf[threadIdx.x]) +
                                                MERGE_SLOT_WORDS + par * POMCP_MAX_PEERS + p.rank;
            *flag = p.merge_seq;
            volatile unsigned long long *own = reinterpret_cast<volatile unsigned long long *>(mine) + MERGE_SLOT_WORDS +
                                               par * POMCP_MAX_PEERS + threadIdx.x;
            const unsigned long long t0 = globaltimer_ns();
            while (*own < p.merge_seq) {                             // peer threadIdx.x has delivered merge `seq`
                __nanosleep(200);
                if (globaltimer_ns() - t0 > p.merge_timeout_ns) { ok = 0; break; }   // a peer never searched: give up
            }
            atomicMax(reinterpret_cast<unsigned long long *>(&S->phase_ns[4]), globaltimer_ns() - t0);   // wait for the slowest peer
        }
        ok = __syncthreads_and(ok);
        __threadfence_system();
        if (threadIdx.x < 64) {
            long long sum = 0;
            for (int r = 0; r < p.world; ++r)
                sum += *((volatile long long *)&mine[(size_t)(par * POMCP_MAX_PEERS + r) * 64 + threadIdx.x]);
            if (threadIdx.x < AS) S->mrg_n[threadIdx.x] = sum; else S->mrg_q[threadIdx.x - AS] = sum;
        }
        if (threadIdx.x == 0) S->mrg_seq = ok ? p.merge_seq : 0ull;
    }
    // counters (warp-reduced) and persistent state
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c_levels += __shfl_xor_sync(FULL, c_levels, o);
        c_rollout += __shfl_xor_sync(FULL, c_rollout, o);
        c_created += __shfl_xor_sync(FULL, c_created, o);
        c_alloc += __shfl_xor_sync(FULL, c_alloc, o);
    }
    if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long *>(&S->counters[0]), (unsigned long long)c_levels);
        atomicAdd(reinterpret_cast<unsigned long long *>(&S->counters[1]), (unsigned long long)c_rollout);
        atomicAdd(reinterpret_cast<unsigned long long *>(&S->counters[2]), (unsigned long long)c_created);
        atomicAdd(reinterpret_cast<unsigned long long *>(&S->counters[4]), (unsigned long long)c_alloc);
    }
#ifdef POMCP_TSEC
    for (int k = 0; k < 8; ++k) {
        long long v = tsec_[k];
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&S->steplog[127][k]), (unsigned long long)v);
    }
#endif
    if (gtid == 0) {
        S->gstep = gstep;
        S->sims_done = done;
        S->stop = 0;
        S->counters[5] += gens;
        S->counters[7] = (long long)(globaltimer_ns() - tclk[5]);
        S->counters[8] = done;
        for (int k = 0; k < 4; ++k) S->phase_ns[k] = (long long)tclk[k];
        if (p.world <= 1) S->phase_ns[4] = 0;
        for (int k = 0; k < 4; ++k) S->steplog[126][k] = (long long)tclk[k];   // profiling aid: phase totals of this search
    }
}

extern "C" __global__ void __launch_bounds__(SEARCH_THREADS, SEARCH_MIN_BLOCKS)
pomcp_search_kernel(const SearchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    search_body<0>(p, smem);
}

extern "C" __global__ void __launch_bounds__(SEARCH_THREADS, SEARCH_MIN_BLOCKS)
pomcp_search_tab_kernel(const SearchParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    search_body<1>(p, smem);
}

// ---------------------------------------------------------------------------------------------------------
// K2: belief update -- batched rejection sampling with an order-preserving (stream-compacted) survivor set,
// then the re-root.  One block: a belief is only max_particle_count (2000) states.
// ---------------------------------------------------------------------------------------------------------
struct AdvanceParams {
    DevState *S;
    const ModelHeader *model;
    const uint64_t *thr53;
    TabHeader tab;
    TabTables tt;
    Tree t;
    uint32_t *bag[2];
    int32_t action, obs_good, need, n_thr;
    uint32_t sampled_after, move, key0, key1;
    long long max_tries;
};

template <int KIND>
__device__ __forceinline__ void advance_body(const AdvanceParams &p, unsigned char *smem)
{
    using MD = Mdl<KIND>;
    typename MD::Ctx mc;
    size_t off = MD::stage(smem, p.model, p.thr53, p.n_thr, p.tab, p.tt, mc);
    const int ZS = MD::zs(mc);
    int *warp_tot = reinterpret_cast<int *>(smem + off);             // [32] + [2] broadcast cells
    __shared__ long long sh_tries;
    DevState *S = p.S;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int root = S->root;
    const uint32_t sampled = p.sampled_after;                        // belief_tree_solver.py:165 model.update first
    if constexpr (KIND == 0) { mc.actual = S->actual; mc.sampled = sampled; }
    const uint32_t *old_bag = p.bag[S->bag_sel];
    uint32_t *new_bag = p.bag[S->bag_sel ^ 1];
    const int n_old = S->n_bag;
    if (tid == 0) sh_tries = -1;
    __syncthreads();
    int got = 0;
    long long base = 0;
    while (got < p.need && base < p.max_tries) {
        uint32_t ns[ADV_TRIES]; uint32_t acc = 0;
#pragma unroll
        for (int r = 0; r < ADV_TRIES; ++r) {
            const long long t = base + (long long)tid * ADV_TRIES + r;
            if (t < p.max_tries) {                                   // model.py:242-251
                const Philox4 x = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), p.move, DOM_UPDATE, p.key0, p.key1);
                const uint32_t st = old_bag[randbelow32(x.x, (uint32_t)n_old)];
                const StepOut so = MD::step(mc, st, p.action, x.y, x.z);
                ns[r] = so.next;
                if (p.obs_good < 0 || so.obs == p.obs_good) acc |= 1u << r;    // obs < 0: unconditioned (degraded mode)
            }
        }
        int cnt = __popc(acc), incl = cnt;                            // block-wide exclusive scan of cnt
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int v = warp_tot[lane], iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, iv, o); if (lane >= o) iv += y; }
            warp_tot[lane] = iv - v;
            if (lane == 31) warp_tot[32] = iv;
        }
        __syncthreads();
        int pos = got + warp_tot[wid] + incl - cnt;
        const int total = warp_tot[32];
#pragma unroll
        for (int r = 0; r < ADV_TRIES; ++r)
            if ((acc >> r) & 1u) {
                if (pos < p.need) new_bag[pos] = ns[r];
                if (pos == p.need - 1) sh_tries = base + (long long)tid * ADV_TRIES + r + 1;
                ++pos;
            }
        got += total;
        base += (long long)ADV_THREADS * ADV_TRIES;
        __syncthreads();
    }
    if (tid == 0) {
        const int accepted = got < p.need ? got : p.need;
        long long tries = sh_tries >= 0 ? sh_tries : (base < p.max_tries ? base : p.max_tries);
        S->adv_tries = tries; S->adv_accepted = accepted;
        S->sampled = sampled;                                        // belief_tree_solver.py:165 happens first
        if (accepted == 0) { S->adv_status = 2; S->adv_root_cell = p.t.cell[root]; }
        else {
            uint32_t c = 0u;                                                          // belief_tree_solver.py:167
            bool hashed = false;
            if constexpr (KIND == 1) hashed = p.t.ckey != nullptr;
            if (p.obs_good >= 0 && hashed) {
                const long long cs = child_slot_hashed<false>(p.t, root, p.action, p.obs_good);
                if (cs >= 0) c = __ldcg(&p.t.child[cs]);
            } else if (p.obs_good >= 0) c = p.t.child[((size_t)root * AS + p.action) * ZS + p.obs_good];
            int status = 0, id;
            if (c == 0) {                                            // never expanded: fresh node at the next position
                id = S->node_count++;
                uint32_t tag = 0u;                                   // tag of the node at the next position
                if constexpr (KIND == 0)
                    tag = rs_step(mc.M, (uint32_t)p.t.cell[root], 0u, p.action, 0ull, mc.actual, sampled).cell;
                p.t.cell[id] = (int32_t)tag;
                uint32_t lm = MD::tag_legal(mc, tag);
                if constexpr (KIND == 0) if (p.t.chance)
                    lm = child_smart_mask(mc.M.hdr, p.t.prob, p.t.chance + (size_t)root * POMCP_RS, p.t.good + (size_t)root * POMCP_RS,
                                          p.t.chance + (size_t)id * POMCP_RS, p.t.good + (size_t)id * POMCP_RS, p.t.cell[root],
                                          p.action, p.obs_good < 0 ? 0 : p.obs_good, (int)tag);
                p.t.legal[id] = lm;
                status = 1;
            } else id = CHILD_ID(c);
            S->root = id; S->n_bag = accepted; S->bag_sel ^= 1;                        // :210
            S->adv_status = status; S->adv_root_cell = p.t.cell[id];
        }
    }
    __syncthreads();
    publish_root(S, p.t, S->root, tid);
}

extern "C" __global__ void __launch_bounds__(ADV_THREADS) pomcp_advance_kernel(const AdvanceParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    advance_body<0>(p, smem);
}

extern "C" __global__ void __launch_bounds__(ADV_THREADS) pomcp_advance_tab_kernel(const AdvanceParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    advance_body<1>(p, smem);
}

// generate_step over a batch (parity harness for rock_model.py:451-465)
extern "C" __global__ void pomcp_step_batch_kernel(const ModelHeader *model, const uint64_t *thr53, int n_thr,
                                                   long long n, const uint32_t *state, const uint8_t *action,
                                                   const uint64_t *u53, const uint32_t *actual,
                                                   const uint32_t *sampled, uint32_t *next_state, uint8_t *flags,
                                                   double *reward)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Mdl<0>::Ctx mc;
    Mdl<0>::stage(smem, model, thr53, n_thr, TabHeader{}, TabTables{}, mc);
    const SmemModel &M = mc.M;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t st = state[i];
        const StepResult sr = rs_step(M, st & 0xFFu, st >> 8, action[i], u53[i], actual[i], sampled[i]);
        next_state[i] = sr.cell | (sr.rocks << 8);
        flags[i] = (uint8_t)(sr.good | (sr.empty << 1) | (sr.terminal << 2) | (sr.legal << 3));
        reward[i] = M.hdr->rew[sr.rew];
    }
}

// generate_step of the tabular model over a batch (parity harness)
__global__ void pomcp_tab_step_batch_kernel(TabHeader th, TabTables tt, long long n, const uint32_t *state,
                                            const uint8_t *action, const uint32_t *w1, const uint32_t *w2,
                                            uint32_t *next_state, int32_t *obs, double *reward, uint8_t *terminal)
{
    Mdl<1>::Ctx mc{th, tt};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const StepOut so = Mdl<1>::step(mc, state[i], action[i], w1[i], w2[i]);
        next_state[i] = so.next; obs[i] = so.obs; reward[i] = so.reward; terminal[i] = (uint8_t)so.terminal;
    }
}

// Plant a root (fresh tree).  With --preferred_actions the root gets fresh rock statistics (rock_model.py:302-304)
// and the smart action set; `keep_stats` re-plants the current root's statistics (tree reset within an episode).
__global__ void pomcp_plant_root_kernel(DevState *S, Tree t, const ModelHeader *H, int root_cell, uint32_t legal,
                                        int n_bag, int bag_sel, int keep_stats)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int old_root = S->root;
        S->root = 0; S->node_count = 1; S->n_bag = n_bag; S->bag_sel = bag_sel;
        t.cell[0] = root_cell;
        if (t.chance) {
            double ch[POMCP_RS]; int16_t gd[POMCP_RS];
            for (int k = 0; k < H->n_rocks; ++k) {
                ch[k] = keep_stats ? t.chance[(size_t)old_root * POMCP_RS + k] : 0.5;
                gd[k] = keep_stats ? t.good[(size_t)old_root * POMCP_RS + k] : (int16_t)0;
            }
            legal = child_smart_mask(H, t.prob, ch, gd, t.chance, t.good, root_cell, 0, 0, root_cell);   // a = 0: no update
        }
        t.legal[0] = legal;
        S->res_legal = legal; S->res_valid = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x < AS) { S->res_n[threadIdx.x] = 0; S->res_q[threadIdx.x] = 0; }
}

// Zero the statistics and child tables of the arena prefix that earlier searches touched.  The exact node count is
// read on the device (the host only knows an upper bound unless it has read the state block back), so a tree reset
// clears what was used, not what could have been used.
__global__ void pomcp_clear_arena_kernel(const DevState *S, Tree t, int zs, long long upper)
{
    long long n = S->node_count;
    if (n > upper) n = upper;
    const long long n4 = n * (AS * 4 / 16), q4 = n * (AS * 8 / 16);
    // direct table: the rows of the used nodes; hashed table: every slot, keys and values (entries sit anywhere)
    const long long c4 = t.ckey ? ((long long)t.cmask + 1) * 4 / 16 : n * ((long long)AS * zs * 4 / 16);
    const long long k4 = t.ckey ? ((long long)t.cmask + 1) * 8 / 16 : 0;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4 *pn = reinterpret_cast<uint4 *>(t.n), *pq = reinterpret_cast<uint4 *>(t.qfix), *pc = reinterpret_cast<uint4 *>(t.child);
    uint4 *pk = reinterpret_cast<uint4 *>(t.ckey);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q4; i += stride) {
        pq[i] = z;
        if (i < n4) pn[i] = z;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c4; i += stride) pc[i] = z;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < k4; i += stride) pk[i] = z;
}

// The child entry of edge (X, a, o) -- id + 1, 0 = none, ~0 = being created -- into *out (pomcp_node_stats walks a
// path with it; direct or hashed table alike).
__global__ void pomcp_find_child_kernel(Tree t, int zs, int X, int a, int o, uint32_t *out)
{
    uint32_t c = 0u;
    if (t.ckey) {
        const long long cs = child_slot_hashed<false>(t, X, a, o);
        if (cs >= 0) c = t.child[cs];
    } else c = t.child[((size_t)X * AS + a) * (size_t)zs + o];
    *out = c;
}

// Root edge statistics as 2*AS 64-bit integers (n[0..AS), qfix[0..AS)) in a caller-provided device buffer: the
// operand of the root-parallel merge (one all-reduce(sum) over NVLink per move).
__global__ void pomcp_root_stats_kernel(const DevState *S, Tree t, long long *out)
{
    const int a = threadIdx.x;
    if (a < AS) {
        const int root = S->root;
        out[a] = t.n[(size_t)root * AS + a];
        out[AS + a] = t.qfix[(size_t)root * AS + a];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Host side: the handle and the C ABI
// ---------------------------------------------------------------------------------------------------------
struct pomcp_handle {
    pomcp_config cfg;
    int device = 0;
    std::string err;
    DevState *dS = nullptr;
    ModelHeader hmodel;
    ModelHeader *dmodel = nullptr;
    uint64_t *dthr = nullptr;
    int n_thr = 0;
    int kind = 0;                   // 0 RockSample, 1 tabular
    int zs = 2;                     // observation stride of the child table
    bool hashed = false;            // ... or a hashed child table (tabular models with more than 32 observations)
    int pc_mult = 16;               // a node is hot (plan cached per block, computed in the prologue) from pc_mult expected
                                    // arrivals per block (POMCP_PC_MULT; results do not depend on it)
    int n_actions = 0;
    TabHeader tab{};
    TabTables tt{};
    bool have_model = false, have_root = false;
    Tree t{};
    SimBuf b{};
    uint32_t *bag[2] = {nullptr, nullptr};
    int bag_cap = 0;
    long long cap = 0;
    long long nodes_used = 0;       // host mirror (upper bound) of DevState.node_count
    long long zeroed_upto = 0;      // arena prefix that may hold stale data
    int root_cell = 0;
    int n_bag = 0;
    int bag_sel = 0;               // host mirror of DevState.bag_sel
    int coop_blocks = 0;
    size_t search_smem = 0;
    long long launches = 0;
    int pref_rollout = 0;           // preferred actions also steer the rollouts (pomcp_set_preferred_actions(2, ...))
    int world = 1, rank = 0;        // fused root merge (pomcp_set_peer_merge)
    unsigned long long peer_buf[POMCP_MAX_PEERS] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long merge_seq = 0;
    unsigned long long merge_timeout_ns = 2000000000ull;   // how long the fused merge waits for a peer (POMCP_MERGE_TIMEOUT_S)
    void *stage = nullptr;          // pinned host staging
    cudaStream_t last_stream = nullptr;
};

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

static int fail(pomcp_handle *h, const char *msg) { h->err = msg; return -1; }

extern "C" const char *pomcp_last_error(const pomcp_handle *h) { return h ? h->err.c_str() : "null handle"; }

static void free_all(pomcp_handle *h)
{
    cudaFree(h->dS); cudaFree(h->dmodel); cudaFree(h->dthr);
    cudaFree(h->t.n); cudaFree(h->t.qfix); cudaFree(h->t.child); cudaFree(h->t.legal); cudaFree(h->t.cell);
    cudaFree(h->t.thr); cudaFree(h->t.plan); cudaFree(h->t.chance); cudaFree(h->t.good); cudaFree((void *)h->t.prob);
    cudaFree(h->b.flag); cudaFree(h->b.depth); cudaFree(h->b.path_node);
    cudaFree(h->b.path_ar); cudaFree(h->b.ro_sum); cudaFree(h->b.path_rew);
    cudaFree((void *)h->tt.Tc); cudaFree((void *)h->tt.Oc); cudaFree((void *)h->tt.R); cudaFree((void *)h->tt.term_state);
    cudaFree((void *)h->tt.Tg);
    cudaFree(h->bag[0]); cudaFree(h->bag[1]);
    if (h->stage) cudaFreeHost(h->stage);
}

extern "C" void pomcp_destroy(pomcp_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_all(h);
    delete h;
}

extern "C" int pomcp_create(const pomcp_config *cfg, int device, pomcp_handle **out)
{
    if (!cfg || !out) return -1;
    pomcp_handle *h = new pomcp_handle();
    *out = h;                          // returned even on failure so that pomcp_last_error() can be read
    h->cfg = *cfg; h->device = device;
    if (cfg->max_depth < 1 || cfg->max_depth > 4096) return fail(h, "max_depth out of range [1, 4096]");
    if (cfg->child_table < 0 || cfg->child_table > 1) return fail(h, "child_table must be 0 (auto) or 1 (hashed)");
    if (cfg->tree_depth_cap < 2 || cfg->tree_depth_cap > 64) return fail(h, "tree_depth_cap out of range [2, 64]");
    if (cfg->gen_w0 < 1 || cfg->gen_growth < 1 || cfg->gen_wmax < cfg->gen_w0) return fail(h, "bad generation schedule");
    if (cfg->node_capacity < 2 || cfg->node_capacity > (1ll << 25)) return fail(h, "node_capacity out of range [2, 2^25]");
    if (cfg->max_particle_count < 1) return fail(h, "max_particle_count < 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(h, "no CUDA device: the B200 POMCP planner has no CPU fallback");
    CK(cudaSetDevice(device));
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) return fail(h, "device lacks cooperative launch");
    h->cap = cfg->node_capacity;
    if (const char *e = getenv("POMCP_PC_MULT")) { const int v = atoi(e); if (v >= 1 && v <= 4096) h->pc_mult = v; }
    if (const char *mt = getenv("POMCP_MERGE_TIMEOUT_S")) {
        const double v = atof(mt);
        if (v > 0) h->merge_timeout_ns = (unsigned long long)(v * 1e9);
    }
    const size_t cap = (size_t)h->cap, wmax = (size_t)cfg->gen_wmax, dcap = (size_t)cfg->tree_depth_cap;
    CK(cudaMalloc(&h->dS, sizeof(DevState)));
    CK(cudaMemset(h->dS, 0, sizeof(DevState)));
    CK(cudaMalloc(&h->dmodel, sizeof(ModelHeader)));
    CK(cudaMalloc(&h->t.n, cap * AS * sizeof(int32_t)));
    CK(cudaMalloc(&h->t.qfix, cap * AS * sizeof(long long)));
    CK(cudaMalloc(&h->t.child, cap * AS * 2 * sizeof(uint32_t)));
    CK(cudaMalloc(&h->t.legal, cap * sizeof(uint32_t)));
    CK(cudaMalloc(&h->t.cell, cap * sizeof(int32_t)));
    CK(cudaMalloc(&h->t.thr, cap * 32 * sizeof(uint32_t)));
    CK(cudaMalloc(&h->t.plan, cap * 2 * sizeof(uint4)));
    CK(cudaMemset(h->t.plan, 0, cap * 2 * sizeof(uint4)));           // stamp 0: no plan yet (stamps only grow)
    CK(cudaMemset(h->t.n, 0, cap * AS * sizeof(int32_t)));
    CK(cudaMemset(h->t.qfix, 0, cap * AS * sizeof(long long)));
    CK(cudaMemset(h->t.child, 0, cap * AS * 2 * sizeof(uint32_t)));
    CK(cudaMalloc(&h->b.flag, wmax));
    CK(cudaMalloc(&h->b.depth, wmax));
    CK(cudaMalloc(&h->b.path_node, wmax * dcap * sizeof(int32_t)));
    CK(cudaMalloc(&h->b.path_ar, wmax * dcap * sizeof(uint16_t)));
    CK(cudaMalloc(&h->b.ro_sum, wmax * sizeof(double)));
    CK(cudaMallocHost(&h->stage, 1 << 16));
    return 0;
}

// Drop the device tables of a tabular model (a handle can be re-targeted at another model).
static int release_tabular(pomcp_handle *h)
{
    cudaFree((void *)h->tt.Tc); cudaFree((void *)h->tt.Oc); cudaFree((void *)h->tt.R); cudaFree((void *)h->tt.term_state);
    cudaFree((void *)h->tt.Tg);
    h->tt = TabTables{};
    return 0;
}

// (Re)allocate the child table: direct-indexed with an observation stride of `zs` slots per (node, action), or
// hashed over (node, action, observation) with a power-of-two slot count >= 2 * node_capacity (pomcp_state.cuh).
static int set_child_stride(pomcp_handle *h, int zs, bool hashed)
{
    if (hashed) {
        size_t slots = 1024;
        while (slots < 2 * (size_t)h->cap) slots *= 2;            // cap <= 2^25: at most 2^26 slots (pending refs fit 31 bits)
        if (!h->hashed || h->t.cmask != (uint32_t)(slots - 1)) {
            CK(cudaDeviceSynchronize());
            cudaFree(h->t.child); h->t.child = nullptr;
            cudaFree(h->t.ckey); h->t.ckey = nullptr;
            CK(cudaMalloc(&h->t.child, slots * sizeof(uint32_t)));
            CK(cudaMalloc(&h->t.ckey, slots * sizeof(unsigned long long)));
            CK(cudaMemset(h->t.child, 0, slots * sizeof(uint32_t)));
            CK(cudaMemset(h->t.ckey, 0, slots * sizeof(unsigned long long)));
            h->t.cmask = (uint32_t)(slots - 1);
            h->hashed = true; h->zs = zs;
        }
        h->zs = zs;
        return 0;
    }
    // a child slot that is still being created is referred to as -1 - slot index in a 32-bit path entry
    if ((unsigned long long)h->cap * AS * (unsigned long long)zs > (1ull << 31))
        return fail(h, "node_capacity * 32 * observation stride exceeds 2^31 child slots: lower node_capacity");
    if (zs != h->zs || h->hashed) {
        CK(cudaDeviceSynchronize());
        cudaFree(h->t.child); h->t.child = nullptr;
        cudaFree(h->t.ckey); h->t.ckey = nullptr; h->t.cmask = 0u; h->hashed = false;
        CK(cudaMalloc(&h->t.child, (size_t)h->cap * AS * (size_t)zs * sizeof(uint32_t)));
        CK(cudaMemset(h->t.child, 0, (size_t)h->cap * AS * (size_t)zs * sizeof(uint32_t)));
        h->zs = zs;
    }
    return 0;
}

extern "C" int pomcp_set_model_rocksample(pomcp_handle *h, int rows, int cols, const int8_t *cell, int start_cell,
                                          const uint64_t *thr53, const double rewards[4])
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    if (rows < 1 || cols < 1 || rows * cols > POMCP_MAX_CELLS) return fail(h, "grid larger than 256 cells");
    ModelHeader &m = h->hmodel;
    memset(&m, 0, sizeof(m));
    m.rows = rows; m.cols = cols; m.n_cells = rows * cols; m.start_cell = start_cell;
    int nr = 0;
    for (int c = 0; c < m.n_cells; ++c) { m.cell[c] = cell[c]; if (cell[c] + 1 > nr) nr = cell[c] + 1; }
    if (nr < 1 || nr > POMCP_MAX_ROCKS || 5 + nr > POMCP_MAX_ACTIONS) return fail(h, "unsupported rock count");
    m.n_rocks = nr; m.n_actions = 5 + nr;
    m.rew[REW_NONE] = 0.0; m.rew[REW_ILLEGAL] = rewards[0]; m.rew[REW_EXIT] = rewards[1];
    m.rew[REW_GOOD] = rewards[2]; m.rew[REW_BAD] = rewards[3];
    m.delta[0] = -cols; m.delta[1] = 1; m.delta[2] = cols; m.delta[3] = -1;
    const uint32_t checks = ((nr + 5 >= 32) ? 0xFFFFFFFFu : ((1u << (nr + 5)) - 1u)) & ~31u;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            const int c = i * cols + j;
            if (m.cell[c] == POMCP_CELL_OBSTACLE) { m.legal[c] = 0; continue; }
            uint32_t mask = checks;
            if (i > 0 && m.cell[c - cols] != POMCP_CELL_OBSTACLE) mask |= 1u;
            if (j + 1 < cols && m.cell[c + 1] != POMCP_CELL_OBSTACLE) mask |= 2u;
            if (i + 1 < rows && m.cell[c + cols] != POMCP_CELL_OBSTACLE) mask |= 4u;
            if (j > 0 && m.cell[c - 1] != POMCP_CELL_OBSTACLE) mask |= 8u;
            if (m.cell[c] >= 0) mask |= 16u;
            m.legal[c] = mask;
        }
    for (int c = 0; c < m.n_cells; ++c) if (m.cell[c] >= 0) m.rock_cell[m.cell[c]] = c;
    for (int mk = 0; mk < 32; ++mk) {
        int j = 0;
        for (int bit = 0; bit < 5; ++bit) if ((mk >> bit) & 1) m.nth5[mk * 8 + j++] = (uint8_t)bit;
    }
    h->n_thr = m.n_cells * nr;
    cudaFree(h->dthr); h->dthr = nullptr;
    CK(cudaMalloc(&h->dthr, (size_t)h->n_thr * sizeof(uint64_t)));
    CK(cudaMemcpy(h->dthr, thr53, (size_t)h->n_thr * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->dmodel, &m, sizeof(m), cudaMemcpyHostToDevice));
    if (release_tabular(h)) return -1;
    h->kind = 0; h->n_actions = m.n_actions;
    if (set_child_stride(h, 2, false)) return -1;
    // launch geometry of the cooperative search kernel
    const int A = m.n_actions;
    size_t smem = ((sizeof(ModelHeader) + 15) & ~size_t(15)) + (((size_t)h->n_thr * 4 + 15) & ~size_t(15)) +
                  (((size_t)h->cfg.max_depth * 8 + 15) & ~size_t(15)) + (size_t)(1 + 2 * A) * A * 12 + (size_t)ARR_HT * 16 + (size_t)PC_SLOTS * PC_WORDS * 4 + (size_t)(4 + 2 * SQ_CAP) * 4 + 48;
    h->search_smem = smem;
    CK(cudaFuncSetAttribute(pomcp_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pomcp_search_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(pomcp_advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pomcp_step_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pomcp_search_kernel, SEARCH_THREADS, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    if (per_sm < 1) return fail(h, "search kernel does not fit on an SM");
    h->coop_blocks = per_sm * sms;
    h->have_model = true; h->have_root = false;
    return 0;
}

extern "C" int pomcp_set_model_tabular(pomcp_handle *h, int n_states, int n_actions, int n_obs, const uint32_t *Tc,
                                       const uint32_t *Oc, const double *R, uint32_t term_action_mask,
                                       const uint8_t *term_state)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    if (n_states < 1 || n_states > (1 << 16)) return fail(h, "n_states out of range [1, 65536]");
    if (n_actions < 1 || n_actions > POMCP_MAX_ACTIONS) return fail(h, "n_actions out of range [1, 32]");
    if (n_obs < 1 || n_obs > POMCP_MAX_OBS) return fail(h, "n_obs out of range [1, 2048]");
    if (!Tc || !Oc || !R) return fail(h, "null model table");
    const size_t S = (size_t)n_states, A = (size_t)n_actions, Z = (size_t)n_obs;
    for (size_t r = 0; r < A * S; ++r) {                    // every cumulative row must end saturated
        if (Tc[r * S + S - 1] != 0xFFFFFFFFu) return fail(h, "transition row does not end at 0xFFFFFFFF");
        if (Oc[r * Z + Z - 1] != 0xFFFFFFFFu) return fail(h, "observation row does not end at 0xFFFFFFFF");
    }
    CK(cudaDeviceSynchronize());
    if (release_tabular(h)) return -1;
    uint32_t *dT, *dO; double *dR; uint8_t *dterm;
    CK(cudaMalloc(&dT, A * S * S * 4)); h->tt.Tc = dT;
    CK(cudaMalloc(&dO, A * S * Z * 4)); h->tt.Oc = dO;
    CK(cudaMalloc(&dR, A * S * 8)); h->tt.R = dR;
    CK(cudaMalloc(&dterm, S)); h->tt.term_state = dterm;
    CK(cudaMemcpy(dT, Tc, A * S * S * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dO, Oc, A * S * Z * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, R, A * S * 8, cudaMemcpyHostToDevice));
    if (term_state) CK(cudaMemcpy(dterm, term_state, S, cudaMemcpyHostToDevice));
    else CK(cudaMemset(dterm, 0, S));
    int zs = 2; while (zs < n_obs) zs *= 2;
    TabHeader &t = h->tab;
    memset(&t, 0, sizeof(t));
    t.S = n_states; t.A = n_actions; t.Z = n_obs; t.ZS = zs; t.term_action_mask = term_action_mask;
    // child table: direct-indexed [node][action][zs] up to POMCP_MAX_OBS_DIRECT observations, hashed above (or on request)
    const bool hashed = n_obs > POMCP_MAX_OBS_DIRECT || h->cfg.child_table == 1;
    t.stage_d1 = (!hashed && (size_t)(1 + zs * n_actions) * n_actions * 12 <= (size_t)(24 << 10)) ? 1 : 0;
    // Guide table of the transition rows (|S| >= 64): 2^bits buckets per row, about four thresholds per bucket, at
    // most 256 MB in total.  guide[j] = first index whose threshold exceeds j << (32 - bits); guide[2^bits] = S - 1.
    int gbits = 0;
    if (n_states >= 64) {
        while ((1 << (gbits + 1)) <= n_states / 4) ++gbits;
        while (gbits > 0 && A * S * ((size_t)(1u << gbits) + 1) * 2 > ((size_t)256 << 20)) --gbits;
    }
    if (gbits > 0) {
        const size_t G = (size_t)1 << gbits;
        std::vector<uint16_t> guide(A * S * (G + 1));
        for (size_t r = 0; r < A * S; ++r) {
            const uint32_t *cdf = Tc + r * S;
            uint16_t *g = guide.data() + r * (G + 1);
            size_t k = 0;
            for (size_t j = 0; j < G; ++j) {
                const uint32_t w0 = (uint32_t)(j << (32 - gbits));
                while (k + 1 < S && !(cdf[k] > w0)) ++k;
                g[j] = (uint16_t)k;
            }
            g[G] = (uint16_t)(S - 1);
        }
        uint16_t *dG;
        CK(cudaMalloc(&dG, guide.size() * sizeof(uint16_t))); h->tt.Tg = dG;
        CK(cudaMemcpy(dG, guide.data(), guide.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    t.guide_bits = gbits;
    h->kind = 1; h->n_actions = n_actions;
    memset(&h->hmodel, 0, sizeof(h->hmodel));
    h->hmodel.n_actions = n_actions; h->hmodel.n_cells = 1;
    h->hmodel.legal[0] = n_actions >= 32 ? 0xFFFFFFFFu : ((1u << n_actions) - 1u);
    CK(cudaMemcpy(h->dmodel, &h->hmodel, sizeof(h->hmodel), cudaMemcpyHostToDevice));
    if (set_child_stride(h, zs, hashed)) return -1;
    if (!h->b.path_rew)
        CK(cudaMalloc(&h->b.path_rew, (size_t)h->cfg.gen_wmax * (size_t)h->cfg.tree_depth_cap * sizeof(double)));
    const size_t n_acc = t.stage_d1 ? (size_t)(1 + zs * n_actions) * n_actions : (size_t)n_actions;
    size_t smem = (((size_t)h->cfg.max_depth * 8 + 15) & ~size_t(15)) + n_acc * 12 + (size_t)ARR_HT * 16 + (size_t)PC_SLOTS * PC_WORDS * 4 + (size_t)(4 + 2 * SQ_CAP) * 4 + 48;
    if (smem < 512) smem = 512;                              // the belief update keeps 34 ints there
    h->search_smem = smem;
    CK(cudaFuncSetAttribute(pomcp_search_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pomcp_advance_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pomcp_search_tab_kernel, SEARCH_THREADS, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    if (per_sm < 1) return fail(h, "search kernel does not fit on an SM");
    h->coop_blocks = per_sm * sms;
    h->have_model = true; h->have_root = false;
    return 0;
}

extern "C" int pomcp_set_preferred_actions(pomcp_handle *h, int enabled, const double *prob)
{
    if (!h) return -1;
    if (!h->have_model) return fail(h, "set the model first");
    if (h->kind != 0) return enabled ? fail(h, "preferred actions are a RockSample heuristic") : 0;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    cudaFree(h->t.chance); cudaFree(h->t.good); cudaFree((void *)h->t.prob);
    h->t.chance = nullptr; h->t.good = nullptr; h->t.prob = nullptr;
    h->have_root = false;                                    // the tree must be re-planted
    h->pref_rollout = enabled == 2;
    if (!enabled) return 0;
    if (!prob) return fail(h, "preferred actions need the sensor probability table");
    double *dp;
    CK(cudaMalloc(&dp, sizeof(double) * (size_t)h->n_thr));
    CK(cudaMemcpy(dp, prob, sizeof(double) * (size_t)h->n_thr, cudaMemcpyHostToDevice));
    h->t.prob = dp;
    CK(cudaMalloc(&h->t.chance, sizeof(double) * (size_t)h->cap * POMCP_RS));
    CK(cudaMalloc(&h->t.good, sizeof(int16_t) * (size_t)h->cap * POMCP_RS));
    return 0;
}

extern "C" int pomcp_root_rock_data(pomcp_handle *h, double *chance, int32_t *good)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    if (!h->t.chance) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    int32_t root;
    CK(cudaMemcpy(&root, &h->dS->root, sizeof(root), cudaMemcpyDeviceToHost));
    int16_t gd[POMCP_RS];
    CK(cudaMemcpy(chance, h->t.chance + (size_t)root * POMCP_RS, sizeof(double) * h->hmodel.n_rocks, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gd, h->t.good + (size_t)root * POMCP_RS, sizeof(gd), cudaMemcpyDeviceToHost));
    for (int k = 0; k < h->hmodel.n_rocks; ++k) good[k] = gd[k];
    return 0;
}

extern "C" int pomcp_set_world(pomcp_handle *h, uint32_t actual, uint32_t sampled)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    uint32_t v[2] = {actual, sampled};
    CK(cudaMemcpy(&h->dS->actual, v, sizeof(v), cudaMemcpyHostToDevice));
    return 0;
}

// Zero the part of the arena that earlier searches touched.
static int clear_arena(pomcp_handle *h, cudaStream_t s)
{
    const long long n = h->zeroed_upto < h->cap ? h->zeroed_upto : h->cap;
    if (n > 0) {                                             // `arrive` keeps its stamps: they are never reused
        long long blocks = (n * (AS * 8 / 16) + 255) / 256;
        if (blocks > 148 * 8 || h->hashed) blocks = 148 * 8;    // (a hashed child table is cleared as a whole)
        pomcp_clear_arena_kernel<<<(int)blocks, 256, 0, s>>>(h->dS, h->t, h->zs, n);
        CK(cudaGetLastError());
        ++h->launches;
    }
    h->zeroed_upto = 0;
    return 0;
}

static int plant_root(pomcp_handle *h, int root_cell, int n_bag, int bag_sel, cudaStream_t s, int keep_stats)
{
    if (clear_arena(h, s)) return -1;
    pomcp_plant_root_kernel<<<1, 32, 0, s>>>(h->dS, h->t, h->dmodel, root_cell, h->hmodel.legal[root_cell], n_bag, bag_sel,
                                             keep_stats);
    CK(cudaGetLastError());
    ++h->launches;
    h->nodes_used = 1; h->zeroed_upto = 1; h->root_cell = root_cell; h->n_bag = n_bag;
    return 0;
}

extern "C" int pomcp_set_root_particles(pomcp_handle *h, const uint32_t *packed, int n, int root_cell)
{
    if (!h) return -1;
    if (!h->have_model) return fail(h, "set the model first");
    if (n < 1) return fail(h, "empty belief");
    if (root_cell < 0 || root_cell >= h->hmodel.n_cells) return fail(h, "root cell outside the grid");
    if (h->kind == 1)
        for (int i = 0; i < n; ++i) if (packed[i] >= (uint32_t)h->tab.S) return fail(h, "particle outside the state space");
    CK(cudaSetDevice(h->device));
    const int need = n > h->cfg.max_particle_count ? n : h->cfg.max_particle_count;
    if (need > h->bag_cap) {
        cudaFree(h->bag[0]); cudaFree(h->bag[1]); h->bag[0] = h->bag[1] = nullptr;
        CK(cudaMalloc(&h->bag[0], (size_t)need * sizeof(uint32_t)));
        CK(cudaMalloc(&h->bag[1], (size_t)need * sizeof(uint32_t)));
        h->bag_cap = need;
    }
    CK(cudaMemcpy(h->bag[0], packed, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (plant_root(h, root_cell, n, 0, 0, 0)) return -1;
    CK(cudaDeviceSynchronize());
    h->have_root = true; h->bag_sel = 0;
    return 0;
}

extern "C" int pomcp_reset_tree(pomcp_handle *h, void *stream)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s;
    return plant_root(h, h->root_cell, h->n_bag, h->bag_sel, s, 1);
}

extern "C" int pomcp_root_stats_device(pomcp_handle *h, int64_t *dev_out, void *stream)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    CK(cudaSetDevice(h->device));
    pomcp_root_stats_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(h->dS, h->t, reinterpret_cast<long long *>(dev_out));
    CK(cudaGetLastError());
    ++h->launches;
    return 0;
}

extern "C" int pomcp_search(pomcp_handle *h, int64_t n_sims, uint32_t sim_offset, uint32_t move, double timeout_s,
                            void *stream)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief: call pomcp_set_root_particles first");
    if (n_sims < 0) return fail(h, "n_sims < 0");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    h->last_stream = s;
    int rc = 0;
    if (h->nodes_used + n_sims + 2 > h->cap) {          // every simulation creates at most one node
        if (n_sims + 3 > h->cap) return fail(h, "node_capacity smaller than n_sims + 3");
        // drop the subtree, keep the belief (documented degradation; status 1)
        if (plant_root(h, h->root_cell, h->n_bag, h->bag_sel, s, 1)) return -1;
        rc = 1;
    }
    SearchParams p;
    memset(&p, 0, sizeof(p));
    p.S = h->dS; p.model = h->dmodel; p.thr53 = h->dthr; p.t = h->t; p.b = h->b;
    p.tab = h->tab; p.tt = h->tt;
    p.bag[0] = h->bag[0]; p.bag[1] = h->bag[1];
    p.n_sims = n_sims; p.sim_offset = sim_offset; p.move = move;
    p.key0 = (uint32_t)h->cfg.seed; p.key1 = (uint32_t)(h->cfg.seed >> 32);
    for (uint32_t r = 0; r < 10; ++r) { p.ks[2 * r] = p.key0 + r * 0x9E3779B9u; p.ks[2 * r + 1] = p.key1 + r * 0xBB67AE85u; }
    p.w0 = h->cfg.gen_w0; p.growth = h->cfg.gen_growth; p.wmax = h->cfg.gen_wmax;
    p.max_depth = h->cfg.max_depth; p.dcap = h->cfg.tree_depth_cap; p.n_thr = h->n_thr; p.node_cap = h->cap;
    p.ucb_c = h->cfg.ucb_coefficient; p.discount = h->cfg.discount;
    p.timeout_ns = timeout_s > 0 ? (unsigned long long)(timeout_s * 1e9) : 0ull;   // checked between generations
    p.merge_timeout_ns = h->merge_timeout_ns;
    p.pref_rollout = h->pref_rollout;
    p.pc_mult = h->pc_mult;
    p.world = h->world; p.rank = h->rank;
    if (h->world > 1) {
        for (int r = 0; r < h->world; ++r) p.peer_buf[r] = h->peer_buf[r];
        p.merge_seq = ++h->merge_seq;
    }
    long long widest = n_sims < (long long)p.wmax ? n_sims : (long long)p.wmax;
    int blocks = (int)((widest + SEARCH_THREADS - 1) / SEARCH_THREADS);
    if (blocks < 1) blocks = 1;
    if (blocks > h->coop_blocks) blocks = h->coop_blocks;
    CK(cudaMemsetAsync(h->dS->cursor, 0, 6 * sizeof(int32_t), s));   // cursor[4], stop, barrier
    void *args[] = {&p};
    CK(cudaLaunchCooperativeKernel(h->kind == 0 ? (void *)pomcp_search_kernel : (void *)pomcp_search_tab_kernel,
                                   dim3(blocks), dim3(SEARCH_THREADS), args, h->search_smem, s));
    ++h->launches;
    h->nodes_used += n_sims;
    h->zeroed_upto = h->nodes_used;
    return rc;
}

extern "C" int pomcp_root_stats(pomcp_handle *h, int64_t *n, int64_t *qfix, uint32_t *legal_mask, int64_t *sims_done)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    CK(cudaSetDevice(h->device));
    DevState *st = reinterpret_cast<DevState *>(h->stage);           // pinned staging: one asynchronous copy + sync
    CK(cudaMemcpyAsync(st, h->dS, DEVSTATE_HEAD, cudaMemcpyDeviceToHost, h->last_stream));
    CK(cudaStreamSynchronize(h->last_stream));
    h->nodes_used = st->node_count;                                  // exact now (the copy is ordered after the search):
    h->zeroed_upto = st->node_count;                                 // the next tree reset clears only what was touched
    const int A = h->n_actions;
    for (int a = 0; a < A; ++a) { n[a] = st->res_n[a]; qfix[a] = st->res_q[a]; }
    if (legal_mask) *legal_mask = st->res_legal;
    if (sims_done) *sims_done = st->sims_done;
    return 0;
}

extern "C" int pomcp_set_peer_merge(pomcp_handle *h, int world, int rank, const uint64_t *peer_buffers)
{
    if (!h) return -1;
    if (world <= 1) { h->world = 1; h->rank = 0; return 0; }
    if (world > POMCP_MAX_PEERS || rank < 0 || rank >= world || !peer_buffers) return fail(h, "bad peer-merge arguments");
    for (int r = 0; r < world; ++r) {
        if (!peer_buffers[r] || (peer_buffers[r] & 15u)) return fail(h, "peer buffer null or misaligned");
        h->peer_buf[r] = peer_buffers[r];
    }
    h->world = world; h->rank = rank; h->merge_seq = 0;
    return 0;
}

extern "C" int pomcp_merged_root_stats(pomcp_handle *h, int64_t *n, int64_t *qfix, uint32_t *legal_mask,
                                       int64_t *sims_done)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    if (h->world <= 1) return pomcp_root_stats(h, n, qfix, legal_mask, sims_done);
    CK(cudaSetDevice(h->device));
    DevState *st = reinterpret_cast<DevState *>(h->stage);
    CK(cudaMemcpyAsync(st, h->dS, DEVSTATE_HEAD, cudaMemcpyDeviceToHost, h->last_stream));
    CK(cudaStreamSynchronize(h->last_stream));
    h->nodes_used = st->node_count;
    h->zeroed_upto = st->node_count;
    if (h->merge_seq == 0 || st->mrg_seq != h->merge_seq) return fail(h, "root merge: a peer did not deliver its statistics");
    for (int a = 0; a < h->n_actions; ++a) { n[a] = st->mrg_n[a]; qfix[a] = st->mrg_q[a]; }
    if (legal_mask) *legal_mask = st->res_legal;
    if (sims_done) *sims_done = st->sims_done;
    return 0;
}

extern "C" int pomcp_node_stats(pomcp_handle *h, const int32_t *path_ao, int len, int64_t *n, int64_t *qfix,
                                uint32_t *legal_mask, int32_t *cell)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    int32_t X;
    CK(cudaMemcpy(&X, &h->dS->root, sizeof(X), cudaMemcpyDeviceToHost));
    for (int i = 0; i < len; ++i) {
        const int a = path_ao[2 * i], o = path_ao[2 * i + 1];
        if (a < 0 || a >= h->n_actions || o < 0 || o >= h->zs) return fail(h, "bad path entry");
        uint32_t c;
        uint32_t *cell_c = reinterpret_cast<uint32_t *>(&h->dS->pad_c[0]);
        pomcp_find_child_kernel<<<1, 1, 0, h->last_stream>>>(h->t, h->zs, X, a, o, cell_c);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->last_stream));
        CK(cudaMemcpy(&c, cell_c, sizeof(c), cudaMemcpyDeviceToHost));
        if (c == 0 || c == CHILD_LOCKED) return 1;
        X = CHILD_ID(c);
    }
    int32_t nn[AS]; long long qq[AS];
    CK(cudaMemcpy(nn, h->t.n + (size_t)X * AS, sizeof(nn), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(qq, h->t.qfix + (size_t)X * AS, sizeof(qq), cudaMemcpyDeviceToHost));
    for (int a = 0; a < h->n_actions; ++a) { n[a] = nn[a]; qfix[a] = qq[a]; }
    if (legal_mask) CK(cudaMemcpy(legal_mask, h->t.legal + X, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (cell) CK(cudaMemcpy(cell, h->t.cell + X, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pomcp_advance(pomcp_handle *h, int action, int obs_good, uint32_t sampled_after, uint32_t move,
                             int64_t max_tries, int32_t *status, int64_t *tries_used, int32_t *accepted)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    if (action < 0 || action >= h->n_actions) return fail(h, "action out of range");
    if (obs_good < -1 || (h->kind == 1 && obs_good >= h->tab.Z)) return fail(h, "observation out of range");
    CK(cudaSetDevice(h->device));
    if (h->nodes_used + 2 > h->cap) return fail(h, "node arena full");
    AdvanceParams p;
    memset(&p, 0, sizeof(p));
    p.S = h->dS; p.model = h->dmodel; p.thr53 = h->dthr; p.t = h->t;
    p.tab = h->tab; p.tt = h->tt;
    p.bag[0] = h->bag[0]; p.bag[1] = h->bag[1];
    p.action = action; p.obs_good = obs_good < 0 ? -1 : h->kind == 1 ? obs_good : (obs_good ? 1 : 0); p.need = h->cfg.max_particle_count; p.n_thr = h->n_thr;
    p.sampled_after = sampled_after; p.move = move;
    p.key0 = (uint32_t)h->cfg.seed; p.key1 = (uint32_t)(h->cfg.seed >> 32);
    p.max_tries = max_tries > 0 ? max_tries : 256ll * p.need;
    cudaStream_t s = h->last_stream;
    if (h->kind == 0) pomcp_advance_kernel<<<1, ADV_THREADS, h->search_smem, s>>>(p);
    else pomcp_advance_tab_kernel<<<1, ADV_THREADS, h->search_smem, s>>>(p);
    CK(cudaGetLastError());
    ++h->launches;
    CK(cudaStreamSynchronize(s));
    DevState st;
    CK(cudaMemcpy(&st, h->dS, DEVSTATE_HEAD, cudaMemcpyDeviceToHost));
    h->nodes_used = st.node_count;
    if (h->zeroed_upto < h->nodes_used) h->zeroed_upto = h->nodes_used;
    h->root_cell = st.adv_root_cell; h->n_bag = st.n_bag; h->bag_sel = st.bag_sel;
    if (status) *status = st.adv_status;
    if (tries_used) *tries_used = st.adv_tries;
    if (accepted) *accepted = st.adv_accepted;
    return st.adv_status == 2 ? 1 : 0;
}

extern "C" int pomcp_root_particles(pomcp_handle *h, uint32_t *out, int capacity, int32_t *n, int32_t *root_cell)
{
    if (!h) return -1;
    if (!h->have_root) return fail(h, "no root belief");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    DevState st;
    CK(cudaMemcpy(&st, h->dS, DEVSTATE_HEAD, cudaMemcpyDeviceToHost));
    const int m = st.n_bag < capacity ? st.n_bag : capacity;
    if (out && m > 0) CK(cudaMemcpy(out, h->bag[st.bag_sel], (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (n) *n = st.n_bag;
    if (root_cell) CK(cudaMemcpy(root_cell, h->t.cell + st.root, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pomcp_generate_steps(pomcp_handle *h, int64_t n, const uint32_t *state, const uint8_t *action,
                                    const uint64_t *u53, const uint32_t *actual, const uint32_t *sampled,
                                    uint32_t *next_state, uint8_t *flags, double *reward)
{
    if (!h) return -1;
    if (!h->have_model) return fail(h, "set the model first");
    if (h->kind != 0) return fail(h, "pomcp_generate_steps is the RockSample harness; use pomcp_tabular_steps");
    if (n <= 0) return 0;
    CK(cudaSetDevice(h->device));
    struct Scratch {                                      // device scratch, released on every exit path
        void *p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Scratch() { for (void *q : p) cudaFree(q); }
    } sc;
    const size_t bytes[8] = {(size_t)n * 4, (size_t)n * 4, (size_t)n, (size_t)n, (size_t)n * 8, (size_t)n * 8, (size_t)n * 4,
                             (size_t)n * 4};
    for (int i = 0; i < 8; ++i) CK(cudaMalloc(&sc.p[i], bytes[i]));
    uint32_t *ds = (uint32_t *)sc.p[0], *dn = (uint32_t *)sc.p[1], *dw0 = (uint32_t *)sc.p[6], *dw1 = (uint32_t *)sc.p[7];
    uint8_t *da = (uint8_t *)sc.p[2], *df = (uint8_t *)sc.p[3];
    uint64_t *du = (uint64_t *)sc.p[4];
    double *dr = (double *)sc.p[5];
    CK(cudaMemcpy(dw0, actual, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw1, sampled, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ds, state, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, action, n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(du, u53, n * 8, cudaMemcpyHostToDevice));
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    pomcp_step_batch_kernel<<<blocks, 256, h->search_smem>>>(h->dmodel, h->dthr, h->n_thr, n, ds, da, du, dw0, dw1, dn, df, dr);
    CK(cudaGetLastError());
    ++h->launches;
    CK(cudaMemcpy(next_state, dn, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(flags, df, n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(reward, dr, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pomcp_tabular_steps(pomcp_handle *h, int64_t n, const uint32_t *state, const uint8_t *action,
                                   const uint32_t *w1, const uint32_t *w2, uint32_t *next_state, int32_t *obs,
                                   double *reward, uint8_t *terminal)
{
    if (!h) return -1;
    if (!h->have_model || h->kind != 1) return fail(h, "set a tabular model first");
    if (n <= 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (state[i] >= (uint32_t)h->tab.S || action[i] >= h->tab.A) return fail(h, "state or action out of range");
    CK(cudaSetDevice(h->device));
    struct Scratch {                                      // device scratch, released on every exit path
        void *p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Scratch() { for (void *q : p) cudaFree(q); }
    } sc;
    const size_t bytes[8] = {(size_t)n * 4, (size_t)n, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4,
                             (size_t)n * 8, (size_t)n};
    for (int i = 0; i < 8; ++i) CK(cudaMalloc(&sc.p[i], bytes[i]));
    CK(cudaMemcpy(sc.p[0], state, bytes[0], cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sc.p[1], action, bytes[1], cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sc.p[2], w1, bytes[2], cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sc.p[3], w2, bytes[3], cudaMemcpyHostToDevice));
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    pomcp_tab_step_batch_kernel<<<blocks, 256>>>(h->tab, h->tt, n, (const uint32_t *)sc.p[0], (const uint8_t *)sc.p[1],
                                                 (const uint32_t *)sc.p[2], (const uint32_t *)sc.p[3], (uint32_t *)sc.p[4],
                                                 (int32_t *)sc.p[5], (double *)sc.p[6], (uint8_t *)sc.p[7]);
    CK(cudaGetLastError());
    ++h->launches;
    CK(cudaMemcpy(next_state, sc.p[4], bytes[4], cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(obs, sc.p[5], bytes[5], cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(reward, sc.p[6], bytes[6], cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(terminal, sc.p[7], bytes[7], cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pomcp_debug_steplog(pomcp_handle *h, int64_t *out /* [128][4] */)
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    CK(cudaMemcpy(out, h->dS->steplog, sizeof(long long) * 128 * 8, cudaMemcpyDeviceToHost));

    return 0;
}

extern "C" int pomcp_counters(pomcp_handle *h, int64_t out[16])
{
    if (!h) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    DevState st;
    CK(cudaMemcpy(&st, h->dS, DEVSTATE_HEAD, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; ++i) out[i] = st.counters[i];
    out[3] = st.node_count;
    out[9] = h->launches;
    out[10] = h->coop_blocks;
    out[11] = SEARCH_THREADS;
    out[13] = st.phase_ns[1];                                       // descent + rollout phases of the last search
    out[14] = st.phase_ns[4];                                       // fused merge: wait for the slowest peer (last search)
    out[15] = st.check_failed;                                      // POMCP_CHECKS builds
    out[12] = (int64_t)DEVSTATE_HEAD;                               // D2H bytes of pomcp_root_stats
    return 0;
}
