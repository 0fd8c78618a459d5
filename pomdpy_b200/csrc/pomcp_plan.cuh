// pomcp_plan.cuh -- grid barrier and small helpers, and the per-node action plans of a generation: batch-UCB
// allocation (many expected arrivals), virtual-pull sequence (few), and the lookup a simulation makes.
#pragma once
#include "pomcp_state.cuh"

// Grid-wide barrier of the cooperative search kernel: one arrival counter that only grows (zeroed by the host before
// each launch); block b has passed `epoch` barriers, so barrier e completes when the counter reaches (e+1)*blocks.
// Waiting threads back off with nanosleep instead of hammering L2, which matters in the phases where a handful of
// warps work and ~600 blocks wait.  The cooperative launch guarantees that all blocks are resident.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int nblocks, unsigned int &epoch)
{
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) {
        __threadfence();                                   // release this block's writes
        atomicAdd(counter, 1u);
        const unsigned int target = epoch * nblocks;
        while (*((volatile unsigned int *)counter) < target) __nanosleep(32);
        __threadfence();                                   // acquire (also invalidates this SM's L1)
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// mean return of an edge in single precision (the UCB arithmetic is fp32; sums stay exact integers)
__device__ __forceinline__ float edge_mean(long long qfix, int n)
{
    return ((float)qfix * (1.0f / 16777216.0f)) / (float)n;
}

// Publish the root's edge statistics in the state block (threads 0..AS-1 of one block): pomcp_root_stats is then a
// single device-to-host copy.
__device__ __forceinline__ void publish_root(DevState *S, const Tree &t, int root, int tid)
{
    if (tid < AS) {
        S->res_n[tid] = t.n[(size_t)root * AS + tid];
        S->res_q[tid] = t.qfix[(size_t)root * AS + tid];
    }
    if (tid == 0) { S->res_legal = t.legal[root]; S->res_valid = 1; }
}

// ---------------------------------------------------------------------------------------------------------
// Action plans.  Within a generation the statistics are read-only, so the action distribution of a
// node is a pure function of (its statistics, k), k = the number of simulations of the generation EXPECTED there:
//     k(root) = generation width;   k(child of X under (a, o)) = k(X) * P_plan(a | X) * N(child) / n(X, a).
// Nothing depends on how many simulations really meet at a node, so every simulation descends the tree on its own --
// no level barriers -- and whichever warp reaches a node first computes its plan (lane = action) and publishes it.
// ---------------------------------------------------------------------------------------------------------
struct PlanRec { uint4 head, data; uint32_t th; };   // th: this lane's threshold (kind 2, warp-computed)

// Batch-UCB allocation, k >= 16.5: water-filling of k pulls over the UCB1 indices mean + c*sqrt(ln(N+1)/n); output:
// 32-bit cumulative thresholds in t.thr[X], every G-th of them in the plan record.
__device__ __forceinline__ PlanRec warp_allocation(const SearchParams &p, int X, float kd, int lane, int A, uint32_t stamp)
{
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t mask = p.t.legal[X];
    const bool legal = lane < A && ((mask >> lane) & 1u);
    const int n = legal ? p.t.n[(size_t)X * AS + lane] : 0;
    const long long qf = legal ? p.t.qfix[(size_t)X * AS + lane] : 0;
    int N = n;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) N += __shfl_xor_sync(full, N, off);
    const int untried = __popc(__ballot_sync(full, legal && n == 0));
    const int nlegal = __popc(__ballot_sync(full, legal));
    const float cf = (float)p.ucb_c;
    float v = 0.0f;
    if (N == 0) v = legal ? 1.0f : 0.0f;
    else if ((float)untried >= kd) v = (legal && n == 0) ? 1.0f : 0.0f;
    else {
        const float L = det_logf_u32((uint32_t)N + 1u);
        const float c2L = (cf * cf) * L;
        const float qbar = legal ? (n == 0 ? 0.0f : edge_mean(qf, n)) : -INFINITY;
        float qmax = qbar;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) qmax = fmaxf(qmax, __shfl_xor_sync(full, qmax, off));
        const unsigned best_lanes = __ballot_sync(full, legal && qbar == qmax);
        const int nbest = __shfl_sync(full, n, __ffs((int)best_lanes) - 1);
        if (!(c2L > 0.0f)) v = (legal && qbar == qmax) ? 1.0f : 0.0f;
        else {
            float lo = qmax + 0.999f * sqrtf(c2L / (kd + (float)nbest));
            float hi = qmax + 1.001f * sqrtf((c2L * (float)nlegal) / kd) + 1.0f;
            for (int it = 0; it <= BISECT_ITERS; ++it) {
                const float mid = it < BISECT_ITERS ? 0.5f * (lo + hi) : lo;
                float t = 0.0f;
                if (legal) {
                    const float d = mid - qbar;
                    const float mm = c2L / (d * d);
                    if (n == 0) t = mm > 1.0f ? mm : 1.0f;
                    else { const float e = mm - (float)n; t = e > 0.0f ? e : 0.0f; }
                }
                if (it == BISECT_ITERS) { v = t; break; }
                float s = t;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(full, s, off);
                if (s >= kd) lo = mid; else hi = mid;
            }
        }
    }
    float x = v;                                    // inclusive scan, Hillis-Steele order
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float y = __shfl_up_sync(full, x, off);
        if (lane >= off) x = x + y;
    }
    const float V = __shfl_sync(full, x, 31);
    const float f = (x / V) * 4294967296.0f;
    // running maximum over the actions with positive weight: non-decreasing thresholds, and an action without
    // weight repeats its predecessor's (prefix sums of different lanes round differently), so it is never drawn
    uint32_t th = v > 0.0f ? (f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f) : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(full, th, off);
        if (lane >= off && y > th) th = y;
    }
    __stcg(&p.t.thr[(size_t)X * 32 + lane], th);
    const unsigned pos = __ballot_sync(full, v > 0.0f);
    // every G-th threshold goes into the node's plan record: a simulation finds its group of G actions there and
    // needs one more access for the thresholds inside the group
    const int G = (A + 5) / 6;
    uint32_t cg[6];
#pragma unroll
    for (int g = 0; g < 6; ++g) { const int src = g * G + G - 1; cg[g] = __shfl_sync(full, th, src < 31 ? src : 31); }
    PlanRec r;
    r.head = make_uint4((uint32_t)((31 - __clz((int)pos)) | (PLAN_THRESHOLDS << 16)), stamp, cg[4], cg[5]);
    r.data = make_uint4(cg[0], cg[1], cg[2], cg[3]);
    r.th = th;
    return r;
}

// Plan of a node for k = round(expected arrivals) <= SEQ_ALLOC_MAX (warp-cooperative, lane = action).
//   kind 0  uniform draw over an action mask:  at least k untried actions (n == 0 scores +inf, pomcp.py:64-65), or
//           k == 1: the UCB1 arg-max set (action_selectors.py:6-37; the draw breaks ties), or c = 0: the best means
//   kind 1  the virtual-pull sequence "pull j takes the UCB1 arg-max (lowest id on ties) given the statistics plus
//           the virtual visits of pulls 0..j-1"; a simulation takes pull j = floor(x0*k/2^32)
__device__ __forceinline__ PlanRec warp_plan_few(const SearchParams &p, int X, uint32_t k, int lane, int A, uint32_t stamp)
{
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t mask = p.t.legal[X];
    const bool legal = lane < A && ((mask >> lane) & 1u);
    const int n = legal ? p.t.n[(size_t)X * AS + lane] : 0;
    const long long qf = legal ? p.t.qfix[(size_t)X * AS + lane] : 0;
    int N = n;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) N += __shfl_xor_sync(full, N, off);
    PlanRec r;
    r.head = make_uint4(PLAN_MASK << 16, stamp, 0u, 0u);
    r.data = make_uint4(0u, 0u, 0u, 0u);
    r.th = 0u;
    const unsigned untried = __ballot_sync(full, legal && n == 0);
    if ((uint32_t)__popc(untried) >= k) { r.data.x = untried; return r; }
    const float L = det_logf_u32((uint32_t)N + 1u);
    const float cf = (float)p.ucb_c;
    if (k == 1) {                                                    // no untried action here: every legal n > 0
        const float sc = legal ? edge_mean(qf, n) + cf * sqrtf(L / (float)n) : -INFINITY;
        float best = sc;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) best = fmaxf(best, __shfl_xor_sync(full, best, off));
        r.data.x = __ballot_sync(full, legal && sc == best);
        return r;
    }
    const float c2L = (cf * cf) * L;
    const float qb = n == 0 ? 0.0f : edge_mean(qf, n);
    if (!(c2L > 0.0f)) {
        float qmax = legal ? qb : -INFINITY;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) qmax = fmaxf(qmax, __shfl_xor_sync(full, qmax, off));
        r.data.x = __ballot_sync(full, legal && qb == qmax);
        return r;
    }
    float sc = legal ? (n == 0 ? INFINITY : qb + cf * sqrtf(L / (float)n)) : -INFINITY;
    int va = 0;
    uint32_t seq[4] = {0u, 0u, 0u, 0u};
    for (uint32_t pull = 0; pull < k; ++pull) {
        float bs = sc;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) bs = fmaxf(bs, __shfl_xor_sync(full, bs, off));
        const int w = __ffs((int)__ballot_sync(full, sc == bs)) - 1;
        seq[pull >> 2] |= (uint32_t)w << (8 * (pull & 3u));
        if (lane == w) { va += 1; sc = qb + cf * sqrtf(L / (float)(n + va)); }
    }
    r.head.x = (PLAN_PULLS << 16) | (k << 8);
    r.data = make_uint4(seq[0], seq[1], seq[2], seq[3]);
    return r;
}

// The same plan computed by ONE thread (the fringe of the tree: most nodes there expect one or two arrivals, and the
// 32 simulations of a warp stand at 32 different nodes -- lane = simulation beats lane = action).  Identical
// arithmetic, hence identical plans: sums and maxima do not depend on the order, ties go to the lowest action.
__device__ __forceinline__ PlanRec thread_plan_few(const SearchParams &p, int X, uint32_t k, int A, uint32_t stamp)
{
    const uint32_t mask = p.t.legal[X];
    const int32_t *nrow = p.t.n + (size_t)X * AS;
    const long long *qrow = p.t.qfix + (size_t)X * AS;
    int N = 0; uint32_t untried = 0u;
#pragma unroll 1
    for (int a = 0; a < A; ++a) {
        if (!((mask >> a) & 1u)) continue;
        const int n = nrow[a];
        N += n;
        if (n == 0) untried |= 1u << a;
    }
    PlanRec r;
    r.head = make_uint4(PLAN_MASK << 16, stamp, 0u, 0u);
    r.data = make_uint4(0u, 0u, 0u, 0u);
    r.th = 0u;
    if ((uint32_t)__popc(untried) >= k) { r.data.x = untried; return r; }
    const float L = det_logf_u32((uint32_t)N + 1u);
    const float cf = (float)p.ucb_c;
    if (k == 1) {                                                    // UCB1 arg-max set
        float best = -INFINITY; uint32_t ties = 0u;
    #pragma unroll 1
    for (int a = 0; a < A; ++a) {
            if (!((mask >> a) & 1u)) continue;
            const int n = nrow[a];
            const float sc = edge_mean(qrow[a], n) + cf * sqrtf(L / (float)n);
            if (sc > best) { best = sc; ties = 1u << a; } else if (sc == best) ties |= 1u << a;
        }
        r.data.x = ties;
        return r;
    }
    const float c2L = (cf * cf) * L;
    float sc[AS];
    float qmax = -INFINITY; uint32_t qties = 0u;
#pragma unroll 1
    for (int a = 0; a < A; ++a) {
        sc[a] = -INFINITY;
        if (!((mask >> a) & 1u)) continue;
        const int n = nrow[a];
        const float qb = n == 0 ? 0.0f : edge_mean(qrow[a], n);
        if (qb > qmax) { qmax = qb; qties = 1u << a; } else if (qb == qmax) qties |= 1u << a;
        sc[a] = n == 0 ? INFINITY : qb + cf * sqrtf(L / (float)n);
    }
    if (!(c2L > 0.0f)) { r.data.x = qties; return r; }
    unsigned long long seq_lo = 0ull, seq_hi = 0ull;                 // the pull bytes (no indexed local array)
#pragma unroll 1
    for (uint32_t pull = 0; pull < k; ++pull) {
        int w = 0; float bs = sc[0];
#pragma unroll 1
        for (int a = 1; a < A; ++a) if (sc[a] > bs) { bs = sc[a]; w = a; }
        if (pull < 8u) seq_lo |= (unsigned long long)w << (8 * pull); else seq_hi |= (unsigned long long)w << (8 * (pull - 8u));
        int va = 0;                                                  // pulls of w so far, this one included
#pragma unroll 1
        for (uint32_t j = 0; j <= pull; ++j)
            va += (int)(((j < 8u ? seq_lo >> (8 * j) : seq_hi >> (8 * (j - 8u))) & 255ull) == (unsigned long long)w);
        const int n = nrow[w];
        const float qb = n == 0 ? 0.0f : edge_mean(qrow[w], n);
        sc[w] = qb + cf * sqrtf(L / (float)(n + va));
    }
    r.head.x = (PLAN_PULLS << 16) | (k << 8);
    r.data = make_uint4((uint32_t)seq_lo, (uint32_t)(seq_lo >> 32), (uint32_t)seq_hi, (uint32_t)(seq_hi >> 32));
    return r;
}

// The action of one simulation from its node's plan and its 32-bit draw x0, and the plan's probability of that
// action.  Thresholds come from the block's shared-memory copy (sthr) if the node is cached there, else from L2.
__device__ __forceinline__ int plan_action(const uint4 head, const uint4 data, const uint32_t *sthr, const uint32_t *gthr,
                                           uint32_t x0, int A, float &prob)
{
    const int meta = (int)head.x, kind = meta >> 16;
    if (kind == PLAN_THRESHOLDS) {                       // first action whose threshold exceeds x0
        const int G = (A + 5) / 6;                       // thresholds are non-decreasing: find the group of G ...
        const int g = x0 < data.x ? 0 : x0 < data.y ? 1 : x0 < data.z ? 2 : x0 < data.w ? 3 : x0 < head.z ? 4 :
                      x0 < head.w ? 5 : 6;
        int a = meta & 255;                              // the last action with weight, if no threshold exceeds x0
        uint32_t hi_t, lo_t;
        if (g < 6) {
            // ... then the action: the (at most 6) thresholds of the group, as scalars (no indexed local array)
            const int b0 = g * G;
#define PLAN_TQ(q) (((q) < G && b0 + (q) < 32) ? (sthr ? sthr[b0 + (q)] : __ldcg(gthr + b0 + (q))) : 0xFFFFFFFFu)
            const uint32_t t0 = PLAN_TQ(0), t1 = PLAN_TQ(1), t2 = PLAN_TQ(2), t3 = PLAN_TQ(3), t4 = PLAN_TQ(4), t5 = PLAN_TQ(5);
#undef PLAN_TQ
            const uint32_t below = g == 0 ? 0u : g == 1 ? data.x : g == 2 ? data.y : g == 3 ? data.z : g == 4 ? data.w : head.z;
            const int q = x0 < t0 ? 0 : x0 < t1 ? 1 : x0 < t2 ? 2 : x0 < t3 ? 3 : x0 < t4 ? 4 : x0 < t5 ? 5 : 6;
            if (q < 6) {
                hi_t = q == 0 ? t0 : q == 1 ? t1 : q == 2 ? t2 : q == 3 ? t3 : q == 4 ? t4 : t5;
                lo_t = q == 0 ? below : q == 1 ? t0 : q == 2 ? t1 : q == 3 ? t2 : q == 4 ? t3 : t4;
                prob = (float)(hi_t - lo_t) * (1.0f / 4294967296.0f);
                return b0 + q;
            }
        }
        hi_t = sthr ? sthr[a] : __ldcg(gthr + a);
        lo_t = a == 0 ? 0u : (sthr ? sthr[a - 1] : __ldcg(gthr + a - 1));
        prob = (float)(hi_t - lo_t) * (1.0f / 4294967296.0f);
        return a;
    }
    if (kind == PLAN_MASK) {
        const uint32_t m = (uint32_t)__popc(data.x);
        prob = 1.0f / (float)m;
        return nth_bit(data.x, randbelow32(x0, m));
    }
    const uint32_t k = (uint32_t)(meta >> 8) & 255u;
    const uint32_t j = randbelow32(x0, k);
    const uint32_t word = (j >> 2) == 0 ? data.x : (j >> 2) == 1 ? data.y : (j >> 2) == 2 ? data.z : data.w;
    const uint32_t a = (word >> (8 * (j & 3u))) & 255u;
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const uint32_t wd = (t >> 2) == 0 ? data.x : (t >> 2) == 1 ? data.y : (t >> 2) == 2 ? data.z : data.w;
        cnt += ((uint32_t)t < k && ((wd >> (8 * (t & 3))) & 255u) == a) ? 1 : 0;
    }
    prob = (float)cnt / (float)k;
    return (int)a;
}
