// pomcp_models.cuh -- model policies of the planner kernels (RockSample, tabular POMDPs): generative step, action set
// of a node, rollout inner loop.
#pragma once
#include "pomcp_state.cuh"

// ---------------------------------------------------------------------------------------------------------
// Model policies.  The planner kernels are templates over KIND: 0 = RockSample (tables staged in shared memory, the
// reference's flagship and the benchmark), 1 = tabular POMDP (tables in global memory).  A policy provides the
// generative step, the action set of a node tag and the rollout inner loop; everything else is model-agnostic.
// ---------------------------------------------------------------------------------------------------------
struct StepOut { uint32_t next, tag; int obs, rew, terminal; double reward; };

template <int KIND> struct Mdl;

template <> struct Mdl<0> {
    struct Ctx { SmemModel M; uint32_t actual, sampled; };
    static __device__ __forceinline__ size_t stage(unsigned char *smem, const ModelHeader *gm, const uint64_t *gthr,
                                                   int n_thr, const TabHeader &, const TabTables &, Ctx &c)
    {   // stage the model tables into shared memory; returns the first free byte offset (16-byte aligned)
        static_assert(sizeof(ModelHeader) % 16 == 0, "ModelHeader is staged in 16-byte pieces");
        const int nw = (int)(sizeof(ModelHeader) / 16);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        const uint4 *src = reinterpret_cast<const uint4 *>(gm);
        for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
        // The rollouts draw only the top 32 bits of the sensor uniform ((w << 21) <= thr53  <=>  w <= thr53 >> 21;
        // a certain sensor, thr53 = 2^53, saturates), so the shared-memory copy is the 32-bit table; tree levels and
        // the belief update read the 53-bit thresholds from global memory (a few per simulation, L1-resident).
        size_t off = (sizeof(ModelHeader) + 15) & ~size_t(15);
        uint32_t *thr32 = reinterpret_cast<uint32_t *>(smem + off);
        for (int i = threadIdx.x; i < n_thr; i += blockDim.x) {
            const uint64_t tq = gthr[i] >> 21;
            thr32[i] = tq > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)tq;
        }
        off += ((size_t)n_thr * 4 + 15) & ~size_t(15);
        c.M.hdr = reinterpret_cast<const ModelHeader *>(smem);
        c.M.thr53 = gthr;
        c.M.thr32 = thr32;
        return off;
    }
    static __device__ __forceinline__ int n_actions(const Ctx &c, const ModelHeader *gm) { return gm->n_actions; }
    static __device__ __forceinline__ int zs(const Ctx &) { return 2; }
    static __device__ __forceinline__ int n_obs(const Ctx &) { return 2; }
    static __device__ __forceinline__ bool stage_d1(const Ctx &) { return true; }
    static __device__ __forceinline__ uint32_t state_tag(uint32_t st) { return st & 0xFFu; }
    static __device__ __forceinline__ uint32_t tag_legal(const Ctx &c, uint32_t tag) { return c.M.hdr->legal[tag]; }
    static __device__ __forceinline__ bool tag_ok(const Ctx &c, uint32_t tag) { return tag < (uint32_t)c.M.hdr->n_cells; }
    static __device__ __forceinline__ StepOut step(const Ctx &c, uint32_t st, int a, uint32_t wy, uint32_t wz)
    {
        const StepResult sr = rs_step(c.M, st & 0xFFu, st >> 8, a, u53_of(wy, wz), c.actual, c.sampled);
        return StepOut{sr.cell | (sr.rocks << 8), sr.cell, sr.good, sr.rew, sr.terminal, 0.0};
    }
    static __device__ __forceinline__ uint16_t path_code(int a, const StepOut &o)
    { return (uint16_t)(a | (o.obs << 5) | (o.rew << 8)); }
    static __device__ __forceinline__ int path_obs(uint32_t ar) { return (int)((ar >> 5) & 1u); }
    static __device__ __forceinline__ double path_reward(const Ctx &c, uint32_t ar, const double *) { return c.M.hdr->rew[ar >> 8]; }
    // Rollout steps t .. t_end-1 from packed state st (uniform legal actions); true when the episode ended.
    // One Philox block per two steps: (x, y) then (z, w); word 0 picks the action, word 1 is the top 32 bits of the
    // 53-bit sensor uniform.
    // Everything by value.  (-DROLL_NOINLINE gives the 100-step loop a register allocation of its own; measured: the
    // caller then saves its live registers around the call and the local-memory traffic doubles -- not the default.)
    struct RollOut { double sum; int t; };
#ifdef ROLL_NOINLINE
#define ROLL_INLINE __noinline__
#else
#define ROLL_INLINE __forceinline__
#endif
    static __device__ ROLL_INLINE RollOut rollout(const ModelHeader *H, const uint32_t *thr32, uint32_t actual, uint32_t sampled,
                                                   uint32_t sim, uint32_t move, const uint32_t *ks, int depth,
                                                   const double *disc, uint32_t st, int t_end)
    {
        int t = 0; double sum = 0.0;
        uint32_t cell = st & 0xFFu, rocks = st >> 8;
        bool finished = false;
        // one step with the draws (w0, w1); true when the episode ended
        auto step = [&](uint32_t w0, uint32_t w1) -> bool {
            const uint32_t mask = H->legal[cell];
            const int a = nth_legal(H, mask, randbelow32(w0, (uint32_t)__popc(mask)));
            if (a < 4) {
                cell = (uint32_t)((int)cell + H->delta[a]);
                if (H->cell[cell] == POMCP_CELL_GOAL) { sum += H->rew[REW_EXIT] * disc[t]; return true; }
            } else if (a == 4) {
                const int code = H->cell[cell];
                const int rew = (((actual & rocks) >> code) & 1u) ? REW_GOOD : REW_BAD;
                rocks &= ~(1u << code);
                sum += H->rew[rew] * disc[t];
            } else {
                const int k = a - 5;
                if (!((sampled >> k) & 1u)) {
                    const uint32_t truth = (actual >> k) & 1u;
                    const uint32_t correct = w1 <= thr32[cell * H->n_rocks + k];
                    const uint32_t good = correct ? truth : (truth ^ 1u);
                    rocks = (rocks & ~(1u << k)) | (good << k);
                }
            }
            return false;
        };
        // t is even on entry (slices are even, a finished rollout is never resumed): two steps per Philox block
        while (t < t_end) {
            const Philox4 r = philox4x32_10_ks<POMCP_ROLL_ROUNDS>((uint32_t)(depth + 1 + (t >> 1)), sim, move, DOM_SIM, ks);
            if (step(r.x, r.y)) { ++t; finished = true; break; }
            if (++t >= t_end) break;                                 // odd horizon
            if (step(r.z, r.w)) { ++t; finished = true; break; }
            ++t;
        }
        (void)finished;
        return RollOut{sum, t};
    }
    // Preferred-action rollout: the rollout carries the rock statistics of its own history -- the
    // leaf's, updated with every (action, observation) it generates (rock_position_history.py:96-137) -- and draws
    // uniformly from generate_smart_actions (:170-234) instead of the legal set.  ch / gd: the statistics after the
    // last tree step; mask: the smart-action set there.  Runs to the end (no slices).
    struct PrefOut { uint32_t st; int t; double sum; int finished; };
    static __device__ __noinline__ PrefOut rollout_preferred(SmemModel M, uint32_t actual, uint32_t sampled,
                                                             const double *prob, const double *leaf_chance,
                                                             const int16_t *leaf_good, int leaf_cell, int leaf_action,
                                                             int leaf_obs, uint32_t sim, uint32_t move, uint32_t key0,
                                                             uint32_t key1, int depth, const double *disc, uint32_t st,
                                                             int t, int t_end, double sum)
    {   // everything by value: a reference to the caller's locals (or to the kernel parameters) would push them into
        // local memory on the hot path as well
        const ModelHeader *H = M.hdr;
        // The leaf may have been created in this very phase by another SM: its rows are read from L2 (its neighbours'
        // rows share cache lines with it, and an earlier read of those may have left a stale copy in this SM's L1)
        double ch[POMCP_RS]; int16_t gd[POMCP_RS];
        for (int k = 0; k < H->n_rocks; ++k) { ch[k] = __ldcg(leaf_chance + k); gd[k] = __ldcg(leaf_good + k); }
        uint32_t mask = child_smart_mask(H, prob, ch, gd, ch, gd, leaf_cell, leaf_action, leaf_obs, (int)(st & 0xFFu));
        Philox4 r{0u, 0u, 0u, 0u};
        for (; t < t_end; ++t) {
            uint32_t w0, w1;
            if ((t & 1) == 0) {
                r = philox4x32_10<POMCP_ROLL_ROUNDS>((uint32_t)(depth + 1 + (t >> 1)), sim, move, DOM_SIM, key0, key1);
                w0 = r.x; w1 = r.y;
            } else { w0 = r.z; w1 = r.w; }
            const int a = nth_bit(mask, randbelow32(w0, (uint32_t)__popc(mask)));
            const StepResult sr = rs_step(M, st & 0xFFu, st >> 8, a, (uint64_t)w1 << 21, actual, sampled);
            const double rew = H->rew[sr.rew];
            if (rew != 0.0) sum += rew * disc[t];
            mask = child_smart_mask(H, prob, ch, gd, ch, gd, (int)(st & 0xFFu), a, sr.good, (int)sr.cell);
            st = sr.cell | (sr.rocks << 8);
            if (sr.terminal) return PrefOut{st, t + 1, sum, 1};
        }
        return PrefOut{st, t, sum, 0};
    }
};

template <> struct Mdl<1> {
    struct Ctx { TabHeader h; TabTables t; };
    static __device__ __forceinline__ size_t stage(unsigned char *, const ModelHeader *, const uint64_t *, int,
                                                   const TabHeader &h, const TabTables &t, Ctx &c)
    { c.h = h; c.t = t; return 0; }
    static __device__ __forceinline__ int n_actions(const Ctx &c, const ModelHeader *) { return c.h.A; }
    static __device__ __forceinline__ int zs(const Ctx &c) { return c.h.ZS; }
    static __device__ __forceinline__ int n_obs(const Ctx &c) { return c.h.Z; }
    static __device__ __forceinline__ bool stage_d1(const Ctx &c) { return c.h.stage_d1 != 0; }
    static __device__ __forceinline__ uint32_t state_tag(uint32_t) { return 0u; }
    static __device__ __forceinline__ uint32_t tag_legal(const Ctx &c, uint32_t)
    { return c.h.A >= 32 ? 0xFFFFFFFFu : ((1u << c.h.A) - 1u); }
    static __device__ __forceinline__ bool tag_ok(const Ctx &, uint32_t tag) { return tag == 0u; }
    static __device__ __forceinline__ int next_state(const Ctx &c, size_t row, uint32_t w)
    {
        if (c.h.guide_bits)
            return tab_sample_guided(c.t.Tc + row * c.h.S, c.t.Tg + row * ((size_t)(1u << c.h.guide_bits) + 1), c.h.guide_bits, w);
        return tab_sample(c.t.Tc + row * c.h.S, c.h.S, w);
    }
    static __device__ __forceinline__ StepOut step(const Ctx &c, uint32_t st, int a, uint32_t wy, uint32_t wz)
    {
        const size_t row = (size_t)a * c.h.S + st;
        const int ns = next_state(c, row, wy);
        const int ob = tab_sample(c.t.Oc + ((size_t)a * c.h.S + ns) * c.h.Z, c.h.Z, wz);
        const int term = (int)((c.h.term_action_mask >> a) & 1u) | (int)__ldg(c.t.term_state + ns);
        return StepOut{(uint32_t)ns, 0u, ob, 0, term, __ldg(c.t.R + row)};
    }
    static __device__ __forceinline__ uint16_t path_code(int a, const StepOut &o) { return (uint16_t)(a | (o.obs << 5)); }
    static __device__ __forceinline__ int path_obs(uint32_t ar) { return (int)(ar >> 5); }
    static __device__ __forceinline__ double path_reward(const Ctx &, uint32_t, const double *rew) { return *rew; }
    // word 0 picks the action (all actions are legal), word 1 the next state; observations are not drawn
    struct RollOut { double sum; int t; };
    static __device__ ROLL_INLINE RollOut rollout(Ctx c, uint32_t sim, uint32_t move, const uint32_t *ks, int depth,
                                                   const double *disc, uint32_t st, int t_end)
    {
        Philox4 r{0u, 0u, 0u, 0u};
        int t = 0; double sum = 0.0;
        for (; t < t_end; ++t) {
            uint32_t w0, w1;
            if ((t & 1) == 0) {
                r = philox4x32_10_ks<POMCP_ROLL_ROUNDS>((uint32_t)(depth + 1 + (t >> 1)), sim, move, DOM_SIM, ks);
                w0 = r.x; w1 = r.y;
            } else { w0 = r.z; w1 = r.w; }
            const int a = (int)randbelow32(w0, (uint32_t)c.h.A);
            const size_t row = (size_t)a * c.h.S + st;
            const double rew = __ldg(c.t.R + row);
            st = (uint32_t)next_state(c, row, w1);
            if (rew != 0.0) sum += rew * disc[t];
            if (((c.h.term_action_mask >> a) & 1u) | __ldg(c.t.term_state + st)) { ++t; break; }
        }
        return RollOut{sum, t};
    }
};
