// pomcp_device.cuh -- device-side building blocks of the B200 POMCP planner (sm_100a).
//
// Everything here is integer work or IEEE-exact floating point (+,-,*,/,sqrt; no FMA contraction: the file is built
// with -fmad=false) so that a search is a deterministic function of its inputs.
#pragma once
#include <stdint.h>

#define POMCP_AS 32             // action stride of the SoA node arena: one warp lane per action
#define POMCP_CELL_GOAL (-2)
#define POMCP_CELL_OBSTACLE (-3)

enum { DOM_SIM = 0, DOM_UPDATE = 1, DOM_FINAL = 2, DOM_AGENT = 3 };
enum { REW_NONE = 0, REW_ILLEGAL = 1, REW_EXIT = 2, REW_GOOD = 3, REW_BAD = 4 };

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator: key = seed, counter = (block k, simulation id, move, domain).
// The rollout stream (the blocks that drive the random steps below a leaf) runs POMCP_ROLL_ROUNDS = 7 rounds, the
// fewest Salmon et al. found Crush-resistant (oracle: ORC_ROLLOUT_ROUNDS; Random123 known answers in
// tests/test_cpu_oracle.py): the rollouts are 85 % of a search's instructions, Philox 26 of the 88 warp instructions of
// a rollout step (2 IMAD.WIDE + 2 LOP3 per round).  Measured: a fresh-tree search of 10^6 simulations 0.483 -> 0.451 ms.
// ---------------------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
#define POMCP_ROLL_ROUNDS 7

template <int ROUNDS = 10>
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
// The same generator with the key schedule precomputed (ks[2r], ks[2r+1] = key of round r): in the rollout loop the
// schedule comes from the kernel's constant bank instead of 20 uniform adds per block.
template <int ROUNDS = 10>
__device__ __forceinline__ Philox4 philox4x32_10_ks(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    const uint32_t *ks)
{
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ ks[2 * r], n2 = hi0 ^ c3 ^ ks[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ uint32_t randbelow32(uint32_t x, uint32_t n) { return __umulhi(x, n); }
__device__ __forceinline__ uint64_t u53_of(uint32_t a, uint32_t b)
{ return ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6); }

// ---------------------------------------------------------------------------------------------------------
// Deterministic single-precision ln of a positive integer: exponent by clz, 2*atanh series in Horner form, ln2 split
// in two floats.  IEEE +,-,*,/ only (no FMA contraction), so the CPU oracle reproduces it bit for bit.  The UCB
// arithmetic is fp32 because fp64 issue is a scarce resource on this chip (profiles/README.md).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float det_logf_u32(uint32_t x)
{
    int e = 31 - __clz((int)x);
    float m = (float)x / (float)(1u << e);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    const float f = (m - 1.0f) / (m + 1.0f);
    const float f2 = f * f;
    float s = 1.0f / 11.0f;
    s = s * f2 + 1.0f / 9.0f; s = s * f2 + 1.0f / 7.0f; s = s * f2 + 1.0f / 5.0f; s = s * f2 + 1.0f / 3.0f;
    s = s * f2 + 1.0f;
    const float lnm = 2.0f * f * s;
    return (float)e * 0.693145751953125f + ((float)e * 1.42860682030941723212e-6f + lnm);
}

// ---------------------------------------------------------------------------------------------------------
// RockSample tables as staged in shared memory (reference examples/rock_sample/rock_model.py).
// ---------------------------------------------------------------------------------------------------------
struct ModelHeader {            // lives in global memory; copied verbatim into shared memory
    int32_t rows, cols, n_rocks, n_actions, n_cells, start_cell, pad0, pad1;
    double rew[8];              // indexed by REW_*: 0, illegal, exit, good, bad
    int32_t delta[4];           // cell offset of N, E, S, W
    int8_t cell[256];           // rock k >= 0 | EMPTY -1 | GOAL -2 | OBSTACLE -3
    uint32_t legal[256];        // legal-action bit mask of each cell (rock_model.py:218-247)
    uint8_t nth5[32 * 8];       // nth5[m*8 + j] = index of the j-th set bit of the 5-bit mask m
    int32_t rock_cell[24];      // cell of rock k (preferred-action heuristic)
};

struct SmemModel {
    const ModelHeader *hdr;
    const uint64_t *thr53;      // [n_cells][n_rocks] (global memory)
    const uint32_t *thr32;      // thr53 >> 21, saturated: the rollouts' copy in shared memory
};

struct StepResult {
    uint32_t cell, rocks;
    int good, empty, terminal, legal, rew;
};

// generate_step (rock_model.py:451-465) as a pure function of (state, action, world, u53).
__device__ __forceinline__ StepResult rs_step(const SmemModel &M, uint32_t cell, uint32_t rocks, int a,
                                              uint64_t u53, uint32_t actual, uint32_t sampled)
{
    const ModelHeader *H = M.hdr;
    StepResult r;
    const int lg = (a >= 5) ? 1 : (int)((H->legal[cell] >> a) & 1u);          // :323-344
    uint32_t ncell = cell;
    if (a < 4 && lg) ncell = (uint32_t)((int)cell + H->delta[a]);
    uint32_t nr = rocks;
    const int code = H->cell[cell];
    if (a == 4 && lg) nr &= ~(1u << code);                                      // :346-368
    int good = 0;
    if (a >= 5) {                                                               // :370-412
        const int k = a - 5;
        if (!((sampled >> k) & 1u)) {
            const int truth = (actual >> k) & 1u;
            const int correct = u53 <= M.thr53[ncell * H->n_rocks + k];
            good = correct ? truth : !truth;
            nr = good ? (nr | (1u << k)) : (nr & ~(1u << k));
        }
    }
    const int terminal = H->cell[ncell] == POMCP_CELL_GOAL;
    int rew = REW_NONE;                                                         // :417-445
    if (!lg) rew = REW_ILLEGAL;
    else if (terminal) rew = REW_EXIT;
    else if (a == 4) rew = (((actual & rocks) >> code) & 1u) ? REW_GOOD : REW_BAD;
    r.cell = ncell; r.rocks = nr; r.good = good; r.empty = (a < 4); r.terminal = terminal; r.legal = lg; r.rew = rew;
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Tabular POMDP (the reference's Model ABC with explicit matrices, e.g. examples/tiger/tiger_model.py:102-139):
// T[a][s][s'], O[a][s'][z] as 32-bit cumulative thresholds floor(cumsum * 2^32) (last entry 0xFFFFFFFF), R[a][s].
// The tables stay in global memory (L1/L2 resident for small models); a packed state is the state index.
// ---------------------------------------------------------------------------------------------------------
struct TabHeader {
    int32_t S, A, Z, ZS;        // ZS: observation stride of the child table (power of two >= max(Z, 2))
    uint32_t term_action_mask;  // actions that end the episode (Tiger: opening a door)
    int32_t stage_d1;           // depth-1 edges fit the per-block shared-memory staging
    int32_t guide_bits;         // log2 of the guide-table buckets per transition row (0: no guide table)
    int32_t pad1;
};
struct TabTables {
    const uint32_t *Tc;         // [A][S][S]
    const uint32_t *Oc;         // [A][S][Z]  (row: action, next state)
    const double *R;            // [A][S]
    const uint8_t *term_state;  // [S]
    const uint16_t *Tg;         // [A][S][2^guide_bits + 1] guide table: first index whose threshold exceeds the bucket start
};

// first index k with cdf[k] > min(w, 2^32 - 2): zero-probability entries are never drawn
__device__ __forceinline__ int tab_sample(const uint32_t *cdf, int n, uint32_t w)
{
    if (w == 0xFFFFFFFFu) w = 0xFFFFFFFEu;
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(cdf + mid) > w) hi = mid; else lo = mid + 1; }
    return lo;
}

// The same draw with a guide table (an accelerator, not a different rule): bucket j = w >> (32 - bits) can only draw
// an index in [guide[j], guide[j + 1]], so the bisection runs over a handful of entries instead of the whole row.
__device__ __forceinline__ int tab_sample_guided(const uint32_t *cdf, const uint16_t *guide, int bits, uint32_t w)
{
    if (w == 0xFFFFFFFFu) w = 0xFFFFFFFEu;
    const uint32_t j = w >> (32 - bits);
    int lo = __ldg(guide + j), hi = __ldg(guide + j + 1);
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(cdf + mid) > w) hi = mid; else lo = mid + 1; }
    return lo;
}

// index of the j-th (0-based) set bit of a legal mask whose bits >= 5 are all set up to n_actions
__device__ __forceinline__ int nth_legal(const ModelHeader *H, uint32_t mask, uint32_t j)
{
    const uint32_t low = mask & 31u;
    const uint32_t m = (uint32_t)__popc(low);
    return (j < m) ? (int)H->nth5[low * 8 + j] : (int)(5 + (j - m));
}

// j-th set bit of an arbitrary 32-bit mask
__device__ __forceinline__ int nth_bit(uint32_t mask, uint32_t j)
{
    return (int)__fns(mask, 0u, (int)j + 1);
}

// ---------------------------------------------------------------------------------------------------------
// --preferred_actions: per-node rock statistics (examples/rock_sample/rock_position_history.py).  A node keeps, per
// rock, the Bayes estimate chance_good (fp64) and the goodness number; a child copies its parent's and updates the
// one rock the (action, observation) pair concerns (create_child, :96-137); the node's in-tree action set is
// generate_smart_actions (:170-234).  Clean per-node copies (the reference shares one list tree-wide).
// ---------------------------------------------------------------------------------------------------------
#define POMCP_RS 24            // rock stride of the statistics arrays

// Statistics of the child of node X under (a, og) -> its smart-action mask; stored for node `store_id` if >= 0.
static __device__ __noinline__ uint32_t child_smart_mask(const ModelHeader *H, const double *prob, const double *chance,
                                                  const int16_t *good, double *chance_out, int16_t *good_out,
                                                  int pcell, int a, int og, int ccell)
{
    const int K = H->n_rocks;
    int kmod = -1, gdk = 0; double chk = 0.0;
    if (a == 4) {
        const int k = H->cell[pcell];
        if (k >= 0 && k < K) { kmod = k; chk = 0.0; gdk = -10; }
    } else if (a >= 5) {
        kmod = a - 5;
        const double p_ok = prob[pcell * K + kmod], p_no = 1 - p_ok;
        double lg = chance[kmod], lb = 1 - lg;
        gdk = good[kmod];
        if (og) { gdk += 1; lg *= p_ok; lb *= p_no; }
        else { gdk -= 1; lg *= p_no; lb *= p_ok; }
        chk = (fabs(lg) < 0.01 && fabs(lb) < 0.01) ? chance[kmod] : lg / (lg + lb);
    }
    const int code = H->cell[ccell];
    int target = -1; uint32_t checks = 0; bool sample_only = false;
    for (int k = 0; k < K; ++k) {
        const double ch = (k == kmod) ? chk : chance[k];
        const int gd = (k == kmod) ? gdk : (int)good[k];
        if (chance_out) { chance_out[k] = ch; good_out[k] = (int16_t)gd; }
        if (k == code && (ch == 1.0 || gd > 0)) sample_only = true;
        if (target < 0 && ch != 0.0 && gd >= 0) target = k;
        if (ch != 0.0 && ch != 1.0 && gd > -2 && gd < 2) checks |= 1u << (5 + k);
    }
    if (sample_only) return 1u << 4;
    if (target < 0) return 1u << 1;
    const int i = ccell / H->cols, j = ccell - i * H->cols;
    const int ri = H->rock_cell[target] / H->cols, rj = H->rock_cell[target] - ri * H->cols;
    uint32_t mask = checks;
    if (ri > i) mask |= 1u << 2; else if (ri < i) mask |= 1u << 0;
    if (rj > j) mask |= 1u << 1; else if (rj < j) mask |= 1u << 3;
    return mask;
}
