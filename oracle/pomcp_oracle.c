/*
 * oracle/pomcp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the POMCP hot path of
 * pemami4911/POMDPy (the reference is pure Python; citations below are
 * file:line in that tree).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library;
 * nothing under pomdpy_b200/ does.
 *
 * Two algorithms live here:
 *
 *  (1) "ref" -- the reference's *sequential* POMCP, quirks included
 *      (per-level particle resampling, rollout from the pre-action node,
 *      stale mean, input-state mutation on a rewarded SAMPLE; SURVEY.md
 *      Appendix A.1-A.5, A.12).  Random draws come from one sequential
 *      Philox stream so that the executed reference, with random.choice /
 *      random.shuffle / np.random.binomial patched to the same stream
 *      (oracle/refharness/), must agree with it bit for bit.  That
 *      comparison is what pins this file to the reference.
 *
 *  (2) "gs"  -- the generation-synchronous tree-parallel POMCP that the CUDA
 *      product implements: identical model function,
 *      UCB1 rule, expansion rule, rollout policy and backup arithmetic, but
 *      the simulations of a GENERATION -- clamp(growth x visits(root), w0,
 *      wmax) of them -- all descend against the statistics as they were at
 *      the generation's start and are backed up together.  Inside a
 *      generation every node hands out its actions by a PLAN that is a pure
 *      function of its statistics and of the number of arrivals it expects
 *      (k(root) = width; k(child) = k * P_plan(a) * N(child) / n(a), fp32,
 *      fixed operation order): batch-UCB water-filling thresholds for
 *      k >= 16.5, a virtual-pull sequence for 2..16 arrivals, a uniform draw
 *      over the untried actions otherwise.  No simulation depends on what
 *      another one of its generation did, so the result is a deterministic
 *      function of (seed, inputs) for any degree of parallelism.  With
 *      generation width 1 it is the textbook sequential POMCP ("canonical"
 *      in SURVEY.md Appendix D).  The CUDA kernels must reproduce its root
 *      statistics, counters and belief updates exactly.
 *      (orc_gs_set_kmode(1) switches k to the history-only estimate
 *      w * N(X) / N(root): an experiment reported in
 *      profiles/r02_quality.md section 4, never what the kernels compute.)
 *      RNG: Philox4x32-10 counter blocks (seed, block, simulation, move,
 *      domain); the blocks that drive a rollout's random steps run 7 rounds
 *      (ORC_ROLLOUT_ROUNDS).  "gs" has no reference to be pinned to beyond
 *      its model function and its width-1 behaviour; tests/golden/gs_pins.json
 *      guards the specification itself.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, the
 * floating-point sequences below are part of the specification).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_ACTIONS 32
#define ORC_MAX_CELLS 256
#define ORC_MAX_ROCKS 24
#define ORC_CELL_EMPTY (-1)
#define ORC_CELL_GOAL (-2)
#define ORC_CELL_OBSTACLE (-3)
#define ORC_QFIX_SCALE 16777216.0 /* 2^24 fixed-point scale of summed returns (gs mode) */
#define ORC_SEQ_ALLOC_MAX 16      /* arrivals <= this: exact virtual-pull allocation; more: water-filling */
#define ORC_BISECT_ITERS 14

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) -- counter-based RNG shared by spec.  */
/* ------------------------------------------------------------------------ */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds)
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc_philox4x32(ctr, key, out, 10); }
/* The ROLLOUT stream of the "gs" planner (the blocks that drive the <= max_depth random steps below a leaf) uses the
 * 7-round variant -- the fewest rounds Salmon et al. found Crush-resistant; Random123 ships known answers for it
 * (tests/test_cpu_oracle.py).  The rollouts are 85 % of a search's instructions and Philox a third of a rollout step;
 * every draw that decides something in the tree (root particle, descent, belief update) keeps the 10 rounds. */
#define ORC_ROLLOUT_ROUNDS 7

/* integer in [0, n): multiply-shift of one 32-bit word */
static inline uint32_t randbelow32(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }
/* 53-bit uniform numerator, same bit recipe as MT19937 genrand_res53 (a>>5, b>>6) */
static inline uint64_t u53_of(uint32_t a, uint32_t b) { return ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6); }

/* stream domains (ctr[3]) */
enum { DOM_SIM = 0, DOM_UPDATE = 1, DOM_FINAL = 2, DOM_AGENT = 3, DOM_SEQ = 7 };

/* ------------------------------------------------------------------------ */
/* RockSample model: examples/rock_sample/rock_model.py                      */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t rows, cols, n_rocks, n_actions, start_cell;
    int8_t cell[ORC_MAX_CELLS];                  /* rock k>=0 | EMPTY -1 | GOAL -2 | OBSTACLE -3 (rock_model.py:24-32) */
    uint32_t legal[ORC_MAX_CELLS];               /* per-cell legal-action bit mask (rock_model.py:218-247) */
    uint64_t thr53[ORC_MAX_CELLS * ORC_MAX_ROCKS]; /* floor(thr * 2^53); CHECK is correct iff u53 <= thr53 */
    double rew_illegal, rew_exit, rew_good, rew_bad; /* signed: -100, +10, +10, -10 (rock_sample_config.json) */
    int32_t rock_cell[ORC_MAX_ROCKS];            /* cell of rock k */
    double prob[ORC_MAX_CELLS * ORC_MAX_ROCKS];  /* sensor correctness probability [cell][rock] (rock_model.py:148-150) */
    int32_t have_prob;
} orc_model;

typedef struct {
    uint32_t actual_mask;  /* model.actual_rock_states (the REAL world; rock_model.py:263-265, A.5) */
    uint32_t sampled_mask; /* model.unique_rocks_sampled (rock_model.py:267-272) */
} orc_world;

static int valid_cell(const orc_model *m, int i, int j)
{ /* rock_model.py:213-216 is_valid_pos */
    return i >= 0 && i < m->rows && j >= 0 && j < m->cols && m->cell[i * m->cols + j] != ORC_CELL_OBSTACLE;
}

/* rock_model.py:218-247 get_legal_actions == rock_position_history.py:139-168 */
static uint32_t legal_mask_of_cell(const orc_model *m, int cell)
{
    static const int DI[4] = {-1, 0, 1, 0}, DJ[4] = {0, 1, 0, -1}; /* N E S W (rock_action.py:6-15) */
    int i = cell / m->cols, j = cell % m->cols;
    uint32_t mask = 0;
    for (int a = 0; a < 4; ++a)
        if (valid_cell(m, i + DI[a], j + DJ[a])) mask |= 1u << a;
    if (m->cell[cell] >= 0 && m->cell[cell] < m->n_rocks) mask |= 1u << 4;
    for (int k = 0; k < m->n_rocks; ++k) mask |= 1u << (5 + k);
    return mask;
}

orc_model *orc_model_create(int rows, int cols, const int8_t *cell, int start_cell,
                            const uint64_t *thr53 /* [rows*cols][n_rocks] */, const double rewards[4])
{
    if (rows * cols > ORC_MAX_CELLS) return NULL;
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->rows = rows; m->cols = cols; m->start_cell = start_cell;
    int nr = 0;
    for (int c = 0; c < rows * cols; ++c) { m->cell[c] = cell[c]; if (cell[c] + 1 > nr) nr = cell[c] + 1; }
    if (nr > ORC_MAX_ROCKS || 5 + nr > ORC_MAX_ACTIONS) { free(m); return NULL; }
    m->n_rocks = nr; m->n_actions = 5 + nr;                    /* rock_model.py:290-297 */
    for (int c = 0; c < rows * cols; ++c) {
        if (m->cell[c] >= 0) m->rock_cell[m->cell[c]] = c;
        m->legal[c] = legal_mask_of_cell(m, c);
        for (int k = 0; k < nr; ++k) m->thr53[c * ORC_MAX_ROCKS + k] = thr53[c * nr + k];
    }
    m->rew_illegal = rewards[0]; m->rew_exit = rewards[1]; m->rew_good = rewards[2]; m->rew_bad = rewards[3];
    return m;
}
void orc_model_destroy(orc_model *m) { free(m); }
void orc_model_set_prob(orc_model *m, const double *prob /* [cells][n_rocks] */)
{
    for (int c = 0; c < m->rows * m->cols; ++c)
        for (int k = 0; k < m->n_rocks; ++k) m->prob[c * ORC_MAX_ROCKS + k] = prob[c * m->n_rocks + k];
    m->have_prob = 1;
}
uint32_t orc_legal_mask(const orc_model *m, int cell) { return m->legal[cell]; }
int orc_model_n_actions(const orc_model *m) { return m->n_actions; }

typedef struct {
    int32_t next_cell; uint32_t next_rocks;
    int32_t obs_good, obs_empty, terminal, legal, used_uniform, rewarded_sample;
    double reward;
} orc_step_out;

/* rock_model.py:451-465 generate_step, as a pure function of (state, action, world, u53).
 * make_next_position :323-344, make_next_state :346-368, make_observation :370-412,
 * make_reward :417-445.  `rewarded_sample` reports the A.4 side effect (the reference clears
 * the *input* state's rock bit on a +10 SAMPLE, rock_model.py:433); callers decide whether to apply it. */
void orc_step(const orc_model *m, const orc_world *w, int cell, uint32_t rocks, int a, uint64_t u53, orc_step_out *o)
{
    static const int DI[4] = {-1, 0, 1, 0}, DJ[4] = {0, 1, 0, -1};
    int i = cell / m->cols, j = cell % m->cols, ni = i, nj = j, legal = 1;
    if (a < 4) {
        ni = i + DI[a]; nj = j + DJ[a];
        if (!valid_cell(m, ni, nj)) { ni = i; nj = j; legal = 0; }
    } else if (a == 4) {
        int code = m->cell[cell];
        legal = (code >= 0 && code < m->n_rocks);
    }
    int ncell = ni * m->cols + nj;
    uint32_t nr = rocks;
    if (legal && a == 4) nr &= ~(1u << m->cell[cell]);
    int good = 0, empty = (a < 4), used = 0;
    if (a >= 5) {
        int k = a - 5;
        if (!((w->sampled_mask >> k) & 1u)) {
            int truth = (w->actual_mask >> k) & 1u;
            int correct = (u53 <= m->thr53[ncell * ORC_MAX_ROCKS + k]);
            used = 1;
            good = correct ? truth : !truth;
            if (good) nr |= (1u << k); else nr &= ~(1u << k);   /* particle overwritten by the observation */
        }
    }
    int terminal = (m->cell[ncell] == ORC_CELL_GOAL);
    double r = 0.0; int rs = 0;
    if (!legal) r = m->rew_illegal;
    else if (terminal) r = m->rew_exit;
    else if (a == 4) {
        int k = m->cell[cell];
        if (((w->actual_mask >> k) & 1u) && ((rocks >> k) & 1u)) { r = m->rew_good; rs = 1; }
        else r = m->rew_bad;
    }
    o->next_cell = ncell; o->next_rocks = nr; o->obs_good = good; o->obs_empty = empty;
    o->terminal = terminal; o->legal = legal; o->used_uniform = used; o->rewarded_sample = rs; o->reward = r;
}

/* ------------------------------------------------------------------------ */
/* Preferred-action heuristic: per-node rock statistics (clean per-node copies; */
/* the reference shares one list across the tree)                  */
/* examples/rock_sample/rock_position_history.py:96-137 and :170-234            */
/* ------------------------------------------------------------------------ */
typedef struct { double chance[ORC_MAX_ROCKS]; int32_t good[ORC_MAX_ROCKS]; } orc_rockstats;

void orc_rockstats_init(const orc_model *m, orc_rockstats *s)
{ for (int k = 0; k < m->n_rocks; ++k) { s->chance[k] = 0.5; s->good[k] = 0; } }

/* create_child (:96-137): statistics of the node reached from a node at `cell` by (action, obs_good) */
void orc_rockstats_child(const orc_model *m, const orc_rockstats *parent, int cell, int a, int obs_good, orc_rockstats *child)
{
    *child = *parent;
    if (a == 4) {
        int k = m->cell[cell];
        if (k >= 0 && k < m->n_rocks) { child->chance[k] = 0.0; child->good[k] = -10; }
    } else if (a >= 5) {
        int k = a - 5;
        double p_ok = m->prob[cell * ORC_MAX_ROCKS + k], p_no = 1 - p_ok;
        double lg = child->chance[k], lb = 1 - lg;
        if (obs_good) { child->good[k] += 1; lg *= p_ok; lb *= p_no; }
        else { child->good[k] -= 1; lg *= p_no; lb *= p_ok; }
        if (!(fabs(lg) < 0.01 && fabs(lb) < 0.01)) child->chance[k] = lg / (lg + lb);   /* :131-135 (the reset is a no-op) */
    }
}

/* generate_smart_actions (:170-234) as a bit mask */
uint32_t orc_smart_mask(const orc_model *m, const orc_rockstats *s, int cell)
{
    int code = m->cell[cell];
    if (code >= 0 && code < m->n_rocks && (s->chance[code] == 1.0 || s->good[code] > 0)) return 1u << 4;
    int target = -1;
    for (int k = 0; k < m->n_rocks; ++k) if (s->chance[k] != 0.0 && s->good[k] >= 0) { target = k; break; }
    if (target < 0) return 1u << 1;                                /* "just head east" */
    int i = cell / m->cols, j = cell % m->cols, ri = m->rock_cell[target] / m->cols, rj = m->rock_cell[target] % m->cols;
    uint32_t mask = 0;
    if (ri > i) mask |= 1u << 2; else if (ri < i) mask |= 1u << 0;
    if (rj > j) mask |= 1u << 1; else if (rj < j) mask |= 1u << 3;
    for (int k = 0; k < m->n_rocks; ++k)
        if (s->chance[k] != 0.0 && s->chance[k] != 1.0 && abs(s->good[k]) < 2) mask |= 1u << (5 + k);
    return mask;
}

/* ======================================================================== */
/* (1) "ref": the reference's sequential POMCP                                */
/* ======================================================================== */
typedef struct { int32_t cell; uint32_t rocks; } ref_particle;

typedef struct {
    int32_t cell; uint32_t legal;
    int64_t N;                               /* DiscreteActionMapping.total_visit_count */
    int64_t n[ORC_MAX_ACTIONS];              /* entry.visit_count   */
    double total[ORC_MAX_ACTIONS];           /* entry.total_q_value */
    double mean[ORC_MAX_ACTIONS];            /* entry.mean_q_value  */
    int32_t child[ORC_MAX_ACTIONS][2];       /* belief child keyed (action, is_good); -1 = none */
    int32_t first_obs[ORC_MAX_ACTIONS];      /* insertion order of the action node's child_map */
    int32_t *bag; int32_t bag_n, bag_cap;    /* state_particles: indices into the particle pool (references!) */
} ref_node;

typedef struct {
    const orc_model *m; orc_world w;
    ref_node *nodes; int32_t n_nodes, cap_nodes;
    ref_particle *pool; int64_t n_pool, cap_pool;
    int32_t root;
    int32_t disable_tree;
    /* parameters (pomcp.py:13-43) */
    int32_t max_depth, max_particle_count; double ucb_c, discount;
    /* quirk switches (all 1 = reference behaviour) */
    int32_t stale_mean, mutate_input, libm_log;
    /* sequential draw stream */
    uint32_t key[2]; uint64_t draw_idx; uint32_t stream_id;
    /* counters */
    int64_t n_steps, n_levels, n_rollout_steps, n_zero_backups, n_backups;
} orc_ref;

static void ref_draw(orc_ref *h, uint32_t x[4])
{
    uint32_t ctr[4] = {(uint32_t)h->draw_idx, (uint32_t)(h->draw_idx >> 32), h->stream_id, DOM_SEQ};
    orc_philox4x32_10(ctr, h->key, x);
    h->draw_idx++;
}
static uint32_t ref_randbelow(orc_ref *h, uint32_t n) { uint32_t x[4]; ref_draw(h, x); return randbelow32(x[0], n); }
static uint64_t ref_u53(orc_ref *h) { uint32_t x[4]; ref_draw(h, x); return u53_of(x[1], x[2]); }

static int32_t ref_new_particle(orc_ref *h, int32_t cell, uint32_t rocks)
{
    if (h->n_pool == h->cap_pool) {
        h->cap_pool = h->cap_pool ? h->cap_pool * 2 : 4096;
        h->pool = (ref_particle *)realloc(h->pool, sizeof(ref_particle) * h->cap_pool);
    }
    h->pool[h->n_pool].cell = cell; h->pool[h->n_pool].rocks = rocks;
    return (int32_t)h->n_pool++;
}
static void ref_bag_push(ref_node *nd, int32_t p)
{
    if (nd->bag_n == nd->bag_cap) {
        nd->bag_cap = nd->bag_cap ? nd->bag_cap * 2 : 8;
        nd->bag = (int32_t *)realloc(nd->bag, sizeof(int32_t) * nd->bag_cap);
    }
    nd->bag[nd->bag_n++] = p;
}
/* discrete_action_mapping.py:16-34: zeroed entries, legal flags from the node's position */
static int32_t ref_new_node(orc_ref *h, int32_t cell)
{
    if (h->n_nodes == h->cap_nodes) {
        h->cap_nodes = h->cap_nodes ? h->cap_nodes * 2 : 1024;
        h->nodes = (ref_node *)realloc(h->nodes, sizeof(ref_node) * h->cap_nodes);
    }
    ref_node *nd = &h->nodes[h->n_nodes];
    memset(nd, 0, sizeof(*nd));
    nd->cell = cell; nd->legal = h->m->legal[cell];
    for (int a = 0; a < ORC_MAX_ACTIONS; ++a) { nd->child[a][0] = nd->child[a][1] = -1; nd->first_obs[a] = -1; }
    return h->n_nodes++;
}

orc_ref *orc_ref_create(const orc_model *m, uint32_t actual_mask, uint32_t sampled_mask,
                        const uint32_t *bag_rocks, int n_bag, /* belief_tree_solver.py:36-38 */
                        int max_depth, int max_particle_count, double ucb_c, double discount,
                        uint64_t seed, uint32_t stream_id)
{
    orc_ref *h = (orc_ref *)calloc(1, sizeof(orc_ref));
    h->m = m; h->w.actual_mask = actual_mask; h->w.sampled_mask = sampled_mask;
    h->max_depth = max_depth; h->max_particle_count = max_particle_count; h->ucb_c = ucb_c; h->discount = discount;
    h->stale_mean = 1; h->mutate_input = 1; h->libm_log = 1;
    h->key[0] = (uint32_t)seed; h->key[1] = (uint32_t)(seed >> 32); h->stream_id = stream_id;
    h->root = ref_new_node(h, m->start_cell);
    for (int i = 0; i < n_bag; ++i) ref_bag_push(&h->nodes[h->root], ref_new_particle(h, m->start_cell, bag_rocks[i]));
    return h;
}
void orc_ref_destroy(orc_ref *h)
{
    if (!h) return;
    for (int i = 0; i < h->n_nodes; ++i) free(h->nodes[i].bag);
    free(h->nodes); free(h->pool); free(h);
}
void orc_ref_set_quirks(orc_ref *h, int stale_mean, int mutate_input, int libm_log)
{ h->stale_mean = stale_mean; h->mutate_input = mutate_input; h->libm_log = libm_log; }
void orc_ref_set_world(orc_ref *h, uint32_t actual, uint32_t sampled) { h->w.actual_mask = actual; h->w.sampled_mask = sampled; }
uint64_t orc_ref_draws(const orc_ref *h) { return h->draw_idx; }
int orc_ref_disabled(const orc_ref *h) { return h->disable_tree; }

/* generate_step on a pooled particle (references: A.4 mutation applies to the pool entry) */
static void ref_step(orc_ref *h, int32_t p, int a, orc_step_out *o)
{
    ref_particle *s = &h->pool[p];
    uint64_t u = 0;
    if (a >= 5 && !((h->w.sampled_mask >> (a - 5)) & 1u)) u = ref_u53(h);   /* np.random.binomial, rock_model.py:395 */
    orc_step(h->m, &h->w, s->cell, s->rocks, a, u, o);
    if (o->rewarded_sample && h->mutate_input) s->rocks &= ~(1u << h->m->cell[s->cell]);  /* rock_model.py:433 */
    h->n_steps++;
}

/* action_selectors.py:6-37 ucb_action + solvers/pomcp.py:52-67 find_fast_ucb (closed form == table, A.11) */
static int ref_ucb_action(orc_ref *h, ref_node *nd, int greedy)
{
    int A = h->m->n_actions;
    int order[ORC_MAX_ACTIONS];
    for (int a = 0; a < A; ++a) order[a] = a;
    for (int i = A - 1; i >= 1; --i) {           /* random.shuffle: Fisher-Yates from the top */
        int j = (int)ref_randbelow(h, (uint32_t)(i + 1));
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    double log_n = log((double)(nd->N + 1));     /* np.log(N + 1) */
    double best_q = -INFINITY; int best[ORC_MAX_ACTIONS], nbest = 0;
    for (int t = 0; t < A; ++t) {
        int a = order[t];
        if (!((nd->legal >> a) & 1u)) continue;
        double q = nd->mean[a];
        if (!greedy) {
            if (nd->n[a] == 0) q += INFINITY;
            else q += h->ucb_c * sqrt(log_n / (double)nd->n[a]);
        }
        if (q >= best_q) {
            if (q > best_q) nbest = 0;
            best_q = q; best[nbest++] = a;
        }
    }
    return best[ref_randbelow(h, (uint32_t)nbest)];   /* random.choice(best_actions) */
}

/* belief_tree_solver.py:123-152 rollout: starts from a FRESH particle of `node` (A.2) */
static double ref_rollout(orc_ref *h, int32_t node)
{
    ref_node *nd = &h->nodes[node];
    uint32_t mask = nd->legal;                                   /* node.data.generate_legal_actions() */
    int32_t p = nd->bag[ref_randbelow(h, (uint32_t)nd->bag_n)];  /* node.sample_particle() */
    int terminal = 0, steps = 0; double sum = 0.0, disc = 1.0;
    int32_t cur = p; int cur_is_temp = 0;
    ref_particle tmp;
    while (steps < h->max_depth && !terminal) {
        int cnt = __builtin_popcount(mask);
        int idx = (int)ref_randbelow(h, (uint32_t)cnt);          /* random.choice(legal_actions), ascending ids */
        int a = -1;
        for (int b = 0, seen = 0; b < ORC_MAX_ACTIONS; ++b)
            if ((mask >> b) & 1u) { if (seen == idx) { a = b; break; } ++seen; }
        orc_step_out o;
        if (!cur_is_temp) ref_step(h, cur, a, &o);
        else {                                                   /* later states are private objects */
            uint64_t u = 0;
            if (a >= 5 && !((h->w.sampled_mask >> (a - 5)) & 1u)) u = ref_u53(h);
            orc_step(h->m, &h->w, tmp.cell, tmp.rocks, a, u, &o);
            h->n_steps++;
        }
        terminal = o.terminal;
        sum += o.reward * disc;
        disc *= h->discount;
        tmp.cell = o.next_cell; tmp.rocks = o.next_rocks; cur_is_temp = 1;
        mask = h->m->legal[tmp.cell];                            /* model.get_legal_actions(state) */
        ++steps; h->n_rollout_steps++;
    }
    return sum;
}

/* solvers/pomcp.py:87-137 traverse */
static double ref_traverse(orc_ref *h, int32_t node, int depth)
{
    double delayed = 0.0;
    ref_node *nd = &h->nodes[node];
    int32_t p = nd->bag[ref_randbelow(h, (uint32_t)nd->bag_n)];  /* :90 sample_particle (A.1) */
    int a = ref_ucb_action(h, nd, 0);                            /* :97 */
    if (depth >= h->max_depth) return 0.0;                       /* :100 */
    h->n_levels++;
    orc_step_out o; ref_step(h, p, a, &o);                       /* :104 */
    int og = o.obs_good ? 1 : 0;
    int32_t child = nd->child[a][og];                            /* :106 */
    if (child < 0 && !o.terminal && nd->N > 0) {                 /* :107-108 */
        int32_t c = ref_new_node(h, o.next_cell);                /* child position: rock_position_history.py:96-99 */
        nd = &h->nodes[node];                                    /* realloc may have moved the array */
        nd->child[a][og] = c; if (nd->first_obs[a] < 0) nd->first_obs[a] = og;
        child = c;
    }
    if (!o.terminal || !o.legal) {                               /* :110 */
        if (child >= 0) {
            ref_node *cn = &h->nodes[child];
            if (cn->bag_n < h->max_particle_count)               /* :115-116 */
                ref_bag_push(cn, ref_new_particle(h, o.next_cell, o.next_rocks));
            delayed = ref_traverse(h, child, depth + 1);
        } else delayed = ref_rollout(h, node);                   /* :119 */
        nd = &h->nodes[node];
    }
    double q = nd->mean[a];                                      /* :128 */
    q += (o.reward + (h->discount * delayed) - q);               /* :131 */
    nd->n[a] += 1; nd->N += 1;                                   /* discrete_action_mapping.py:137-144 */
    h->n_backups++;
    if (q == 0) h->n_zero_backups++;
    if (!(q == 0 && h->stale_mean)) {                            /* :146-172 (A.3: `if delta_total_q == 0: return`) */
        nd->total[a] += q;
        nd->mean[a] = nd->total[a] / (double)nd->n[a];
    }
    return q;
}

/* belief_tree_solver.py:90-121 rollout_search (degraded mode) */
static void ref_rollout_search(orc_ref *h)
{
    int32_t node = h->root;
    uint32_t mask = h->nodes[node].legal;
    for (int a = 0; a < h->m->n_actions; ++a) {
        if (!((mask >> a) & 1u)) continue;
        ref_node *nd = &h->nodes[node];
        int32_t p = nd->bag[ref_randbelow(h, (uint32_t)nd->bag_n)];
        orc_step_out o; ref_step(h, p, a, &o);
        double delayed = 0.0;
        if (!o.terminal) {
            int og = o.obs_good ? 1 : 0;
            int32_t child = nd->child[a][og];
            if (child < 0) {
                child = ref_new_node(h, o.next_cell);
                nd = &h->nodes[node];
                nd->child[a][og] = child; if (nd->first_obs[a] < 0) nd->first_obs[a] = og;
            }
            ref_bag_push(&h->nodes[child], ref_new_particle(h, o.next_cell, o.next_rocks));
            delayed = ref_rollout(h, child);
            nd = &h->nodes[node];
        }
        double q = nd->mean[a];
        q += (o.reward + h->discount * delayed - q);
        nd->n[a] += 1; nd->N += 1;
        if (!(q == 0 && h->stale_mean)) { nd->total[a] += q; nd->mean[a] = nd->total[a] / (double)nd->n[a]; }
    }
}

/* solvers/pomcp.py:69-78 select_eps_greedy_action; belief_tree_solver.py:42-56 monte_carlo_approx */
int orc_ref_search(orc_ref *h, int n_sims)
{
    if (h->disable_tree) ref_rollout_search(h);
    else for (int i = 0; i < n_sims; ++i) ref_traverse(h, h->root, 0);
    return ref_ucb_action(h, &h->nodes[h->root], 1);
}
void orc_ref_root_stats(const orc_ref *h, int64_t *n, double *total, double *mean, uint32_t *legal, int64_t *N)
{
    const ref_node *nd = &h->nodes[h->root];
    for (int a = 0; a < h->m->n_actions; ++a) { n[a] = nd->n[a]; total[a] = nd->total[a]; mean[a] = nd->mean[a]; }
    *legal = nd->legal; *N = nd->N;
}
void orc_ref_counters(const orc_ref *h, int64_t out[6])
{ out[0] = h->n_steps; out[1] = h->n_levels; out[2] = h->n_rollout_steps; out[3] = h->n_backups; out[4] = h->n_zero_backups; out[5] = h->n_nodes; }

/* agent.py:157 solver.belief_tree_index.sample_particle(): returns a pool reference */
int32_t orc_ref_sample_root_particle(orc_ref *h)
{
    ref_node *nd = &h->nodes[h->root];
    return nd->bag[ref_randbelow(h, (uint32_t)nd->bag_n)];
}
void orc_ref_particle(const orc_ref *h, int32_t p, int32_t *cell, uint32_t *rocks) { *cell = h->pool[p].cell; *rocks = h->pool[p].rocks; }

/* agent.py:174 the real environment step; returns the pool index of next_state */
int32_t orc_ref_env_step(orc_ref *h, int32_t p, int a, orc_step_out *o)
{
    ref_step(h, p, a, o);
    if (!o->legal) return p;   /* make_next_state returns state.copy() sharing the rock list (:354-356) */
    return ref_new_particle(h, o->next_cell, o->next_rocks);
}

/* belief_tree_solver.py:154-212 update + model.py:221-252 generate_particles.
 * `sampled_mask_after` is the world after model.update(step_result) (rock_model.py:267-272). */
int orc_ref_update(orc_ref *h, int action, int obs_good, uint32_t sampled_mask_after)
{
    h->w.sampled_mask = sampled_mask_after;                      /* :165 */
    ref_node *root = &h->nodes[h->root];
    int32_t child = root->child[action][obs_good ? 1 : 0];       /* :167 */
    if (child < 0) {                                             /* :171-185 */
        if (root->first_obs[action] < 0) { h->disable_tree = 1; return 1; }
        child = root->child[action][root->first_obs[action]];
        if (child < 0) child = root->child[action][1 - root->first_obs[action]];
    }
    /* generate_particles matches against obs_map.get_belief(obs): the observed child, or None when it was
     * never created -- in which case no candidate can match an existing node ... the reference then loops
     * forever (model.py:242-251, A.15); the oracle reports it instead. */
    int32_t target = root->child[action][obs_good ? 1 : 0];
    ref_node *cn = &h->nodes[child];
    if (cn->bag_n < h->max_particle_count) {                     /* :188-195 */
        if (target < 0) return -2;
        int need = h->max_particle_count - cn->bag_n;
        int32_t *fresh = (int32_t *)malloc(sizeof(int32_t) * (size_t)need); int got = 0;
        int64_t tries = 0;
        while (got < need) {
            root = &h->nodes[h->root];
            int32_t p = root->bag[ref_randbelow(h, (uint32_t)root->bag_n)];   /* random.choice(prev_particles) */
            orc_step_out o; ref_step(h, p, action, &o);
            if ((o.obs_good ? 1 : 0) == (obs_good ? 1 : 0)) fresh[got++] = ref_new_particle(h, o.next_cell, o.next_rocks);
            if (++tries > 50000000) { free(fresh); return -3; }
        }
        for (int i = 0; i < got; ++i) ref_bag_push(&h->nodes[child], fresh[i]);
        free(fresh);
    }
    if (h->nodes[child].bag_n == 0) { h->disable_tree = 1; return 1; }   /* :205-208 */
    h->root = child;                                             /* :210; prune_siblings only frees memory */
    return 0;
}

/* ======================================================================== */
/* (2) "gs": generation-synchronous tree-parallel POMCP (the CUDA algorithm) */
/* ======================================================================== */

/* Deterministic natural log of a positive integer, IEEE +,-,*,/ only, fixed operation order.
 * x = 2^e * m, m in [1/sqrt2*..., sqrt2); ln m = 2 atanh(f), f = (m-1)/(m+1); 13-term odd series. */
double orc_det_log_u64(uint64_t x)
{
    int e = 63 - __builtin_clzll(x);
    double m = (double)x / (double)(1ull << e);                  /* exact for x < 2^53 */
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double f = (m - 1.0) / (m + 1.0);
    double f2 = f * f;
    double s = 1.0 / 25.0;
    s = s * f2 + 1.0 / 23.0; s = s * f2 + 1.0 / 21.0; s = s * f2 + 1.0 / 19.0; s = s * f2 + 1.0 / 17.0;
    s = s * f2 + 1.0 / 15.0; s = s * f2 + 1.0 / 13.0; s = s * f2 + 1.0 / 11.0; s = s * f2 + 1.0 / 9.0;
    s = s * f2 + 1.0 / 7.0;  s = s * f2 + 1.0 / 5.0;  s = s * f2 + 1.0 / 3.0;  s = s * f2 + 1.0;
    double lnm = 2.0 * f * s;
    return (double)e * 0.6931471803691238 + ((double)e * 1.9082149292705877e-10 + lnm);
}

/* Single-precision twin used by the gs-mode UCB arithmetic (the CUDA kernels compute UCB in fp32). */
float orc_det_logf_u32(uint32_t x)
{
    int e = 31 - __builtin_clz(x);
    float m = (float)x / (float)(1u << e);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float f = (m - 1.0f) / (m + 1.0f);
    float f2 = f * f;
    float s = 1.0f / 11.0f;
    s = s * f2 + 1.0f / 9.0f; s = s * f2 + 1.0f / 7.0f; s = s * f2 + 1.0f / 5.0f; s = s * f2 + 1.0f / 3.0f;
    s = s * f2 + 1.0f;
    float lnm = 2.0f * f * s;
    return (float)e * 0.693145751953125f + ((float)e * 1.42860682030941723212e-6f + lnm);
}

/* ------------------------------------------------------------------------------------------------------------
 * Tabular POMDP (the reference's Model ABC with explicit matrices, e.g. examples/tiger/tiger_model.py:102-139):
 * T[a][s][s'], O[a][s'][z], R[a][s], as 32-bit cumulative thresholds floor(cumsum * 2^32) (last entry 0xFFFFFFFF).
 * A draw w picks the first index k with cdf[k] > min(w, 2^32 - 2): zero-probability entries are never picked.
 * An episode ends after an action of term_action_mask (Tiger: opening a door) or on entering a term_state.
 * ---------------------------------------------------------------------------------------------------------- */
#define ORC_Z_MAX 2048                   /* an observation index travels in 11 bits of a path entry on the device */
typedef struct {
    int32_t S, A, Z, ZS;                 /* ZS: power of two >= Z (the device's observation stride; unused here) */
    uint32_t term_action_mask;
    uint32_t *Tc, *Oc; double *R; uint8_t *term_state;
} orc_tab;

orc_tab *orc_tab_create(int S, int A, int Z, const uint32_t *Tc, const uint32_t *Oc, const double *R,
                        uint32_t term_action_mask, const uint8_t *term_state)
{
    if (S < 1 || A < 1 || A > ORC_MAX_ACTIONS || Z < 1 || Z > ORC_Z_MAX) return NULL;
    orc_tab *m = (orc_tab *)calloc(1, sizeof(orc_tab));
    m->S = S; m->A = A; m->Z = Z; m->ZS = 2; while (m->ZS < Z) m->ZS *= 2;
    m->term_action_mask = term_action_mask;
    size_t nt = (size_t)A * S * S, no = (size_t)A * S * Z;
    m->Tc = (uint32_t *)malloc(4 * nt); memcpy(m->Tc, Tc, 4 * nt);
    m->Oc = (uint32_t *)malloc(4 * no); memcpy(m->Oc, Oc, 4 * no);
    m->R = (double *)malloc(8 * (size_t)A * S); memcpy(m->R, R, 8 * (size_t)A * S);
    m->term_state = (uint8_t *)calloc((size_t)S, 1);
    if (term_state) memcpy(m->term_state, term_state, (size_t)S);
    return m;
}
void orc_tab_destroy(orc_tab *m) { if (!m) return; free(m->Tc); free(m->Oc); free(m->R); free(m->term_state); free(m); }

static int tab_sample(const uint32_t *cdf, int n, uint32_t w)
{
    if (w == 0xFFFFFFFFu) w = 0xFFFFFFFEu;
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] > w) hi = mid; else lo = mid + 1; }
    return lo;
}

/* generate_step of a tabular model: w1 draws the next state, w2 the observation (not drawn in rollouts) */
void orc_tab_step(const orc_tab *m, int s, int a, uint32_t w1, uint32_t w2, int with_obs,
                  int32_t *next, int32_t *obs, double *reward, int32_t *terminal)
{
    int ns = tab_sample(m->Tc + ((size_t)a * m->S + s) * m->S, m->S, w1);
    *next = ns;
    *obs = with_obs ? tab_sample(m->Oc + ((size_t)a * m->S + ns) * m->Z, m->Z, w2) : 0;
    *reward = m->R[(size_t)a * m->S + s];
    *terminal = (int)((m->term_action_mask >> a) & 1u) | (int)m->term_state[ns];
}

typedef struct {
    int32_t cell; uint32_t legal;
    int64_t n[ORC_MAX_ACTIONS];
    int64_t qfix[ORC_MAX_ACTIONS];             /* sum of llrint(q * 2^24) */
    orc_rockstats st;                          /* only meaningful with preferred actions */
} gs_node;

typedef struct {
    const orc_model *m; orc_world w;           /* RockSample ... */
    const orc_tab *tm;                         /* ... or a tabular model (m == NULL) */
    int32_t A, ZS;                             /* action count, observation count rounded up to a power of two */
    gs_node *nodes; int32_t n_nodes, cap_nodes; int32_t root;
    /* Observation map of every action node, as the reference keeps it: a mapping observation -> child belief node
     * that grows with the observations actually seen (discrete_observation_mapping.py:12-47, a dict per action
     * node), so |Z| is unbounded.  Here: a chain of entries per (node, action). */
    int32_t *omap_head;                        /* [cap_nodes][32] first entry of the chain, -1 = empty */
    struct gs_oentry { int32_t obs, child, next; } *omap; int32_t n_omap, cap_omap;
    uint32_t *bag; int32_t n_bag;              /* packed states: cell | rocks << 8 */
    int32_t max_depth, dcap; double ucb_c, discount;
    uint32_t key[2];
    int64_t n_levels, n_rollout_steps, n_created, n_alloc_nodes, n_generations;
    int32_t max_tree_depth;
    int32_t preferred;                         /* --preferred_actions: in-tree action sets from the rock statistics */
    int32_t preferred_rollout;                 /* rollouts follow generate_smart_actions of the rollout's own history */
} orc_gs;

/* Model dispatch of the gs planner.  A packed state is cell | rocks << 8 (RockSample) or the state index (tabular);
 * a node's tag is what its action set depends on: the robot cell (RockSample), nothing (tabular: 0). */
typedef struct { uint32_t next; int32_t tag; int32_t obs; double reward; int32_t terminal; } gs_step_out;

static uint32_t gs_tag_legal(const orc_gs *h, int32_t tag)
{ return h->m ? h->m->legal[tag] : (h->A >= 32 ? 0xFFFFFFFFu : ((1u << h->A) - 1u)); }
static int32_t gs_state_tag(const orc_gs *h, uint32_t st) { return h->m ? (int32_t)(st & 0xFFu) : 0; }

static void gs_step(const orc_gs *h, uint32_t st, int a, uint32_t w1, uint32_t w2, int rollout, gs_step_out *o)
{
    if (h->m) {
        orc_step_out r;
        orc_step(h->m, &h->w, (int)(st & 0xFFu), st >> 8, a, rollout ? (uint64_t)w1 << 21 : u53_of(w1, w2), &r);
        o->next = (uint32_t)r.next_cell | (r.next_rocks << 8); o->tag = r.next_cell; o->obs = r.obs_good ? 1 : 0;
        o->reward = r.reward; o->terminal = r.terminal;
    } else {
        int32_t ns;
        orc_tab_step(h->tm, (int)st, a, w1, w2, !rollout, &ns, &o->obs, &o->reward, &o->terminal);
        o->next = (uint32_t)ns; o->tag = 0;
    }
}

/* get_belief / get_entry (discrete_observation_mapping.py:18-23, 43-47: a scan over the entries): the child of
 * action node (X, a) under observation o, or -1 */
static int32_t gs_child(const orc_gs *h, int32_t X, int a, int o)
{
    for (int32_t e = h->omap_head[(size_t)X * ORC_MAX_ACTIONS + (size_t)a]; e >= 0; e = h->omap[e].next)
        if (h->omap[e].obs == o) return h->omap[e].child;
    return -1;
}
/* create_belief (discrete_observation_mapping.py:25-31): a new entry; the caller has checked that there is none */
static void gs_set_child(orc_gs *h, int32_t X, int a, int o, int32_t child)
{
    if (h->n_omap == h->cap_omap) {
        h->cap_omap = h->cap_omap ? h->cap_omap * 2 : 1024;
        h->omap = (struct gs_oentry *)realloc(h->omap, sizeof(struct gs_oentry) * (size_t)h->cap_omap);
    }
    int32_t *head = &h->omap_head[(size_t)X * ORC_MAX_ACTIONS + (size_t)a];
    h->omap[h->n_omap].obs = o; h->omap[h->n_omap].child = child; h->omap[h->n_omap].next = *head;
    *head = h->n_omap++;
}

static int32_t gs_new_node(orc_gs *h, int32_t cell)
{
    if (h->n_nodes == h->cap_nodes) {
        h->cap_nodes = h->cap_nodes ? h->cap_nodes * 2 : 1024;
        h->nodes = (gs_node *)realloc(h->nodes, sizeof(gs_node) * (size_t)h->cap_nodes);
        h->omap_head = (int32_t *)realloc(h->omap_head, sizeof(int32_t) * (size_t)h->cap_nodes * ORC_MAX_ACTIONS);
    }
    gs_node *nd = &h->nodes[h->n_nodes];
    memset(nd, 0, sizeof(*nd));
    nd->cell = cell; nd->legal = gs_tag_legal(h, cell);
    for (int a = 0; a < ORC_MAX_ACTIONS; ++a) h->omap_head[(size_t)h->n_nodes * ORC_MAX_ACTIONS + a] = -1;
    return h->n_nodes++;
}

orc_gs *orc_gs_create(const orc_model *m, uint32_t actual_mask, uint32_t sampled_mask,
                      const uint32_t *bag_packed, int n_bag, int root_cell,
                      int max_depth, int dcap, double ucb_c, double discount, uint64_t seed)
{
    orc_gs *h = (orc_gs *)calloc(1, sizeof(orc_gs));
    h->m = m; h->w.actual_mask = actual_mask; h->w.sampled_mask = sampled_mask;
    h->A = m->n_actions; h->ZS = 2;
    h->max_depth = max_depth; h->dcap = dcap; h->ucb_c = ucb_c; h->discount = discount;
    h->key[0] = (uint32_t)seed; h->key[1] = (uint32_t)(seed >> 32);
    h->bag = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_bag > 0 ? n_bag : 1));
    memcpy(h->bag, bag_packed, sizeof(uint32_t) * (size_t)n_bag); h->n_bag = n_bag;
    h->root = gs_new_node(h, root_cell);
    return h;
}
/* The same planner over a tabular model; bag_packed holds state indices. */
orc_gs *orc_gs_create_tab(const orc_tab *tm, const uint32_t *bag_packed, int n_bag,
                          int max_depth, int dcap, double ucb_c, double discount, uint64_t seed)
{
    orc_gs *h = (orc_gs *)calloc(1, sizeof(orc_gs));
    h->tm = tm; h->A = tm->A; h->ZS = tm->ZS;
    h->max_depth = max_depth; h->dcap = dcap; h->ucb_c = ucb_c; h->discount = discount;
    h->key[0] = (uint32_t)seed; h->key[1] = (uint32_t)(seed >> 32);
    h->bag = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_bag > 0 ? n_bag : 1));
    memcpy(h->bag, bag_packed, sizeof(uint32_t) * (size_t)n_bag); h->n_bag = n_bag;
    h->root = gs_new_node(h, 0);
    return h;
}
void orc_gs_destroy(orc_gs *h) { if (!h) return; free(h->nodes); free(h->omap_head); free(h->omap); free(h->bag); free(h); }
void orc_gs_set_world(orc_gs *h, uint32_t actual, uint32_t sampled) { h->w.actual_mask = actual; h->w.sampled_mask = sampled; }

/* --preferred_actions (rock_position_history.py:52-55): call right after orc_gs_create; the root gets fresh
 * statistics (rock_model.py:302-304) and its action set becomes generate_smart_actions(). */
void orc_gs_set_preferred(orc_gs *h, int enabled)
{
    h->preferred_rollout = enabled == 2;       /* 2: the rollout policy uses the heuristic too */
    h->preferred = enabled != 0;
    gs_node *r = &h->nodes[h->root];
    orc_rockstats_init(h->m, &r->st);
    r->legal = h->preferred ? orc_smart_mask(h->m, &r->st, r->cell) : h->m->legal[r->cell];
}

/* data.create_child + create_action_mapping for a new node (belief_node.py:96-124) */
static void gs_init_child(orc_gs *h, int32_t child, int32_t parent, int a, int og)
{
    if (!h->preferred) return;
    gs_node *c = &h->nodes[child]; const gs_node *p = &h->nodes[parent];
    orc_rockstats_child(h->m, &p->st, p->cell, a, og, &c->st);
    c->legal = orc_smart_mask(h->m, &c->st, c->cell);
}

static void gs_block(const orc_gs *h, uint32_t k, uint32_t sim, uint32_t move, uint32_t dom, uint32_t x[4])
{
    uint32_t ctr[4] = {k, sim, move, dom};
    orc_philox4x32_10(ctr, h->key, x);
}
static int g_gs_rollout_rounds = ORC_ROLLOUT_ROUNDS;   /* experiment switch (tests of the statistics only) */
void orc_gs_set_rollout_rounds(int r) { g_gs_rollout_rounds = r; }
static void gs_rollout_block(const orc_gs *h, uint32_t k, uint32_t sim, uint32_t move, uint32_t x[4])
{
    uint32_t ctr[4] = {k, sim, move, DOM_SIM};
    orc_philox4x32(ctr, h->key, x, g_gs_rollout_rounds);
}

/* Mean of an edge from its integer statistics (0 for an unvisited edge: entry.mean_q_value starts at 0). */
static float gs_mean(const gs_node *nd, int a)
{
    if (nd->n[a] == 0) return 0.0f;
    return ((float)nd->qfix[a] * (1.0f / 16777216.0f)) / (float)nd->n[a];
}

/* sum of 32 lane values in butterfly order (xor 16, 8, 4, 2, 1) -- the order a warp shuffle reduction uses */
static float butterfly_sum32(const float v[32])
{
    float x[32], y[32];
    memcpy(x, v, sizeof(x));
    for (int off = 16; off >= 1; off >>= 1) {
        for (int i = 0; i < 32; ++i) y[i] = x[i] + x[i ^ off];
        memcpy(x, y, sizeof(x));
    }
    return x[0];
}

/* ---------------------------------------------------------------------------------------------------------------
 * Action plans.  Within a generation the statistics are read-only, so the action distribution of a node is a pure
 * function of (its statistics, k) where k is the number of simulations of this generation EXPECTED at the node:
 *     k(root) = generation width;   k(child of X under (a, o)) = k(X) * P_plan(a | X) * N(child) / n(X, a)
 * (N(child) / n(X, a): the share of the edge's past visits that saw observation o).  A plan is one of
 *     thresholds  k >= 16.5: batch-UCB allocation -- a water-filling of k pulls over the UCB1 indices
 *     pulls       round(k) in 2..16: the virtual-pull sequence "pull j takes the UCB1 arg-max given the virtual
 *                 visits of pulls 0..j-1"; a simulation takes pull floor(x0 * round(k) / 2^32)
 *     mask        uniform draw over an action set: the untried actions (if at least round(k) of them), else for
 *                 round(k) == 1 the UCB1 arg-max set (action_selectors.py:6-37), or for c = 0 the best means
 * Nothing depends on how many simulations really meet at the node, so every simulation descends the tree on its own.
 * --------------------------------------------------------------------------------------------------------------- */
enum { GS_PLAN_MASK = 0, GS_PLAN_PULLS = 1, GS_PLAN_THRESHOLDS = 2 };
typedef struct { int32_t kind, kq, last; uint32_t mask; uint8_t pulls[16]; uint32_t thr[32]; } gs_plan;

/* Batch-UCB allocation: weights v[a] >= 0 from a water-filling of kd pulls over the UCB1 indices,
 * turned into 32-bit cumulative thresholds.  A simulation with 32-bit draw x takes the first action a with
 * x < thr[a], or `*last` if there is none. */
static void gs_allocation(const orc_gs *h, const gs_node *nd, float kd, uint32_t thr[32], int *last)
{
    int A = h->A;
    int64_t N = 0; int untried = 0, nlegal = 0;
    float v[32]; const float cf = (float)h->ucb_c;
    for (int a = 0; a < 32; ++a) v[a] = 0.0f;
    for (int a = 0; a < A; ++a) if ((nd->legal >> a) & 1u) { N += nd->n[a]; nlegal++; if (nd->n[a] == 0) untried++; }
    if (N == 0) { for (int a = 0; a < A; ++a) if ((nd->legal >> a) & 1u) v[a] = 1.0f; }
    else if ((float)untried >= kd) { for (int a = 0; a < A; ++a) if (((nd->legal >> a) & 1u) && nd->n[a] == 0) v[a] = 1.0f; }
    else {
        float L = orc_det_logf_u32((uint32_t)N + 1u);
        float c2L = (cf * cf) * L;
        float qbar[32]; float qmax = -INFINITY; int64_t nbest = 0;
        for (int a = 0; a < A; ++a) {
            if (!((nd->legal >> a) & 1u)) continue;
            qbar[a] = gs_mean(nd, a);
            if (qbar[a] > qmax) { qmax = qbar[a]; nbest = nd->n[a]; }
        }
        if (!(c2L > 0.0f)) { for (int a = 0; a < A; ++a) if (((nd->legal >> a) & 1u) && qbar[a] == qmax) v[a] = 1.0f; }
        else {
            float lo = qmax + 0.999f * sqrtf(c2L / (kd + (float)nbest));
            float hi = qmax + 1.001f * sqrtf((c2L * (float)nlegal) / kd) + 1.0f;
            for (int it = 0; it <= ORC_BISECT_ITERS; ++it) {          /* bisection steps, then the weights at `lo` */
                float mid = it < ORC_BISECT_ITERS ? 0.5f * (lo + hi) : lo;
                float t[32];
                for (int a = 0; a < 32; ++a) {
                    t[a] = 0.0f;
                    if (a < A && ((nd->legal >> a) & 1u)) {
                        float d = mid - qbar[a];
                        float mm = c2L / (d * d);
                        if (nd->n[a] == 0) t[a] = mm > 1.0f ? mm : 1.0f;
                        else { float e = mm - (float)nd->n[a]; t[a] = e > 0.0f ? e : 0.0f; }
                    }
                }
                if (it == ORC_BISECT_ITERS) { memcpy(v, t, sizeof(t)); break; }
                if (butterfly_sum32(t) >= kd) lo = mid; else hi = mid;
            }
        }
    }
    /* inclusive scan in Hillis-Steele order (what __shfl_up_sync steps 1,2,4,8,16 compute) */
    float x[32], y[32];
    memcpy(x, v, sizeof(x));
    for (int off = 1; off < 32; off <<= 1) {
        for (int i = 0; i < 32; ++i) y[i] = i >= off ? x[i] + x[i - off] : x[i];
        memcpy(x, y, sizeof(x));
    }
    float V = x[31];
    *last = 0;
    uint32_t run = 0;                      /* running maximum over the actions with positive weight: thresholds are
                                            * non-decreasing and an action without weight is never drawn */
    for (int a = 0; a < 32; ++a) {
        float f = (x[a] / V) * 4294967296.0f;
        uint32_t t = f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f;
        if (v[a] > 0.0f) { *last = a; if (t > run) run = t; }
        thr[a] = run;
    }
}

static int nth_set_bit(uint32_t mask, uint32_t j)
{
    for (int b = 0, seen = 0; b < 32; ++b) if ((mask >> b) & 1u) { if ((uint32_t)seen == j) return b; ++seen; }
    return -1;
}

/* Plan of a node with statistics (N > 0) for kq = round(k) <= ORC_SEQ_ALLOC_MAX expected arrivals.
 * Untried actions score +inf (solvers/pomcp.py:64-65): with at least kq of them the draw is uniform over them. */
static void gs_plan_few(const orc_gs *h, const gs_node *nd, int kq, gs_plan *pl)
{
    int A = h->A; const float cf = (float)h->ucb_c;
    int64_t N = 0; uint32_t untried = 0;
    pl->kq = kq; pl->kind = GS_PLAN_MASK;
    for (int a = 0; a < A; ++a) if ((nd->legal >> a) & 1u) { N += nd->n[a]; if (nd->n[a] == 0) untried |= 1u << a; }
    if (__builtin_popcount(untried) >= kq) { pl->mask = untried; return; }
    float L = orc_det_logf_u32((uint32_t)N + 1u);
    if (kq == 1) {                             /* UCB1 arg-max set; the draw breaks ties */
        float best = -INFINITY; uint32_t ties = 0;
        for (int a = 0; a < A; ++a) {
            if (!((nd->legal >> a) & 1u)) continue;
            float s = gs_mean(nd, a) + cf * sqrtf(L / (float)nd->n[a]);
            if (s > best) { best = s; ties = 1u << a; } else if (s == best) ties |= 1u << a;
        }
        pl->mask = ties; return;
    }
    float c2L = (cf * cf) * L;
    float sc[32]; int64_t va[32]; float qmax = -INFINITY; uint32_t qties = 0;
    for (int a = 0; a < 32; ++a) { va[a] = 0; sc[a] = -INFINITY; }
    for (int a = 0; a < A; ++a) {
        if (!((nd->legal >> a) & 1u)) continue;
        float qb = gs_mean(nd, a);
        if (qb > qmax) { qmax = qb; qties = 1u << a; } else if (qb == qmax) qties |= 1u << a;
        sc[a] = nd->n[a] == 0 ? INFINITY : qb + cf * sqrtf(L / (float)nd->n[a]);
    }
    if (!(c2L > 0.0f)) { pl->mask = qties; return; }
    pl->kind = GS_PLAN_PULLS;
    for (int pull = 0; pull < kq; ++pull) {
        int w = 0;
        for (int a = 1; a < A; ++a) if (sc[a] > sc[w]) w = a;
        pl->pulls[pull] = (uint8_t)w;
        va[w] += 1;
        sc[w] = gs_mean(nd, w) + cf * sqrtf(L / (float)(nd->n[w] + va[w]));
    }
}

static void gs_make_plan(const orc_gs *h, const gs_node *nd, float k, gs_plan *pl)
{
    if (k >= 16.5f) {
        pl->kind = GS_PLAN_THRESHOLDS; pl->kq = 0;
        gs_allocation(h, nd, k, pl->thr, &pl->last);
    } else {
        int kq = (int)(k + 0.5f);
        gs_plan_few(h, nd, kq < 1 ? 1 : kq, pl);
    }
}

/* The action a simulation with 32-bit draw x0 takes under a plan, and the plan's probability of that action. */
static int gs_plan_action(const gs_plan *pl, uint32_t x0, float *prob)
{
    if (pl->kind == GS_PLAN_THRESHOLDS) {
        int a = pl->last;
        for (int b = 0; b < 32; ++b) if (x0 < pl->thr[b]) { a = b; break; }
        *prob = (float)(pl->thr[a] - (a > 0 ? pl->thr[a - 1] : 0u)) * (1.0f / 4294967296.0f);
        return a;
    }
    if (pl->kind == GS_PLAN_MASK) {
        uint32_t m = (uint32_t)__builtin_popcount(pl->mask);
        *prob = 1.0f / (float)m;
        return nth_set_bit(pl->mask, randbelow32(x0, m));
    }
    int a = pl->pulls[randbelow32(x0, (uint32_t)pl->kq)], cnt = 0;
    for (int j = 0; j < pl->kq; ++j) if (pl->pulls[j] == a) ++cnt;
    *prob = (float)cnt / (float)pl->kq;
    return a;
}

typedef struct {
    uint32_t sim;          /* global simulation index (stream id) */
    uint32_t st;           /* packed state */
    int32_t depth;         /* number of path entries */
    int32_t status;        /* 1 leaf: rollout pending, 2 done (terminal or horizon) */
    int32_t *pnode; uint8_t *pact; double *prew;
    int32_t last_obs;      /* observation of the last tree step (start of the rollout's history) */
    double delayed;
} gs_sim;

/* belief_tree_solver.py:123-152, but from the post-action state (canonical) */
static double gs_rollout(orc_gs *h, gs_sim *s, uint32_t move)
{
    double sum = 0.0, disc = 1.0; int t = 0, terminal = 0;
    uint32_t st = s->st;
    uint32_t x[4] = {0, 0, 0, 0};
    /* preferred-action rollout: the rollout keeps the rock statistics of its own history -- the leaf node's, updated
     * with every (action, observation) it generates (rock_position_history.py:96-137) -- and draws uniformly from
     * generate_smart_actions (:170-234) instead of the legal set */
    orc_rockstats rst; int pref = h->m && h->preferred_rollout;
    if (pref) {
        const gs_node *ln = &h->nodes[s->pnode[s->depth - 1]];
        orc_rockstats_child(h->m, &ln->st, ln->cell, s->pact[s->depth - 1], s->last_obs, &rst);
    }
    while (t < h->max_depth && !terminal) {
        /* one Philox block per two rollout steps: words (0,1) then (2,3); the first word picks the action, the
         * second drives the model (RockSample: top 32 bits of the 53-bit sensor uniform; tabular: next state) */
        if ((t & 1) == 0) gs_rollout_block(h, (uint32_t)(s->depth + 1 + (t >> 1)), s->sim, move, x);
        uint32_t w0 = x[(t & 1) * 2], w1 = x[(t & 1) * 2 + 1];
        uint32_t mask = pref ? orc_smart_mask(h->m, &rst, (int)(st & 0xFFu)) : gs_tag_legal(h, gs_state_tag(h, st));
        int idx = (int)randbelow32(w0, (uint32_t)__builtin_popcount(mask)), a = -1;
        for (int b = 0, seen = 0; b < 32; ++b) if ((mask >> b) & 1u) { if (seen == idx) { a = b; break; } ++seen; }
        gs_step_out o; gs_step(h, st, a, w1, 0u, 1, &o);
        if (pref) { orc_rockstats nx; orc_rockstats_child(h->m, &rst, (int)(st & 0xFFu), a, o.obs, &nx); rst = nx; }
        terminal = o.terminal;
        if (o.reward != 0.0) sum += o.reward * disc;
        disc *= h->discount;
        st = o.next;
        ++t; h->n_rollout_steps++;
    }
    return sum;
}

static int g_gs_kmode = 0;   /* experiment switch: 0 = expected arrivals carried down the plans, 1 = w * N(X) / N(root) */
void orc_gs_set_kmode(int m) { g_gs_kmode = m; }
static int64_t gs_visits(const gs_node *nd) { int64_t N = 0; for (int b = 0; b < 32; ++b) N += nd->n[b]; return N; }

/* A simulation standing at a node without statistics (N = 0): every legal action is untried, so UCB1 is a uniform
 * draw (action_selectors.py:6-37 with +inf bonuses), the node cannot be expanded (solvers/pomcp.py:107) and the
 * simulation ends in a rollout. */
static void gs_virgin_step(orc_gs *h, gs_sim *s, int32_t Y, uint32_t move)
{
    int d = s->depth;
    if (d >= h->max_depth) { s->status = 2; s->delayed = 0.0; return; }              /* solvers/pomcp.py:100 */
    uint32_t x[4]; gs_block(h, (uint32_t)(d + 1), s->sim, move, DOM_SIM, x);
    uint32_t mask = h->nodes[Y].legal;
    int a = nth_set_bit(mask, randbelow32(x[0], (uint32_t)__builtin_popcount(mask)));
    gs_step_out o; gs_step(h, s->st, a, x[1], x[2], 0, &o);
    h->n_levels++;
    s->pnode[d] = Y; s->pact[d] = (uint8_t)a; s->prew[d] = o.reward; s->depth = d + 1; s->last_obs = o.obs;
    if (s->depth > h->max_tree_depth) h->max_tree_depth = s->depth;
    s->st = o.next;
    s->status = o.terminal ? 2 : 1; s->delayed = 0.0;
}

/* One search: n_sims simulations with global ids [sim_offset, sim_offset + n_sims), in generations of width
 * w0, w0*growth, ... capped at wmax.  Statistics are read-only within a generation (the backups of a generation are
 * applied after its last simulation), so the order of the simulations inside a generation does not matter. */
int orc_gs_search(orc_gs *h, int64_t n_sims, uint32_t sim_offset, uint32_t move, int w0, int growth, int wmax)
{
    if (w0 < 1) w0 = 1; if (growth < 1) growth = 1; if (wmax < w0) wmax = w0;
    int dcap = h->dcap;
    int64_t done = 0;
    gs_sim *S = (gs_sim *)calloc((size_t)wmax, sizeof(gs_sim));
    int32_t *pn = (int32_t *)malloc(sizeof(int32_t) * (size_t)wmax * (size_t)dcap);
    uint8_t *pa = (uint8_t *)malloc((size_t)wmax * (size_t)dcap);
    double *pr = (double *)malloc(sizeof(double) * (size_t)wmax * (size_t)dcap);
    gs_plan *cache = NULL; int64_t *cache_stamp = NULL; int32_t cache_cap = 0;   /* plans of this generation, by node */
    gs_plan fresh;
    while (done < n_sims) {
        /* width of the generation: growth x the visits the root has seen, within [w0, wmax] */
        int64_t W = (int64_t)growth * gs_visits(&h->nodes[h->root]);
        if (W < w0) W = w0; if (W > wmax) W = wmax;
        int64_t w = W; if (w > n_sims - done) w = n_sims - done;
        h->n_generations++;
        const int64_t stamp = h->n_generations;
        const int32_t nn0 = h->n_nodes;            /* nodes created in this generation have no statistics: never planned */
        if (cache_cap < nn0) {
            cache_cap = nn0 * 2;
            cache = (gs_plan *)realloc(cache, sizeof(gs_plan) * (size_t)cache_cap);
            cache_stamp = (int64_t *)realloc(cache_stamp, sizeof(int64_t) * (size_t)cache_cap);
            memset(cache_stamp, 0, sizeof(int64_t) * (size_t)cache_cap);
        }
        const int root_virgin = gs_visits(&h->nodes[h->root]) == 0;
        for (int64_t i = 0; i < w; ++i) {
            gs_sim *s = &S[i];
            s->sim = sim_offset + (uint32_t)(done + i);
            s->pnode = pn + i * dcap; s->pact = pa + i * dcap; s->prew = pr + i * dcap;
            uint32_t x[4]; gs_block(h, 0, s->sim, move, DOM_SIM, x);           /* belief_node.py:47-48 */
            s->st = h->bag[randbelow32(x[0], (uint32_t)h->n_bag)];
            s->depth = 0; s->status = 0; s->delayed = 0.0;
            if (root_virgin) { gs_virgin_step(h, s, h->root, move); continue; }
            int32_t X = h->root; float k = (float)w;
            for (;;) {                                                         /* solvers/pomcp.py:87-137 */
                int d = s->depth;
                if (d >= h->max_depth) { s->status = 2; break; }               /* :100 */
                const gs_plan *pl;
                if (k < 1.5f) { gs_make_plan(h, &h->nodes[X], k, &fresh); pl = &fresh; }
                else {
                    if (cache_stamp[X] != stamp) {
                        gs_make_plan(h, &h->nodes[X], k, &cache[X]); cache_stamp[X] = stamp;
                        if (cache[X].kind == GS_PLAN_THRESHOLDS) h->n_alloc_nodes++;
                    }
                    pl = &cache[X];
                }
                uint32_t x4[4]; gs_block(h, (uint32_t)(d + 1), s->sim, move, DOM_SIM, x4);
                float pa_; const int a = gs_plan_action(pl, x4[0], &pa_);
                gs_step_out o; gs_step(h, s->st, a, x4[1], x4[2], 0, &o);      /* :104 */
                h->n_levels++;
                s->pnode[d] = X; s->pact[d] = (uint8_t)a; s->prew[d] = o.reward; s->depth = d + 1; s->last_obs = o.obs;
                if (s->depth > h->max_tree_depth) h->max_tree_depth = s->depth;
                s->st = o.next;
                if (o.terminal) { s->status = 2; break; }
                int32_t child = gs_child(h, X, a, o.obs);                      /* :106 */
                if (child < 0) {
                    if (d + 1 >= dcap) { s->status = 1; break; }              /* depth cap of the tree: leaf (:119) */
                    child = gs_new_node(h, o.tag);                             /* :107-108 (N(X) > 0 here) */
                    gs_init_child(h, child, X, a, o.obs);
                    gs_set_child(h, X, a, o.obs, child); h->n_created++;
                }
                const int64_t Nc = child < nn0 ? gs_visits(&h->nodes[child]) : 0;
                if (Nc == 0) { gs_virgin_step(h, s, child, move); break; }
                if (g_gs_kmode == 0) k = (k * pa_) * ((float)Nc / (float)h->nodes[X].n[a]);   /* expected arrivals at the child */
                else k = (float)w * ((float)Nc / (float)gs_visits(&h->nodes[h->root]));
                X = child;
            }
        }
        /* rollouts, then backup (commutative integer updates) */
        for (int64_t i = 0; i < w; ++i) {
            gs_sim *s = &S[i];
            if (s->status == 1) s->delayed = gs_rollout(h, s, move);
        }
        for (int64_t i = 0; i < w; ++i) {
            gs_sim *s = &S[i];
            double delayed = s->delayed;
            for (int d = s->depth - 1; d >= 0; --d) {
                double q = s->prew[d] + h->discount * delayed;                 /* solvers/pomcp.py:131 */
                gs_node *nd = &h->nodes[s->pnode[d]];
                nd->n[s->pact[d]] += 1;                                        /* discrete_action_mapping.py:137-144 */
                nd->qfix[s->pact[d]] += (int64_t)llrint(q * ORC_QFIX_SCALE);   /* :146-172, as an exact integer sum */
                delayed = q;
            }
        }
        done += w;
    }
    free(S); free(pn); free(pa); free(pr); free(cache); free(cache_stamp);
    return 0;
}

void orc_gs_root_stats(const orc_gs *h, int64_t *n, int64_t *qfix, uint32_t *legal)
{
    const gs_node *nd = &h->nodes[h->root];
    for (int a = 0; a < h->A; ++a) { n[a] = nd->n[a]; qfix[a] = nd->qfix[a]; }
    *legal = nd->legal;
}
void orc_gs_counters(const orc_gs *h, int64_t out[8])
{
    out[0] = h->n_levels; out[1] = h->n_rollout_steps; out[2] = h->n_created; out[3] = h->n_nodes;
    out[4] = h->n_alloc_nodes; out[5] = h->n_generations; out[6] = h->max_tree_depth; out[7] = h->n_bag;
}

/* Statistics of the node reached from the root by a sequence of (action, obs) pairs; returns 0 if it exists. */
int orc_gs_node_stats(const orc_gs *h, const int32_t *path_ao, int len, int64_t *n, int64_t *qfix, uint32_t *legal, int32_t *cell)
{
    int32_t X = h->root;
    for (int i = 0; i < len; ++i) { X = gs_child(h, X, path_ao[2 * i], path_ao[2 * i + 1]); if (X < 0) return 1; }
    const gs_node *nd = &h->nodes[X];
    for (int a = 0; a < h->A; ++a) { n[a] = nd->n[a]; qfix[a] = nd->qfix[a]; }
    *legal = nd->legal; *cell = nd->cell;
    return 0;
}

/* Greedy root action (solvers/pomcp.py:78): arg-max of the mean over legal actions (0 for unvisited), uniform
 * tie-break from the DOM_FINAL stream.  n/qfix may be the all-reduced sums of several ranks. */
int orc_gs_greedy(const int64_t *n, const int64_t *qfix, uint32_t legal, int n_actions, uint64_t seed, uint32_t move)
{
    double best = -INFINITY; int cand[32], nc = 0;
    for (int a = 0; a < n_actions; ++a) {
        if (!((legal >> a) & 1u)) continue;
        double q = n[a] == 0 ? 0.0 : ((double)qfix[a] * (1.0 / ORC_QFIX_SCALE)) / (double)n[a];
        if (q > best) { best = q; nc = 0; cand[nc++] = a; } else if (q == best) cand[nc++] = a;
    }
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, ctr[4] = {0, 0, move, DOM_FINAL}, x[4];
    orc_philox4x32_10(ctr, key, x);
    return cand[randbelow32(x[0], (uint32_t)nc)];
}

/* Belief update (belief_tree_solver.py:154-212 + model.py:221-252), deterministic form: try t draws particle
 * bag[randbelow] and one uniform from block (t, 0, move, DOM_UPDATE); the new bag is the first `need`
 * accepted next-states in try order.  Returns the number accepted (< need only if max_tries was hit);
 * status: 0 re-rooted at an existing child, 1 fresh root created (child never expanded). */
int64_t orc_gs_update(orc_gs *h, int action, int obs_good, uint32_t sampled_mask_after, uint32_t move,
                      int need, int64_t max_tries, int32_t *status, int64_t *tries_used)
{
    const int obs = h->m ? (obs_good ? 1 : 0) : obs_good;      /* tabular: the observation index */
    h->w.sampled_mask = sampled_mask_after;
    uint32_t *nb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)need);
    int64_t got = 0, t = 0;
    for (; t < max_tries && got < need; ++t) {
        uint32_t x[4]; gs_block(h, (uint32_t)t, (uint32_t)(t >> 32), move, DOM_UPDATE, x);
        uint32_t st = h->bag[randbelow32(x[0], (uint32_t)h->n_bag)];
        gs_step_out o; gs_step(h, st, action, x[1], x[2], 0, &o);
        if (o.obs == obs) nb[got++] = o.next;
    }
    *tries_used = t;
    if (got == 0) { free(nb); *status = 2; return 0; }
    /* child position: make_next_position on the node's cell (rock_position_history.py:96-99) */
    int32_t child = gs_child(h, h->root, action, obs);
    *status = 0;
    if (child < 0) {
        int32_t tag = 0;
        if (h->m) { orc_step_out o; orc_step(h->m, &h->w, h->nodes[h->root].cell, 0, action, 0, &o); tag = o.next_cell; }
        child = gs_new_node(h, tag); *status = 1;
        gs_init_child(h, child, h->root, action, obs);
    }
    h->root = child;
    free(h->bag); h->bag = nb; h->n_bag = (int32_t)got;
    return got;
}
void orc_gs_bag(const orc_gs *h, uint32_t *out) { memcpy(out, h->bag, sizeof(uint32_t) * (size_t)h->n_bag); }
int orc_gs_bag_size(const orc_gs *h) { return h->n_bag; }
int orc_gs_root_cell(const orc_gs *h) { return h->nodes[h->root].cell; }
