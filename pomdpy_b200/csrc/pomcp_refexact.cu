// pomcp_refexact.cu -- the reference's SEQUENTIAL POMCP, quirks included, as a single-thread CUDA kernel.
//
// Purpose: parity, not speed.  With the same random draws (one sequential Philox stream, the one
// oracle/refharness/ injects into the executed reference) this path reproduces the reference bit for bit ON THE
// DEVICE: root (visit_count, total_q_value, mean_q_value), the chosen action, the real environment step and the
// belief update -- pomdpy/solvers/pomcp.py:69-137, pomdpy/solvers/belief_tree_solver.py:42-212,
// pomdpy/action_selection/action_selectors.py:6-37, pomdpy/discrete_pomdp/discrete_action_mapping.py:137-172,
// pomdpy/pomdp/model.py:221-252, examples/rock_sample/rock_model.py:451-465.  It shares the device generate_step
// (rs_step) and the model tables with the fast planner, so it also ties that code to the executed reference.
//
// Kept from the reference: a fresh particle from the node's bag at every level (A.1), rollout
// from a fresh particle of the current node (A.2), stale mean when the sampled return is exactly 0 (A.3), the
// rewarded SAMPLE clearing the rock bit of the *input* particle, which lives in the bag (A.4), observations from the
// true world (A.5), expansion only if the parent has visits (A.12), random.shuffle + random.choice tie-breaking.
// ln(N+1) comes from a host table (numpy's values), sqrt and / are IEEE on both sides, no FMA contraction.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pomcp_b200.h"
#include "pomcp_device.cuh"

#define RX_AS 32
#define RX_DOM_SEQ 7

struct RxNode {
    int32_t cell; uint32_t legal;
    long long N;
    long long n[RX_AS];
    double total[RX_AS], mean[RX_AS];
    int32_t child[RX_AS][2];
    int32_t first_obs[RX_AS];
    int32_t bag_off, bag_n, bag_cap;          // slice of the bag pool (indices into the particle pool)
};

struct RxState {
    int32_t root, n_nodes, disable_tree, pad;
    long long n_particles, bag_used;
    unsigned long long draw_idx;
    uint32_t actual, sampled;
    // outputs of the last call
    int32_t out_action, out_obs_good, out_terminal, out_legal, out_next_cell, out_particle, out_status, out_pad;
    uint32_t out_next_rocks, out_pad2;
    double out_reward;
};

struct RxParams {
    RxState *S;
    const ModelHeader *model;
    const uint64_t *thr53;
    RxNode *nodes; long long cap_nodes;
    int32_t *p_cell; uint32_t *p_rocks; long long cap_particles;      // particle pool (particles are references)
    int32_t *bags; long long cap_bag;                                 // bag pool
    const double *log_table; int32_t n_log;                           // ln(N + 1), N < n_log, numpy's values
    int32_t max_depth, max_particle_count;
    double ucb_c, discount;
    uint32_t key0, key1, stream_id;
};

struct Rx {
    const RxParams &p;
    SmemModel M;
    __device__ Philox4 draw() const
    {
        RxState *S = p.S;
        const Philox4 r = philox4x32_10((uint32_t)S->draw_idx, (uint32_t)(S->draw_idx >> 32), p.stream_id, RX_DOM_SEQ,
                                        p.key0, p.key1);
        S->draw_idx++;
        return r;
    }
    __device__ uint32_t randbelow(uint32_t n) const { return randbelow32(draw().x, n); }
    __device__ uint64_t u53() const { const Philox4 r = draw(); return u53_of(r.y, r.z); }

    __device__ int new_particle(int cell, uint32_t rocks) const
    {
        const long long i = p.S->n_particles++;
        if (i >= p.cap_particles) { p.S->out_status = -10; return 0; }
        p.p_cell[i] = cell; p.p_rocks[i] = rocks;
        return (int)i;
    }
    __device__ void bag_push(RxNode &nd, int particle) const
    {
        if (nd.bag_n == nd.bag_cap) {                       // grow by doubling inside the pool (single thread)
            const int ncap = nd.bag_cap ? nd.bag_cap * 2 : 8;
            const long long off = p.S->bag_used;
            if (off + ncap > p.cap_bag) { p.S->out_status = -11; return; }
            p.S->bag_used += ncap;
            for (int k = 0; k < nd.bag_n; ++k) p.bags[off + k] = p.bags[nd.bag_off + k];
            nd.bag_off = (int32_t)off; nd.bag_cap = ncap;
        }
        p.bags[nd.bag_off + nd.bag_n++] = particle;
    }
    __device__ int new_node(int cell) const                 // discrete_action_mapping.py:16-34
    {
        const int id = p.S->n_nodes++;
        if (id >= p.cap_nodes) { p.S->out_status = -12; return 0; }
        RxNode &nd = p.nodes[id];
        nd.cell = cell; nd.legal = M.hdr->legal[cell]; nd.N = 0;
        for (int a = 0; a < RX_AS; ++a) {
            nd.n[a] = 0; nd.total[a] = 0.0; nd.mean[a] = 0.0; nd.child[a][0] = nd.child[a][1] = -1; nd.first_obs[a] = -1;
        }
        nd.bag_off = 0; nd.bag_n = 0; nd.bag_cap = 0;
        return id;
    }
    // generate_step on a pooled particle: the A.4 mutation applies to the pool entry (rock_model.py:433)
    __device__ StepResult step(int particle, int a) const
    {
        uint64_t u = 0;
        if (a >= 5 && !((p.S->sampled >> (a - 5)) & 1u)) u = u53();          // np.random.binomial, rock_model.py:395
        const uint32_t rocks = p.p_rocks[particle];
        const StepResult sr = rs_step(M, (uint32_t)p.p_cell[particle], rocks, a, u, p.S->actual, p.S->sampled);
        if (sr.rew == REW_GOOD) p.p_rocks[particle] = rocks & ~(1u << M.hdr->cell[p.p_cell[particle]]);
        return sr;
    }
    __device__ double ln_np1(long long N) const { return N < p.n_log ? p.log_table[N] : log((double)(N + 1)); }

    // action_selectors.py:6-37 (+ solvers/pomcp.py:52-67: the table equals the closed form)
    __device__ int ucb_action(const RxNode &nd, bool greedy) const
    {
        const int A = M.hdr->n_actions;
        int order[RX_AS];
        for (int a = 0; a < A; ++a) order[a] = a;
        for (int i = A - 1; i >= 1; --i) {                                   // random.shuffle
            const int j = (int)randbelow((uint32_t)(i + 1));
            const int t = order[i]; order[i] = order[j]; order[j] = t;
        }
        const double log_n = ln_np1(nd.N);
        double best_q = -INFINITY; int best[RX_AS], nbest = 0;
        for (int t = 0; t < A; ++t) {
            const int a = order[t];
            if (!((nd.legal >> a) & 1u)) continue;
            double q = nd.mean[a];
            if (!greedy) q += nd.n[a] == 0 ? INFINITY : p.ucb_c * sqrt(log_n / (double)nd.n[a]);
            if (q >= best_q) {
                if (q > best_q) nbest = 0;
                best_q = q; best[nbest++] = a;
            }
        }
        return best[randbelow((uint32_t)nbest)];                             // random.choice(best_actions)
    }

    // belief_tree_solver.py:123-152: from a FRESH particle of `node` (A.2)
    __device__ double rollout(int node) const
    {
        const RxNode &nd = p.nodes[node];
        uint32_t mask = nd.legal;
        const int particle = p.bags[nd.bag_off + randbelow((uint32_t)nd.bag_n)];
        int terminal = 0, steps = 0; double sum = 0.0, disc = 1.0;
        uint32_t cell = 0, rocks = 0; bool pooled = true;
        while (steps < p.max_depth && !terminal) {
            const int a = nth_bit(mask, randbelow((uint32_t)__popc(mask)));   // random.choice(legal_actions)
            StepResult sr;
            if (pooled) sr = step(particle, a);
            else {
                uint64_t u = 0;
                if (a >= 5 && !((p.S->sampled >> (a - 5)) & 1u)) u = u53();
                sr = rs_step(M, cell, rocks, a, u, p.S->actual, p.S->sampled);
            }
            terminal = sr.terminal;
            sum += M.hdr->rew[sr.rew] * disc;
            disc *= p.discount;
            cell = sr.cell; rocks = sr.rocks; pooled = false;
            mask = M.hdr->legal[cell];
            ++steps;
        }
        return sum;
    }

    // solvers/pomcp.py:87-137, the recursion unrolled: descend, then back up from the bottom
    __device__ void simulate() const
    {
        int pnode[128], pact[128]; double prew[128];
        int depth = 0, node = p.S->root;
        double delayed = 0.0;
        for (;;) {
            RxNode &nd = p.nodes[node];
            const int particle = p.bags[nd.bag_off + randbelow((uint32_t)nd.bag_n)];     // :90 (A.1)
            const int a = ucb_action(nd, false);                                           // :97
            if (depth >= p.max_depth) { delayed = 0.0; break; }                            // :100
            const StepResult sr = step(particle, a);                                       // :104
            const int og = sr.good ? 1 : 0;
            int child = nd.child[a][og];                                                   // :106
            if (child < 0 && !sr.terminal && nd.N > 0) {                                   // :107-108
                child = new_node((int)sr.cell);
                nd.child[a][og] = child; if (nd.first_obs[a] < 0) nd.first_obs[a] = og;
            }
            pnode[depth] = node; pact[depth] = a; prew[depth] = M.hdr->rew[sr.rew];
            ++depth;
            if (!sr.terminal || !sr.legal) {                                               // :110
                if (child >= 0) {
                    RxNode &cn = p.nodes[child];
                    if (cn.bag_n < p.max_particle_count) bag_push(cn, new_particle((int)sr.cell, sr.rocks));   // :115-116
                    node = child;
                    continue;                                                              // traverse(child, depth + 1)
                }
                delayed = rollout(node);                                                   // :119
            } else delayed = 0.0;
            break;
        }
        for (int d = depth - 1; d >= 0; --d) {                                             // :126-137
            RxNode &nd = p.nodes[pnode[d]];
            const int a = pact[d];
            double q = nd.mean[a];
            q += (prew[d] + (p.discount * delayed) - q);
            nd.n[a] += 1; nd.N += 1;                                                       // discrete_action_mapping.py:137-144
            if (!(q == 0)) { nd.total[a] += q; nd.mean[a] = nd.total[a] / (double)nd.n[a]; } // :146-172 (A.3)
            delayed = q;
        }
    }
};

enum { RX_OP_SEARCH = 0, RX_OP_SAMPLE_ROOT = 1, RX_OP_ENV_STEP = 2, RX_OP_UPDATE = 3 };

extern "C" __global__ void pomcp_refexact_kernel(const RxParams p, int op, int arg0, int arg1, uint32_t arg2)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Rx rx{p, SmemModel{p.model, p.thr53}};
    RxState *S = p.S;
    S->out_status = 0;
    if (op == RX_OP_SEARCH) {                                   // solvers/pomcp.py:69-78
        if (!S->disable_tree) for (int i = 0; i < arg0; ++i) rx.simulate();
        S->out_action = rx.ucb_action(p.nodes[S->root], true);
    } else if (op == RX_OP_SAMPLE_ROOT) {                       // agent.py:157
        const RxNode &nd = p.nodes[S->root];
        S->out_particle = p.bags[nd.bag_off + rx.randbelow((uint32_t)nd.bag_n)];
    } else if (op == RX_OP_ENV_STEP) {                          // agent.py:174 (the real step; arg0 = particle, arg1 = action)
        const StepResult sr = rx.step(arg0, arg1);
        S->out_obs_good = sr.good; S->out_terminal = sr.terminal; S->out_legal = sr.legal;
        S->out_reward = p.model->rew[sr.rew]; S->out_next_cell = (int32_t)sr.cell; S->out_next_rocks = sr.rocks;
        S->out_particle = sr.legal ? rx.new_particle((int)sr.cell, sr.rocks) : arg0;
    } else if (op == RX_OP_UPDATE) {                            // belief_tree_solver.py:154-212 (arg0 action, arg1 obs, arg2 sampled)
        S->sampled = arg2;                                      // :165 model.update first
        RxNode &root = p.nodes[S->root];
        const int og = arg1 ? 1 : 0;
        int child = root.child[arg0][og];                       // :167
        if (child < 0) {                                        // :171-185
            if (root.first_obs[arg0] < 0) { S->disable_tree = 1; S->out_status = 1; return; }
            child = root.child[arg0][root.first_obs[arg0]];
        }
        const int target = root.child[arg0][og];
        if (p.nodes[child].bag_n < p.max_particle_count) {      // :188-195, model.py:221-252
            if (target < 0) { S->out_status = -2; return; }     // the reference would loop forever (A.15)
            int need = p.max_particle_count - p.nodes[child].bag_n;
            long long tries = 0;
            while (need > 0) {
                const int particle = p.bags[root.bag_off + rx.randbelow((uint32_t)root.bag_n)];
                const StepResult sr = rx.step(particle, arg0);
                if ((sr.good ? 1 : 0) == og) { rx.bag_push(p.nodes[child], rx.new_particle((int)sr.cell, sr.rocks)); --need; }
                if (++tries > 50000000ll || S->out_status < 0) { S->out_status = -3; return; }
            }
        }
        if (p.nodes[child].bag_n == 0) { S->disable_tree = 1; S->out_status = 1; return; }     // :205-208
        S->root = child;                                        // :210
    }
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
struct pomcp_refexact {
    RxParams p{};
    ModelHeader hmodel;
    std::string err;
    int device = 0;
    int n_actions = 0;
};

#define RCK(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } \
    } while (0)

extern "C" const char *pomcp_refexact_last_error(const pomcp_refexact *h) { return h ? h->err.c_str() : "null handle"; }

extern "C" void pomcp_refexact_destroy(pomcp_refexact *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->p.S); cudaFree((void *)h->p.model); cudaFree((void *)h->p.thr53); cudaFree(h->p.nodes);
    cudaFree(h->p.p_cell); cudaFree(h->p.p_rocks); cudaFree(h->p.bags); cudaFree((void *)h->p.log_table);
    delete h;
}

extern "C" int pomcp_refexact_create(int device, int rows, int cols, const int8_t *cell, int start_cell,
                                     const uint64_t *thr53, const double rewards[4], uint32_t actual_mask,
                                     const uint32_t *bag_rocks, int n_bag, int max_depth, int max_particle_count,
                                     double ucb_c, double discount, uint64_t seed, uint32_t stream_id,
                                     const double *log_table, int n_log, int64_t cap_nodes, int64_t cap_particles,
                                     pomcp_refexact **out)
{
    pomcp_refexact *h = new pomcp_refexact();
    *out = h;
    h->device = device;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { h->err = "no CUDA device: no CPU fallback"; return -1; }
    if (max_depth < 1 || max_depth > 127) { h->err = "max_depth out of range [1, 127]"; return -1; }
    RCK(cudaSetDevice(device));
    ModelHeader &m = h->hmodel;
    memset(&m, 0, sizeof(m));
    m.rows = rows; m.cols = cols; m.n_cells = rows * cols; m.start_cell = start_cell;
    int nr = 0;
    for (int c = 0; c < m.n_cells; ++c) { m.cell[c] = cell[c]; if (cell[c] + 1 > nr) nr = cell[c] + 1; }
    m.n_rocks = nr; m.n_actions = 5 + nr; h->n_actions = 5 + nr;
    m.rew[REW_ILLEGAL] = rewards[0]; m.rew[REW_EXIT] = rewards[1]; m.rew[REW_GOOD] = rewards[2]; m.rew[REW_BAD] = rewards[3];
    m.delta[0] = -cols; m.delta[1] = 1; m.delta[2] = cols; m.delta[3] = -1;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            const int c = i * cols + j;
            if (m.cell[c] == POMCP_CELL_OBSTACLE) continue;
            uint32_t mask = ((nr + 5 >= 32) ? 0xFFFFFFFFu : ((1u << (nr + 5)) - 1u)) & ~31u;
            if (i > 0 && m.cell[c - cols] != POMCP_CELL_OBSTACLE) mask |= 1u;
            if (j + 1 < cols && m.cell[c + 1] != POMCP_CELL_OBSTACLE) mask |= 2u;
            if (i + 1 < rows && m.cell[c + cols] != POMCP_CELL_OBSTACLE) mask |= 4u;
            if (j > 0 && m.cell[c - 1] != POMCP_CELL_OBSTACLE) mask |= 8u;
            if (m.cell[c] >= 0) mask |= 16u;
            m.legal[c] = mask;
        }
    RxParams &p = h->p;
    ModelHeader *dm; uint64_t *dt; double *dl;
    RCK(cudaMalloc(&dm, sizeof(m))); RCK(cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice));
    RCK(cudaMalloc(&dt, sizeof(uint64_t) * m.n_cells * nr));
    RCK(cudaMemcpy(dt, thr53, sizeof(uint64_t) * m.n_cells * nr, cudaMemcpyHostToDevice));
    RCK(cudaMalloc(&dl, sizeof(double) * n_log)); RCK(cudaMemcpy(dl, log_table, sizeof(double) * n_log, cudaMemcpyHostToDevice));
    p.model = dm; p.thr53 = dt; p.log_table = dl; p.n_log = n_log;
    p.cap_nodes = cap_nodes; p.cap_particles = cap_particles; p.cap_bag = cap_particles * 2 + cap_nodes * 16;
    RCK(cudaMalloc(&p.nodes, sizeof(RxNode) * cap_nodes));
    RCK(cudaMalloc(&p.p_cell, sizeof(int32_t) * cap_particles)); RCK(cudaMalloc(&p.p_rocks, sizeof(uint32_t) * cap_particles));
    RCK(cudaMalloc(&p.bags, sizeof(int32_t) * p.cap_bag));
    RCK(cudaMalloc(&p.S, sizeof(RxState)));
    p.max_depth = max_depth; p.max_particle_count = max_particle_count; p.ucb_c = ucb_c; p.discount = discount;
    p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32); p.stream_id = stream_id;
    // root node and its bag, built on the host (belief_tree_solver.py:30-40)
    RxState st; memset(&st, 0, sizeof(st));
    st.root = 0; st.n_nodes = 1; st.n_particles = n_bag; st.bag_used = n_bag; st.actual = actual_mask; st.sampled = 0;
    RxNode root; memset(&root, 0, sizeof(root));
    root.cell = start_cell; root.legal = m.legal[start_cell];
    for (int a = 0; a < RX_AS; ++a) { root.child[a][0] = root.child[a][1] = -1; root.first_obs[a] = -1; }
    root.bag_off = 0; root.bag_n = n_bag; root.bag_cap = n_bag;
    if (n_bag > cap_particles) { h->err = "particle capacity too small"; return -1; }
    int32_t *cells = new int32_t[n_bag], *idx = new int32_t[n_bag];
    for (int i = 0; i < n_bag; ++i) { cells[i] = start_cell; idx[i] = i; }
    RCK(cudaMemcpy(p.p_cell, cells, sizeof(int32_t) * n_bag, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(p.p_rocks, bag_rocks, sizeof(uint32_t) * n_bag, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(p.bags, idx, sizeof(int32_t) * n_bag, cudaMemcpyHostToDevice));
    delete[] cells; delete[] idx;
    RCK(cudaMemcpy(p.nodes, &root, sizeof(root), cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(p.S, &st, sizeof(st), cudaMemcpyHostToDevice));
    return 0;
}

static int rx_run(pomcp_refexact *h, int op, int a0, int a1, uint32_t a2, RxState *st)
{
    RCK(cudaSetDevice(h->device));
    pomcp_refexact_kernel<<<1, 32>>>(h->p, op, a0, a1, a2);
    RCK(cudaGetLastError());
    RCK(cudaDeviceSynchronize());
    RCK(cudaMemcpy(st, h->p.S, sizeof(*st), cudaMemcpyDeviceToHost));
    if (st->out_status < 0) { h->err = "refexact kernel status " + std::to_string(st->out_status); return st->out_status; }
    return st->out_status;
}

// select_eps_greedy_action: n_sims sequential simulations, then the greedy action; root statistics out
extern "C" int pomcp_refexact_search(pomcp_refexact *h, int n_sims, int32_t *action, int64_t *n, double *total,
                                     double *mean, uint32_t *legal, int64_t *N, uint64_t *draws)
{
    RxState st;
    const int rc = rx_run(h, RX_OP_SEARCH, n_sims, 0, 0, &st);
    if (rc < 0) return rc;
    RxNode root;
    RCK(cudaMemcpy(&root, h->p.nodes + st.root, sizeof(root), cudaMemcpyDeviceToHost));
    for (int a = 0; a < h->n_actions; ++a) { n[a] = root.n[a]; total[a] = root.total[a]; mean[a] = root.mean[a]; }
    *legal = root.legal; *N = root.N; *action = st.out_action; *draws = st.draw_idx;
    return rc;
}

extern "C" int pomcp_refexact_sample_root_particle(pomcp_refexact *h, int32_t *particle)
{
    RxState st;
    const int rc = rx_run(h, RX_OP_SAMPLE_ROOT, 0, 0, 0, &st);
    *particle = st.out_particle;
    return rc;
}

// the real environment step on the true state (agent.py:174); out[0..5] = obs_good, terminal, legal, next_cell,
// next_rocks, next particle
extern "C" int pomcp_refexact_env_step(pomcp_refexact *h, int particle, int action, int64_t out[6], double *reward,
                                       uint64_t *draws)
{
    RxState st;
    const int rc = rx_run(h, RX_OP_ENV_STEP, particle, action, 0, &st);
    out[0] = st.out_obs_good; out[1] = st.out_terminal; out[2] = st.out_legal; out[3] = st.out_next_cell;
    out[4] = st.out_next_rocks; out[5] = st.out_particle;
    *reward = st.out_reward; *draws = st.draw_idx;
    return rc;
}

extern "C" int pomcp_refexact_update(pomcp_refexact *h, int action, int obs_good, uint32_t sampled_after, int32_t *disable_tree,
                                     uint64_t *draws)
{
    RxState st;
    const int rc = rx_run(h, RX_OP_UPDATE, action, obs_good, sampled_after, &st);
    *disable_tree = st.disable_tree; *draws = st.draw_idx;
    return rc;
}
