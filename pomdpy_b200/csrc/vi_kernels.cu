// vi_kernels.cu -- value-iteration alpha-vector backup and pointwise-dominance prune (sm_100a), fp64.
//
// Secondary path: pomdpy/solvers/value_iteration.py:24-142 of the reference.
// Built with -fmad=false: products and sums are evaluated in the reference's order, so the alpha vectors equal the
// numpy oracle's bit for bit.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/vi_b200.h"
#include "pomcp_device.cuh"     // Philox4x32-10

#define VCK(call)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_));          \
            return -1;                                                               \
        }                                                                            \
    } while (0)

// candidate c = (u * Z + z) * G + g ; element i.   value_iteration.py:51-67
__global__ void vi_backup_kernel(int S, int A, int Z, int G, const double *__restrict__ T, const double *__restrict__ O,
                                 const double *__restrict__ R, const double *__restrict__ gamma, double discount,
                                 double *__restrict__ cand, int32_t *__restrict__ cand_action)
{
    const long long total = (long long)A * Z * G * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % S);
        const long long c = idx / S;
        const int g = (int)(c % G);
        const int z = (int)((c / G) % Z);
        const int u = (int)(c / ((long long)G * Z));
        const double vo = gamma[(size_t)g * S + i] * O[((size_t)u * S + i) * Z + z];    // v.v[i] * o[u][i][z]
        double acc = 0.0;
        for (int j = 0; j < S; ++j) acc = acc + vo * T[((size_t)u * S + j) * S + i];     // ... * t[u][j][i], summed over j
        cand[idx] = discount * (R[(size_t)u * S + i] + acc);
        if (i == 0) cand_action[c] = u;
    }
}

// value_iteration.py:88-142 as the pointwise test its LP computes.  One thread per vector i.
__global__ void vi_prune_kernel(int S, int n, const double *__restrict__ v, double delta, int32_t *__restrict__ keep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int beaten = 0;
        for (int j = 0; j < n && !beaten; ++j) {
            if (j == i) continue;
            bool dom = true;                               // max_s (v_i - v_j) < delta
            for (int s = 0; s < S; ++s)
                if (!(v[(size_t)i * S + s] - v[(size_t)j * S + s] < delta)) { dom = false; break; }
            if (!dom) continue;
            bool mutual = true;                            // max_s (v_j - v_i) < delta
            for (int s = 0; s < S; ++s)
                if (!(v[(size_t)j * S + s] - v[(size_t)i * S + s] < delta)) { mutual = false; break; }
            if (!mutual || j < i) beaten = 1;
        }
        keep[i] = !beaten;
    }
}

// Order-preserving compaction of the kept vectors (one block; n is a few thousand at most on this path).
__global__ void vi_compact_kernel(int S, int n, const double *__restrict__ v, const int32_t *__restrict__ act,
                                  const int32_t *__restrict__ keep, double *__restrict__ out, int32_t *__restrict__ out_act,
                                  int32_t *__restrict__ n_out)
{
    __shared__ int base, warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + tid;
        const int k = (i < n) ? keep[i] : 0;
        int incl = k;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int t = warp_tot[lane]; int it = t;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, it, o); if (lane >= o) it += y; }
            warp_tot[lane] = it - t;
        }
        __syncthreads();
        const int pos = base + warp_tot[wid] + incl - k;
        if (k) {
            for (int s = 0; s < S; ++s) out[(size_t)pos * S + s] = v[(size_t)i * S + s];
            out_act[pos] = act[i];
        }
        __syncthreads();
        if (tid == blockDim.x - 1) base = pos + k;
        __syncthreads();
    }
    if (tid == 0) *n_out = base;
}

extern "C" int vi_solve_reference(int device, int S, int A, int Z, const double *T, const double *O, const double *R,
                                  double discount, int horizon, int capacity, double *alpha_out, int32_t *action_out,
                                  int32_t *n_out, char *err, int errlen)
{
    if (S < 1 || A < 1 || Z < 1 || horizon < 0 || capacity < 1) { snprintf(err, errlen, "bad sizes"); return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(err, errlen, "no CUDA device: no CPU fallback"); return -1; }
    VCK(cudaSetDevice(device));
    double *dT, *dO, *dR, *dG, *dAll, *dNext; int32_t *dAct, *dAllAct, *dNextAct, *dKeep, *dN;
    const size_t cap = (size_t)capacity, work = cap * (1 + (size_t)A * Z);       // gamma + candidates
    VCK(cudaMalloc(&dT, sizeof(double) * A * S * S)); VCK(cudaMalloc(&dO, sizeof(double) * A * S * Z));
    VCK(cudaMalloc(&dR, sizeof(double) * A * S));
    VCK(cudaMalloc(&dG, sizeof(double) * cap * S)); VCK(cudaMalloc(&dAct, sizeof(int32_t) * cap));
    VCK(cudaMalloc(&dAll, sizeof(double) * work * S)); VCK(cudaMalloc(&dAllAct, sizeof(int32_t) * work));
    VCK(cudaMalloc(&dNext, sizeof(double) * work * S)); VCK(cudaMalloc(&dNextAct, sizeof(int32_t) * work));
    VCK(cudaMalloc(&dKeep, sizeof(int32_t) * work)); VCK(cudaMalloc(&dN, sizeof(int32_t)));
    VCK(cudaMemcpy(dT, T, sizeof(double) * A * S * S, cudaMemcpyHostToDevice));
    VCK(cudaMemcpy(dO, O, sizeof(double) * A * S * Z, cudaMemcpyHostToDevice));
    VCK(cudaMemcpy(dR, R, sizeof(double) * A * S, cudaMemcpyHostToDevice));
    VCK(cudaMemset(dG, 0, sizeof(double) * S));                                  // the dummy zero vector (:38-40)
    int minus1 = -1;
    VCK(cudaMemcpy(dAct, &minus1, sizeof(int32_t), cudaMemcpyHostToDevice));
    int G = 1, rc = 0;
    for (int k = 0; k < horizon; ++k) {
        const int ncand = A * Z * G;
        const int keep_old = (k == 0) ? 0 : G;                                   // the dummy is dropped after pass 0 (:69-72)
        const int n = keep_old + ncand;
        if ((size_t)n > work) { snprintf(err, errlen, "capacity %d too small at horizon %d", capacity, k + 1); rc = -2; break; }
        if (keep_old) {
            VCK(cudaMemcpy(dAll, dG, sizeof(double) * (size_t)G * S, cudaMemcpyDeviceToDevice));
            VCK(cudaMemcpy(dAllAct, dAct, sizeof(int32_t) * G, cudaMemcpyDeviceToDevice));
        }
        const long long total = (long long)ncand * S;
        int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
        vi_backup_kernel<<<blocks, 256>>>(S, A, Z, G, dT, dO, dR, dG, discount, dAll + (size_t)keep_old * S, dAllAct + keep_old);
        int pb = (n + 127) / 128; if (pb > 148 * 8) pb = 148 * 8;
        vi_prune_kernel<<<pb, 128>>>(S, n, dAll, 1e-10, dKeep);
        vi_compact_kernel<<<1, 1024>>>(S, n, dAll, dAllAct, dKeep, dNext, dNextAct, dN);
        VCK(cudaGetLastError());
        int newG = 0;
        VCK(cudaMemcpy(&newG, dN, sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (newG > capacity) { snprintf(err, errlen, "capacity %d too small: %d vectors at horizon %d", capacity, newG, k + 1); rc = -2; break; }
        VCK(cudaMemcpy(dG, dNext, sizeof(double) * (size_t)newG * S, cudaMemcpyDeviceToDevice));
        VCK(cudaMemcpy(dAct, dNextAct, sizeof(int32_t) * newG, cudaMemcpyDeviceToDevice));
        G = newG;
    }
    if (rc == 0) {
        VCK(cudaMemcpy(alpha_out, dG, sizeof(double) * (size_t)G * S, cudaMemcpyDeviceToHost));
        VCK(cudaMemcpy(action_out, dAct, sizeof(int32_t) * G, cudaMemcpyDeviceToHost));
        *n_out = G;
    }
    cudaFree(dT); cudaFree(dO); cudaFree(dR); cudaFree(dG); cudaFree(dAct); cudaFree(dAll); cudaFree(dAllAct);
    cudaFree(dNext); cudaFree(dNextAct); cudaFree(dKeep); cudaFree(dN);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Textbook projection for large |S|: B[s'][o*G+g] = O[a][s'][o] * gamma[g][s'], then C = T_a . B with the fp64
// tensor cores (mma.sync m8n8k4 f64 -- tcgen05 has no fp64 kind; DMMA is the fp64 tensor path on Blackwell).
// ---------------------------------------------------------------------------------------------------------
__global__ void vi_scale_kernel(int S, int Z, int G, const double *__restrict__ Oa, const double *__restrict__ gamma,
                                double *__restrict__ B)
{
    const long long total = (long long)S * Z * G;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(idx % ((long long)Z * G));
        const int sp = (int)(idx / ((long long)Z * G));
        const int o = col / G, g = col - o * G;
        B[idx] = Oa[(size_t)sp * Z + o] * gamma[(size_t)g * S + sp];
    }
}

// C[M x N] = A[M x K] . B[K x N], row-major fp64, on the fp64 tensor cores.
// Block tile 128 x 64, k-step 16, 3-stage cp.async pipeline; 8 warps as 4 (M) x 2 (N), each owning a 32 x 32 sub-tile
// = 4 x 4 DMMA tiles (mma.sync.m8n8k4.f64: 32 accumulator doubles per thread).  Shared-memory rows are padded so that
// the 64-bit fragment loads of each half-warp hit 16 distinct bank pairs.  M % 128 == 0, N % 64 == 0, K % 16 == 0.
#define GBM 128
#define GBN 64
#define GBK 16
#define GSTAGES 3
#define SA_LD 20               // doubles per A row in shared memory (16 + 4 pad)
#define SB_LD 68               // doubles per B row in shared memory (64 + 4 pad)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}

__global__ void __launch_bounds__(256) vi_dgemm_dmma_kernel(int M, int N, int K, const double *__restrict__ A,
                                                            const double *__restrict__ B, double *__restrict__ C)
{
    extern __shared__ __align__(16) double gsm[];
    double *sA = gsm;                                   // [GSTAGES][GBM][SA_LD]
    double *sB = gsm + GSTAGES * GBM * SA_LD;           // [GSTAGES][GBK][SB_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bm = blockIdx.y * GBM, bn = blockIdx.x * GBN;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int gid = lane >> 2, tig = lane & 3;                  // mma.m8n8k4 fragment coordinates
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int k0) {
        double *a = sA + stage * GBM * SA_LD, *b = sB + stage * GBK * SB_LD;
#pragma unroll
        for (int e = 0; e < 4; ++e) {                           // A: 128 rows x 16 k = 1024 chunks of 2 doubles
            const int ch = tid + e * 256, r = ch >> 3, c2 = (ch & 7) * 2;
            cp_async16(a + r * SA_LD + c2, A + (size_t)(bm + r) * K + k0 + c2);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {                           // B: 16 k x 64 cols = 512 chunks
            const int ch = tid + e * 256, r = ch >> 5, c2 = (ch & 31) * 2;
            cp_async16(b + r * SB_LD + c2, B + (size_t)(k0 + r) * N + bn + c2);
        }
        asm volatile("cp.async.commit_group;");
    };

    const int nk = K / GBK;
    for (int s = 0; s < GSTAGES - 1; ++s) { if (s < nk) load_stage(s, s * GBK); else asm volatile("cp.async.commit_group;"); }
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(GSTAGES - 2));
        __syncthreads();
        const int nxt = kt + GSTAGES - 1;
        if (nxt < nk) load_stage(nxt % GSTAGES, nxt * GBK); else asm volatile("cp.async.commit_group;");
        const double *a = sA + (kt % GSTAGES) * GBM * SA_LD, *b = sB + (kt % GSTAGES) * GBK * SB_LD;
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a[(wm + i * 8 + gid) * SA_LD + kk + tig];   // A fragment: row gid, col tig
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b[(kk + tig) * SB_LD + wn + j * 8 + gid];   // B fragment: row tig, col gid
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(af[i]), "d"(bf[j]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {                           // C fragment: row gid, cols 2*tig, 2*tig+1
            const int r = bm + wm + i * 8 + gid, c = bn + wn + j * 8 + 2 * tig;
            *reinterpret_cast<double2 *>(&C[(size_t)r * N + c]) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
}
#define DGEMM_SMEM ((GSTAGES * GBM * SA_LD + GSTAGES * GBK * SB_LD) * sizeof(double))

extern "C" int vi_project_textbook_dev(int device, int S, int A, int Z, int G, const double *T_dev, const double *O_dev,
                                       const double *gamma_dev, double *scratch_dev, double *proj_dev, void *stream,
                                       char *err, int errlen)
{
    const int N = Z * G;
    VCK(cudaSetDevice(device));
    if (S % GBM || N % GBN || S % GBK) { snprintf(err, errlen, "S must be a multiple of 128 and Z*G of 64"); return -1; }
    VCK(cudaFuncSetAttribute(vi_dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DGEMM_SMEM));
    cudaStream_t st = (cudaStream_t)stream;
    for (int a = 0; a < A; ++a) {
        vi_scale_kernel<<<148 * 8, 256, 0, st>>>(S, Z, G, O_dev + (size_t)a * S * Z, gamma_dev, scratch_dev);
        dim3 grid(N / GBN, S / GBM);
        vi_dgemm_dmma_kernel<<<grid, 256, DGEMM_SMEM, st>>>(S, N, S, T_dev + (size_t)a * S * S, scratch_dev,
                                                   proj_dev + (size_t)a * S * N);
    }
    VCK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Point-based backup for large |S| (the cross-sum of the exact backup, taken at belief points):
//   proj[a][s][o,g]  = sum_s' T[a][s][s'] O[a][s'][o] gamma_g[s']               (DMMA GEMMs, above)
//   score[a][p][o,g] = b_p . proj[a][:, o,g]                                    (DMMA GEMMs: B [P x S] . proj[a])
//   g*(a,p,o) = argmax_g score (lowest g on ties);   v[a][p] = b_p . R[a] + discount * sum_o score[a][p][o, g*]
//   a*(p) = argmax_a v (lowest a on ties);   alpha_p = R[a*] + discount * sum_o proj[a*][:, o, g*(a*,p,o)]
// one new vector per belief point, then the pointwise-dominance prune.  All fp64, sums over o in index order.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vi_pb_choose_kernel(int S, int Z, int G, int P, const double *__restrict__ score,
                                                           const double *__restrict__ R, const double *__restrict__ B,
                                                           double discount, int32_t *__restrict__ best,
                                                           double *__restrict__ value)
{
    extern __shared__ double pb_sm[];                          // [Z] best scores, then [8] warp partials
    const int a = blockIdx.y, p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *sc = score + ((size_t)a * P + p) * (size_t)Z * G;
    for (int o = wid; o < Z; o += 8) {                         // one warp per observation: arg-max over g
        double bv = -INFINITY; int bg = 0;
        for (int g = lane; g < G; g += 32) { const double v = sc[(size_t)o * G + g]; if (v > bv) { bv = v; bg = g; } }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, off);
            const int og = __shfl_xor_sync(0xFFFFFFFFu, bg, off);
            if (ov > bv || (ov == bv && og < bg)) { bv = ov; bg = og; }
        }
        if (lane == 0) { pb_sm[o] = bv; best[((size_t)a * P + p) * Z + o] = bg; }
    }
    double part = 0.0;                                         // b_p . R[a]
    for (int s = tid; s < S; s += 256) part += B[(size_t)p * S + s] * R[(size_t)a * S + s];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, off);
    if (lane == 0) pb_sm[Z + wid] = part;
    __syncthreads();
    if (tid == 0) {
        double br = 0.0, fut = 0.0;
        for (int w = 0; w < 8; ++w) br += pb_sm[Z + w];
        for (int o = 0; o < Z; ++o) fut += pb_sm[o];
        value[(size_t)a * P + p] = br + discount * fut;
    }
}

__global__ void vi_pb_pick_kernel(int A, int P, const double *__restrict__ value, int32_t *__restrict__ action)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double bv = -INFINITY; int ba = 0;
    for (int a = 0; a < A; ++a) { const double v = value[(size_t)a * P + p]; if (v > bv) { bv = v; ba = a; } }
    action[p] = ba;
}

__global__ void __launch_bounds__(256) vi_pb_alpha_kernel(int S, int Z, int G, int P, const double *__restrict__ proj,
                                                          const double *__restrict__ R, const int32_t *__restrict__ best,
                                                          const int32_t *__restrict__ action, double discount,
                                                          double *__restrict__ alpha)
{
    extern __shared__ int pb_cols[];                           // [Z] column o*G + g* of the chosen action
    const int p = blockIdx.y, a = action[p];
    for (int o = threadIdx.x; o < Z; o += blockDim.x) pb_cols[o] = o * G + best[((size_t)a * P + p) * Z + o];
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double *row = proj + ((size_t)a * S + s) * (size_t)Z * G;
    double fut = 0.0;
    for (int o = 0; o < Z; ++o) fut += row[pb_cols[o]];
    alpha[(size_t)p * S + s] = R[(size_t)a * S + s] + discount * fut;
}

// the part of the point-based backup that follows the projection (proj_dev given, by either projection kernel)
extern "C" int vi_point_based_select_dev(int device, int S, int A, int Z, int G, int P, const double *R_dev,
                                         const double *beliefs_dev, double discount, const double *proj_dev,
                                         double *score_dev, int32_t *best_dev, double *value_dev, double *alpha_dev,
                                         int32_t *action_dev, int prune, double *pruned_dev, int32_t *pruned_action_dev,
                                         int32_t *keep_dev, int32_t *n_out_dev, void *stream, char *err, int errlen)
{
    const int N = Z * G;
    VCK(cudaSetDevice(device));
    if (P % GBM) { snprintf(err, errlen, "the number of belief points must be a multiple of 128"); return -1; }
    if (S % GBK || N % GBN) { snprintf(err, errlen, "S must be a multiple of 16 and Z*G of 64"); return -1; }
    if (Z > 2048) { snprintf(err, errlen, "|Z| > 2048"); return -1; }
    VCK(cudaFuncSetAttribute(vi_dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DGEMM_SMEM));
    cudaStream_t st = (cudaStream_t)stream;
    for (int a = 0; a < A; ++a) {                              // score[a] = B . proj[a]
        dim3 grid(N / GBN, P / GBM);
        vi_dgemm_dmma_kernel<<<grid, 256, DGEMM_SMEM, st>>>(P, N, S, beliefs_dev, proj_dev + (size_t)a * S * N,
                                                   score_dev + (size_t)a * P * N);
    }
    vi_pb_choose_kernel<<<dim3(P, A), 256, (Z + 8) * sizeof(double), st>>>(S, Z, G, P, score_dev, R_dev, beliefs_dev,
                                                                             discount, best_dev, value_dev);
    vi_pb_pick_kernel<<<(P + 255) / 256, 256, 0, st>>>(A, P, value_dev, action_dev);
    vi_pb_alpha_kernel<<<dim3((S + 255) / 256, P), 256, Z * sizeof(int), st>>>(S, Z, G, P, proj_dev, R_dev, best_dev,
                                                                               action_dev, discount, alpha_dev);
    if (prune) {
        vi_prune_kernel<<<(P + 127) / 128, 128, 0, st>>>(S, P, alpha_dev, 1e-10, keep_dev);
        vi_compact_kernel<<<1, 1024, 0, st>>>(S, P, alpha_dev, action_dev, keep_dev, pruned_dev, pruned_action_dev, n_out_dev);
    }
    VCK(cudaGetLastError());
    return 0;
}

extern "C" int vi_point_based_backup_dev(int device, int S, int A, int Z, int G, int P, const double *T_dev,
                                         const double *O_dev, const double *R_dev, const double *gamma_dev,
                                         const double *beliefs_dev, double discount, double *scratch_dev,
                                         double *proj_dev, double *score_dev, int32_t *best_dev, double *value_dev,
                                         double *alpha_dev, int32_t *action_dev, int prune, double *pruned_dev,
                                         int32_t *pruned_action_dev, int32_t *keep_dev, int32_t *n_out_dev, void *stream,
                                         char *err, int errlen)
{
    VCK(cudaSetDevice(device));
    if (P % GBM) { snprintf(err, errlen, "the number of belief points must be a multiple of 128"); return -1; }
    if (Z > 2048) { snprintf(err, errlen, "|Z| > 2048"); return -1; }
    int rc = vi_project_textbook_dev(device, S, A, Z, G, T_dev, O_dev, gamma_dev, scratch_dev, proj_dev, stream, err, errlen);
    if (rc) return rc;
    return vi_point_based_select_dev(device, S, A, Z, G, P, R_dev, beliefs_dev, discount, proj_dev, score_dev, best_dev,
                                     value_dev, alpha_dev, action_dev, prune, pruned_dev, pruned_action_dev, keep_dev,
                                     n_out_dev, stream, err, errlen);
}

// ---------------------------------------------------------------------------------------------------------
// (f).4: policy execution and the point-based value-iteration LOOP on the device.
// ---------------------------------------------------------------------------------------------------------
#define VI_RT 256                 // threads per block of the rollout / expansion kernels
#define VI_DOM_POLICY 5u          // Philox domains of these kernels (pomcp_device.cuh uses 0..3)
#define VI_DOM_EXPAND 6u

__device__ __forceinline__ double vi_u53(uint32_t a, uint32_t b)
{ return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0); }

// sum of one value per thread over the block, in a fixed order (pairwise halving): every thread gets the result
__device__ __forceinline__ double vi_block_sum(double v, double *red)
{
    red[threadIdx.x] = v;
    __syncthreads();
    for (int off = VI_RT / 2; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] = red[threadIdx.x] + red[threadIdx.x + off];
        __syncthreads();
    }
    const double s = red[0];
    __syncthreads();
    return s;
}

// first index j with u < p[0] + ... + p[j] (sequential, index order; the last index if rounding leaves u above the sum)
__device__ __forceinline__ int vi_sample_index(const double *p, int n, int stride, double u)
{
    double acc = 0.0;
    for (int j = 0; j < n; ++j) { acc = acc + p[(size_t)j * stride]; if (u < acc) return j; }
    return n - 1;
}

// nb[j] = O[a][j][z] * sum_i T[a][i][j] * b[i], then normalised (Model.belief_update / tiger_model.py:109-131 for any
// tabular model): thread j sums over i in index order, the normaliser is a fixed-order block sum.  Returns the
// normaliser = P(z | b, a) on every thread; if it is 0 (impossible observation) the belief is left unchanged.
__device__ __forceinline__ double vi_belief_update(int S, int Z, const double *Ta, const double *Oa, int z, double *b, double *nb,
                                                   double *red)
{
    double part = 0.0;
    for (int j = threadIdx.x; j < S; j += VI_RT) {
        double acc = 0.0;
        for (int i = 0; i < S; ++i) acc = acc + Ta[(size_t)i * S + j] * b[i];
        acc = acc * Oa[(size_t)j * Z + z];
        nb[j] = acc;
        part = part + acc;
    }
    const double tot = vi_block_sum(part, red);
    if (tot > 0.0) for (int j = threadIdx.x; j < S; j += VI_RT) b[j] = nb[j] / tot;
    __syncthreads();
    return tot;
}

// Policy execution (agent.py:215-264 run_value_iteration: arg-max alpha vector -> action -> environment step -> belief
// update, until a terminal step or max_steps), one block per epoch, the whole episode in one launch: no host round trip
// per step.  The environment is the tabular model itself (hidden state s ~ b0, s' ~ T[a][s], z ~ O[a][s'], r = R[a][s];
// actions in term_action_mask end the episode, as opening a door does in the reference's Tiger).
// out[epoch] = {discounted return, undiscounted return, steps, last action}.
__global__ void __launch_bounds__(VI_RT) vi_policy_rollout_kernel(int S, int A, int Z, int G, const double *__restrict__ alpha,
                                                                  const int32_t *__restrict__ act, const double *__restrict__ T,
                                                                  const double *__restrict__ O, const double *__restrict__ R,
                                                                  const double *__restrict__ b0, uint32_t term_action_mask,
                                                                  double discount, int max_steps, uint32_t key0, uint32_t key1,
                                                                  double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char vsm[];
    double *b = reinterpret_cast<double *>(vsm), *nb = b + S, *red = nb + S;
    double *wbest = red + VI_RT; int *wbest_g = reinterpret_cast<int *>(wbest + VI_RT / 32);
    __shared__ int sh_i[4];
    const int epoch = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < S; j += VI_RT) b[j] = b0[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        const Philox4 r = philox4x32_10(0xFFFFFFFFu, (uint32_t)epoch, 0u, VI_DOM_POLICY, key0, key1);
        sh_i[0] = vi_sample_index(b, S, 1, vi_u53(r.x, r.y));                 // the hidden true state
    }
    __syncthreads();
    int s = sh_i[0];
    double ret = 0.0, dret = 0.0, disc = 1.0;
    int steps = 0, last_a = -1;
    for (int t = 0; t < max_steps; ++t) {
        // arg-max_g alpha_g . b (value_iteration.py:161-181; lowest g on ties): a warp per vector, lane-strided sums
        double best = -INFINITY; int best_g = -1;
        for (int g = wid; g < G; g += VI_RT / 32) {
            double acc = 0.0;
            for (int j = lane; j < S; j += 32) acc = acc + alpha[(size_t)g * S + j] * b[j];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xFFFFFFFFu, acc, off);
            if (acc > best) { best = acc; best_g = g; }
        }
        if (lane == 0) { wbest[wid] = best; wbest_g[wid] = best_g; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double bv = -INFINITY; int bg = -1;
            for (int w = 0; w < VI_RT / 32; ++w)
                if (wbest_g[w] >= 0 && (wbest[w] > bv || (wbest[w] == bv && wbest_g[w] < bg))) { bv = wbest[w]; bg = wbest_g[w]; }
            const int a = act[bg];
            const Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)epoch, 0u, VI_DOM_POLICY, key0, key1);
            const int s2 = vi_sample_index(T + ((size_t)a * S + s) * S, S, 1, vi_u53(r.x, r.y));
            const int z = vi_sample_index(O + ((size_t)a * S + s2) * Z, Z, 1, vi_u53(r.z, r.w));
            sh_i[0] = a; sh_i[1] = s2; sh_i[2] = z;
        }
        __syncthreads();
        const int a = sh_i[0], s2 = sh_i[1], z = sh_i[2];
        const double rew = R[(size_t)a * S + s];
        ret = ret + rew; dret = dret + disc * rew; disc = disc * discount;
        s = s2; ++steps; last_a = a;
        if ((term_action_mask >> a) & 1u) break;                                // agent.py:244-246
        vi_belief_update(S, Z, T + (size_t)a * S * S, O + (size_t)a * S * Z, z, b, nb, red);   // :233-234
    }
    if (threadIdx.x == 0) {
        out[(size_t)epoch * 4 + 0] = dret; out[(size_t)epoch * 4 + 1] = ret;
        out[(size_t)epoch * 4 + 2] = (double)steps; out[(size_t)epoch * 4 + 3] = (double)last_a;
    }
}

extern "C" int vi_policy_rollout_dev(int device, int S, int A, int Z, int G, const double *alpha_dev, const int32_t *action_dev,
                                     const double *T_dev, const double *O_dev, const double *R_dev, const double *b0_dev,
                                     uint32_t term_action_mask, double discount, int max_steps, int n_epochs, uint64_t seed,
                                     double *out_dev, void *stream, char *err, int errlen)
{
    VCK(cudaSetDevice(device));
    if (S < 1 || G < 1 || n_epochs < 1 || A < 1 || A > 32) { snprintf(err, errlen, "bad sizes"); return -1; }
    const size_t smem = (size_t)(2 * S + VI_RT + VI_RT / 32) * sizeof(double) + (VI_RT / 32) * sizeof(int) + 16;
    if (smem > 200 * 1024) { snprintf(err, errlen, "|S| too large for the belief in shared memory"); return -1; }
    VCK(cudaFuncSetAttribute(vi_policy_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vi_policy_rollout_kernel<<<n_epochs, VI_RT, smem, (cudaStream_t)stream>>>(S, A, Z, G, alpha_dev, action_dev, T_dev, O_dev, R_dev,
                                                                                b0_dev, term_action_mask, discount, max_steps,
                                                                                (uint32_t)seed, (uint32_t)(seed >> 32), out_dev);
    VCK(cudaGetLastError());
    return 0;
}

// Belief-point expansion of point-based value iteration (stochastic simulation with a random action): belief p gets
// the successor b' = update(b_p, a, z), a uniform in [0, A), z ~ P(z | b_p, a); it is stored as point P + p.
__global__ void __launch_bounds__(VI_RT) vi_expand_beliefs_kernel(int S, int A, int Z, int P, const double *__restrict__ T,
                                                                  const double *__restrict__ O, double *__restrict__ beliefs,
                                                                  uint32_t round, uint32_t key0, uint32_t key1)
{
    extern __shared__ __align__(16) unsigned char vsm[];
    double *b = reinterpret_cast<double *>(vsm), *nb = b + S, *red = nb + S, *pz = red + VI_RT;
    __shared__ int sh_z;
    const int p = blockIdx.x;
    for (int j = threadIdx.x; j < S; j += VI_RT) b[j] = beliefs[(size_t)p * S + j];
    __syncthreads();
    const Philox4 r = philox4x32_10(round, (uint32_t)p, 0u, VI_DOM_EXPAND, key0, key1);
    const int a = (int)__umulhi(r.x, (uint32_t)A);
    const double *Ta = T + (size_t)a * S * S, *Oa = O + (size_t)a * S * Z;
    // pred[j] = sum_i T[a][i][j] b[i] (kept in nb), then P(z) = sum_j pred[j] O[a][j][z]
    for (int j = threadIdx.x; j < S; j += VI_RT) {
        double acc = 0.0;
        for (int i = 0; i < S; ++i) acc = acc + Ta[(size_t)i * S + j] * b[i];
        nb[j] = acc;
    }
    __syncthreads();
    for (int z = 0; z < Z; ++z) {
        double part = 0.0;
        for (int j = threadIdx.x; j < S; j += VI_RT) part = part + nb[j] * Oa[(size_t)j * Z + z];
        const double tot = vi_block_sum(part, red);
        if (threadIdx.x == 0) pz[z] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int z = 0; z < Z; ++z) tot = tot + pz[z];
        sh_z = vi_sample_index(pz, Z, 1, vi_u53(r.y, r.z) * tot);
    }
    __syncthreads();
    const int z = sh_z;
    double part = 0.0;
    for (int j = threadIdx.x; j < S; j += VI_RT) { const double v = nb[j] * Oa[(size_t)j * Z + z]; nb[j] = v; part = part + v; }
    const double tot = vi_block_sum(part, red);
    for (int j = threadIdx.x; j < S; j += VI_RT) beliefs[(size_t)(P + p) * S + j] = tot > 0.0 ? nb[j] / tot : b[j];
}

extern "C" int vi_expand_beliefs_dev(int device, int S, int A, int Z, int P, const double *T_dev, const double *O_dev,
                                     double *beliefs_dev, uint32_t round, uint64_t seed, void *stream, char *err, int errlen)
{
    VCK(cudaSetDevice(device));
    const size_t smem = (size_t)(2 * S + VI_RT + Z) * sizeof(double) + 16;
    if (smem > 200 * 1024) { snprintf(err, errlen, "|S| too large for the belief in shared memory"); return -1; }
    VCK(cudaFuncSetAttribute(vi_expand_beliefs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vi_expand_beliefs_kernel<<<P, VI_RT, smem, (cudaStream_t)stream>>>(S, A, Z, P, T_dev, O_dev, beliefs_dev, round,
                                                                        (uint32_t)seed, (uint32_t)(seed >> 32));
    VCK(cudaGetLastError());
    return 0;
}
