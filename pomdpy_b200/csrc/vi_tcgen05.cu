// vi_tcgen05.cu -- the textbook projection of value iteration on the 5th-generation tensor cores.
//
//   proj[a][s][o*G + g] = sum_s' T[a][s][s'] * O[a][s'][o] * gamma[g][s']            (value_iteration.py:51-67, per action
//                                                                                     a GEMM [S x S] . [S x Z*G])
// tcgen05 has no fp64 kind, so this is the FAST mode next to the exact fp64 DMMA kernel of vi_kernels.cu: every fp64
// operand x is split into two TF32 numbers hi = tf32(x), lo = tf32(x - hi) (22 mantissa bits together) and the product is
// taken as hi.hi + hi.lo + lo.hi -- three kind::tf32 MMAs per k-step into one fp32 accumulator in tensor memory
// ("TF32 x 3").  What is dropped is lo.lo (2^-22 relative) and the fp32 rounding of the accumulation over K; the result
// is fp32-grade (measured against numpy float64 in tests/test_vi.py and bench_vi.py), not the bit-exact fp64 of the
// default path.
//
// Kernel anatomy (the default kernel: one CTA per 128 x TN output tile, 128 threads):
//   thread 0   TMA producer: per k-block of 32 (= one 128-byte swizzle row) four cp.async.bulk.tensor.2d loads
//              (A hi, A lo, B hi, B lo; K-major, SWIZZLE_128B) into a STAGES-deep ring, completion on an mbarrier
//   thread 32  MMA issuer: 4 k-steps (UMMA_K = 8) x 3 products tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = TN,
//              NACC accumulators [128 lanes x TN columns] fp32 in TMEM, each over a contiguous 1/NACC of K (all 512
//              columns at TN = 128); tcgen05.commit frees the stage / signals the epilogue
//   all 4 warps  epilogue: tcgen05.ld.32x32b (a thread owns one row of the tile), the NACC partial sums added,
//              32-byte stores
// Shared-memory / instruction descriptors follow the bit layouts of CUTLASS's cute/arch/mma_sm100_desc.hpp
// (SmemDescriptor, InstrDescriptor); the tensor maps come from cuTensorMapEncodeTiled, looked up through the runtime
// (no link against libcuda).
// Two kernels: vi_tf32x3_gemm_kernel (one CTA per tile, 4 partial accumulators: the default) and
// vi_tf32x3_gemm_persistent_kernel (one CTA per SM, producer / MMA / four epilogue warps, two accumulator sets so that a
// tile's epilogue runs under the next tile's MMAs; 2 partial accumulators).  Measured on 4096^3: 0.586 / 0.569 ms
// (704 / 724 TFLOP/s of tf32 MMA work).  What mattered most was the MMA thread's own instruction stream: two 64-bit
// divisions per k-block (the accumulator index) in it cost 25 % of the kernel; 128 x 256 tiles and a grouped tile order
// (fewer DRAM reads) gained nothing.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/vi_b200.h"

#define VCK(call)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_));          \
            return -1;                                                               \
        }                                                                            \
    } while (0)

namespace {

constexpr int TC_M = 128;            // rows of an output tile (= TMEM lanes)
constexpr int TC_K = 32;             // fp32 elements per k-block: 128 bytes, one swizzle row
constexpr int TC_UK = 8;             // K of one kind::tf32 MMA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// K-major operand tile, rows of 128 bytes, SWIZZLE_128B (cute SmemDescriptor: start >> 4 at [0,14), leading byte offset
// >> 4 at [16,30) = 1 (unused for swizzled K-major), stride byte offset >> 4 at [32,46) = 8 rows x 128 B = 1024 B,
// version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64))
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute InstrDescriptor: c_format F32 = 1 at [4,6), a/b_format TF32 = 2 at [7,10) / [10,13), both K-major (bits 15, 16 = 0),
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{ return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

template <int TN, int STAGES>
struct TcSmem {
    static constexpr int A_BYTES = TC_M * TC_K * 4;                  // one of hi / lo
    static constexpr int B_BYTES = TN * TC_K * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024;        // + slack to align the ring to 1024 B
};

template <int TN, int STAGES, int NACC>
__global__ void __launch_bounds__(128, 1)
vi_tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                      const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                      double *__restrict__ C, int ldc, int K)
{
    using L = TcSmem<TN, STAGES>;
    extern __shared__ uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t ring = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // (a grouped tile order that keeps a wave's A and B panels inside the L2 cuts the DRAM reads from 1.1 GB to ~0.5 GB
    // per 4096^3 GEMM and changes nothing in time: the kernel is not DRAM-bound)
    const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * TN;
    const int nkb = K / TC_K;
    const int nacc = nkb < NACC ? nkb : NACC;                        // accumulators in use (every one gets k-blocks)

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&acc_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 1) {                                                 // one warp allocates the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TN * NACC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_slot;

    if (tid == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES; const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
            const uint32_t st = ring + (uint32_t)s * L::STAGE_BYTES, fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, (uint32_t)L::STAGE_BYTES);
            tma_load_2d(st, &mapAhi, kb * TC_K, m0, fb);
            tma_load_2d(st + L::A_BYTES, &mapAlo, kb * TC_K, m0, fb);
            tma_load_2d(st + 2 * L::A_BYTES, &mapBhi, kb * TC_K, n0, fb);
            tma_load_2d(st + 2 * L::A_BYTES + L::B_BYTES, &mapBlo, kb * TC_K, n0, fb);
        }
    } else if (tid == 32) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = umma_idesc_tf32(TC_M, TN);
        // NACC accumulators, each over a contiguous 1/NACC of K (summed by the epilogue): the tensor core's fp32
        // accumulation error grows with the length and the magnitude of a running sum.  (Counters, no division: this
        // thread's instruction stream is the issue path of the tensor pipe.)
        const uint64_t dring = umma_desc_k128(ring);
        const int part = nkb / nacc, extra = nkb - part * nacc;      // the first `extra` parts are one k-block longer
        int acc = 0, left = part + (extra > 0 ? 1 : 0);
        bool fresh = true;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES; const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(smem_u32(&full_bar[s]), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors: the start-address field counts 16-byte units, so tiles and k-steps are plain additions
            const uint64_t dst = dring + (uint64_t)((uint32_t)s * (L::STAGE_BYTES >> 4));
            const uint32_t td = tmem_acc + (uint32_t)(acc * TN);
#pragma unroll
            for (int k = 0; k < TC_K / TC_UK; ++k) {
                const uint64_t ko = (uint64_t)(k * TC_UK * 4 >> 4);  // 32 bytes along K inside the swizzle row
                const uint64_t dAh = dst + ko, dAl = dst + (L::A_BYTES >> 4) + ko;
                const uint64_t dBh = dst + (2 * L::A_BYTES >> 4) + ko, dBl = dst + ((2 * L::A_BYTES + L::B_BYTES) >> 4) + ko;
                umma_tf32(td, dAl, dBh, idesc, (fresh && k == 0) ? 0u : 1u);         // small terms first
                umma_tf32(td, dAh, dBl, idesc, 1u);
                umma_tf32(td, dAh, dBh, idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[s]));                     // the stage is free once these MMAs have read it
            fresh = false;
            if (--left == 0) { ++acc; left = part + (acc < extra ? 1 : 0); fresh = true; }
        }
        umma_commit(smem_u32(&acc_bar));                              // ... and the accumulator complete
    }
    __syncwarp();
    // ---- epilogue: TMEM -> registers -> fp64 -> global ----
    mbar_wait(smem_u32(&acc_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    double *crow = C + (size_t)(m0 + warp * 32 + lane) * ldc + n0;
#pragma unroll 1
    for (int c = 0; c < TN; c += 32) {
        // all NACC loads of the column group are issued before the one wait (a load + wait round trip per partial
        // sum serialises the epilogue); the partial sums are added in fp32 -- one rounding each, far below the
        // accumulation error -- ...
        uint32_t v[NACC][32];
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)((q < nacc ? q : 0) * TN + c);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]), "=r"(v[q][7]),
                  "=r"(v[q][8]), "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]), "=r"(v[q][14]), "=r"(v[q][15]),
                  "=r"(v[q][16]), "=r"(v[q][17]), "=r"(v[q][18]), "=r"(v[q][19]), "=r"(v[q][20]), "=r"(v[q][21]), "=r"(v[q][22]), "=r"(v[q][23]),
                  "=r"(v[q][24]), "=r"(v[q][25]), "=r"(v[q][26]), "=r"(v[q][27]), "=r"(v[q][28]), "=r"(v[q][29]), "=r"(v[q][30]), "=r"(v[q][31])
                : "r"(taddr) : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float sum[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            sum[j] = __uint_as_float(v[0][j]);
#pragma unroll
            for (int q = 1; q < NACC; ++q) sum[j] += q < nacc ? __uint_as_float(v[q][j]) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2)                              // ... and widened once
            *reinterpret_cast<double2 *>(crow + c + j) = make_double2((double)sum[j], (double)sum[j + 1]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TN * NACC) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// Persistent variant: one CTA per SM walks the tiles; 6 warps -- warp 0 TMA producer, warp 1 MMA issuer (and TMEM
// owner), warps 2..5 epilogue -- and TWO accumulator sets in TMEM (2 x NACC x TN columns = all 512 at TN = 128,
// NACC = 2), so the epilogue of a tile (TMEM -> registers -> fp64 -> global) runs under the MMAs of the next one and
// the operand ring never drains between tiles.
template <int TN, int STAGES, int NACC>
__global__ void __launch_bounds__(192, 1)
vi_tf32x3_gemm_persistent_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                                 const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                                 double *__restrict__ C, int ldc, int K, int tiles_m, int tiles_n)
{
    using L = TcSmem<TN, STAGES>;
    static_assert(2 * NACC * TN <= 512, "two accumulator sets must fit the 512 TMEM columns");
    extern __shared__ uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t ring = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = K / TC_K;
    const int nacc = nkb < NACC ? nkb : NACC;
    const int n_tiles = tiles_m * tiles_n;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&acc_full[b]), 1); mbar_init(smem_u32(&acc_empty[b]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(2 * NACC * TN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer: the ring runs on across tiles ----
            int s = 0; uint32_t ph = 0u;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * TC_M, n0 = (tile % tiles_n) * TN;   // as the per-tile kernel's grid: n fastest
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                    const uint32_t st = ring + (uint32_t)s * L::STAGE_BYTES, fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, (uint32_t)L::STAGE_BYTES);
                    tma_load_2d(st, &mapAhi, kb * TC_K, m0, fb);
                    tma_load_2d(st + L::A_BYTES, &mapAlo, kb * TC_K, m0, fb);
                    tma_load_2d(st + 2 * L::A_BYTES, &mapBhi, kb * TC_K, n0, fb);
                    tma_load_2d(st + 2 * L::A_BYTES + L::B_BYTES, &mapBlo, kb * TC_K, n0, fb);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer (counters only: this instruction stream is the issue path of the tensor pipe) ----
            constexpr uint32_t idesc = umma_idesc_tf32(TC_M, TN);
            const uint64_t dring = umma_desc_k128(ring);
            const int part = nkb / nacc, extra = nkb - part * nacc;
            int s = 0, it = 0; uint32_t ph = 0u;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1; const uint32_t bph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(smem_u32(&acc_empty[buf]), bph ^ 1u);      // the epilogue has drained this accumulator set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tset = tmem_acc + (uint32_t)(buf * NACC * TN);
                int acc = 0, left = part + (extra > 0 ? 1 : 0);
                bool fresh = true;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t dst = dring + (uint64_t)((uint32_t)s * (L::STAGE_BYTES >> 4));
                    const uint32_t td = tset + (uint32_t)(acc * TN);
#pragma unroll
                    for (int k = 0; k < TC_K / TC_UK; ++k) {
                        const uint64_t ko = (uint64_t)(k * TC_UK * 4 >> 4);
                        const uint64_t dAh = dst + ko, dAl = dst + (L::A_BYTES >> 4) + ko;
                        const uint64_t dBh = dst + (2 * L::A_BYTES >> 4) + ko, dBl = dst + ((2 * L::A_BYTES + L::B_BYTES) >> 4) + ko;
                        umma_tf32(td, dAl, dBh, idesc, (fresh && k == 0) ? 0u : 1u);
                        umma_tf32(td, dAh, dBl, idesc, 1u);
                        umma_tf32(td, dAh, dBh, idesc, 1u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    fresh = false;
                    if (--left == 0) { ++acc; left = part + (acc < extra ? 1 : 0); fresh = true; }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(&acc_full[buf]));
            }
        }
    } else {
        // ---- epilogue warps: TMEM lanes [32 * (warp % 4), +32) ----
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1; const uint32_t bph = (uint32_t)(it >> 1) & 1u;
            const int m0 = (tile / tiles_n) * TC_M, n0 = (tile % tiles_n) * TN;   // as the per-tile kernel's grid: n fastest
            mbar_wait(smem_u32(&acc_full[buf]), bph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            double *crow = C + (size_t)(m0 + quarter * 32 + lane) * ldc + n0;
            const uint32_t tset = tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * NACC * TN);
#pragma unroll 1
            for (int c = 0; c < TN; c += 32) {
                uint32_t v[NACC][32];
#pragma unroll
                for (int q = 0; q < NACC; ++q) {
                    const uint32_t taddr = tset + (uint32_t)((q < nacc ? q : 0) * TN + c);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]), "=r"(v[q][7]),
                          "=r"(v[q][8]), "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]), "=r"(v[q][14]), "=r"(v[q][15]),
                          "=r"(v[q][16]), "=r"(v[q][17]), "=r"(v[q][18]), "=r"(v[q][19]), "=r"(v[q][20]), "=r"(v[q][21]), "=r"(v[q][22]), "=r"(v[q][23]),
                          "=r"(v[q][24]), "=r"(v[q][25]), "=r"(v[q][26]), "=r"(v[q][27]), "=r"(v[q][28]), "=r"(v[q][29]), "=r"(v[q][30]), "=r"(v[q][31])
                        : "r"(taddr) : "memory");
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float s0 = __uint_as_float(v[0][j]), s1 = __uint_as_float(v[0][j + 1]);
#pragma unroll
                    for (int q = 1; q < NACC; ++q) {
                        s0 += q < nacc ? __uint_as_float(v[q][j]) : 0.0f;
                        s1 += q < nacc ? __uint_as_float(v[q][j + 1]) : 0.0f;
                    }
                    *reinterpret_cast<double2 *>(crow + c + j) = make_double2((double)s0, (double)s1);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));   // 4 arrivals (one per epilogue warp) free the set
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(2 * NACC * TN) : "memory");
}

__device__ __forceinline__ float tf32_rna(float x)
{ uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// x (fp64) -> hi = tf32(x), lo = tf32(x - hi)
__global__ void vi_split_tf32_kernel(size_t n, const double *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        const float h = tf32_rna((float)v);
        hi[i] = h; lo[i] = tf32_rna((float)(v - (double)h));
    }
}
// Bt[o*G + g][s'] = O[a][s'][o] * gamma[g][s']  (fp64 product, then split): the GEMM's B operand, K-major
__global__ void vi_scale_split_kernel(int S, int Z, int G, const double *__restrict__ Oa, const double *__restrict__ gamma,
                                      float *__restrict__ hi, float *__restrict__ lo)
{
    const size_t total = (size_t)Z * G * S;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int sp = (int)(i % S); const int col = (int)(i / S);
        const int o = col / G, g = col - o * G;
        const double v = Oa[(size_t)sp * Z + o] * gamma[(size_t)g * S + sp];
        const float h = tf32_rna((float)v);
        hi[i] = h; lo[i] = tf32_rna((float)(v - (double)h));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [rows x K] fp32 row-major, box = 32 (K) x box_rows, SWIZZLE_128B
static int make_map(CUtensorMap *m, const float *base, int rows, int K, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

template <int TN, int STAGES, int NACC>
static int launch_gemm(int M, int N, int K, const float *Ahi, const float *Alo, const float *Bhi, const float *Blo,
                       double *C, cudaStream_t st, char *err, int errlen)
{
    using L = TcSmem<TN, STAGES>;
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, Ahi, M, K, TC_M) || make_map(&mAl, Alo, M, K, TC_M) || make_map(&mBh, Bhi, N, K, TN) ||
        make_map(&mBl, Blo, N, K, TN)) { snprintf(err, errlen, "cuTensorMapEncodeTiled failed"); return -1; }
    VCK(cudaFuncSetAttribute(vi_tf32x3_gemm_kernel<TN, STAGES, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    dim3 grid(N / TN, M / TC_M);
    vi_tf32x3_gemm_kernel<TN, STAGES, NACC><<<grid, 128, L::TOTAL, st>>>(mAh, mAl, mBh, mBl, C, N, K);
    VCK(cudaGetLastError());
    return 0;
}

template <int TN, int STAGES, int NACC>
static int launch_gemm_persistent(int device, int M, int N, int K, const float *Ahi, const float *Alo, const float *Bhi,
                                  const float *Blo, double *C, cudaStream_t st, char *err, int errlen)
{
    using L = TcSmem<TN, STAGES>;
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, Ahi, M, K, TC_M) || make_map(&mAl, Alo, M, K, TC_M) || make_map(&mBh, Bhi, N, K, TN) ||
        make_map(&mBl, Blo, N, K, TN)) { snprintf(err, errlen, "cuTensorMapEncodeTiled failed"); return -1; }
    VCK(cudaFuncSetAttribute(vi_tf32x3_gemm_persistent_kernel<TN, STAGES, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    int sms = 0;
    VCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int tiles_m = M / TC_M, tiles_n = N / TN, n_tiles = tiles_m * tiles_n;
    vi_tf32x3_gemm_persistent_kernel<TN, STAGES, NACC><<<n_tiles < sms ? n_tiles : sms, 192, L::TOTAL, st>>>(
        mAh, mAl, mBh, mBl, C, N, K, tiles_m, tiles_n);
    VCK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" int vi_split_tf32_dev(int device, long long n, const double *x_dev, float *hi_dev, float *lo_dev, void *stream,
                                 char *err, int errlen)
{
    VCK(cudaSetDevice(device));
    vi_split_tf32_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((size_t)n, x_dev, hi_dev, lo_dev);
    VCK(cudaGetLastError());
    return 0;
}

extern "C" int vi_gemm_tf32x3_dev(int device, int M, int N, int K, const float *Ahi_dev, const float *Alo_dev,
                                  const float *Bthi_dev, const float *Btlo_dev, double *C_dev, int variant, void *stream,
                                  char *err, int errlen)
{
    VCK(cudaSetDevice(device));
    if (M % TC_M || K % TC_K || K < TC_K) { snprintf(err, errlen, "M must be a multiple of 128 and K of 32"); return -1; }
    if (N % 128) { snprintf(err, errlen, "N must be a multiple of 128"); return -1; }
    if (variant == VI_TC_PERSISTENT)
        return launch_gemm_persistent<128, 3, 2>(device, M, N, K, Ahi_dev, Alo_dev, Bthi_dev, Btlo_dev, C_dev, (cudaStream_t)stream, err, errlen);
    return launch_gemm<128, 3, 4>(M, N, K, Ahi_dev, Alo_dev, Bthi_dev, Btlo_dev, C_dev, (cudaStream_t)stream, err, errlen);
}

extern "C" int vi_project_textbook_tf32x3_dev(int device, int S, int A, int Z, int G, const float *Thi_dev,
                                              const float *Tlo_dev, const double *O_dev, const double *gamma_dev,
                                              float *scratch_dev, double *proj_dev, int variant, void *stream, char *err,
                                              int errlen)
{
    const int N = Z * G;
    VCK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    float *bhi = scratch_dev, *blo = scratch_dev + (size_t)N * S;
    for (int a = 0; a < A; ++a) {
        vi_scale_split_kernel<<<148 * 8, 256, 0, st>>>(S, Z, G, O_dev + (size_t)a * S * Z, gamma_dev, bhi, blo);
        const int rc = vi_gemm_tf32x3_dev(device, S, N, S, Thi_dev + (size_t)a * S * S, Tlo_dev + (size_t)a * S * S, bhi, blo,
                                          proj_dev + (size_t)a * S * N, variant, stream, err, errlen);
        if (rc) return rc;
    }
    VCK(cudaGetLastError());
    return 0;
}
