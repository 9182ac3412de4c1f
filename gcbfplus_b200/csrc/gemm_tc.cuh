// gemm_tc.cuh -- Blackwell tensor-core GEMM for the dense MLP layers: tcgen05.mma (kind::tf32)
// with fp32 operands split as x = hi + lo (3xTF32: hi*hi + hi*lo + lo*hi, fp32 accumulation in
// TMEM), operands staged by TMA (SWIZZLE_128B, K-major) through an mbarrier pipeline.
//
//   C[M,N] = epi( A[M,K] @ Bt[N,K]^T )        A, Bt row-major fp32 (both "K-major")
//
// Accuracy: each product carries a relative error ~2^-21 (the dropped lo*lo term and the tf32
// truncation of lo), i.e. fp32-class results (tests: <= 2e-6 relative to |A||B|), which keeps the
// <= 1e-5 parity bar of the GNN outputs -- a single-pass TF32/BF16 MMA (2^-11 / 2^-8) would not.
//
// Warp roles of gemm_tc_kernel (320 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM alloc + MMA
// issuer (1 lane), warps 2-5 = operand split of A (hi/lo in shared memory; the weight planes arrive
// pre-split), warps 6-9 = epilogue (TMEM -> registers -> global, bias / ReLU / ReLU-mask / accumulate /
// row dot products) on a double-buffered accumulator.  Persistent tile loop; M may come from a device
// counter (edge count) so the launch is CUDA-graph friendly.  gemm_tn_tc_kernel (weight gradient, both
// operands MN-major and split in shared memory, column sums fused) is described at its definition.
// Measured (tools/gemm_tile_prof.py): a 128 x 256 tile's epilogue takes 10.5 us against
// 4.4 us of MMAs at K = 128 -- the kernel runs at the pace of its 4 epilogue warps.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"

namespace gcbf {
namespace tc {

constexpr int BM = 128;        // rows per tile (= UMMA M, TMEM lanes)
constexpr int BK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32: 32 bytes per MMA
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 256-bit global store (sm_100: STG.E.ENL2.256): one full 32-byte sector per lane.  The epilogue owns one output row
// per lane, so a 128-bit store touches half a sector and the row needs twice the LSU instructions.
__device__ __forceinline__ void st_global_v8(float* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

__device__ __forceinline__ void ld_global_v8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p)
                 : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024 B | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int THREADS_NN = 320;   // warp 0 TMA, warp 1 MMA, warps 2-5 operand split (A), warps 6-9 epilogue

template <int BN>
struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 2 : 3;
    static constexpr int A_BYTES = BM * BK * 4;     // 16 KB
    static constexpr int B_BYTES = BN * BK * 4;     // 16 / 32 KB
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi + lo of both operands
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// k-blocks of 16 floats (64-byte rows, SWIZZLE_64B) halve the stage size: twice the stages in flight for the same
// shared memory, which is what hides the TMA latency behind the 3 MMAs per k-step (the 32-float ring had 2 stages).
template <int BN, int BKc>
struct CfgK {
    static constexpr int A_BYTES = BM * BKc * 4;
    static constexpr int B_BYTES = BN * BKc * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;          // 2 / 3 (BKc = 32), 4 / 6 (BKc = 16)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};
// K-major descriptor for a row of BKc floats: SWIZZLE_128B (layout 2, SBO 1024) or SWIZZLE_64B (layout 4, SBO 512)
template <int BKc>
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
    constexpr uint64_t sbo = (BKc == 32) ? 1024 : 512;
    constexpr uint64_t layout = (BKc == 32) ? 2 : 4;
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// hi = tf32(x) rounded to nearest (13 low mantissa bits cleared, so the tensor core's own fp32->tf32
// conversion is exact), lo = tf32(x - hi): |x - hi - lo| <= 2^-24 |x|.
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// Weights are split once per parameter update (gcbf_prepare_params); activations are split in shared memory.
static __global__ void split_tf32_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = in[i];
        const float h = rn_tf32(x);
        hi[i] = h;
        lo[i] = rn_tf32(x - h);
    }
}

template <int BN, int EPI, bool ACCUM, int BKc>
__global__ void __launch_bounds__(THREADS_NN, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
               const __grid_constant__ CUtensorMap tmBl, const float* __restrict__ bias,
               const float* __restrict__ bias2, float* __restrict__ C, const float* __restrict__ aux,
               const int32_t* __restrict__ m_ptr, const int m_fixed, const int m_cap, const int K, const int N,
               const int ndot) {
    using CF = CfgK<BN, BKc>;
    constexpr int STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE_BYTES);
    uint64_t* full = bars;                     // [STAGES] TMA bytes landed
    uint64_t* conv = bars + STAGES;            // [STAGES] A hi/lo split done (128 arrivals)
    uint64_t* empty = bars + 2 * STAGES;       // [STAGES] MMAs of the stage retired
    uint64_t* tmem_full = bars + 3 * STAGES;   // [2] accumulator ready for the epilogue
    uint64_t* tmem_empty = bars + 3 * STAGES + 2;   // [2] accumulator drained (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = N / BN;                // 1, or 2 when a 256-wide layer is split to fill more SMs
    const int nkb = K / BKc;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 128);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    int M = m_ptr ? *m_ptr : m_fixed;
    M = min(M, m_cap);
    const int n_tiles = ((M + BM - 1) / BM) * tiles_n;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * CF::STAGE_BYTES;
                    mbar_expect_tx(&full[s], CF::A_BYTES + 2 * CF::B_BYTES);
                    tma_load_2d(st, &tmA, &full[s], kb * BKc, m0);
                    tma_load_2d(st + 2 * CF::A_BYTES, &tmBh, &full[s], kb * BKc, n0);
                    tma_load_2d(st + 2 * CF::A_BYTES + CF::B_BYTES, &tmBl, &full[s], kb * BKc, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                const uint32_t tmem_d = tmem_base + acc * BN;
                mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&conv[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * CF::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + CF::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * CF::A_BYTES;
                    const uint32_t b_lo = b_hi + CF::B_BYTES;
#pragma unroll
                    for (int k = 0; k < BKc / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;  // bytes inside the 128-byte swizzle row
                        const uint64_t dah = make_desc_k<BKc>(a_hi + koff), dal = make_desc_k<BKc>(a_lo + koff);
                        const uint64_t dbh = make_desc_k<BKc>(b_hi + koff), dbl = make_desc_k<BKc>(b_lo + koff);
                        umma_tf32(tmem_d, dal, dbh, idesc, (kb | k) != 0);   // small terms first
                        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp < 6) {
        // ================= A-operand split (warps 2..5, 128 threads) =================
        const int et = threadIdx.x - 64;          // 0..127
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                float4* h4 = reinterpret_cast<float4*>(smem + s * CF::STAGE_BYTES);
                float4* l4 = reinterpret_cast<float4*>(smem + s * CF::STAGE_BYTES + CF::A_BYTES);
                constexpr int NV = BKc / 4;                                  // A tile = 128 NV float4
                float4 v[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) v[j] = h4[et + 128 * j];
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 h, l;
                    h.x = rn_tf32(v[j].x); h.y = rn_tf32(v[j].y); h.z = rn_tf32(v[j].z); h.w = rn_tf32(v[j].w);
                    l.x = rn_tf32(v[j].x - h.x); l.y = rn_tf32(v[j].y - h.y);
                    l.z = rn_tf32(v[j].z - h.z); l.w = rn_tf32(v[j].w - h.w);
                    h4[et + 128 * j] = h;
                    l4[et + 128 * j] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&conv[s]);
            }
        }
    } else {
        // ================= epilogue (warps 6..9): TMEM -> registers -> global =================
        const int quarter = warp & 3;             // TMEM lane quarter this warp may access
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            const uint32_t acc = tcount & 1;
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = quarter * 32 + lane;
            const int m = m0 + row;
            float dot = 0.f;
            float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                // ReLU-mask epilogue: the 32 mask values of this row chunk are requested BEFORE the accumulator read, so
                // the global-load latency overlaps the TMEM load instead of sitting between it and the stores
                float4 mkp[8];
                if (EPI == EPI_RELU_MASK && m < M) {
                    const float* mrow = aux + (size_t)m * N + n0 + c0;
                    if ((reinterpret_cast<uintptr_t>(mrow) & 31) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) ld_global_v8(mrow + 8 * j, mkp[2 * j], mkp[2 * j + 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) mkp[j] = *reinterpret_cast<const float4*>(mrow + 4 * j);
                    }
                }
                tmem_ld32(tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                if (EPI == EPI_RELU_DOTN) {  // output layer partial sums over this column tile
                    const float* hw = aux + (size_t)(n0 + c0) * ndot;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float h = fmaxf(__uint_as_float(v[j]) + bias[n0 + c0 + j], 0.f);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < ndot) dq[q] = fmaf(h, hw[j * ndot + q], dq[q]);
                    }
                    if (c0 + 32 == BN && m < M)
                        *reinterpret_cast<float4*>(C + ((size_t)(n0 / BN) * m_cap + m) * 4) = make_float4(dq[0], dq[1], dq[2], dq[3]);
                    continue;
                }
                if (EPI == EPI_RELU_DOT) {   // gate logit: relu(acc + bias) . aux  (the row never leaves the SM)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        dot = fmaf(fmaxf(__uint_as_float(v[j]) + bias[c0 + j], 0.f), aux[c0 + j], dot);
                    if (c0 + 32 == BN && m < M) C[m] = dot + bias2[0];
                    continue;
                }
                if (m < M) {
                    const int n = n0 + c0;
                    float* crow = C + (size_t)m * N + n;
                    const float* mrow = (EPI == EPI_RELU_MASK) ? aux + (size_t)m * N + n : nullptr;
                    // 256-bit accesses when the row is 32-byte aligned (warp-uniform: N % 8 == 0)
                    const bool wide = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(mrow)) & 31) == 0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float4 o[2];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh)
                            o[hh] = make_float4(__uint_as_float(v[j + 4 * hh]), __uint_as_float(v[j + 4 * hh + 1]),
                                                __uint_as_float(v[j + 4 * hh + 2]), __uint_as_float(v[j + 4 * hh + 3]));
                        if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const float4 bb = *reinterpret_cast<const float4*>(bias + n + j + 4 * hh);
                                o[hh].x += bb.x; o[hh].y += bb.y; o[hh].z += bb.z; o[hh].w += bb.w;
                                if (bias2) {
                                    const float4 b2 = *reinterpret_cast<const float4*>(bias2 + n + j + 4 * hh);
                                    o[hh].x += b2.x; o[hh].y += b2.y; o[hh].z += b2.z; o[hh].w += b2.w;
                                }
                                if (EPI == EPI_BIAS_RELU) {
                                    o[hh].x = fmaxf(o[hh].x, 0.f); o[hh].y = fmaxf(o[hh].y, 0.f);
                                    o[hh].z = fmaxf(o[hh].z, 0.f); o[hh].w = fmaxf(o[hh].w, 0.f);
                                }
                            }
                        } else if (EPI == EPI_RELU_MASK) {
                            const float4 mk[2] = {mkp[j / 4], mkp[j / 4 + 1]};
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                o[hh].x = mk[hh].x > 0.f ? o[hh].x : 0.f; o[hh].y = mk[hh].y > 0.f ? o[hh].y : 0.f;
                                o[hh].z = mk[hh].z > 0.f ? o[hh].z : 0.f; o[hh].w = mk[hh].w > 0.f ? o[hh].w : 0.f;
                            }
                        }
                        if (ACCUM) {
                            float4 old[2];
                            if (wide) {
                                ld_global_v8(crow + j, old[0], old[1]);
                            } else {
                                old[0] = *reinterpret_cast<const float4*>(crow + j);
                                old[1] = *reinterpret_cast<const float4*>(crow + j + 4);
                            }
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                o[hh].x += old[hh].x; o[hh].y += old[hh].y; o[hh].z += old[hh].z; o[hh].w += old[hh].w;
                            }
                        }
                        if (wide) {
                            st_global_v8(crow + j, o[0], o[1]);
                        } else {
                            *reinterpret_cast<float4*>(crow + j) = o[0];
                            *reinterpret_cast<float4*>(crow + j + 4) = o[1];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&tmem_empty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                     : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, bk cols], 128-byte (bk = 32) or 64-byte (bk = 16) swizzle.
inline int32_t make_map(CUtensorMap* map, const float* ptr, int rows, int cols, int box_rows, int bk = BK) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return -2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d", (int)r, rows, cols);
        return -2;
    }
    return 0;
}

// k-block size of the NN kernel: 32 (default: 128-byte rows, 2 / 3 stages) or 16 (GCBF_TC_BK=16: 64-byte rows,
// SWIZZLE_64B, 4 / 6 stages).  Measured on the train step: 9.39 ms (32) vs 9.90 ms (16) -- the kernel is not short of
// pipeline depth (it is paced by its epilogue, see the file header); kept as an option.
inline int tc_bk() {
    static const int bk = [] {
        const char* e = getenv("GCBF_TC_BK");
        return (e && atoi(e) == 16) ? 16 : 32;
    }();
    return bk;
}

template <int BN, int BKc>
inline int32_t launch_bn(int epi, bool accum, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBl,
                         const float* bias,
                         const float* bias2, float* C, const float* aux, RowCount rc, int K, int N, int grid,
                         cudaStream_t st, int ndot = 0) {
    constexpr int smem = CfgK<BN, BKc>::SMEM_BYTES;
#define GCBF_TC_CASE(E, ACC)                                                                                      \
    do {                                                                                                          \
        auto kern = gemm_tc_kernel<BN, E, ACC, BKc>;                                                                 \
        static bool attr_done = false;                                                                            \
        if (!attr_done) {                                                                                         \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                        \
            attr_done = true;                                                                                     \
        }                                                                                                         \
        kern<<<grid, THREADS_NN, smem, st>>>(tmA, tmB, tmBl, bias, bias2, C, aux, rc.ptr, rc.fixed, rc.cap, K, N, \
                                             ndot);                                                               \
    } while (0)
    if (!accum) {
        switch (epi) {
            case EPI_BIAS: GCBF_TC_CASE(EPI_BIAS, false); break;
            case EPI_BIAS_RELU: GCBF_TC_CASE(EPI_BIAS_RELU, false); break;
            case EPI_NONE: GCBF_TC_CASE(EPI_NONE, false); break;
            case EPI_RELU_MASK: GCBF_TC_CASE(EPI_RELU_MASK, false); break;
            case EPI_RELU_DOT: GCBF_TC_CASE(EPI_RELU_DOT, false); break;
            case EPI_RELU_DOTN: GCBF_TC_CASE(EPI_RELU_DOTN, false); break;
            default: set_error("bad epilogue"); return -1;
        }
    } else {
        switch (epi) {
            case EPI_NONE: GCBF_TC_CASE(EPI_NONE, true); break;
            case EPI_RELU_MASK: GCBF_TC_CASE(EPI_RELU_MASK, true); break;
            default: set_error("bad accumulate epilogue"); return -1;
        }
    }
#undef GCBF_TC_CASE
    count_launch();
    return check_launch("gemm_tc_kernel");
}

// C[M,N] = epi(A[M,K] @ Bt[N,K]^T) with Bt given as its tf32 split (Bt_hi + Bt_lo, split_tf32_kernel).
// A must be backed by at least rc.cap rows.
inline int32_t launch_gemm_tc(int epi, bool accum, const float* A, const float* Bt_hi, const float* Bt_lo,
                              const float* bias, const float* bias2, float* C, const float* aux, RowCount rc, int K,
                              int N, cudaStream_t st, int ndot = 0, int* parts_out = nullptr) {
    if (K % BK != 0 || (N != 128 && N != 256)) {
        set_error("gemm_tc: K=%d N=%d unsupported", K, N);
        return -1;
    }
    const int rows = rc.ptr ? rc.cap : min(rc.fixed, rc.cap);
    if (rows <= 0) return 0;
    const int tiles_m = (rows + BM - 1) / BM;
    // a 256-wide layer over few row tiles is split into two 128-wide column tiles: twice the CTAs (more SMs
    // busy), half the weight traffic and MMA time per CTA, 3 pipeline stages instead of 2
    static const bool force128 = [] { const char* e = getenv("GCBF_TC_BN128"); return e && e[0] == '1'; }();   // A/B switch
    const int bn = (N == 256 && (2 * tiles_m <= sm_count() || force128) && epi != EPI_RELU_DOT) ? 128 : N;
    CUtensorMap tmA, tmB, tmBl;
    const int bk = tc_bk();
    if (int32_t r = make_map(&tmA, A, rc.cap, K, BM, bk)) return r;
    if (int32_t r = make_map(&tmB, Bt_hi, N, K, bn, bk)) return r;
    if (int32_t r = make_map(&tmBl, Bt_lo, N, K, bn, bk)) return r;
    const int grid = min(tiles_m * (N / bn), sm_count());
    if (parts_out) *parts_out = N / bn;
    if (bk == 32) {
        if (bn == 256) return launch_bn<256, 32>(epi, accum, tmA, tmB, tmBl, bias, bias2, C, aux, rc, K, N, grid, st, ndot);
        return launch_bn<128, 32>(epi, accum, tmA, tmB, tmBl, bias, bias2, C, aux, rc, K, N, grid, st, ndot);
    }
    if (bn == 256) return launch_bn<256, 16>(epi, accum, tmA, tmB, tmBl, bias, bias2, C, aux, rc, K, N, grid, st, ndot);
    return launch_bn<128, 16>(epi, accum, tmA, tmB, tmBl, bias, bias2, C, aux, rc, K, N, grid, st, ndot);
}


// =====================================================================================================
// backward-weight on tensor cores:  C[K1, N] += sum_m w(m) X[m, k1] dY[m, n]
// D[M_mma = 128 k1-rows, N_mma = N] += A B^T with A = X^T and B = dY^T, both "MN-major" (the MMA-K index is
// the row index m of X / dY).  TMA boxes of [32 rows(m) x 32 floats] with SWIZZLE_128B land exactly in the
// MN-major tf32 swizzle atoms.  For MN-major 32-bit operands the only legal shared-memory layout is
// SWIZZLE_128B_BASE32B (CUTLASS sm100_common.inl:92; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of
// 4 m-rows x 128 B with the 32-byte chunks XOR-ed by (row % 4); atoms along MN at LBO = 4096 B (one box),
// along K at SBO = 512 B.
// Split over M: CTA (tile, split) walks 32-row chunks {split, split + S, ...}; result red.add'ed to C.
// =====================================================================================================
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t box_bytes = 4096) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(box_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
    return make_idesc(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

constexpr int THREADS_TN = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 operand split + column sums + epilogue

// `cs1` / `cs2` (optional): column sums sum_m w(m) dY[m, n] -- the bias gradient of the same layer (and the agent
// one-hot row of the update layer) -- accumulated by the operand-split warps from the very values they convert, so no
// separate pass over dY is needed.  Only the CTAs of k1-tile 0 contribute.
// RCH = rows (MMA-K extent) per pipeline stage: 16 (default; 4 / 6 stages) or 32 (2 / 3 stages)
template <int BN, int RCH>
__global__ void __launch_bounds__(THREADS_TN, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                  float* __restrict__ C, const float* __restrict__ roww, const int32_t* __restrict__ row2agent,
                  const int32_t* __restrict__ m_ptr, const int m_fixed, const int m_cap, const int N, const int splits,
                  const int n_agents_total, float* __restrict__ cs1, float* __restrict__ cs2) {
    constexpr int A_BYTES = 128 * RCH * 4;           // 4 boxes of [RCH m x 32 k1]
    constexpr int B_BYTES = BN * RCH * 4;            // BN/32 boxes of [RCH m x 32 n]
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
    constexpr int BOX = RCH * 128;                   // bytes of one box
    constexpr int BOX4 = BOX / 16;                   // float4 per box
    constexpr int NB = BN / 32;
    constexpr int NJA = (4 * BOX4) / 256, NJB = (NB * BOX4) / 256;   // float4 per thread and plane
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* conv = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* tmem_full = bars + 3 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int M = m_ptr ? *m_ptr : m_fixed;
    M = min(M, m_cap);
    const int tile = blockIdx.x / splits, split = blockIdx.x % splits;
    const int k1_0 = tile * 128;
    const int n_chunks = (M + RCH - 1) / RCH;
    const int n_my = (split < n_chunks) ? (n_chunks - split + splits - 1) / splits : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 256);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int m0 = (split + it * splits) * RCH;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
#pragma unroll
                for (int i = 0; i < 4; ++i) tma_load_2d(st + i * BOX, &tmX, &full[s], k1_0 + 32 * i, m0);
#pragma unroll
                for (int j = 0; j < NB; ++j) tma_load_2d(st + 2 * A_BYTES + j * BOX, &tmY, &full[s], 32 * j, m0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_mn(128, BN);
            for (int it = 0; it < n_my; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&conv[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_hi + 2 * A_BYTES;
                const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < RCH / 8; ++k) {
                    const uint32_t koff = k * 1024;      // 8 m-rows
                    const uint64_t dah = make_desc_mn(a_hi + koff, BOX), dal = make_desc_mn(a_lo + koff, BOX);
                    const uint64_t dbh = make_desc_mn(b_hi + koff, BOX), dbl = make_desc_mn(b_lo + koff, BOX);
                    umma_tf32(tmem_base, dal, dbh, idesc, (it | k) != 0);
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(&empty[s]);
            }
            if (n_my > 0) umma_commit(tmem_full);
        }
    } else {
        // 256 operand-split threads: float4 index i = et + 256 j of a plane lies in box i / BOX4, box row (i % BOX4) / 8
        // (the same for every j: 256 % BOX4 == 0), 16-byte unit et % 8
        const int et = threadIdx.x - 64;
        const int quarter = warp & 3;
        const int brow = (et % BOX4) >> 3;
        const bool do_cs = (cs1 != nullptr) && tile == 0;
        auto rn = [](float x) { return rn_tf32(x); };
        float4 csum[NJB];
#pragma unroll
        for (int j = 0; j < NJB; ++j) csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < n_my; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const int m = (split + it * splits) * RCH + brow;
            const bool vrow = m < M;
            float wrow = 1.f;
            if (roww && vrow) {
                int ag = row2agent ? row2agent[m] : m;
                ag = min(max(ag, 0), n_agents_total - 1);
                wrow = roww[ag];
            }
            mbar_wait(&full[s], ph);
            uint8_t* st = smem + s * STAGE_BYTES;
            {   // X: rows >= M hold stale data (possibly NaN): select, do not multiply
                float4* h4 = reinterpret_cast<float4*>(st);
                float4* l4 = reinterpret_cast<float4*>(st + A_BYTES);
#pragma unroll
                for (int j = 0; j < NJA; ++j) {
                    const int i = et + 256 * j;
                    float4 v = h4[i];
                    if (!vrow) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 h, l;
                    h.x = rn(v.x); h.y = rn(v.y); h.z = rn(v.z); h.w = rn(v.w);
                    l.x = rn(v.x - h.x); l.y = rn(v.y - h.y); l.z = rn(v.z - h.z); l.w = rn(v.w - h.w);
                    h4[i] = h;
                    l4[i] = l;
                }
            }
            {   // dY, scaled by the row weight
                float4* h4 = reinterpret_cast<float4*>(st + 2 * A_BYTES);
                float4* l4 = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int j = 0; j < NJB; ++j) {
                    const int i = et + 256 * j;
                    float4 v = h4[i];
                    if (vrow) { v.x *= wrow; v.y *= wrow; v.z *= wrow; v.w *= wrow; }
                    else v = make_float4(0.f, 0.f, 0.f, 0.f);
                    csum[j].x += v.x; csum[j].y += v.y; csum[j].z += v.z; csum[j].w += v.w;
                    float4 h, l;
                    h.x = rn(v.x); h.y = rn(v.y); h.z = rn(v.z); h.w = rn(v.w);
                    l.x = rn(v.x - h.x); l.y = rn(v.y - h.y); l.z = rn(v.z - h.z); l.w = rn(v.w - h.w);
                    h4[i] = h;
                    l4[i] = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&conv[s]);
        }
        if (n_my > 0) {
            mbar_wait(tmem_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (do_cs) {
                // every MMA has retired: stage 0 is free.  SWIZZLE_128B_ATOM_32B: the 32-byte chunk q of box row r sits at
                // position q ^ (r % 4), so 16-byte unit u of row r holds columns 8 ((u / 2) ^ (r % 4)) + 4 (u % 2) .. + 3
                float* s_cs = reinterpret_cast<float*>(smem);
                for (int c = et; c < BN; c += 256) s_cs[c] = 0.f;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int u = et & 7;
                const int col0 = 8 * ((u >> 1) ^ (brow & 3)) + 4 * (u & 1);
#pragma unroll
                for (int j = 0; j < NJB; ++j) {
                    const int box = (et + 256 * j) / BOX4;
                    atomicAdd(&s_cs[32 * box + col0 + 0], csum[j].x);
                    atomicAdd(&s_cs[32 * box + col0 + 1], csum[j].y);
                    atomicAdd(&s_cs[32 * box + col0 + 2], csum[j].z);
                    atomicAdd(&s_cs[32 * box + col0 + 3], csum[j].w);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int c = et; c < BN; c += 256) {
                    const float v = s_cs[c];
                    atomicAdd(cs1 + c, v);
                    if (cs2) atomicAdd(cs2 + c, v);
                }
            }
            const int row = quarter * 32 + lane;
            float* crow = C + (size_t)(k1_0 + row) * N;
            const int chalf = (warp - 2) >> 2;          // two warps per TMEM lane quarter: each takes half of the columns
#pragma unroll 1
            for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                // 128-bit vector reductions (REDG.E.ADD.F32x4): 8 per 32 columns instead of 32 scalar atomics -- the
                // lane owns 32 consecutive floats of one C row (16-byte aligned: N % 4 == 0, c0 % 32 == 0)
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(crow + c0 + j),
                                 "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                 "f"(__uint_as_float(v[j + 3]))
                                 : "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

inline int32_t make_map_box(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows, int box_cols,
                            CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return -2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
        return -2;
    }
    return 0;
}

// C[K1, N] += sum_m w(m) X[m, :K1] dY[m, :N]; X row stride ldx (>= K1, multiple of 4), K1 % 128 == 0, N in {128, 256}.
// cs1 / cs2 (optional): += sum_m w(m) dY[m, :N] (bias gradient fused into the same pass over dY).
inline int32_t launch_gemm_tn_tc(const float* X, int ldx, const float* dY, float* C, const float* roww,
                                 const int32_t* row2agent, RowCount rc, int K1, int N, int n_agents_total,
                                 cudaStream_t st, float* cs1 = nullptr, float* cs2 = nullptr) {
    if (K1 % 128 != 0 || (N != 128 && N != 256) || ldx % 4 != 0) {
        set_error("gemm_tn_tc: K1=%d N=%d ldx=%d unsupported", K1, N, ldx);
        return -1;
    }
    const int rows = rc.ptr ? rc.cap : min(rc.fixed, rc.cap);
    if (rows <= 0) return 0;
    static const int rch = [] {
        const char* e = getenv("GCBF_TC_TN_ROWS");
        return (e && atoi(e) == 32) ? 32 : 16;
    }();
    CUtensorMap tmX, tmY;
    if (int32_t r = make_map_box(&tmX, X, rc.cap, K1, ldx, rch, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return r;
    if (int32_t r = make_map_box(&tmY, dY, rc.cap, N, N, rch, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return r;
    const int tiles = K1 / 128;
    const int chunks = (rows + rch - 1) / rch;
    int splits = max(1, sm_count() / tiles);
    splits = min(splits, max(1, chunks * rch / 128));       // >= 128 rows per CTA
    const int grid = tiles * splits;
    const int smem = 192 * 1024 + 1024 + 256;
#define GCBF_TN_CASE(BN_, R_)                                                                                         \
    do {                                                                                                              \
        static bool done = false;                                                                                     \
        if (!done) { cudaFuncSetAttribute(gemm_tn_tc_kernel<BN_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); done = true; } \
        gemm_tn_tc_kernel<BN_, R_><<<grid, THREADS_TN, smem, st>>>(tmX, tmY, C, roww, row2agent, rc.ptr, rc.fixed, rc.cap, N,  \
                                                                   splits, n_agents_total, cs1, cs2);                \
    } while (0)
    if (N == 256) { if (rch == 32) GCBF_TN_CASE(256, 32); else GCBF_TN_CASE(256, 16); }
    else { if (rch == 32) GCBF_TN_CASE(128, 32); else GCBF_TN_CASE(128, 16); }
#undef GCBF_TN_CASE
    count_launch();
    return check_launch("gemm_tn_tc_kernel");
}

}  // namespace tc
}  // namespace gcbf
