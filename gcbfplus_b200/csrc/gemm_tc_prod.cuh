// gemm_tc_prod.cuh -- tensor-core GEMMs whose A operand is PRODUCED inside the kernel (rollout / inference
// path), so the per-edge and per-agent inputs never round-trip through HBM and two launches disappear:
//
//   PROD_EDGE : A[e, :] = relu(feat_e @ W1[:ed] + W1[ed + sender_type] + W1[ed+3+2] + b1)   (edge_l1_kernel)
//               feat_e from agent / goal / hit states through the receiver-grouped edge lists
//   PROD_ATTN : A[a, :] = sum_e softmax_e(logit) * MSG[e, :]                                 (attn_aggregate_kernel)
//
// The four "operand" warps compute their row's 32 columns per k-block, split them into the tf32 hi / lo
// planes and store them straight into the SWIZZLE_128B K-major layout the MMA descriptor expects
// (16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) * 16)); B (weights, pre-split) still arrives by TMA.
#pragma once
#include "gemm_tc.cuh"
#include "gnn.cuh"

namespace gcbf {
namespace tc {

enum { PROD_EDGE = 1, PROD_ATTN = 2 };

struct ProdArgs {
    // PROD_EDGE
    gcbf_env_desc d;
    const float *W1, *b1, *agent, *goal, *hits;
    const int32_t *edge_recv, *edge_src;
    int clip_all;
    // PROD_ATTN
    const float *logits, *msg;
    const int32_t *row_start, *row_deg;
    int edge_cap;
};

// CHAIN (PROD_EDGE, BN = 128 only): a second GEMM is chained onto the tile while it is still on the SM --
//   logit[m] = relu(MSG[m, :128] @ A1 + b_g) . avec + c            (gate layer + folded gate vector, gnn.py:64-67)
// The epilogue warps hand the message tile over in shared memory (tf32 hi / lo planes in the K-major SWIZZLE_128B
// layout, 4 k-blocks of 32 = 128 KB over the drained pipeline stages 0-1), the gate weights stream through the third
// stage (two 32 KB slots), the accumulator lives in the second half of the CTA's TMEM.  One launch (and one
// round trip of MSG through L2) less per env-step; the fixed ~6 us of a GEMM launch is paid once for both layers.
struct ChainArgs {
    const float* bias_g;   // [128] gate hidden-layer bias
    const float* avec;     // [128] folded gate vector
    const float* cst;      // [1]   folded gate constant
    float* logits;         // [M]   out
};

template <int BN, int EPI, int PROD, int KIND, bool CHAIN>
__global__ void __launch_bounds__(THREADS_NN, 1)
gemm_tc_prod_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    const __grid_constant__ ProdArgs pa, const float* __restrict__ bias,
                    const float* __restrict__ bias2, float* __restrict__ C, const float* __restrict__ aux,
                    const int32_t* __restrict__ m_ptr, const int m_fixed, const int m_cap, const int K, const int N,
                    const int ndot, const __grid_constant__ CUtensorMap tmB2h, const __grid_constant__ CUtensorMap tmB2l,
                    const ChainArgs ch) {
    using CF = Cfg<BN>;
    using T = EnvTraits<KIND>;
    constexpr int ED = T::ED, SD = T::SD;
    constexpr int STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* conv = bars + STAGES;
    uint64_t* empty = bars + 2 * STAGES;
    uint64_t* tmem_full = bars + 3 * STAGES;
    uint64_t* tmem_empty = bars + 3 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
    uint64_t* b2_full = bars + 16;             // [2] CHAIN: gate-weight slot landed
    uint64_t* b2_empty = bars + 18;            // [2] CHAIN: gate-weight slot consumed
    uint64_t* a2_ready = bars + 20;            // CHAIN: message tile written to shared memory (128 arrivals)
    uint64_t* tmem2_full = bars + 21;          // CHAIN: gate accumulator ready / chained GEMM retired
    float* sW = reinterpret_cast<float*>(smem + STAGES * CF::STAGE_BYTES + 256);   // PROD_EDGE: [ED + 3][256]
    static_assert(!CHAIN || (BN == 128 && PROD == PROD_EDGE && STAGES == 3 && CF::STAGE_BYTES == 65536),
                  "chained gate GEMM: 128-wide message tile over three 64 KB stages");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = N / BN;                // 1, or 2 when a 256-wide layer is split to fill more SMs
    const int nkb = K / BK;

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
        if (CHAIN) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&b2_full[i], 1);
                mbar_init(&b2_empty[i], 1);
            }
            mbar_init(a2_ready, 128);
            mbar_init(tmem2_full, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (PROD == PROD_EDGE) {   // first message layer weights: W1[:ED] and the per-sender-type bias table
        for (int i = threadIdx.x; i < ED * 256; i += blockDim.x) sW[i] = pa.W1[i];
        for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
            const int t = i / 256, c = i % 256;
            sW[(ED + t) * 256 + c] = pa.W1[(ED + t) * 256 + c] + pa.W1[(ED + 3 + 2) * 256 + c] + pa.b1[c];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    int M = m_ptr ? *m_ptr : m_fixed;
    M = min(M, m_cap);
    const int n_tiles = ((M + BM - 1) / BM) * tiles_n;

    if (warp == 0) {
        // ================= TMA producer (weights only) =================
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const int n0 = (tile % tiles_n) * BN;
                // CHAIN: the previous tile's chained GEMM still reads stages 0-2 until it retires
                if (CHAIN && tcount > 0) mbar_wait(tmem2_full, (tcount - 1) & 1);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * CF::STAGE_BYTES;
                    mbar_expect_tx(&full[s], 2 * CF::B_BYTES);
                    tma_load_2d(st + 2 * CF::A_BYTES, &tmBh, &full[s], kb * BK, n0);
                    tma_load_2d(st + 2 * CF::A_BYTES + CF::B_BYTES, &tmBl, &full[s], kb * BK, n0);
                }
                if (CHAIN) {
                    mbar_wait(&tmem_full[0], tcount & 1);      // every main-loop MMA retired: stage 2 is free
                    for (int kb2 = 0; kb2 < 4; ++kb2) {
                        const uint32_t j2 = tcount * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                        mbar_wait(&b2_empty[slot], (use & 1) ^ 1);
                        uint8_t* sl = smem + 2 * CF::STAGE_BYTES + slot * 32768;
                        mbar_expect_tx(&b2_full[slot], 32768);
                        tma_load_2d(sl, &tmB2h, &b2_full[slot], kb2 * BK, 0);
                        tma_load_2d(sl + 16384, &tmB2l, &b2_full[slot], kb2 * BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = CHAIN ? 0u : (tcount & 1);
                const uint32_t tmem_d = tmem_base + acc * BN;
                mbar_wait(&tmem_empty[acc], (CHAIN ? (tcount & 1) : ((tcount >> 1) & 1)) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&conv[s], ph);     // A planes written by the operand warps
                    mbar_wait(&full[s], ph);     // B planes landed (TMA)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * CF::STAGE_BYTES);
                    const uint32_t a_lo = a_hi + CF::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * CF::A_BYTES;
                    const uint32_t b_lo = b_hi + CF::B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;
                        const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                        const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
                        umma_tf32(tmem_d, dal, dbh, idesc, (kb | k) != 0);
                        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[acc]);
                if (CHAIN) {
                    // ---- chained gate GEMM: D2[128 x 128] = MSG tile (shared memory) x A1, accumulator in columns BN..
                    mbar_wait(&tmem_empty[1], (tcount & 1) ^ 1);
                    mbar_wait(a2_ready, tcount & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb2 = 0; kb2 < 4; ++kb2) {
                        const uint32_t j2 = tcount * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                        mbar_wait(&b2_full[slot], use & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_hi = smem_u32(smem + kb2 * 32768);
                        const uint32_t a_lo = a_hi + 16384;
                        const uint32_t b_hi = smem_u32(smem + 2 * CF::STAGE_BYTES + slot * 32768);
                        const uint32_t b_lo = b_hi + 16384;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint32_t koff = k * UMMA_K * 4;
                            const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                            const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
                            umma_tf32(tmem_base + BN, dal, dbh, idesc, (kb2 | k) != 0);
                            umma_tf32(tmem_base + BN, dah, dbl, idesc, 1u);
                            umma_tf32(tmem_base + BN, dah, dbh, idesc, 1u);
                        }
                        umma_commit(&b2_empty[slot]);
                    }
                    umma_commit(tmem2_full);
                }
            }
        }
    } else if (warp < 6) {
        // ================= operand warps: produce the A tile (thread = row) =================
        const int r = threadIdx.x - 64;           // row inside the tile, 0..127
        const int A_tot = pa.d.n_graphs * pa.d.n_agents;
        uint32_t it = 0, tcount_p = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount_p) {
            const int m = (tile / tiles_n) * BM + r;
            const bool row_ok = m < M;
            if (CHAIN && tcount_p > 0) mbar_wait(tmem2_full, (tcount_p - 1) & 1);   // stages still feed the chained GEMM
            // ---- per-row setup
            float f[ED];
            int stype = 0;
            float att[8];
            int rs = 0, rd = 0;
            float mx = 0.f, inv_den = 0.f;
            if (PROD == PROD_EDGE) {
#pragma unroll
                for (int c = 0; c < ED; ++c) f[c] = 0.f;
                if (row_ok) {
                    const int a = min(max(pa.edge_recv[m], 0), A_tot - 1);
                    const int code = min(pa.edge_src[m], A_tot - 1);
                    float er[ED], es[ED], coef, nrm;
                    edge_state_dev<KIND>(pa.agent + (size_t)a * SD, er);
                    sender_state_dev<KIND>(code, a, pa.d.n_hits, pa.agent, pa.goal, pa.hits, es);
                    edge_feat_dev<KIND>(er, es, pa.clip_all || code == -1, pa.d.comm_radius, f, &coef, &nrm);
                    stype = (code >= 0) ? 2 : ((code == -1) ? 1 : 0);
                }
            } else {
                if (row_ok) {
                    rs = pa.row_start[m];
                    rd = pa.row_deg[m];
                    if (rs < 0 || rs + rd > pa.edge_cap) rd = 0;
                    mx = -INFINITY;
                    for (int e = rs; e < rs + rd; ++e) mx = fmaxf(mx, pa.logits[e]);
                    float den = 0.f;
                    for (int e = rs; e < rs + rd; ++e) den += expf(pa.logits[e] - mx);
                    inv_den = 1.f / den;
#pragma unroll
                    for (int q = 0; q < 8; ++q) att[q] = (q < rd) ? expf(pa.logits[rs + q] - mx) / den : 0.f;
                }
            }
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);     // stage free (its previous MMAs retired)
                uint8_t* hi_row = smem + s * CF::STAGE_BYTES + r * 128;
                uint8_t* lo_row = hi_row + CF::A_BYTES;
#pragma unroll
                for (int c = 0; c < 8; ++c) {     // 16-byte chunk c = columns kb*32 + 4c .. +3
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (row_ok) {
                        const int n = kb * BK + c * 4;
                        if (PROD == PROD_EDGE) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float y = sW[(ED + stype) * 256 + n + j];
#pragma unroll
                                for (int q = 0; q < ED; ++q) y = fmaf(f[q], sW[q * 256 + n + j], y);
                                v[j] = fmaxf(y, 0.f);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (q < rd) {
                                    const float4 mv = *reinterpret_cast<const float4*>(pa.msg + (size_t)(rs + q) * 128 + n);
                                    v[0] = fmaf(att[q], mv.x, v[0]);
                                    v[1] = fmaf(att[q], mv.y, v[1]);
                                    v[2] = fmaf(att[q], mv.z, v[2]);
                                    v[3] = fmaf(att[q], mv.w, v[3]);
                                }
                            }
                            for (int q = 8; q < rd; ++q) {
                                const float w = expf(pa.logits[rs + q] - mx) * inv_den;
                                const float4 mv = *reinterpret_cast<const float4*>(pa.msg + (size_t)(rs + q) * 128 + n);
                                v[0] = fmaf(w, mv.x, v[0]);
                                v[1] = fmaf(w, mv.y, v[1]);
                                v[2] = fmaf(w, mv.z, v[2]);
                                v[3] = fmaf(w, mv.w, v[3]);
                            }
                        }
                    }
                    float4 h, l;
                    h.x = rn_tf32(v[0]); h.y = rn_tf32(v[1]); h.z = rn_tf32(v[2]); h.w = rn_tf32(v[3]);
                    l.x = rn_tf32(v[0] - h.x); l.y = rn_tf32(v[1] - h.y);
                    l.z = rn_tf32(v[2] - h.z); l.w = rn_tf32(v[3] - h.w);
                    const int off = ((c ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(hi_row + off) = h;
                    *reinterpret_cast<float4*>(lo_row + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&conv[s]);
            }
        }
    } else {
        // ================= epilogue (warps 6..9) =================
        const int quarter = warp & 3;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            const uint32_t acc = CHAIN ? 0u : (tcount & 1);
            mbar_wait(&tmem_full[acc], CHAIN ? (tcount & 1) : ((tcount >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = quarter * 32 + lane;
            const int m = m0 + row;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                if (CHAIN) {
                    // message tile -> global (the aggregation kernel reads it) and -> shared memory as the A operand
                    // of the chained GEMM: k-block c0 / 32, row `row`, 16-byte chunks XOR-swizzled like the TMA does
                    uint8_t* hi_row = smem + (c0 >> 5) * 32768 + row * 128;
                    uint8_t* lo_row = hi_row + 16384;
                    float* crow = C + (size_t)m * N + n0 + c0;
                    const bool wide = (reinterpret_cast<uintptr_t>(crow) & 31) == 0;
                    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + c0 + j);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        if (m < M) {
                            if (!wide) *reinterpret_cast<float4*>(crow + j) = o;
                            else if (j & 4) st_global_v8(crow + j - 4, prev, o);
                            else prev = o;
                        }
                        float4 h, l;
                        h.x = rn_tf32(o.x); h.y = rn_tf32(o.y); h.z = rn_tf32(o.z); h.w = rn_tf32(o.w);
                        l.x = rn_tf32(o.x - h.x); l.y = rn_tf32(o.y - h.y);
                        l.z = rn_tf32(o.z - h.z); l.w = rn_tf32(o.w - h.w);
                        const int off = (((j >> 2) ^ (row & 7)) << 4);
                        *reinterpret_cast<float4*>(hi_row + off) = h;
                        *reinterpret_cast<float4*>(lo_row + off) = l;
                    }
                    continue;
                }
                if (m < M) {
                    const int n = n0 + c0;
                    float* crow = C + (size_t)m * N + n;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        const float4 bb = *reinterpret_cast<const float4*>(bias + n + j);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        if (bias2) {
                            const float4 b2 = *reinterpret_cast<const float4*>(bias2 + n + j);
                            o.x += b2.x; o.y += b2.y; o.z += b2.z; o.w += b2.w;
                        }
                        if (EPI == EPI_BIAS_RELU) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        *reinterpret_cast<float4*>(crow + j) = o;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (CHAIN) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> MMA (async proxy)
                mbar_arrive(a2_ready);
                mbar_arrive(&tmem_empty[0]);
                // ---- gate logit from the chained accumulator: relu(acc2 + b_g) . avec + c
                mbar_wait(tmem2_full, tcount & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float dot = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        dot = fmaf(fmaxf(__uint_as_float(v[j]) + ch.bias_g[c0 + j], 0.f), ch.avec[c0 + j], dot);
                }
                if (m < M) ch.logits[m] = dot + ch.cst[0];
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&tmem_empty[1]);
                continue;
            }
            mbar_arrive(&tmem_empty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                     : "memory");
    }
    (void)aux;
    (void)ndot;
}

// =====================================================================================================
// edge_chain_kernel: the rollout's edge kernel (PROD_EDGE + CHAIN of gemm_tc_prod_kernel, same arithmetic and the
// same bits) with EIGHT worker warps instead of 4 + 4: in-kernel %globaltimer stamps showed the tile bound by its 4
// producer warps (0.93 us per k-block against 0.55 us of MMA issue) and by a 4.1 us single-pass drain of the
// message tile.  Warps 2..9 first PRODUCE the A operand (two threads per edge row, 16 of the 32 columns of a k-block
// each), then DRAIN accumulator 0 (two warps per TMEM lane quarter, 64 columns each: global store + tf32 hi / lo
// hand-over planes for the chained gate GEMM); the gate logit stays one sequential fmaf chain per row on the four
// quarter-owning warps (the summation order is part of the result).  320 threads, 1 CTA / SM, persistent tile loop.
// =====================================================================================================
template <int KIND>
__global__ void __launch_bounds__(THREADS_NN, 1)
edge_chain_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const __grid_constant__ CUtensorMap tmB2h, const __grid_constant__ CUtensorMap tmB2l,
                  const __grid_constant__ ProdArgs pa, const float* __restrict__ bias, float* __restrict__ C,
                  const int32_t* __restrict__ m_ptr, const int m_cap, const ChainArgs ch) {
    using T = EnvTraits<KIND>;
    constexpr int ED = T::ED, SD = T::SD;
    constexpr int STG = 65536, A_BYTES = 16384, B_BYTES = 16384, BN = 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * STG);
    uint64_t* full = bars;                 // [3] weight planes landed (TMA)
    uint64_t* conv = bars + 3;             // [3] A planes written (256 arrivals)
    uint64_t* empty = bars + 6;            // [3] MMAs of the stage retired
    uint64_t* tf0 = bars + 9;              // message accumulator ready
    uint64_t* te0 = bars + 10;             // message accumulator drained (256 arrivals)
    uint64_t* te1 = bars + 11;             // gate accumulator drained (128 arrivals)
    uint64_t* b2_full = bars + 12;         // [2] gate-weight slot landed
    uint64_t* b2_empty = bars + 14;        // [2] gate-weight slot consumed
    uint64_t* a2_ready = bars + 16;        // message tile handed over in shared memory (256 arrivals)
    uint64_t* t2f = bars + 17;             // chained GEMM retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* sW = reinterpret_cast<float*>(smem + 3 * STG + 256);   // [ED + 3][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 3; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 256);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tf0, 1);
        mbar_init(te0, 256);
        mbar_init(te1, 128);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&b2_full[i], 1);
            mbar_init(&b2_empty[i], 1);
        }
        mbar_init(a2_ready, 256);
        mbar_init(t2f, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ED * 256; i += blockDim.x) sW[i] = pa.W1[i];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
        const int t = i / 256, c = i % 256;
        sW[(ED + t) * 256 + c] = pa.W1[(ED + t) * 256 + c] + pa.W1[(ED + 3 + 2) * 256 + c] + pa.b1[c];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int M = min(*m_ptr, m_cap);
    const int n_tiles = (M + BM - 1) / BM;
    constexpr uint32_t idesc = make_idesc(BM, BN);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, tc_ = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_) {
                if (tc_ > 0) mbar_wait(t2f, (tc_ - 1) & 1);       // the previous tile's chained GEMM still reads the stages
                for (int kb = 0; kb < 8; ++kb, ++it) {
                    const int s = it % 3;
                    mbar_wait(&empty[s], ((it / 3) & 1) ^ 1);
                    uint8_t* st = smem + s * STG;
                    mbar_expect_tx(&full[s], 2 * B_BYTES);
                    tma_load_2d(st + 2 * A_BYTES, &tmBh, &full[s], kb * BK, 0);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmBl, &full[s], kb * BK, 0);
                }
                mbar_wait(tf0, tc_ & 1);                          // main-loop MMAs retired: stage 2 is free
                for (int kb2 = 0; kb2 < 4; ++kb2) {
                    const uint32_t j2 = tc_ * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                    mbar_wait(&b2_empty[slot], (use & 1) ^ 1);
                    uint8_t* sl = smem + 2 * STG + slot * 32768;
                    mbar_expect_tx(&b2_full[slot], 32768);
                    tma_load_2d(sl, &tmB2h, &b2_full[slot], kb2 * BK, 0);
                    tma_load_2d(sl + 16384, &tmB2l, &b2_full[slot], kb2 * BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tc_ = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_) {
                mbar_wait(te0, (tc_ & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < 8; ++kb, ++it) {
                    const int s = it % 3;
                    const uint32_t ph = (it / 3) & 1;
                    mbar_wait(&conv[s], ph);
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * STG);
                    const uint32_t a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;
                        const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                        const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
                        umma_tf32(tmem_base, dal, dbh, idesc, (kb | k) != 0);
                        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(tf0);
                mbar_wait(te1, (tc_ & 1) ^ 1);
                mbar_wait(a2_ready, tc_ & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb2 = 0; kb2 < 4; ++kb2) {
                    const uint32_t j2 = tc_ * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                    mbar_wait(&b2_full[slot], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + kb2 * 32768);
                    const uint32_t a_lo = a_hi + 16384;
                    const uint32_t b_hi = smem_u32(smem + 2 * STG + slot * 32768);
                    const uint32_t b_lo = b_hi + 16384;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;
                        const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                        const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
                        umma_tf32(tmem_base + BN, dal, dbh, idesc, (kb2 | k) != 0);
                        umma_tf32(tmem_base + BN, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_base + BN, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&b2_empty[slot]);
                }
                umma_commit(t2f);
            }
        }
    } else {
        const int pr_ = threadIdx.x - 64;
        const int r = pr_ & 127;                  // producer: row of the tile
        const int half = pr_ >> 7;                // producer: 16-byte chunks 4 half .. 4 half + 3
        const int quarter = warp & 3;             // drain: TMEM lane quarter of this warp
        const int chalf = (warp - 2) >> 2;        // drain: message columns [64 chalf, 64 chalf + 64)
        const int row = quarter * 32 + lane;
        const int A_tot = pa.d.n_graphs * pa.d.n_agents;
        uint32_t it = 0, tc_ = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_) {
            {
                const int m = tile * BM + r;
                const bool row_ok = m < M;
                if (tc_ > 0) mbar_wait(t2f, (tc_ - 1) & 1);
                float f[ED];
                int stype = 0;
#pragma unroll
                for (int c = 0; c < ED; ++c) f[c] = 0.f;
                if (row_ok) {
                    const int a = min(max(pa.edge_recv[m], 0), A_tot - 1);
                    const int code = min(pa.edge_src[m], A_tot - 1);
                    float er[ED], es[ED], coef, nrm;
                    edge_state_dev<KIND>(pa.agent + (size_t)a * SD, er);
                    sender_state_dev<KIND>(code, a, pa.d.n_hits, pa.agent, pa.goal, pa.hits, es);
                    edge_feat_dev<KIND>(er, es, pa.clip_all || code == -1, pa.d.comm_radius, f, &coef, &nrm);
                    stype = (code >= 0) ? 2 : ((code == -1) ? 1 : 0);
                }
                for (int kb = 0; kb < 8; ++kb, ++it) {
                    const int s = it % 3;
                    mbar_wait(&empty[s], ((it / 3) & 1) ^ 1);
                    uint8_t* hi_row = smem + s * STG + r * 128;
                    uint8_t* lo_row = hi_row + A_BYTES;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int c = half * 4 + cc;
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (row_ok) {
                            const int n = kb * BK + c * 4;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float y = sW[(ED + stype) * 256 + n + j];
#pragma unroll
                                for (int q = 0; q < ED; ++q) y = fmaf(f[q], sW[q * 256 + n + j], y);
                                v[j] = fmaxf(y, 0.f);
                            }
                        }
                        float4 h, l;
                        h.x = rn_tf32(v[0]); h.y = rn_tf32(v[1]); h.z = rn_tf32(v[2]); h.w = rn_tf32(v[3]);
                        l.x = rn_tf32(v[0] - h.x); l.y = rn_tf32(v[1] - h.y);
                        l.z = rn_tf32(v[2] - h.z); l.w = rn_tf32(v[3] - h.w);
                        const int off = ((c ^ (r & 7)) << 4);
                        *reinterpret_cast<float4*>(hi_row + off) = h;
                        *reinterpret_cast<float4*>(lo_row + off) = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&conv[s]);
                }
            }
            const int m = tile * BM + row;
            mbar_wait(tf0, tc_ & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                uint8_t* hi_row = smem + (c0 >> 5) * 32768 + row * 128;
                uint8_t* lo_row = hi_row + 16384;
                float* crow = C + (size_t)m * BN + c0;
                const bool wide = (reinterpret_cast<uintptr_t>(crow) & 31) == 0;
                float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
                    const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + j);
                    o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    if (m < M) {
                        if (!wide) *reinterpret_cast<float4*>(crow + j) = o;
                        else if (j & 4) st_global_v8(crow + j - 4, prev, o);
                        else prev = o;
                    }
                    float4 h, l;
                    h.x = rn_tf32(o.x); h.y = rn_tf32(o.y); h.z = rn_tf32(o.z); h.w = rn_tf32(o.w);
                    l.x = rn_tf32(o.x - h.x); l.y = rn_tf32(o.y - h.y);
                    l.z = rn_tf32(o.z - h.z); l.w = rn_tf32(o.w - h.w);
                    const int off = (((j >> 2) ^ (row & 7)) << 4);
                    *reinterpret_cast<float4*>(hi_row + off) = h;
                    *reinterpret_cast<float4*>(lo_row + off) = l;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(a2_ready);
            mbar_arrive(te0);
            if (chalf == 0) {
                mbar_wait(t2f, tc_ & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float dot = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        dot = fmaf(fmaxf(__uint_as_float(v[j]) + ch.bias_g[c0 + j], 0.f), ch.avec[c0 + j], dot);
                }
                if (m < M) ch.logits[m] = dot + ch.cst[0];
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(te1);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

template <int KIND>
inline int32_t launch_edge_chain(const CUtensorMap& tmB, const CUtensorMap& tmBl, const CUtensorMap& tmG, const CUtensorMap& tmGl,
                                 const ProdArgs& pa, const float* bias, float* C, RowCount rc, int grid, cudaStream_t st,
                                 const ChainArgs& chain) {
    constexpr int smem = 3 * 65536 + 256 + 9 * 256 * 4 + 1024;
    auto kern = edge_chain_kernel<KIND>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done = true;
    }
    kern<<<grid, THREADS_NN, smem, st>>>(tmB, tmBl, tmG, tmGl, pa, bias, C, rc.ptr, rc.cap, chain);
    count_launch();
    return check_launch("edge_chain_kernel");
}

template <int BN, int EPI, int PROD, int KIND, bool CHAIN = false>
inline int32_t launch_prod_inst(const CUtensorMap& tmB, const CUtensorMap& tmBl, const ProdArgs& pa, const float* bias,
                                const float* bias2, float* C, RowCount rc, int K, int N, int grid, cudaStream_t st,
                                const CUtensorMap* tmB2h = nullptr, const CUtensorMap* tmB2l = nullptr,
                                const ChainArgs* chain = nullptr) {
    constexpr int smem = Cfg<BN>::SMEM_BYTES + 9 * 256 * 4;
    auto kern = gemm_tc_prod_kernel<BN, EPI, PROD, KIND, CHAIN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done = true;
    }
    ChainArgs ch;
    memset(&ch, 0, sizeof(ch));
    if (chain) ch = *chain;
    kern<<<grid, THREADS_NN, smem, st>>>(tmB, tmBl, pa, bias, bias2, C, nullptr, rc.ptr, rc.fixed, rc.cap, K, N, 0,
                                         tmB2h ? *tmB2h : tmB, tmB2l ? *tmB2l : tmBl, ch);
    count_launch();
    return check_launch("gemm_tc_prod_kernel");
}

// MSG[e, :128] = relu-layer-1(edge e) @ W23 + b23: edge_l1 producer + folded message GEMM (K = 256, N = 128).
inline int32_t launch_edge_msg(const gcbf_env_desc* d, const float* W1, const float* b1, const float* agent,
                               const float* goal, const float* hits, const int32_t* edge_recv,
                               const int32_t* edge_src, const int32_t* counters, int clip_all, const float* Bt_hi,
                               const float* Bt_lo, const float* bias, float* msg, cudaStream_t st,
                               const float* gate_Bt_hi = nullptr, const float* gate_Bt_lo = nullptr,
                               const ChainArgs* chain = nullptr) {
    // chain != nullptr: the gate layer (K = N = 128, weights gate_Bt_*) and its folded gate vector run inside the same
    // kernel and the logits are written to chain->logits
    ProdArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.d = *d;
    pa.W1 = W1; pa.b1 = b1; pa.agent = agent; pa.goal = goal; pa.hits = hits;
    pa.edge_recv = edge_recv; pa.edge_src = edge_src; pa.clip_all = clip_all;
    const int K = 256, N = 128;
    CUtensorMap tmB, tmBl;
    if (int32_t r = make_map(&tmB, Bt_hi, N, K, N)) return r;
    if (int32_t r = make_map(&tmBl, Bt_lo, N, K, N)) return r;
    const RowCount rc{counters, 0, d->edge_cap};
    const int grid = min((d->edge_cap + BM - 1) / BM, sm_count());
    if (chain) {
        CUtensorMap tmG, tmGl;
        if (int32_t r = make_map(&tmG, gate_Bt_hi, 128, 128, 128)) return r;
        if (int32_t r = make_map(&tmGl, gate_Bt_lo, 128, 128, 128)) return r;
        // GCBF_CHAIN8=0: the 4 + 4 warp version (gemm_tc_prod_kernel<..., CHAIN>), kept for A/B measurements
        static const bool eight = [] { const char* e = getenv("GCBF_CHAIN8"); return !(e && e[0] == '0'); }();
        if (eight && rc.ptr != nullptr) {
            GCBF_DISPATCH_ENV(d->env_kind, { return launch_edge_chain<KIND>(tmB, tmBl, tmG, tmGl, pa, bias, msg, rc, grid, st, *chain); });
            return -1;
        }
        GCBF_DISPATCH_ENV(d->env_kind, {
            return launch_prod_inst<128, EPI_BIAS, PROD_EDGE, KIND, true>(tmB, tmBl, pa, bias, nullptr, msg, rc, K, N, grid, st,
                                                                          &tmG, &tmGl, chain);
        });
        return -1;
    }
    GCBF_DISPATCH_ENV(d->env_kind, {
        return launch_prod_inst<128, EPI_BIAS, PROD_EDGE, KIND>(tmB, tmBl, pa, bias, nullptr, msg, rc, K, N, grid, st);
    });
    return -1;
}

// V1[a, :256] = relu(aggregate(a) @ U1' + bias + bias2): attention-aggregate producer + update layer 1 (K = 128).
inline int32_t launch_attn_upd(const gcbf_env_desc* d, const float* logits, const float* msg, const int32_t* row_start,
                               const int32_t* row_deg, const float* Bt_hi, const float* Bt_lo, const float* bias,
                               const float* bias2, float* v1, cudaStream_t st) {
    ProdArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.d = *d;
    pa.logits = logits; pa.msg = msg; pa.row_start = row_start; pa.row_deg = row_deg; pa.edge_cap = d->edge_cap;
    const int K = 128, N = 256;
    const int A = d->n_graphs * d->n_agents;
    const int tiles_m = (A + BM - 1) / BM;
    const int bn = (2 * tiles_m <= sm_count()) ? 128 : 256;
    CUtensorMap tmB, tmBl;
    if (int32_t r = make_map(&tmB, Bt_hi, N, K, bn)) return r;
    if (int32_t r = make_map(&tmBl, Bt_lo, N, K, bn)) return r;
    const RowCount rc{nullptr, A, A};
    const int grid = min(tiles_m * (N / bn), sm_count());
    if (bn == 128)
        return launch_prod_inst<128, EPI_BIAS_RELU, PROD_ATTN, GCBF_ENV_DOUBLE_INTEGRATOR>(tmB, tmBl, pa, bias, bias2, v1, rc, K, N, grid, st);
    return launch_prod_inst<256, EPI_BIAS_RELU, PROD_ATTN, GCBF_ENV_DOUBLE_INTEGRATOR>(tmB, tmBl, pa, bias, bias2, v1, rc, K, N, grid, st);
}

}  // namespace tc
}  // namespace gcbf
