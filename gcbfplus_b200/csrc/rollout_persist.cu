// rollout_persist.cu -- the whole closed-loop rollout of trainer/utils.py:25-55 as ONE persistent kernel: one
// thread-block CLUSTER per environment, a loop over the T env-steps inside the kernel, cluster barriers where the
// 5-launch path (gcbf_rollout_step) has kernel boundaries.
//
// Why: environments are independent graphs, and at BASELINE's configs[2] (16 envs x 512 agents) every kernel of the
// 5-launch env-step is a single wave whose duration is its per-CTA latency chain plus ~4 us of launch / prologue
// (mbarrier init, TMEM allocation, layer-1 table, tensor-map fetch, obstacle / ray tables): ~20 of the 76 us.  Here
// that setup is paid once per ROLLOUT, the 4 kernel boundaries per step become barrier.cluster phases (~0.2 us), and
// an environment's messages / aggregates / hidden rows stay in L2 between phases.
//
// Per env-step (cluster of C CTAs, 512 threads each; per-env receiver-ordered edge lists):
//   E   edge tiles of 128 edges -> CTA (tile % C): edge features + message layer 1 produced in-CTA -> folded message
//       GEMM 256->128 (tcgen05 3xTF32, TMEM acc 0) -> chained gate GEMM 128->128 (acc 1) -> logits        [gemm_tc_prod.cuh]
//   A   warp per receiver: segment softmax + weighted aggregate                                             [gnn.cuh attn_aggregate_kernel]
//   U1  (agent tile, column half) items: update layer 128->256, bias + one-hot row, ReLU                   [gemm_tc.cuh EPI_BIAS_RELU]
//   U2  same items: folded update/head layer 256->256, ReLU, output-layer partial sums                      [gemm_tc.cuh EPI_RELU_DOTN]
//   G   policy tail (tanh, a = 2 pi + u_ref, clip, Euler; record action / next state / reward / cost) + LiDAR +
//       stable top-k + radius neighbour lists of the next state, rows laid out in agent order through a
//       CTA-local prefix sum and a cluster-wide exchange of the CTA totals over distributed shared memory    [geometry.cu graph_build_kernel]
// Every arithmetic step is the one the 5-launch path performs (same operand split, same MMA order, same epilogues,
// same reduction orders), so the two paths give the same bits (tests/test_gpu_rollout.py).  This translation unit is
// compiled with -fmad=false like geometry.cu (the LiDAR / dynamics code must keep one rounding per operation); the
// GEMM-side code uses explicit fmaf wherever the other translation units rely on contraction.
//
// Replaces: gcbfplus/trainer/utils.py:25-55 (rollout scan body), algo/gcbf_plus.py:176-186, env/double_integrator.py:145-198,
// 223-320, env/utils.py:49-131, nn/gnn.py:44-104 -- for the 2-D environments (SingleIntegrator, DoubleIntegrator, DubinsCar)
// with n_agents <= 512; LinearDrone (514 rays, 33 KB of per-warp alpha scratch) stays on the 5-launch path.
#include <stdlib.h>

#include "gemm_tc.cuh"
#include "geometry_dev.cuh"
#include "gnn.cuh"

namespace gcbf {
namespace rp {
using namespace tc;

constexpr int PT = 512;             // threads per CTA
constexpr int PW = PT / 32;         // warps per CTA
constexpr int STG = 65536;          // one pipeline stage: A hi | A lo | B hi | B lo, 16 KB each (BN = 128)
constexpr int MAX_N = 512;
constexpr int MAX_OBS = 32;

// barrier indices
enum { B_FULL = 0, B_CONV = 3, B_EMPTY = 6, B_TF0 = 9, B_TE0 = 10, B_TE1 = 11, B_B2F = 12, B_B2E = 14, B_A2R = 16, B_T2F = 17,
       B_COUNT = 18 };

struct PArgs {
    gcbf_env_desc d;
    int T, cap_env, C;
    // folded policy weights (gcbf_prepare_infer) and raw layer-1 / bias rows
    const float *W1, *b1, *b23, *bias_g, *avec, *cst, *b_u1, *b_u1row, *buh, *ho, *bho;
    const float *goal, *obstacles, *ray_table;
    float *agent, *hits, *actions, *rewards, *costs;     // trajectory record
    int32_t* counters;                                   // [T + 1][4]
    // per-environment scratch (L2)
    int32_t *row_start, *row_deg, *edge_recv, *edge_src; // [2][...] double-buffered lists
    float *msg, *logit, *ag, *v1, *z;
    unsigned long long* prof;                            // optional [T + 1][8] %globaltimer stamps of cluster 0 / CTA 0 (ns)
    // mode 0: one hardware cluster of C CTAs per environment (every barrier is barrier.cluster).
    // mode 1 ("pairs"): B200 keeps at most 15 clusters of 8 CTAs resident and BASELINE's config has 16 environments, so
    // the C CTAs of an environment are C/2 hardware clusters of 2.  A pair owns 2 APC consecutive agents, their edge
    // rows (its own segment of the environment's edge lists), their edge tiles and their agent tile: the phases
    // E -> A -> U1 -> U2 and the graph build only need pair barriers (barrier.cluster, DSMEM for the row prefix);
    // the one environment-wide dependency -- the policy tail needs every agent's next state -- is a software barrier
    // (arrival counter in global memory) once per step.
    int soft;                                            // 0 / 1 = mode
    unsigned* gbar;                                      // [E] arrival counters of the environment barrier (zeroed by the launcher)
    int* gtot;                                           // [E][8] (unused since the pair mode; kept for the layout)
};

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    // bounded wait: a protocol bug must end in a trap (error return), never in a hung GPU
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// software group barrier (soft mode): bar.sync, then one thread publishes this CTA's writes (gpu-scope fence, cumulative
// over the bar.sync), arrives on the environment's counter and spins with acquire loads until all C CTAs of the group
// have arrived for this generation; its trailing gpu-scope fence invalidates the SM's L1 for the loads that follow
__device__ __forceinline__ void soft_group_sync(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v < target && clock64() - t0 > 4000000000ll) __trap();
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// 12 MMAs of one 32-wide k-block of a 3xTF32 product (small terms first), D at tmem_d
__device__ __forceinline__ void mma_kblock(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           bool first) {
    constexpr uint32_t idesc = make_idesc(BM, 128);
#pragma unroll
    for (int k = 0; k < BK / UMMA_K; ++k) {
        const uint32_t koff = k * UMMA_K * 4;
        const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
        const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
        umma_tf32(tmem_d, dal, dbh, idesc, !(first && k == 0));
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
    }
}

template <int KIND>
__global__ void __launch_bounds__(PT, 1)
rollout_persist_kernel(const __grid_constant__ PArgs P, const __grid_constant__ CUtensorMap tmW23h,
                       const __grid_constant__ CUtensorMap tmW23l, const __grid_constant__ CUtensorMap tmA1h,
                       const __grid_constant__ CUtensorMap tmA1l, const __grid_constant__ CUtensorMap tmU1h,
                       const __grid_constant__ CUtensorMap tmU1l, const __grid_constant__ CUtensorMap tmUHh,
                       const __grid_constant__ CUtensorMap tmUHl, const __grid_constant__ CUtensorMap tmAG,
                       const __grid_constant__ CUtensorMap tmV1) {
    using TR = EnvTraits<KIND>;
    constexpr int SD = TR::SD, ED = TR::ED, NU = TR::NU, PD = TR::PD;
    static_assert(PD == 2, "persistent rollout kernel: 2-D environments");
    constexpr int OBW = 16, OBS2 = 24;
    constexpr int A_BYTES = 16384, B_BYTES = 16384;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * STG);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 3 * STG + 256);
    int* s_tot = reinterpret_cast<int*>(smem + 3 * STG + 256 + 16);       // [8] CTA edge totals of the cluster
    float* sW = reinterpret_cast<float*>(smem + 3 * STG + 512);            // [(ED + 3)][256]
    float* sst = sW + 7 * 256;                                             // [N][SD] states of all agents of the environment
    float* sobs = sst + MAX_N * 4;                                         // [O][24]
    float* stab = sobs + MAX_OBS * OBS2;                                   // [32][2]
    unsigned* sbits = reinterpret_cast<unsigned*>(stab + 64);              // [APC][n_words]
    int* s_off = reinterpret_cast<int*>(sbits + 64 * 16);                  // [APC + 1]
    unsigned* s_hb = reinterpret_cast<unsigned*>(s_off + 72);              // [APC]
    float* s_red = reinterpret_cast<float*>(s_hb + 64);                    // [3][PW]

    const gcbf_env_desc& d = P.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = P.C;
    uint32_t rank_u;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
    (void)rank_u;
    const bool soft = P.soft != 0;
    const int rank = (int)(blockIdx.x % C);          // CTA inside the environment
    const int env = blockIdx.x / C;
    const int L = soft ? 2 : C;                      // CTAs per LOCAL group = hardware cluster size
    const int lrank = rank % L, grp = rank / L;      // (cluster rank, group inside the environment)
    unsigned n_sync = 0;          // environment barriers passed so far (pair mode: arrival target = n_sync * C)
    // SYNC_LOCAL: barrier of the hardware cluster (the CTAs that share edge rows / agent tiles).
    // SYNC_ENV: all C CTAs of the environment (mode 0: the same cluster barrier; pair mode: software barrier).
#define SYNC_LOCAL() cluster_sync_all()
#define SYNC_ENV()                                                     \
    do {                                                               \
        if (soft) {                                                    \
            ++n_sync;                                                  \
            soft_group_sync(P.gbar + env, n_sync * (unsigned)C);       \
        } else {                                                       \
            cluster_sync_all();                                        \
        }                                                              \
    } while (0)
    const int N = d.n_agents, E = d.n_graphs, O = d.n_obs, R = d.n_hits, cap = P.cap_env;
    const int A_tot = E * N;
    const int APC = (N + C - 1) / C;                 // agents per CTA
    const int a_lo = min(rank * APC, N), a_hi = min(a_lo + APC, N);
    const int n_words = (N + 31) / 32;
    const size_t env_e0 = (size_t)env * cap;          // first edge slot of this environment
    const int env_a0 = env * N;                       // first global agent id
    const int seg_cap = cap / (C / L);                // edge rows of one local group (mode 0: the whole environment)
    const int seg_off = grp * seg_cap;                // ... and where they start inside the environment's lists
    const int ga_lo = min(grp * L * APC, N), ga_hi = min(ga_lo + L * APC, N);   // agents of my local group

    // ---------------------------------------------------------------- one-time setup
    if (tid == 0) {
        for (int s = 0; s < 3; ++s) {
            mbar_init(&bars[B_FULL + s], 1);
            mbar_init(&bars[B_CONV + s], 256);
            mbar_init(&bars[B_EMPTY + s], 1);
        }
        mbar_init(&bars[B_TF0], 1);
        mbar_init(&bars[B_TE0], 256);
        mbar_init(&bars[B_TE1], 128);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[B_B2F + i], 1);
            mbar_init(&bars[B_B2E + i], 1);
        }
        mbar_init(&bars[B_A2R], 256);
        mbar_init(&bars[B_T2F], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // message layer 1: W1[:ED] and the per-sender-type bias table (gemm_tc_prod.cuh)
    for (int i = tid; i < ED * 256; i += PT) sW[i] = P.W1[i];
    for (int i = tid; i < 3 * 256; i += PT) {
        const int t = i / 256, c = i % 256;
        sW[(ED + t) * 256 + c] = P.W1[(ED + t) * 256 + c] + P.W1[(ED + 3 + 2) * 256 + c] + P.b1[c];
    }
    // obstacles of this environment (+ derived far-skip fields) and the ray table stay resident for the whole rollout
    if (O > 0) {
        const float* ob = P.obstacles + (d.obs_per_graph ? (size_t)env * O * OBW : 0);
        for (int i = tid; i < O * OBW; i += PT) sobs[(i / OBW) * OBS2 + (i % OBW)] = ob[i];
    }
    for (int i = tid; i < d.n_rays * PD; i += PT) stab[i] = P.ray_table[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int o = tid; o < O; o += PT) {
        float* ob = sobs + OBS2 * o;
        const float reach = d.comm_radius + sqrtf(ob[2] * ob[2] + ob[3] * ob[3]) + 2e-3f;
        ob[14] = reach * reach;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int kp = (k + 3) & 3;
            ob[15 + k] = ob[6 + 2 * kp] - ob[6 + 2 * k];
            ob[19 + k] = ob[7 + 2 * kp] - ob[7 + 2 * k];
        }
    }
    const uint32_t tmem_base = *tmem_slot;
    uint32_t it = 0;      // k-blocks pushed through the 3-stage ring so far (all roles advance it identically)
    uint32_t ne = 0;      // edge tiles this CTA has processed (chain barriers)
    uint32_t n0 = 0;      // uses of accumulator 0 (edge tiles + U1 / U2 items)
    int M_cur = 0;        // edge rows of my local group's segment in the current graph
    SYNC_LOCAL();
    if (P.prof != nullptr && rank == 0 && tid == 0) P.prof[(size_t)(P.T + 1) * 8 + 2 * env] = gtime();   // cluster start

    // =================================================================================================
    for (int t = -1; t < P.T; ++t) {
        const int b = t & 1;                       // list half holding the graph of state t (t = -1: none)
        const float* agent_t = P.agent + (size_t)max(t, 0) * A_tot * SD;
        const float* hits_t = P.hits + (size_t)max(t, 0) * A_tot * R * PD;
        const int32_t* rs_t = P.row_start + (size_t)b * A_tot;
        const int32_t* rd_t = P.row_deg + (size_t)b * A_tot;
        const int32_t* er_t = P.edge_recv + (size_t)b * E * cap;
        const int32_t* es_t = P.edge_src + (size_t)b * E * cap;
        const bool stamp = P.prof != nullptr && blockIdx.x == 0 && tid == 0;
        unsigned long long* pr = P.prof + (size_t)(t + 1) * 8;
        if (stamp) pr[0] = gtime();
        if (t >= 0) {
            // ============================================================ phase E: edge tiles
            const int n_tiles = (M_cur + BM - 1) / BM;                  // tiles of my group's edge segment
            const int my_tiles = (n_tiles > lrank) ? (n_tiles - lrank + L - 1) / L : 0;
            const size_t seg_e0 = env_e0 + seg_off;                      // first slot of the segment
            if (warp == 0) {
                if (lane == 0) {
                    uint32_t it_l = it, ne_l = ne, n0_l = n0;
                    for (int tile = lrank; tile < n_tiles; tile += L, ++ne_l, ++n0_l) {
                        if (ne_l > 0) mbar_wait_wd(&bars[B_T2F], (ne_l - 1) & 1);
                        for (int kb = 0; kb < 8; ++kb, ++it_l) {
                            const int s = it_l % 3;
                            mbar_wait_wd(&bars[B_EMPTY + s], ((it_l / 3) & 1) ^ 1);
                            uint8_t* st = smem + s * STG;
                            mbar_expect_tx(&bars[B_FULL + s], 2 * B_BYTES);
                            tma_load_2d(st + 2 * A_BYTES, &tmW23h, &bars[B_FULL + s], kb * BK, 0);
                            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmW23l, &bars[B_FULL + s], kb * BK, 0);
                        }
                        mbar_wait_wd(&bars[B_TF0], n0_l & 1);        // main-loop MMAs retired: stage 2 is free
                        for (int kb2 = 0; kb2 < 4; ++kb2) {
                            const uint32_t j2 = ne_l * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                            mbar_wait_wd(&bars[B_B2E + slot], (use & 1) ^ 1);
                            uint8_t* sl = smem + 2 * STG + slot * 32768;
                            mbar_expect_tx(&bars[B_B2F + slot], 32768);
                            tma_load_2d(sl, &tmA1h, &bars[B_B2F + slot], kb2 * BK, 0);
                            tma_load_2d(sl + 16384, &tmA1l, &bars[B_B2F + slot], kb2 * BK, 0);
                        }
                    }
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    uint32_t it_l = it, ne_l = ne, n0_l = n0;
                    for (int tile = lrank; tile < n_tiles; tile += L, ++ne_l, ++n0_l) {
                        mbar_wait_wd(&bars[B_TE0], (n0_l & 1) ^ 1);
                        tc_fence_after();
                        for (int kb = 0; kb < 8; ++kb, ++it_l) {
                            const int s = it_l % 3;
                            const uint32_t ph = (it_l / 3) & 1;
                            mbar_wait_wd(&bars[B_CONV + s], ph);
                            mbar_wait_wd(&bars[B_FULL + s], ph);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + s * STG);
                            mma_kblock(tmem_base, a_hi, a_hi + A_BYTES, a_hi + 2 * A_BYTES, a_hi + 2 * A_BYTES + B_BYTES, kb == 0);
                            umma_commit(&bars[B_EMPTY + s]);
                        }
                        umma_commit(&bars[B_TF0]);
                        // chained gate GEMM: D2 = MSG tile (shared memory hi / lo planes) x A1, accumulator 1
                        mbar_wait_wd(&bars[B_TE1], (ne_l & 1) ^ 1);
                        mbar_wait_wd(&bars[B_A2R], ne_l & 1);
                        tc_fence_after();
                        for (int kb2 = 0; kb2 < 4; ++kb2) {
                            const uint32_t j2 = ne_l * 4 + kb2, slot = j2 & 1, use = j2 >> 1;
                            mbar_wait_wd(&bars[B_B2F + slot], use & 1);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + kb2 * 32768);
                            const uint32_t b_hi = smem_u32(smem + 2 * STG + slot * 32768);
                            mma_kblock(tmem_base + 128, a_hi, a_hi + 16384, b_hi, b_hi + 16384, kb2 == 0);
                            umma_commit(&bars[B_B2E + slot]);
                        }
                        umma_commit(&bars[B_T2F]);
                    }
                }
            } else if (warp < 10) {
                // worker warps 2..9 (256 threads): first PRODUCE the A operand of the tile (two threads per edge row, 16 of
                // the 32 columns of a k-block each: features + message layer 1, tf32 hi / lo split, swizzled K-major),
                // then drain it (message tile -> global + shared-memory hand-over, two warps per TMEM lane quarter with 64
                // columns each), then warps 2..5 turn the chained accumulator into the gate logits.
                const int pr_ = tid - 64;
                const int r = pr_ & 127;                  // producer: row of the tile
                const int half = pr_ >> 7;                // producer: 16-byte chunks 4 half .. 4 half + 3
                const int quarter = warp & 3;             // epilogue: TMEM lane quarter of this warp
                const int chalf = (warp - 2) >> 2;        // epilogue: message columns [64 chalf, 64 chalf + 64)
                const int row = quarter * 32 + lane;
                uint32_t it_l = it, ne_l = ne, n0_l = n0;
                for (int tile = lrank; tile < n_tiles; tile += L, ++ne_l, ++n0_l) {
                    {
                        const int ml = tile * BM + r;
                        const bool row_ok = ml < M_cur;
                        if (ne_l > 0) mbar_wait_wd(&bars[B_T2F], (ne_l - 1) & 1);
                        float f[ED];
                        int stype = 0;
#pragma unroll
                        for (int c = 0; c < ED; ++c) f[c] = 0.f;
                        if (row_ok) {
                            const int a = min(max(er_t[seg_e0 + ml], env_a0), env_a0 + N - 1);
                            const int code = min(es_t[seg_e0 + ml], env_a0 + N - 1);
                            float er[ED], es[ED], coef, nrm;
                            // agent states come from the CTA's own copy of the environment's states (the values the
                            // record holds; no dependence on another CTA's global stores)
                            edge_state_dev<KIND>(sst + (size_t)(a - env_a0) * SD, er);
                            if (code >= 0) edge_state_dev<KIND>(sst + (size_t)(max(code, env_a0) - env_a0) * SD, es);
                            else sender_state_dev<KIND>(code, a, R, agent_t, P.goal, hits_t, es);
                            edge_feat_dev<KIND>(er, es, code == -1, d.comm_radius, f, &coef, &nrm);
                            stype = (code >= 0) ? 2 : ((code == -1) ? 1 : 0);
                        }
                        for (int kb = 0; kb < 8; ++kb, ++it_l) {
                            const int s = it_l % 3;
                            mbar_wait_wd(&bars[B_EMPTY + s], ((it_l / 3) & 1) ^ 1);
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
                            fence_async_smem();
                            mbar_arrive(&bars[B_CONV + s]);
                        }
                    }
                    // ---- drain accumulator 0: message tile -> global + hi / lo hand-over planes
                    const int ml = tile * BM + row;
                    mbar_wait_wd(&bars[B_TF0], n0_l & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                        uint8_t* hi_row = smem + (c0 >> 5) * 32768 + row * 128;
                        uint8_t* lo_row = hi_row + 16384;
                        float* crow = P.msg + (seg_e0 + ml) * 128 + c0;
                        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                            const float4 bb = *reinterpret_cast<const float4*>(P.b23 + c0 + j);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                            if (ml < M_cur) {
                                if (j & 4) st_global_v8(crow + j - 4, prev, o);
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
                    tc_fence_before();
                    fence_async_smem();
                    mbar_arrive(&bars[B_A2R]);
                    mbar_arrive(&bars[B_TE0]);
                    if (chalf == 0) {
                        // ---- gate logit from the chained accumulator: one sequential fmaf chain over the 128 columns
                        // (the summation order of the 5-launch path), so it stays with the four quarter-owning warps
                        mbar_wait_wd(&bars[B_T2F], ne_l & 1);
                        tc_fence_after();
                        float dot = 0.f;
#pragma unroll 1
                        for (int c0 = 0; c0 < 128; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld32(tmem_base + 128 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                dot = fmaf(fmaxf(__uint_as_float(v[j]) + P.bias_g[c0 + j], 0.f), P.avec[c0 + j], dot);
                        }
                        if (ml < M_cur) P.logit[seg_e0 + ml] = dot + P.cst[0];
                        tc_fence_before();
                        mbar_arrive(&bars[B_TE1]);
                    }
                }
            }
            it += 8u * my_tiles;
            ne += my_tiles;
            n0 += my_tiles;
            SYNC_LOCAL();
            if (stamp) pr[1] = gtime();

            // ============================================================ phase A: segment softmax + aggregate
            for (int il = a_lo + warp; il < a_hi; il += PW) {
                const int a = env_a0 + il;
                const int rs = rs_t[a];
                int rd = rd_t[a];
                if (rs < 0 || rs + rd > cap) rd = 0;
                const float* ATT = P.logit + env_e0;
                const float* MSG = P.msg + env_e0 * 128;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rd <= 4) {
                    float lg[4];
                    float4 mv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = rs + min(q, max(rd - 1, 0));
                        lg[q] = (q < rd) ? ATT[e] : -INFINITY;
                        mv[q] = (q < rd) ? *reinterpret_cast<const float4*>(MSG + (size_t)e * 128 + lane * 4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    float mx = -INFINITY;
#pragma unroll
                    for (int q = 0; q < 4; ++q) mx = fmaxf(mx, lg[q]);
                    float den = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < rd) den += expf(lg[q] - mx);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < rd) {
                            const float att = expf(lg[q] - mx) / den;
                            acc.x = fmaf(att, mv[q].x, acc.x);
                            acc.y = fmaf(att, mv[q].y, acc.y);
                            acc.z = fmaf(att, mv[q].z, acc.z);
                            acc.w = fmaf(att, mv[q].w, acc.w);
                        }
                    }
                } else {
                    float mx = -INFINITY;
                    for (int e = rs; e < rs + rd; ++e) mx = fmaxf(mx, ATT[e]);
                    float den = 0.f;
                    for (int e = rs; e < rs + rd; ++e) den += expf(ATT[e] - mx);
                    for (int e = rs; e < rs + rd; ++e) {
                        const float att = expf(ATT[e] - mx) / den;
                        const float4 m = *reinterpret_cast<const float4*>(MSG + (size_t)e * 128 + lane * 4);
                        acc.x = fmaf(att, m.x, acc.x);
                        acc.y = fmaf(att, m.y, acc.y);
                        acc.z = fmaf(att, m.z, acc.z);
                        acc.w = fmaf(att, m.w, acc.w);
                    }
                }
                *reinterpret_cast<float4*>(P.ag + (size_t)a * 128 + lane * 4) = acc;
            }
            fence_async_global();          // generic-proxy writes of AG -> TMA (async proxy) reads in phase U1
            SYNC_LOCAL();
            if (stamp) pr[2] = gtime();

            // ============================================================ phases U1 / U2: agent-side GEMMs
            const int n_items = ((ga_hi - ga_lo + BM - 1) / BM) * 2;   // (agent tile of my group, 128-column half)
            const int my_items = (n_items > lrank) ? (n_items - lrank + L - 1) / L : 0;
            float* z_t = P.z + (size_t)(t & 1) * 2 * A_tot * 4;         // output partial sums, double-buffered by step parity
#pragma unroll 1
            for (int ph2 = 0; ph2 < 2; ++ph2) {
                const int nkb = ph2 == 0 ? 4 : 8;                      // K = 128 (update layer) / 256 (folded update/head)
                const CUtensorMap* tmA = ph2 == 0 ? &tmAG : &tmV1;
                const CUtensorMap* tmBh = ph2 == 0 ? &tmU1h : &tmUHh;
                const CUtensorMap* tmBl = ph2 == 0 ? &tmU1l : &tmUHl;
                if (warp == 0) {
                    if (lane == 0) {
                        uint32_t it_l = it;
                        fence_async_global();      // consumer side of the generic-store -> TMA-load hand-over (AG / V1)
                        for (int item = lrank; item < n_items; item += L) {
                            const int m0 = ga_lo + (item >> 1) * BM, nc0 = (item & 1) * 128;
                            for (int kb = 0; kb < nkb; ++kb, ++it_l) {
                                const int s = it_l % 3;
                                mbar_wait_wd(&bars[B_EMPTY + s], ((it_l / 3) & 1) ^ 1);
                                uint8_t* st = smem + s * STG;
                                mbar_expect_tx(&bars[B_FULL + s], A_BYTES + 2 * B_BYTES);
                                tma_load_2d(st, tmA, &bars[B_FULL + s], kb * BK, env_a0 + m0);
                                tma_load_2d(st + 2 * A_BYTES, tmBh, &bars[B_FULL + s], kb * BK, nc0);
                                tma_load_2d(st + 2 * A_BYTES + B_BYTES, tmBl, &bars[B_FULL + s], kb * BK, nc0);
                            }
                        }
                    }
                } else if (warp == 1) {
                    if (lane == 0) {
                        uint32_t it_l = it, n0_l = n0;
                        for (int item = lrank; item < n_items; item += L, ++n0_l) {
                            mbar_wait_wd(&bars[B_TE0], (n0_l & 1) ^ 1);
                            tc_fence_after();
                            for (int kb = 0; kb < nkb; ++kb, ++it_l) {
                                const int s = it_l % 3;
                                mbar_wait_wd(&bars[B_CONV + s], (it_l / 3) & 1);
                                tc_fence_after();
                                const uint32_t a_hi = smem_u32(smem + s * STG);
                                mma_kblock(tmem_base, a_hi, a_hi + A_BYTES, a_hi + 2 * A_BYTES, a_hi + 2 * A_BYTES + B_BYTES,
                                           kb == 0);
                                umma_commit(&bars[B_EMPTY + s]);
                            }
                            umma_commit(&bars[B_TF0]);
                        }
                    }
                } else if (warp < 10) {
                    // worker warps 2..9: split the TMA-landed A tile into tf32 hi / lo planes (256 threads, 4 float4 each),
                    // then drain the accumulator: U1 = 8 warps x 64 columns (independent elements); U2's output-layer
                    // partial sums are one sequential fmaf chain per row (the 5-launch order) -> the 4 quarter-owning warps
                    const int et = tid - 64;
                    const int quarter = warp & 3;
                    const int chalf = (warp - 2) >> 2;
                    const int row = quarter * 32 + lane;
                    uint32_t it_l = it, n0_l = n0;
                    for (int item = lrank; item < n_items; item += L, ++n0_l) {
                        for (int kb = 0; kb < nkb; ++kb, ++it_l) {
                            const int s = it_l % 3;
                            mbar_wait_wd(&bars[B_FULL + s], (it_l / 3) & 1);
                            float4* h4 = reinterpret_cast<float4*>(smem + s * STG);
                            float4* l4 = reinterpret_cast<float4*>(smem + s * STG + A_BYTES);
                            float4 v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = h4[et + 256 * j];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 h, l;
                                h.x = rn_tf32(v[j].x); h.y = rn_tf32(v[j].y); h.z = rn_tf32(v[j].z); h.w = rn_tf32(v[j].w);
                                l.x = rn_tf32(v[j].x - h.x); l.y = rn_tf32(v[j].y - h.y);
                                l.z = rn_tf32(v[j].z - h.z); l.w = rn_tf32(v[j].w - h.w);
                                h4[et + 256 * j] = h;
                                l4[et + 256 * j] = l;
                            }
                            fence_async_smem();
                            mbar_arrive(&bars[B_CONV + s]);
                        }
                        const int m0 = ga_lo + (item >> 1) * BM, nc0 = (item & 1) * 128;
                        const int ml = m0 + row;                   // agent inside the environment
                        mbar_wait_wd(&bars[B_TF0], n0_l & 1);
                        tc_fence_after();
                        if (ph2 == 0) {
#pragma unroll 1
                            for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 32) {
                                uint32_t v[32];
                                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                                if (ml < ga_hi) {
                                    const int n = nc0 + c0;
                                    float* crow = P.v1 + (size_t)(env_a0 + ml) * 256 + n;
#pragma unroll
                                    for (int j = 0; j < 32; j += 8) {
                                        float4 o[2];
#pragma unroll
                                        for (int hh = 0; hh < 2; ++hh) {
                                            o[hh] = make_float4(__uint_as_float(v[j + 4 * hh]), __uint_as_float(v[j + 4 * hh + 1]),
                                                                __uint_as_float(v[j + 4 * hh + 2]), __uint_as_float(v[j + 4 * hh + 3]));
                                            const float4 bb = *reinterpret_cast<const float4*>(P.b_u1 + n + j + 4 * hh);
                                            o[hh].x += bb.x; o[hh].y += bb.y; o[hh].z += bb.z; o[hh].w += bb.w;
                                            const float4 b2 = *reinterpret_cast<const float4*>(P.b_u1row + n + j + 4 * hh);
                                            o[hh].x += b2.x; o[hh].y += b2.y; o[hh].z += b2.z; o[hh].w += b2.w;
                                            o[hh].x = fmaxf(o[hh].x, 0.f); o[hh].y = fmaxf(o[hh].y, 0.f);
                                            o[hh].z = fmaxf(o[hh].z, 0.f); o[hh].w = fmaxf(o[hh].w, 0.f);
                                        }
                                        st_global_v8(crow + j, o[0], o[1]);
                                    }
                                }
                            }
                        } else if (chalf == 0) {
                            float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                            for (int c0 = 0; c0 < 128; c0 += 32) {
                                uint32_t v[32];
                                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                                const float* hw = P.ho + (size_t)(nc0 + c0) * NU;
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float h = fmaxf(__uint_as_float(v[j]) + P.buh[nc0 + c0 + j], 0.f);
#pragma unroll
                                    for (int q = 0; q < 4; ++q)
                                        if (q < NU) dq[q] = fmaf(h, hw[j * NU + q], dq[q]);
                                }
                            }
                            if (ml < ga_hi)
                                *reinterpret_cast<float4*>(z_t + ((size_t)(item & 1) * A_tot + env_a0 + ml) * 4) =
                                    make_float4(dq[0], dq[1], dq[2], dq[3]);
                        }
                        tc_fence_before();
                        mbar_arrive(&bars[B_TE0]);
                    }
                }
                it += (uint32_t)nkb * my_items;
                n0 += my_items;
                if (ph2 == 0) fence_async_global();   // V1 rows (generic stores) -> TMA reads of phase U2
                if (ph2 == 0) SYNC_LOCAL();
                else SYNC_ENV();                      // the policy tail needs the output sums of EVERY agent
                if (stamp) pr[3 + ph2] = gtime();
            }
        }

        // ================================================================ phase G: tail + graph of state t + 1
        {
            const int tn = t + 1;
            float* agent_n = P.agent + (size_t)tn * A_tot * SD;
            float* hits_n = P.hits + (size_t)tn * A_tot * R * PD;
            const int bn_ = tn & 1;
            int32_t* rs_n = P.row_start + (size_t)bn_ * A_tot;
            int32_t* rd_n = P.row_deg + (size_t)bn_ * A_tot;
            int32_t* er_n = P.edge_recv + (size_t)bn_ * E * cap;
            int32_t* es_n = P.edge_src + (size_t)bn_ * E * cap;
            if (t < 0) {
                for (int i = tid; i < N; i += PT) {
                    const float* a = agent_n + (size_t)(env_a0 + i) * SD;
#pragma unroll
                    for (int c = 0; c < SD; ++c) sst[i * SD + c] = a[c];
                }
            } else {
                // fused policy tail (geometry.cu graph_build_kernel): every CTA recomputes the next state of all N
                // agents into its position table; the cluster's first CTA records and reduces reward / cost
                const bool rec = rank == 0;
                float acc[3] = {0.f, 0.f, 0.f};
                float* act_t = P.actions + (size_t)t * A_tot * NU;
                for (int i = tid; i < N; i += PT) {
                    const size_t a = (size_t)env_a0 + i;
                    float zz[4] = {0.f, 0.f, 0.f, 0.f};
                    const float* z_t = P.z + (size_t)(t & 1) * 2 * A_tot * 4;
                    for (int p = 0; p < 2; ++p) {
                        const float4 v = *reinterpret_cast<const float4*>(z_t + ((size_t)p * A_tot + a) * 4);
                        zz[0] += v.x; zz[1] += v.y; zz[2] += v.z; zz[3] += v.w;
                    }
                    float x[SD], gl[SD], ur[NU], u[NU], xn[SD];
#pragma unroll
                    for (int c = 0; c < SD; ++c) {
                        x[c] = sst[i * SD + c];              // state t (== agent_t[a], this CTA's copy)
                        gl[c] = P.goal[a * SD + c];
                    }
                    u_ref_dev<KIND>(d, x, gl, ur);
                    float sq = 0.f;
#pragma unroll
                    for (int c = 0; c < NU; ++c) {
                        const float act = 2.f * tanhf(zz[c] + P.bho[c]) + ur[c];
                        if (rec) act_t[a * NU + c] = act;
                        u[c] = isnan(act) ? act : fminf(fmaxf(act, -d.u_lim), d.u_lim);
                        const float df = u[c] - ur[c];
                        sq = (c == 0) ? df * df : sq + df * df;
                    }
                    euler_dev<KIND>(d, x, gl, u, xn);
#pragma unroll
                    for (int c = 0; c < SD; ++c) sst[i * SD + c] = xn[c];
                    if (rec) {
#pragma unroll
                        for (int c = 0; c < SD; ++c) agent_n[a * SD + c] = xn[c];
                        const float nr = sqrtf(sq);
                        bool col = false;
                        const int rs = rs_t[a], rd = rd_t[a];
                        for (int e = rs + 1; e < rs + rd; ++e) {
                            const int sidx = es_t[env_e0 + e];
                            if (sidx < 0) break;
                            float dd = 0.f;
#pragma unroll
                            for (int c = 0; c < PD; ++c) {
                                const float dlt = x[c] - agent_t[(size_t)sidx * SD + c];
                                dd = (c == 0) ? dlt * dlt : dd + dlt * dlt;
                            }
                            col = col || (d.two_r > sqrtf(dd));
                        }
                        bool in_obs = false;
                        if (O > 0) {
                            const float* ob = P.obstacles + (d.obs_per_graph ? (size_t)env * O * OBW : 0);
                            in_obs = inside_any<PD>(ob, O, x, d.radius);
                        }
                        acc[0] += nr * nr;
                        acc[1] += col ? 1.f : 0.f;
                        acc[2] += in_obs ? 1.f : 0.f;
                    }
                }
                if (rec) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float v = warp_sum(acc[q]);
                        if (lane == 0) s_red[q * PW + warp] = v;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        float s3[3] = {0.f, 0.f, 0.f};
                        for (int w = 0; w < PW; ++w) {
                            s3[0] += s_red[0 * PW + w];
                            s3[1] += s_red[1 * PW + w];
                            s3[2] += s_red[2 * PW + w];
                        }
                        P.rewards[(size_t)t * E + env] = -(s3[0] / (float)N);
                        P.costs[(size_t)t * E + env] = s3[1] / (float)N + s3[2] / (float)N;
                    }
                }
            }
            __syncthreads();
            if (stamp) pr[5] = gtime();

            // ---- LiDAR + neighbour bits for this CTA's agents (warp per agent, PW agents per round)
            const int n_slots = a_hi - a_lo;
            for (int slot0 = 0; slot0 < APC; slot0 += PW) {
                const int slot = slot0 + warp;
                const int i = a_lo + slot;
                const bool valid = slot < n_slots;
                const int ii = valid ? i : 0;
                float p[PD];
#pragma unroll
                for (int c = 0; c < PD; ++c) p[c] = sst[ii * SD + c];
                const size_t a_glob = (size_t)env_a0 + ii;
                float* my_hits = hits_n + a_glob * R * PD;
                if (valid) {
                    const bool ray_ok = lane < d.n_rays;
                    const int rl = ray_ok ? lane : 0;
                    const float x1 = p[0], y1 = p[1];
                    const float x2 = x1 + stab[rl * 2 + 0], y2 = y1 + stab[rl * 2 + 1];
                    const float rdx = x1 - x2, rdy = y1 - y2;
                    float alpha;
                    if (O == 0) {
                        alpha = 1.f * NO_HIT;
                    } else {
                        alpha = NO_HIT;
                        bool is_in = false;
                        for (int o = 0; o < O; ++o) {
                            const float* ob = sobs + OBS2 * o;
                            const float cx = x1 - ob[0], cy = y1 - ob[1];
                            const bool far = (cx * cx + cy * cy) > ob[14];
                            bool degenerate = false;
                            if (far) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float det = rdx * ob[19 + k] - rdy * ob[15 + k];
                                    degenerate = degenerate || !(det != 0.f);
                                }
                            }
                            if (!far) is_in = is_in || rect_inside(ob, x1, y1, 0.f);
                            if (!far || __any_sync(0xffffffffu, degenerate))
                                alpha = nanmin(alpha, rect_raytrace(ob, x1, y1, x2, y2));
                        }
                        alpha = alpha * (1.f - (is_in ? 1.f : 0.f));
                    }
                    const float hx = x1 + (x2 - x1) * alpha;
                    const float hy = y1 + (y2 - y1) * alpha;
                    SortKey k;
                    k.flag = ray_ok ? (isnan(alpha) ? 1 : 0) : 2;
                    k.alpha = alpha;
                    k.idx = lane;
                    const bool all_miss = __all_sync(0xffffffffu, !ray_ok || alpha == NO_HIT);
                    if (!all_miss) k = warp_sort32(k, lane);
                    const float shx = __shfl_sync(0xffffffffu, hx, k.idx);
                    const float shy = __shfl_sync(0xffffffffu, hy, k.idx);
                    if (lane < R) {
                        my_hits[lane * 2 + 0] = shx;
                        my_hits[lane * 2 + 1] = shy;
                    }
                }
                __syncwarp();
                unsigned hit_bits = 0u;
                {
                    bool act = false;
                    if (valid && lane < R) {
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < PD; ++c) {
                            const float dlt = p[c] - my_hits[lane * PD + c];
                            acc = (c == 0) ? dlt * dlt : acc + dlt * dlt;
                        }
                        act = acc < d.lidar_sq_thr;
                    }
                    hit_bits = __ballot_sync(0xffffffffu, act);
                }
                int cnt = 0;
                if (valid) {
                    unsigned* my_bits = sbits + slot * n_words;
                    const int n_full = N >> 5;
#pragma unroll 4
                    for (int w = 0; w < n_full; ++w) {
                        const int j = (w << 5) + lane;
                        const float2 q = *reinterpret_cast<const float2*>(sst + j * SD);
                        const float dx = p[0] - q.x, dy = p[1] - q.y;
                        float acc = dx * dx;
                        acc = acc + dy * dy;
                        unsigned bits = __ballot_sync(0xffffffffu, acc < d.comm_sq_thr);
                        if (w == (i >> 5)) bits &= ~(1u << (i & 31));
                        if (lane == 0) my_bits[w] = bits;
                        cnt += __popc(bits);
                    }
                    if (N & 31) {
                        const int j = (n_full << 5) + lane;
                        bool ok = false;
                        if (j < N && j != i) {
                            float acc = 0.f;
#pragma unroll
                            for (int c = 0; c < PD; ++c) {
                                const float dlt = p[c] - sst[j * SD + c];
                                acc = (c == 0) ? dlt * dlt : acc + dlt * dlt;
                            }
                            ok = acc < d.comm_sq_thr;
                        }
                        const unsigned bits = __ballot_sync(0xffffffffu, ok);
                        if (lane == 0) my_bits[n_full] = bits;
                        cnt += __popc(bits);
                    }
                }
                if (lane == 0 && slot < APC) {
                    s_off[slot + 1] = valid ? (1 + cnt + __popc(hit_bits)) : 0;
                    s_hb[slot] = hit_bits;
                }
            }
            __syncthreads();
            if (warp == 0) {                     // exclusive prefix of the (<= 64) row degrees, agent order
                int v0 = (lane < APC) ? s_off[lane + 1] : 0;
                int v1 = (lane + 32 < APC) ? s_off[lane + 33] : 0;
                int inc0 = v0, inc1 = v1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n0_ = __shfl_up_sync(0xffffffffu, inc0, o);
                    const int n1_ = __shfl_up_sync(0xffffffffu, inc1, o);
                    if (lane >= o) { inc0 += n0_; inc1 += n1_; }
                }
                const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
                __syncwarp();
                if (lane == 0) s_off[0] = 0;
                if (lane < APC) s_off[lane + 1] = inc0;
                if (lane + 32 < APC) s_off[lane + 33] = tot0 + inc1;
                __syncwarp();
                const int total = s_off[APC];
                if (lane < L) {                  // my total -> slot [lrank] of every CTA of my hardware cluster (DSMEM)
                    uint32_t ra;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&s_tot[lrank])), "r"(lane));
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(total) : "memory");
                }
            }
            SYNC_LOCAL();
            if (stamp) pr[6] = gtime();
            int base = 0, env_total = 0;         // (rows before mine, rows of my local group's segment)
            for (int r2 = 0; r2 < L; ++r2) {
                const int v = s_tot[r2];
                base += (r2 < lrank) ? v : 0;
                env_total += v;
            }
            // ---- fill pass: rows [goal | agents ascending | active hits ascending], agent order inside the environment
            for (int slot = warp; slot < n_slots; slot += PW) {
                const int i = a_lo + slot;
                const int a_id = env_a0 + i;
                const int deg = s_off[slot + 1] - s_off[slot];
                const bool over = base + s_off[slot] + deg > seg_cap;
                const int rbase = seg_off + base + s_off[slot];       // row offset inside the environment's lists
                if (over) {
                    if (lane == 0) {
                        atomicOr(&P.counters[(size_t)tn * 4 + 1], 1);
                        rs_n[a_id] = 0;
                        rd_n[a_id] = 0;
                    }
                    continue;
                }
                int32_t* er = er_n + env_e0;
                int32_t* es = es_n + env_e0;
                if (lane == 0) {
                    rs_n[a_id] = rbase;
                    rd_n[a_id] = deg;
                    er[rbase] = a_id;
                    es[rbase] = -1;
                }
                int pos = rbase + 1;
                const unsigned lt = (1u << lane) - 1u;
                const unsigned* my_bits = sbits + slot * n_words;
                for (int w = 0; w < n_words; ++w) {
                    const unsigned bits = my_bits[w];
                    if (bits == 0u) continue;
                    if ((bits >> lane) & 1u) {
                        const int e = pos + __popc(bits & lt);
                        er[e] = a_id;
                        es[e] = env_a0 + (w << 5) + lane;
                    }
                    pos += __popc(bits);
                }
                const unsigned hb = s_hb[slot];
                if ((hb >> lane) & 1u) {
                    const int e = pos + __popc(hb & lt);
                    er[e] = a_id;
                    es[e] = -2 - lane;
                }
            }
            if (lrank == 0 && tid == 0) atomicAdd(&P.counters[(size_t)tn * 4 + 0], min(env_total, seg_cap));
            M_cur = min(env_total, seg_cap);
            SYNC_LOCAL();
            if (stamp) pr[7] = gtime();
        }
    }
#undef SYNC_LOCAL
#undef SYNC_ENV
    if (P.prof != nullptr && rank == 0 && tid == 0) P.prof[(size_t)(P.T + 1) * 8 + 2 * env + 1] = gtime();   // cluster end
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

struct WsLayout {
    int64_t msg, logit, ag, v1, z, row_start, row_deg, edge_recv, edge_src, gbar, gtot, total;
};
static WsLayout make_ws_layout(int E, int N, int cap_env) {   // (+ 16 (T + 1) floats of phase stamps appended by the caller)
    WsLayout W;
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += (n + 63) & ~(int64_t)63; return r; };
    const int64_t A = (int64_t)E * N, EC = (int64_t)E * cap_env;
    W.msg = take(EC * 128);
    W.logit = take(EC);
    W.ag = take(A * 128 + 128 * 128);      // + one tile of slack: the last environment's row tile may overhang
    W.v1 = take(A * 256 + 128 * 256);
    W.z = take(2 * 2 * A * 4);             // [step parity][column half][A][4]
    W.row_start = take(2 * A);
    W.row_deg = take(2 * A);
    W.edge_recv = take(2 * EC);
    W.edge_src = take(2 * EC);
    W.gbar = take(E);
    W.gtot = take(8 * (int64_t)E);
    W.total = o;
    return W;
}
static int cluster_size(int N, int cap_env) {
    const int items = ((N + 127) / 128) * 2;
    const int tiles = (min(cap_env, 3 * N) + 127) / 128;      // typical real edge count ~2 N
    int c = 1;
    while (c < 8 && c < max(items, tiles)) c <<= 1;
    return c;
}

}  // namespace rp
}  // namespace gcbf

using namespace gcbf;

extern "C" __attribute__((visibility("default"))) int64_t gcbf_rollout_persistent_workspace_floats(const gcbf_env_desc* desc) {
    if (!desc || desc->edge_cap <= 0 || desc->n_graphs <= 0 || desc->n_agents <= 0) return -1;
    return rp::make_ws_layout(desc->n_graphs, desc->n_agents, desc->edge_cap / desc->n_graphs).total + 64;
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_persistent_max_clusters(int32_t cluster_size);

// 0: unsupported configuration; 1: supported, but the device cannot keep one cluster per environment resident at the
// same time (B200: at most 15 clusters of 8 CTAs with this kernel's 212 KB of shared memory per CTA -- environments
// beyond that wait for a free cluster slot and the rollout takes two rounds); 2: supported and fully co-resident.
extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_persistent_supported(const gcbf_env_desc* desc) {
    if (!desc) return 0;
    const bool ok = desc->env_kind >= 0 && desc->env_kind <= 2 && desc->n_agents >= 1 && desc->n_agents <= rp::MAX_N &&
                    desc->n_obs <= rp::MAX_OBS && desc->n_rays <= 32 && desc->n_hits == desc->n_rays && desc->n_graphs >= 1 &&
                    desc->edge_cap / desc->n_graphs >= desc->n_agents && (desc->obs_per_graph == 1 || desc->n_obs == 0);
    if (!ok) return 0;
    const int C = rp::cluster_size(desc->n_agents, desc->edge_cap / desc->n_graphs);
    static int cached[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // occupancy per cluster size (0 = not queried yet)
    if (cached[C] == 0) {
        const int n = gcbf_rollout_persistent_max_clusters(C);
        cached[C] = n > 0 ? n : -1;
    }
    if (cached[C] > 0 && desc->n_graphs <= cached[C]) return 2;          // one hardware cluster per environment, all resident
    if (cached[2] == 0) {
        const int n = gcbf_rollout_persistent_max_clusters(2);
        cached[2] = n > 0 ? n : -1;
    }
    if (C >= 2 && desc->n_graphs * C <= sm_count() && desc->n_graphs * (C / 2) <= cached[2]) return 2;   // pair mode
    return 1;
}

static int persist_smem_bytes() {
    return 3 * rp::STG + 512 + 7 * 256 * 4 + rp::MAX_N * 4 * 4 + rp::MAX_OBS * 24 * 4 + 64 * 4 + 64 * 16 * 4 + 72 * 4 + 64 * 4 +
           3 * rp::PW * 4 + 1024;
}

/* Co-resident clusters of `cluster_size` CTAs of the persistent rollout kernel on the current device
 * (cudaOccupancyMaxActiveClusters): environments beyond this number wait for a free cluster slot. */
extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_persistent_max_clusters(int32_t cluster_size) {
    auto kern = rp::rollout_persist_kernel<GCBF_ENV_DOUBLE_INTEGRATOR>;
    const int smem = persist_smem_bytes();
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    if (cluster_size > 8 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(cluster_size * 64), 1, 1);
    cfg.blockDim = dim3(rp::PT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_persistent(
    const gcbf_env_desc* desc, int32_t n_steps, const float* actor_params, const float* infer_blob, const float* goal,
    const float* obstacles, const float* ray_table, float* agent_rec, float* hits_rec, float* actions_rec, float* rewards,
    float* costs, int32_t* counters, float* workspace, int64_t workspace_floats, uint64_t* phase_stamps, void* stream) {
    GCBF_REQUIRE(desc && actor_params && infer_blob && goal && ray_table && agent_rec && hits_rec && actions_rec && rewards &&
                     costs && counters && workspace, "gcbf_rollout_persistent: NULL pointer argument");
    GCBF_REQUIRE(gcbf_rollout_persistent_supported(desc) > 0, "gcbf_rollout_persistent: unsupported configuration (2-D envs, "
                 "n_agents <= 512, n_obs <= 32, edge_cap >= n_graphs * n_agents)");
    GCBF_REQUIRE(desc->n_obs == 0 || obstacles, "obstacles is NULL but n_obs > 0");
    GCBF_REQUIRE(n_steps >= 0, "n_steps must be >= 0");
    const int E = desc->n_graphs, N = desc->n_agents;
    const int cap_env = desc->edge_cap / E;
    const rp::WsLayout W = rp::make_ws_layout(E, N, cap_env);
    GCBF_REQUIRE(workspace_floats >= W.total, "workspace too small: %lld < %lld floats", (long long)workspace_floats,
                 (long long)W.total);
    GCBF_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)actor_params & 15) == 0 && ((uintptr_t)infer_blob & 15) == 0,
                 "workspace must be 256-byte aligned, parameters 16-byte aligned");
    const int ed = env_ed(desc->env_kind), nu = env_nu(desc->env_kind);
    const ParamLayout L = make_layout(ed, nu);
    const InferLayout I = make_infer_layout(nu);
    rp::PArgs P;
    memset(&P, 0, sizeof(P));
    P.d = *desc;
    P.T = n_steps;
    P.cap_env = cap_env;
    P.C = rp::cluster_size(N, cap_env);
    const int max_cl = gcbf_rollout_persistent_max_clusters(P.C);
    // mode 0: one hardware cluster per environment when all of them are resident at once; otherwise (B200: 16 environments
    // x 8 CTAs, 15 such clusters fit) mode 1: clusters of 2 + one software barrier per step, when the grid fits on the
    // device (1 CTA / SM).  GCBF_PERSIST_SOFT=1 forces mode 1 (tests).
    static const int force_soft = [] { const char* e = getenv("GCBF_PERSIST_SOFT"); return e ? atoi(e) : -1; }();
    P.soft = (force_soft >= 0) ? (force_soft != 0) : ((E <= max_cl) ? 0 : 1);
    GCBF_REQUIRE(!P.soft || (P.C >= 2 && E * P.C <= sm_count()), "pair mode needs n_graphs * %d <= %d CTAs", P.C, sm_count());
    GCBF_REQUIRE(P.soft || E <= max_cl || force_soft == 0, "more environments (%d) than resident clusters (%d)", E, max_cl);
    P.gbar = reinterpret_cast<unsigned*>(workspace + W.gbar);
    P.gtot = reinterpret_cast<int*>(workspace + W.gtot);
    P.W1 = actor_params + L.w[L_MSG0];
    P.b1 = actor_params + L.b[L_MSG0];
    P.b23 = infer_blob + I.b23;
    P.bias_g = actor_params + L.b[L_ATT0];
    P.avec = infer_blob + I.a23;
    P.cst = infer_blob + I.c23;
    P.b_u1 = actor_params + L.b[L_UPD0];
    P.b_u1row = actor_params + L.w[L_UPD0] + 2 * 256;
    P.buh = infer_blob + I.buh;
    P.ho = infer_blob + I.ho;
    P.bho = infer_blob + I.bho;
    P.goal = goal;
    P.obstacles = obstacles;
    P.ray_table = ray_table;
    P.agent = agent_rec;
    P.hits = hits_rec;
    P.actions = actions_rec;
    P.rewards = rewards;
    P.costs = costs;
    P.counters = counters;
    P.prof = reinterpret_cast<unsigned long long*>(phase_stamps);
    P.msg = workspace + W.msg;
    P.logit = workspace + W.logit;
    P.ag = workspace + W.ag;
    P.v1 = workspace + W.v1;
    P.z = workspace + W.z;
    P.row_start = reinterpret_cast<int32_t*>(workspace + W.row_start);
    P.row_deg = reinterpret_cast<int32_t*>(workspace + W.row_deg);
    P.edge_recv = reinterpret_cast<int32_t*>(workspace + W.edge_recv);
    P.edge_src = reinterpret_cast<int32_t*>(workspace + W.edge_src);
    CUtensorMap tW23h, tW23l, tA1h, tA1l, tU1h, tU1l, tUHh, tUHl, tAG, tV1;
    int32_t rc;
#define RC(x) do { if ((rc = (x))) return rc; } while (0)
    RC(tc::make_map(&tW23h, infer_blob + I.t_w23, 128, 256, 128));
    RC(tc::make_map(&tW23l, infer_blob + I.t_w23 + 256 * 128, 128, 256, 128));
    RC(tc::make_map(&tA1h, infer_blob + I.t_a1, 128, 128, 128));
    RC(tc::make_map(&tA1l, infer_blob + I.t_a1 + 128 * 128, 128, 128, 128));
    RC(tc::make_map(&tU1h, infer_blob + I.t_u1, 256, 128, 128));
    RC(tc::make_map(&tU1l, infer_blob + I.t_u1 + 256 * 128, 256, 128, 128));
    RC(tc::make_map(&tUHh, infer_blob + I.t_uh, 256, 256, 128));
    RC(tc::make_map(&tUHl, infer_blob + I.t_uh + 256 * 256, 256, 256, 128));
    RC(tc::make_map(&tAG, P.ag, E * N + 128, 128, 128));
    RC(tc::make_map(&tV1, P.v1, E * N + 128, 256, 128));
#undef RC
    const int smem = persist_smem_bytes();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(E * P.C), 1, 1);
    cfg.blockDim = dim3(rp::PT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)(P.soft ? 2 : P.C);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // pair mode spins on a counter the other CTAs of the environment must reach: every CTA has to be resident.  The grid
    // was checked against the SM count (1 CTA / SM) and the cluster-of-2 occupancy above; GCBF_PERSIST_COOP=1 additionally
    // asks the driver to guarantee it (cooperative launch attribute).  A protocol failure ends in a trap, not a hang.
    static const bool coop = [] { const char* e = getenv("GCBF_PERSIST_COOP"); return e && e[0] == '1'; }();
    if (P.soft && coop) {
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaSuccess;
    if (P.soft && (e = cudaMemsetAsync(P.gbar, 0, sizeof(unsigned) * E, (cudaStream_t)stream)) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
        return (int32_t)e;
    }
    switch (desc->env_kind) {
#define GCBF_RP_CASE(K)                                                                                               \
    case K: {                                                                                                         \
        auto kern = rp::rollout_persist_kernel<K>;                                                                    \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                            \
        if (e == cudaSuccess)                                                                                         \
            e = cudaLaunchKernelEx(&cfg, kern, P, tW23h, tW23l, tA1h, tA1l, tU1h, tU1l, tUHh, tUHl, tAG, tV1);        \
    } break;
        GCBF_RP_CASE(GCBF_ENV_SINGLE_INTEGRATOR)
        GCBF_RP_CASE(GCBF_ENV_DOUBLE_INTEGRATOR)
        GCBF_RP_CASE(GCBF_ENV_DUBINS_CAR)
#undef GCBF_RP_CASE
        default: set_error("gcbf_rollout_persistent: bad env_kind"); return -1;
    }
    if (e != cudaSuccess) {
        set_error("rollout_persist_kernel launch: %s", cudaGetErrorString(e));
        return (int32_t)e;
    }
    count_launch();
    return check_launch("rollout_persist_kernel");
}
