// gnn.cu -- GNN forward for one network (CBF h(x) or policy pi(x)) over a swarm batch.
#include <stdlib.h>

#include "gemm.cuh"
#include "gemm_tc.cuh"
#include "gnn.cuh"
#include "gemm_tc_prod.cuh"
#include "translayout.cuh"
#include "smalljobs.cuh"

using namespace gcbf;

namespace gcbf {

// Dense layer i of the network: Y = epi(X @ W_i + b_i).  With PT (transposed weights, see
// translayout.cuh) the tcgen05 tensor-core kernel is used, otherwise the strict-fp32 SIMT kernel.
static int32_t dense_fwd(int epi, const ParamLayout& L, const TransLayout& TL, int li, const float* P, const float* PT,
                         const float* X, float* Y, const float* bias2, RowCount rc, cudaStream_t st) {
    const int row_off = (li == L_UPD0) ? 3 * 256 : 0;   // update/Dense_0: rows 3..130 multiply the aggregated message
    const int K = (li == L_UPD0) ? 128 : L.in[li];
    if (PT) {   // PT = prepared parameters (PreparedLayout): tf32-split transposed weights
        const PreparedLayout Q = make_prepared_layout(L, TL);
        return tc::launch_gemm_tc(epi, false, X, PT + Q.pt_hi + TL.w[li], PT + Q.pt_lo + TL.w[li], P + L.b[li], bias2, Y,
                                  nullptr, rc, K, L.out[li], st);
    }
    return launch_gemm_nn(epi, false, X, P + L.w[li] + row_off, P + L.b[li], bias2, Y, nullptr, rc, K, L.out[li], st);
}

int32_t gnn_forward_impl(const gcbf_env_desc* d, int out_dim, const float* P, const float* PT, const float* agent, const float* goal,
                         const float* hits, const int32_t* row_start, const int32_t* row_deg,
                         const int32_t* edge_recv, const int32_t* edge_src, const int32_t* counters, int clip_all,
                         float* out, float* ws, cudaStream_t st) {
    const int ed = env_ed(d->env_kind);
    const ParamLayout L = make_layout(ed, out_dim);
    const TransLayout TL = make_trans_layout(L);
    const int A = d->n_graphs * d->n_agents;
    const int cap = d->edge_cap;
    const GnnWs W = make_ws(cap, A);
    const RowCount re{counters, 0, cap};
    const RowCount ra{nullptr, A, A};
    const int nsm = sm_count();
    int32_t rc;
    // 1. edge features + message layer 1
    {
        const int grid = min((cap + 7) / 8, 4 * nsm);
        GCBF_DISPATCH_ENV(d->env_kind, {
            edge_l1_kernel<KIND><<<grid, 256, 0, st>>>(*d, P + L.w[L_MSG0], P + L.b[L_MSG0], agent, goal, hits,
                                                       edge_recv, edge_src, counters, clip_all, ws + W.feat,
                                                       ws + W.x1);
        });
        count_launch();
        if ((rc = check_launch("edge_l1_kernel"))) return rc;
    }
    // 2-5. message MLP tail + gate MLP
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_MSG1, P, PT, ws + W.x1, ws + W.x2, nullptr, re, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_MSGOUT, P, PT, ws + W.x2, ws + W.msg, nullptr, re, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS_RELU, L, TL, L_ATT0, P, PT, ws + W.msg, ws + W.g1, nullptr, re, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_ATT1, P, PT, ws + W.g1, ws + W.g2, nullptr, re, st))) return rc;
    // 6. attention softmax + aggregation
    {
        const int grid = min((A + 7) / 8, 4 * nsm);
        attn_aggregate_kernel<<<grid, 256, 0, st>>>(A, cap, ws + W.g2, ws + W.msg, P + L.w[L_GATE], P + L.b[L_GATE],
                                                    row_start, row_deg, ws + W.att, ws + W.ag);
        count_launch();
        if ((rc = check_launch("attn_aggregate_kernel"))) return rc;
    }
    // 7-11. update MLP (agent one-hot [0,0,1] folded into the bias: row 2 of update/Dense_0) + head MLP
    if ((rc = dense_fwd(EPI_BIAS_RELU, L, TL, L_UPD0, P, PT, ws + W.ag, ws + W.v1, P + L.w[L_UPD0] + 2 * 256, ra, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_UPD1, P, PT, ws + W.v1, ws + W.v2, nullptr, ra, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_UPDOUT, P, PT, ws + W.v2, ws + W.v3, nullptr, ra, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS_RELU, L, TL, L_HEAD0, P, PT, ws + W.v3, ws + W.h1, nullptr, ra, st))) return rc;
    if ((rc = dense_fwd(EPI_BIAS, L, TL, L_HEAD1, P, PT, ws + W.h1, ws + W.h2, nullptr, ra, st))) return rc;
    // 12. output layer + tanh
    {
        const int grid = min((A + 7) / 8, 4 * nsm);
        head_out_kernel<<<grid, 256, 0, st>>>(A, out_dim, ws + W.h2, P + L.w[L_OUT], P + L.b[L_OUT], out);
        count_launch();
        if ((rc = check_launch("head_out_kernel"))) return rc;
    }
    return 0;
}

}  // namespace gcbf

extern "C" __attribute__((visibility("default"))) int64_t gcbf_gnn_workspace_floats(const gcbf_env_desc* desc, int32_t out_dim) {
    (void)out_dim;
    if (!desc || desc->edge_cap <= 0 || desc->n_graphs <= 0 || desc->n_agents <= 0) return -1;
    return make_ws(desc->edge_cap, (int64_t)desc->n_graphs * desc->n_agents).total;
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_gnn_forward(const gcbf_env_desc* desc, int32_t net_kind, int32_t out_dim, const float* params,
                                    const float* params_t, const float* agent, const float* goal, const float* hits,
                                    const int32_t* row_start, const int32_t* row_deg, const int32_t* edge_recv,
                                    const int32_t* edge_src, const int32_t* counters, int32_t clip_all, float* out,
                                    float* workspace, int64_t workspace_floats, void* stream) {
    GCBF_REQUIRE(desc && params && agent && goal && hits && row_start && row_deg && edge_recv && edge_src && counters &&
                     out && workspace, "gcbf_gnn_forward: NULL pointer argument");
    GCBF_REQUIRE(desc->env_kind >= 0 && desc->env_kind <= 3, "bad env_kind");
    GCBF_REQUIRE(net_kind == GCBF_NET_CBF || net_kind == GCBF_NET_ACTOR, "bad net_kind %d", net_kind);
    GCBF_REQUIRE(out_dim >= 1 && out_dim <= 4, "bad out_dim %d", out_dim);
    GCBF_REQUIRE(net_kind != GCBF_NET_CBF || out_dim == 1, "CBF net has out_dim 1");
    GCBF_REQUIRE(desc->edge_cap > 0, "edge_cap must be positive");
    const int64_t need = make_ws(desc->edge_cap, (int64_t)desc->n_graphs * desc->n_agents).total;
    GCBF_REQUIRE(workspace_floats >= need, "workspace too small: %lld < %lld floats", (long long)workspace_floats,
                 (long long)need);
    GCBF_REQUIRE((((uintptr_t)params | (uintptr_t)workspace) & 15) == 0, "params/workspace must be 16-byte aligned");
    GCBF_REQUIRE(params_t == nullptr || (((uintptr_t)params_t) & 15) == 0, "params_t must be 16-byte aligned");
    return gnn_forward_impl(desc, out_dim, params, params_t, agent, goal, hits, row_start, row_deg, edge_recv, edge_src,
                            counters, clip_all, out, workspace, (cudaStream_t)stream);
}

// ---- building blocks exported for unit tests and for bench.py's isolated kernel timing ----
extern "C" __attribute__((visibility("default"))) int32_t gcbf_gemm_nn(int32_t epi, int32_t accum, const float* A,
                                                                       const float* B, const float* bias,
                                                                       const float* bias2, float* C, const float* aux,
                                                                       const int32_t* m_ptr, int32_t m_fixed,
                                                                       int32_t m_cap, int32_t K, int32_t N,
                                                                       void* stream) {
    GCBF_REQUIRE(A && B && C, "gcbf_gemm_nn: NULL pointer");
    GCBF_REQUIRE((epi != EPI_BIAS && epi != EPI_BIAS_RELU) || bias, "gcbf_gemm_nn: bias required");
    GCBF_REQUIRE(epi != EPI_RELU_MASK || aux, "gcbf_gemm_nn: aux required");
    return launch_gemm_nn(epi, accum != 0, A, B, bias, bias2, C, aux, RowCount{m_ptr, m_fixed, m_cap}, K, N,
                          (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_gemm_tn(const float* X, int32_t ldx, const float* dY,
                                                                       float* C, const float* roww,
                                                                       const int32_t* row2agent, const int32_t* m_ptr,
                                                                       int32_t m_fixed, int32_t m_cap, int32_t K1,
                                                                       int32_t N, int32_t n_agents_total, void* stream) {
    GCBF_REQUIRE(X && dY && C, "gcbf_gemm_tn: NULL pointer");
    return launch_gemm_tn(X, ldx, dY, C, roww, row2agent, RowCount{m_ptr, m_fixed, m_cap}, K1, N, n_agents_total,
                          (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_colsum(const float* dY, float* db, const float* roww,
                                                                      const int32_t* row2agent, const int32_t* m_ptr,
                                                                      int32_t m_fixed, int32_t m_cap, int32_t N,
                                                                      int32_t n_agents_total, void* stream) {
    GCBF_REQUIRE(dY && db, "gcbf_colsum: NULL pointer");
    return launch_colsum(dY, db, roww, row2agent, RowCount{m_ptr, m_fixed, m_cap}, N, n_agents_total,
                         (cudaStream_t)stream);
}

// ---- tensor-core (tcgen05 3xTF32) variant of gcbf_gemm_nn: C = epi(A[M,K] @ Bt[N,K]^T) ----
extern "C" __attribute__((visibility("default"))) int32_t gcbf_gemm_tc(int32_t epi, int32_t accum, const float* A,
                                                                       const float* Bt_hi, const float* Bt_lo,
                                                                       const float* bias,
                                                                       const float* bias2, float* C, const float* aux,
                                                                       const int32_t* m_ptr, int32_t m_fixed,
                                                                       int32_t m_cap, int32_t K, int32_t N,
                                                                       void* stream) {
    GCBF_REQUIRE(A && Bt_hi && Bt_lo && C, "gcbf_gemm_tc: NULL pointer");
    GCBF_REQUIRE((epi != EPI_BIAS && epi != EPI_BIAS_RELU) || bias, "gcbf_gemm_tc: bias required");
    GCBF_REQUIRE(epi != EPI_RELU_MASK || aux, "gcbf_gemm_tc: aux required");
    GCBF_REQUIRE((((uintptr_t)A | (uintptr_t)Bt_hi | (uintptr_t)Bt_lo | (uintptr_t)C) & 15) == 0,
                 "gcbf_gemm_tc: 16-byte alignment required");
    return gcbf::tc::launch_gemm_tc(epi, accum != 0, A, Bt_hi, Bt_lo, bias, bias2, C, aux,
                                    RowCount{m_ptr, m_fixed, m_cap}, K, N, (cudaStream_t)stream);
}

// Transposed GEMM weights of one network (the K-major B operands of the tensor-core path).
extern "C" __attribute__((visibility("default"))) int32_t gcbf_params_t_count(int32_t edge_dim, int32_t out_dim) {
    if (edge_dim < 1 || edge_dim > 6 || out_dim < 1 || out_dim > 4) return -1;
    const ParamLayout L = make_layout(edge_dim, out_dim);
    return make_prepared_layout(L, make_trans_layout(L)).total;
}

namespace gcbf {
// Builds the prepared-parameter blob (PreparedLayout) of one network.
// One launch builds the whole prepared blob: the transposed tf32 planes of the 9 GEMM weights (32 x 32 tiles through
// shared memory, split on the way out) and the split planes of the untransposed parameters (the blocks after the tiles).
struct PrepJobs {
    int n;
    int src[12], dst[12], rows[12], cols[12], tile0[13];
};
static __global__ void __launch_bounds__(256)
prepare_kernel(const PrepJobs J, const float* __restrict__ P, float* __restrict__ pt_hi, float* __restrict__ pt_lo,
               float* __restrict__ p_hi, float* __restrict__ p_lo, const int n_params) {
    const int n_tiles = J.tile0[J.n];
    if ((int)blockIdx.x < n_tiles) {
        __shared__ float tile[32][33];
        int j = 0;
        while (j + 1 < J.n && (int)blockIdx.x >= J.tile0[j + 1]) ++j;
        const int t = blockIdx.x - J.tile0[j];
        const int rows = J.rows[j], cols = J.cols[j];
        const int tiles_c = (cols + 31) / 32;
        const int c0 = (t % tiles_c) * 32, r0 = (t / tiles_c) * 32;
        const float* in = P + J.src[j];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int i = ty; i < 32; i += 8) {
            const int r = r0 + i, c = c0 + tx;
            if (r < rows && c < cols) tile[i][tx] = in[(size_t)r * cols + c];
        }
        __syncthreads();
        for (int i = ty; i < 32; i += 8) {
            const int c = c0 + i, r = r0 + tx;
            if (r < rows && c < cols) {
                const float x = tile[tx][i];
                const float h = tc::rn_tf32(x);
                pt_hi[J.dst[j] + (size_t)c * rows + r] = h;
                pt_lo[J.dst[j] + (size_t)c * rows + r] = tc::rn_tf32(x - h);
            }
        }
        return;
    }
    const int nb = gridDim.x - n_tiles;
    for (int i = (blockIdx.x - n_tiles) * 256 + threadIdx.x; i < n_params; i += nb * 256) {
        const float x = P[i];
        const float h = tc::rn_tf32(x);
        p_hi[i] = h;
        p_lo[i] = tc::rn_tf32(x - h);
    }
}

int32_t build_prepared(const ParamLayout& L, const float* P, float* out, cudaStream_t st) {
    const TransLayout TL = make_trans_layout(L);
    const PreparedLayout Q = make_prepared_layout(L, TL);
    PrepJobs J;
    J.n = 0;
    int tiles = 0;
    for (int i = 0; i < 12; ++i) {
        if (TL.w[i] < 0) continue;
        const int rows = (i == L_UPD0) ? 128 : L.in[i];
        J.src[J.n] = L.w[i] + (i == L_UPD0 ? 3 * 256 : 0);
        J.dst[J.n] = TL.w[i];
        J.rows[J.n] = rows;
        J.cols[J.n] = L.out[i];
        J.tile0[J.n] = tiles;
        tiles += ((rows + 31) / 32) * ((L.out[i] + 31) / 32);
        ++J.n;
    }
    J.tile0[J.n] = tiles;
    const int split_blocks = min((L.total + 255) / 256, 2 * sm_count());
    prepare_kernel<<<tiles + split_blocks, 256, 0, st>>>(J, P, out + Q.pt_hi, out + Q.pt_lo, out + Q.p_hi, out + Q.p_lo,
                                                         L.total);
    count_launch();
    return check_launch("prepare_kernel");
}
}  // namespace gcbf

extern "C" __attribute__((visibility("default"))) int32_t gcbf_split_tf32(const float* in, float* hi, float* lo, int32_t n,
                                                                          void* stream) {
    GCBF_REQUIRE(in && hi && lo && n > 0, "gcbf_split_tf32: bad argument");
    gcbf::tc::split_tf32_kernel<<<min((n + 255) / 256, 4 * sm_count()), 256, 0, (cudaStream_t)stream>>>(in, hi, lo, n);
    count_launch();
    return check_launch("split_tf32_kernel");
}
extern "C" __attribute__((visibility("default"))) int32_t gcbf_prepare_params(int32_t edge_dim, int32_t out_dim,
                                                                              const float* params, float* params_t,
                                                                              void* stream) {
    GCBF_REQUIRE(edge_dim >= 1 && edge_dim <= 6 && out_dim >= 1 && out_dim <= 4 && params && params_t,
                 "gcbf_prepare_params: bad argument");
    return build_prepared(make_layout(edge_dim, out_dim), params, params_t, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_gemm_tn_tc(const float* X, int32_t ldx, const float* dY,
                                                                          float* C, const float* roww,
                                                                          const int32_t* row2agent,
                                                                          const int32_t* m_ptr, int32_t m_fixed,
                                                                          int32_t m_cap, int32_t K1, int32_t N,
                                                                          int32_t n_agents_total, void* stream) {
    GCBF_REQUIRE(X && dY && C, "gcbf_gemm_tn_tc: NULL pointer");
    GCBF_REQUIRE((((uintptr_t)X | (uintptr_t)dY | (uintptr_t)C) & 15) == 0, "gcbf_gemm_tn_tc: 16-byte alignment required");
    return gcbf::tc::launch_gemm_tn_tc(X, ldx, dY, C, roww, row2agent, RowCount{m_ptr, m_fixed, m_cap}, K1, N,
                                       n_agents_total, (cudaStream_t)stream);
}

// =====================================================================================================
// Inference path with folded weights.  Every MLP block of the GNN ends in two linear layers with no
// activation in between (act_final=False, gcbfplus/nn/mlp.py:23-29), so for rollouts
//   msg   = relu1 @ (W2 W3) + (b2 W3 + b3)                      256 -> 128
//   gate  = relu(msg A1 + ba1) . (A2 a3) + (ba2 . a3 + ba3)     128 -> 128 -> 1
//   h1    = relu(relu(aggr U1' + bu1') @ (U2 U3 H1) + ((bu2 U3 + bu3) H1 + bh1))   128 -> 256 -> 256
//   out   = tanh(h1 @ (H2 H3) + (bh2 H3 + bh3))                 256 -> nu
// 4 GEMMs + 3 small kernels per forward instead of 9 + 3, 2.4x fewer FLOPs.  The folded weights are
// rebuilt from the training parameters by gcbf_prepare_infer (once per parameter update).
// =====================================================================================================
namespace gcbf {

// Folded weights + operand planes: two dependency waves of small products, then all planes.  The jobs of several
// networks can share the three launches (prepare_infer_pair: both networks of the train step).
static void prepare_infer_jobs(int ed, int out_dim, const float* P, float* blob, SmallJobList& J1, SmallJobList& J2,
                               PlaneJobList& PJ) {
    const ParamLayout L = make_layout(ed, out_dim);
    const InferLayout I = make_infer_layout(out_dim);
    // message tail: W23 = W2 W3, b23 = b2 W3 + b3
    J1.add(blob + I.w23, 256, 128, 256, P + L.w[L_MSG1], 256, 1, P + L.w[L_MSGOUT], 128, 1, nullptr, nullptr, nullptr, false);
    J1.add(blob + I.b23, 1, 128, 256, P + L.b[L_MSG1], 0, 1, P + L.w[L_MSGOUT], 128, 1, nullptr, nullptr, P + L.b[L_MSGOUT], false);
    // gate tail: a23 = A2 a3 ; c = ba2 . a3 + ba3
    J1.add(blob + I.a23, 128, 1, 128, P + L.w[L_ATT1], 128, 1, P + L.w[L_GATE], 1, 1, nullptr, nullptr, nullptr, false);
    J1.add(blob + I.c23, 1, 1, 128, P + L.b[L_ATT1], 0, 1, P + L.w[L_GATE], 1, 1, nullptr, nullptr, P + L.b[L_GATE], false);
    // update tail: Q = U2 U3, b' = bu2 U3 + bu3
    J1.add(blob + I.q_u12, 256, 128, 256, P + L.w[L_UPD1], 256, 1, P + L.w[L_UPDOUT], 128, 1, nullptr, nullptr, nullptr, false);
    J1.add(blob + I.b_u12, 1, 128, 256, P + L.b[L_UPD1], 0, 1, P + L.w[L_UPDOUT], 128, 1, nullptr, nullptr, P + L.b[L_UPDOUT], false);
    // head tail: HO = H2 H3, bho = bh2 H3 + bh3
    J1.add(blob + I.ho, 256, out_dim, 256, P + L.w[L_HEAD1], 256, 1, P + L.w[L_OUT], out_dim, 1, nullptr, nullptr, nullptr, false);
    J1.add(blob + I.bho, 1, out_dim, 256, P + L.b[L_HEAD1], 0, 1, P + L.w[L_OUT], out_dim, 1, nullptr, nullptr, P + L.b[L_OUT], false);
    // update tail folded into the head's first layer: UH = Q H1, buh = b' H1 + bh1
    J2.add(blob + I.uh, 256, 256, 128, blob + I.q_u12, 128, 1, P + L.w[L_HEAD0], 256, 1, nullptr, nullptr, nullptr, false);
    J2.add(blob + I.buh, 1, 256, 128, blob + I.b_u12, 0, 1, P + L.w[L_HEAD0], 256, 1, nullptr, nullptr, P + L.b[L_HEAD0], false);
    // tf32 planes: transposed (forward B operands) and straight (backward-data B operands) of the 4 GEMM weights
    struct { const float* src; int rows, cols, t, p; } T[4] = {
        {blob + I.w23, 256, 128, I.t_w23, I.p_w23}, {P + L.w[L_ATT0], 128, 128, I.t_a1, I.p_a1},
        {P + L.w[L_UPD0] + 3 * 256, 128, 256, I.t_u1, I.p_u1}, {blob + I.uh, 256, 256, I.t_uh, I.p_uh}};
    for (int i = 0; i < 4; ++i) {
        const int n = T[i].rows * T[i].cols;
        PJ.add(T[i].src, T[i].rows, T[i].cols, true, blob + T[i].t, blob + T[i].t + n);
        PJ.add(T[i].src, T[i].rows, T[i].cols, false, blob + T[i].p, blob + T[i].p + n);
    }
}
static int32_t prepare_infer_launch(SmallJobList& J1, SmallJobList& J2, PlaneJobList& PJ, cudaStream_t st) {
    if (int32_t rc = J1.launch(st)) return rc;
    if (int32_t rc = J2.launch(st)) return rc;
    return PJ.launch(st);
}
int32_t prepare_infer_impl(int ed, int out_dim, const float* P, float* blob, cudaStream_t st) {
    SmallJobList J1, J2;
    PlaneJobList PJ;
    prepare_infer_jobs(ed, out_dim, P, blob, J1, J2, PJ);
    return prepare_infer_launch(J1, J2, PJ, st);
}
int32_t prepare_infer_pair(int ed, int out_a, const float* Pa, float* blob_a, int out_b, const float* Pb, float* blob_b,
                           cudaStream_t st) {
    SmallJobList J1, J2;
    PlaneJobList PJ;
    prepare_infer_jobs(ed, out_a, Pa, blob_a, J1, J2, PJ);
    prepare_infer_jobs(ed, out_b, Pb, blob_b, J1, J2, PJ);
    return prepare_infer_launch(J1, J2, PJ, st);
}

int32_t gnn_infer_impl(const gcbf_env_desc* d, int out_dim, const float* P, const float* blob, int use_tc,
                              const float* agent, const float* goal, const float* hits, const int32_t* row_start,
                              const int32_t* row_deg, const int32_t* edge_recv, const int32_t* edge_src,
                              const int32_t* counters, int clip_all, float* out, float* ws, cudaStream_t st,
                              float* z_out = nullptr, int* z_parts = nullptr, int32_t* zero_counter = nullptr,
                              int select = 0xF, int keep_activations = 0) {
    // keep_activations (folded train step): the unfused launch sequence, every layer output left in the workspace
    // (feat, x1, msg, g1, att, ag, v1, h1) for the backward pass; GEMMs still on the tensor-core path when use_tc
    // select (gcbf_rollout_step_select, measurement hook): bit 0 edge message (+ chained gate) kernel, bit 1 attention
    // aggregate, bit 2 update layer, bit 3 folded update/head layer; a cleared bit skips that launch
    // z_out != nullptr (rollout step): instead of `out`, write the output layer's pre-activation partial sums
    // z[part][A][4] (no bias, no tanh) for the policy tail fused into graph_build_kernel
    const int ed = env_ed(d->env_kind);
    const ParamLayout L = make_layout(ed, out_dim);
    const InferLayout I = make_infer_layout(out_dim);
    const int A = d->n_graphs * d->n_agents, cap = d->edge_cap;
    const GnnWs W = make_ws(cap, A);
    const RowCount re{counters, 0, cap};
    const RowCount ra{nullptr, A, A};
    const int nsm = sm_count();
    int32_t rc;
    auto gemm = [&](int epi, const float* X, const float* Wf, int t_off, int K, int N, const float* bias,
                    const float* bias2, float* Y, RowCount rows) -> int32_t {
        if (use_tc) return tc::launch_gemm_tc(epi, false, X, blob + t_off, blob + t_off + K * N, bias, bias2, Y, nullptr, rows, K, N, st);
        return launch_gemm_nn(epi, false, X, Wf, bias, bias2, Y, nullptr, rows, K, N, st);
    };
    if (use_tc && !keep_activations) {
        // tensor-core path, 4 launches: {edge features + layer 1 produced in-kernel -> folded message GEMM},
        // {gate layer + folded gate vector -> logits}, {softmax-aggregate produced in-kernel -> update layer 1},
        // {update/head folded layer (+ output layer) below}
        // GCBF_CHAIN=0: gate layer as its own GEMM launch (A/B measurements)
        static const bool chain_on = [] { const char* e = getenv("GCBF_CHAIN"); return !(e && e[0] == '0'); }();
        if (!(select & 1)) {
        } else if (chain_on) {
            tc::ChainArgs ch;
            ch.bias_g = P + L.b[L_ATT0];
            ch.avec = blob + I.a23;
            ch.cst = blob + I.c23;
            ch.logits = ws + W.att;
            if ((rc = tc::launch_edge_msg(d, P + L.w[L_MSG0], P + L.b[L_MSG0], agent, goal, hits, edge_recv, edge_src,
                                          counters, clip_all, blob + I.t_w23, blob + I.t_w23 + 256 * 128, blob + I.b23,
                                          ws + W.msg, st, blob + I.t_a1, blob + I.t_a1 + 128 * 128, &ch))) return rc;
        } else {
            if ((rc = tc::launch_edge_msg(d, P + L.w[L_MSG0], P + L.b[L_MSG0], agent, goal, hits, edge_recv, edge_src,
                                          counters, clip_all, blob + I.t_w23, blob + I.t_w23 + 256 * 128, blob + I.b23,
                                          ws + W.msg, st))) return rc;
            if ((rc = tc::launch_gemm_tc(EPI_RELU_DOT, false, ws + W.msg, blob + I.t_a1, blob + I.t_a1 + 128 * 128,
                                         P + L.b[L_ATT0], blob + I.c23, ws + W.att, blob + I.a23, re, 128, 128, st))) return rc;
        }
        // (measured: producing the aggregate inside the update GEMM (tc::launch_attn_upd) is slower than the
        //  separate warp-per-receiver kernel + TMA-fed GEMM: 36.6 us vs 13.2 + 11.7 us -- its N-split repeats
        //  the aggregation and the per-thread MSG gathers are latency-bound; the edge producer above is a win)
        if (select & 2) {
            const int grid = min((A + 7) / 8, 8 * nsm);   // 8 x 256 threads per SM: one receiver per warp in flight (latency-bound kernel)
            attn_aggregate_kernel<<<grid, 256, 0, st>>>(A, cap, nullptr, ws + W.msg, blob + I.a23, blob + I.c23, row_start,
                                                        row_deg, ws + W.att, ws + W.ag, zero_counter);
            count_launch();
            if ((rc = check_launch("attn_aggregate_kernel"))) return rc;
        }
        if ((select & 4) && (rc = gemm(EPI_BIAS_RELU, ws + W.ag, P + L.w[L_UPD0] + 3 * 256, I.t_u1, 128, 256, P + L.b[L_UPD0],
                                       P + L.w[L_UPD0] + 2 * 256, ws + W.v1, ra))) return rc;
    } else {
        {
            const int grid = min((cap + 7) / 8, 4 * nsm);
            GCBF_DISPATCH_ENV(d->env_kind, {
                edge_l1_kernel<KIND><<<grid, 256, 0, st>>>(*d, P + L.w[L_MSG0], P + L.b[L_MSG0], agent, goal, hits,
                                                           edge_recv, edge_src, counters, clip_all, ws + W.feat, ws + W.x1);
            });
            count_launch();
            if ((rc = check_launch("edge_l1_kernel"))) return rc;
        }
        if ((rc = gemm(EPI_BIAS, ws + W.x1, blob + I.w23, I.t_w23, 256, 128, blob + I.b23, nullptr, ws + W.msg, re))) return rc;
        if ((rc = gemm(EPI_BIAS_RELU, ws + W.msg, P + L.w[L_ATT0], I.t_a1, 128, 128, P + L.b[L_ATT0], nullptr, ws + W.g1, re))) return rc;
        {
            const int grid = min((A + 7) / 8, 8 * nsm);
            attn_aggregate_kernel<<<grid, 256, 0, st>>>(A, cap, ws + W.g1, ws + W.msg, blob + I.a23, blob + I.c23, row_start,
                                                        row_deg, ws + W.att, ws + W.ag, zero_counter);
            count_launch();
            if ((rc = check_launch("attn_aggregate_kernel"))) return rc;
        }
        if ((rc = gemm(EPI_BIAS_RELU, ws + W.ag, P + L.w[L_UPD0] + 3 * 256, I.t_u1, 128, 256, P + L.b[L_UPD0],
                       P + L.w[L_UPD0] + 2 * 256, ws + W.v1, ra))) return rc;
    }
    if (z_out != nullptr) {
        if (use_tc) {   // last hidden layer + output layer partial sums in the GEMM epilogue (h1 never leaves the SM)
            *z_parts = (2 * ((A + tc::BM - 1) / tc::BM) <= sm_count()) ? 2 : 1;
            if (!(select & 8)) return 0;
            return tc::launch_gemm_tc(EPI_RELU_DOTN, false, ws + W.v1, blob + I.t_uh, blob + I.t_uh + 256 * 256, blob + I.buh,
                                      nullptr, z_out, blob + I.ho, ra, 256, 256, st, out_dim, z_parts);
        }
        if ((rc = gemm(EPI_BIAS_RELU, ws + W.v1, blob + I.uh, I.t_uh, 256, 256, blob + I.buh, nullptr, ws + W.h1, ra))) return rc;
        const int grid = min((A + 7) / 8, 4 * nsm);
        head_z_kernel<<<grid, 256, 0, st>>>(A, out_dim, ws + W.h1, blob + I.ho, z_out);
        count_launch();
        *z_parts = 1;
        return check_launch("head_z_kernel");
    }
    if ((rc = gemm(EPI_BIAS_RELU, ws + W.v1, blob + I.uh, I.t_uh, 256, 256, blob + I.buh, nullptr, ws + W.h1, ra))) return rc;
    if (out != nullptr) {
        const int grid = min((A + 7) / 8, 4 * nsm);
        head_out_kernel<<<grid, 256, 0, st>>>(A, out_dim, ws + W.h1, blob + I.ho, blob + I.bho, out);
        count_launch();
        if ((rc = check_launch("head_out_kernel"))) return rc;
    }
    return 0;
}

}  // namespace gcbf

extern "C" __attribute__((visibility("default"))) int32_t gcbf_infer_count(int32_t edge_dim, int32_t out_dim) {
    if (edge_dim < 1 || edge_dim > 6 || out_dim < 1 || out_dim > 4) return -1;
    return make_infer_layout(out_dim).total;
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_prepare_infer(int32_t edge_dim, int32_t out_dim,
                                                                             const float* params, float* infer_blob,
                                                                             void* stream) {
    GCBF_REQUIRE(edge_dim >= 1 && edge_dim <= 6 && out_dim >= 1 && out_dim <= 4 && params && infer_blob,
                 "gcbf_prepare_infer: bad argument");
    GCBF_REQUIRE((((uintptr_t)params | (uintptr_t)infer_blob) & 15) == 0, "gcbf_prepare_infer: 16-byte alignment required");
    return prepare_infer_impl(edge_dim, out_dim, params, infer_blob, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_gnn_infer(
    const gcbf_env_desc* desc, int32_t net_kind, int32_t out_dim, const float* params, const float* infer_blob,
    int32_t use_tensor_cores, const float* agent, const float* goal, const float* hits, const int32_t* row_start,
    const int32_t* row_deg, const int32_t* edge_recv, const int32_t* edge_src, const int32_t* counters,
    int32_t clip_all, float* out, float* workspace, int64_t workspace_floats, void* stream) {
    GCBF_REQUIRE(desc && params && infer_blob && agent && goal && hits && row_start && row_deg && edge_recv && edge_src &&
                     counters && out && workspace, "gcbf_gnn_infer: NULL pointer argument");
    GCBF_REQUIRE(desc->env_kind >= 0 && desc->env_kind <= 3, "bad env_kind");
    GCBF_REQUIRE(net_kind == GCBF_NET_CBF || net_kind == GCBF_NET_ACTOR, "bad net_kind %d", net_kind);
    GCBF_REQUIRE(out_dim >= 1 && out_dim <= 4 && (net_kind != GCBF_NET_CBF || out_dim == 1), "bad out_dim %d", out_dim);
    GCBF_REQUIRE(desc->edge_cap > 0, "edge_cap must be positive");
    const int64_t need = make_ws(desc->edge_cap, (int64_t)desc->n_graphs * desc->n_agents).total;
    GCBF_REQUIRE(workspace_floats >= need, "workspace too small: %lld < %lld floats", (long long)workspace_floats,
                 (long long)need);
    GCBF_REQUIRE((((uintptr_t)params | (uintptr_t)workspace | (uintptr_t)infer_blob) & 15) == 0,
                 "params/infer_blob/workspace must be 16-byte aligned");
    return gnn_infer_impl(desc, out_dim, params, infer_blob, use_tensor_cores, agent, goal, hits, row_start, row_deg,
                          edge_recv, edge_src, counters, clip_all, out, workspace, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// One closed-loop rollout step in a single call: policy forward (folded weights) -> a = 2 pi + u_ref,
// clip, Euler, reward / cost terms -> LiDAR + neighbour lists of the next state (+ reward / cost reduction).
// 5 kernel launches (tensor-core path).
// ---------------------------------------------------------------------------------------------------
namespace gcbf {
int32_t graph_build_impl(const gcbf_env_desc* desc, const float* agent, const float* obstacles, const float* ray_table,
                         float* hits, int32_t* row_start, int32_t* row_deg, int32_t* edge_recv, int32_t* edge_src,
                         int32_t* counters, int32_t flags, const TailArgs& tail, float* reward, float* cost, void* stream);
}  // namespace gcbf

extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_step_select(
    const gcbf_env_desc* desc, const float* actor_params, const float* infer_blob, int32_t use_tensor_cores,
    const float* agent, const float* goal, const float* obstacles, const float* ray_table, const float* hits,
    const int32_t* row_start, const int32_t* row_deg, const int32_t* edge_recv, const int32_t* edge_src,
    const int32_t* counters, float* action, float* next_agent, float* next_hits, int32_t* next_row_start,
    int32_t* next_row_deg, int32_t* next_edge_recv, int32_t* next_edge_src, int32_t* next_counters, float* reward,
    float* cost, float* workspace, int64_t workspace_floats, int32_t select, void* stream) {
    GCBF_REQUIRE(desc && actor_params && infer_blob && agent && goal && ray_table && hits && row_start && row_deg &&
                     edge_recv && edge_src && counters && action && next_agent && next_hits && next_row_start &&
                     next_row_deg && next_edge_recv && next_edge_src && next_counters && reward && cost && workspace,
                 "gcbf_rollout_step: NULL pointer argument");
    GCBF_REQUIRE(next_row_start != row_start && next_row_deg != row_deg && next_edge_recv != edge_recv &&
                     next_edge_src != edge_src && next_counters != counters,
                 "gcbf_rollout_step: the next graph must not alias the current one (double-buffer the edge lists)");
    GCBF_REQUIRE(desc->env_kind >= 0 && desc->env_kind <= 3 && desc->edge_cap > 0, "gcbf_rollout_step: bad descriptor");
    GCBF_REQUIRE(desc->n_obs == 0 || obstacles, "obstacles is NULL but n_obs > 0");
    const int nu = env_nu(desc->env_kind);
    const int64_t A = (int64_t)desc->n_graphs * desc->n_agents;
    const GnnWs W = make_ws(desc->edge_cap, A);
    const int64_t need = W.total + 8 * A + 8;
    GCBF_REQUIRE(workspace_floats >= need, "workspace too small: %lld < %lld floats", (long long)workspace_floats,
                 (long long)need);
    GCBF_REQUIRE(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const InferLayout I = make_infer_layout(nu);
    float* z = workspace + ((W.total + 3) & ~(int64_t)3);    // [2][A][4] output-layer partial sums
    int32_t rc;
    int parts = 1;
    // 6 launches, no memset / copy node in between: {edge features + message layer}, {gate layer -> logits},
    // {segment softmax + aggregate; clears the next edge counter}, {update layer}, {folded update/head layer + output
    // layer partial sums}, {policy tail fused into the graph build of the next state}.
    // (Programmatic dependent launch of this chain was built and measured: +1 % (inside a CUDA graph the
    //  kernel-to-kernel gap is already ~1 us) and it was NOT safe as written -- a dependent kernel that starts early
    //  can keep L1 / read-only-cache lines of buffers its predecessor rewrites (DubinsCar rollouts became
    //  non-deterministic with only the edge-message GEMM launched that way).  Removed.)
    GCBF_REQUIRE(select == GCBF_STEP_ALL || use_tensor_cores, "gcbf_rollout_step_select: partial steps need the tensor-core path");
    if ((rc = gnn_infer_impl(desc, nu, actor_params, infer_blob, use_tensor_cores, agent, goal, hits, row_start, row_deg,
                             edge_recv, edge_src, counters, 0, nullptr, workspace, st, z, &parts, next_counters,
                             select & 0xF))) return rc;
    if (!(select & 16)) return 0;
    TailArgs tl;
    tl.z = z;
    tl.parts = parts;
    tl.z_cap = (int)A;
    tl.bHO = infer_blob + I.bho;
    tl.agent_prev = agent;
    tl.goal = goal;
    tl.row_start_prev = row_start;
    tl.row_deg_prev = row_deg;
    tl.edge_src_prev = edge_src;
    tl.action = action;
    tl.next_agent = next_agent;
    // the attention kernel cleared the next edge counter; a partial step without it lets the build clear it itself
    return graph_build_impl(desc, nullptr, obstacles, ray_table, next_hits, next_row_start, next_row_deg, next_edge_recv,
                            next_edge_src, next_counters, 1 | ((select & 2) ? 4 : 0), tl, reward, cost, stream);
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_rollout_step(
    const gcbf_env_desc* desc, const float* actor_params, const float* infer_blob, int32_t use_tensor_cores,
    const float* agent, const float* goal, const float* obstacles, const float* ray_table, const float* hits,
    const int32_t* row_start, const int32_t* row_deg, const int32_t* edge_recv, const int32_t* edge_src,
    const int32_t* counters, float* action, float* next_agent, float* next_hits, int32_t* next_row_start,
    int32_t* next_row_deg, int32_t* next_edge_recv, int32_t* next_edge_src, int32_t* next_counters, float* reward,
    float* cost, float* workspace, int64_t workspace_floats, void* stream) {
    return gcbf_rollout_step_select(desc, actor_params, infer_blob, use_tensor_cores, agent, goal, obstacles, ray_table,
                                    hits, row_start, row_deg, edge_recv, edge_src, counters, action, next_agent, next_hits,
                                    next_row_start, next_row_deg, next_edge_recv, next_edge_src, next_counters, reward, cost,
                                    workspace, workspace_floats, GCBF_STEP_ALL, stream);
}

extern "C" __attribute__((visibility("default"))) int64_t gcbf_rollout_workspace_floats(const gcbf_env_desc* desc) {
    if (!desc || desc->edge_cap <= 0 || desc->n_graphs <= 0 || desc->n_agents <= 0) return -1;
    const int64_t A = (int64_t)desc->n_graphs * desc->n_agents;
    return make_ws(desc->edge_cap, A).total + 8 * A + 16;
}
