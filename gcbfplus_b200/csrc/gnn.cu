// gnn.cu -- GNN forward for one network (CBF h(x) or policy pi(x)) over a swarm batch.
#include "gemm.cuh"
#include "gnn.cuh"

using namespace gcbf;

namespace gcbf {

int32_t gnn_forward_impl(const gcbf_env_desc* d, int out_dim, const float* P, const float* agent, const float* goal,
                         const float* hits, const int32_t* row_start, const int32_t* row_deg,
                         const int32_t* edge_recv, const int32_t* edge_src, const int32_t* counters, int clip_all,
                         float* out, float* ws, cudaStream_t st) {
    const int ed = env_ed(d->env_kind);
    const ParamLayout L = make_layout(ed, out_dim);
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
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.x1, P + L.w[L_MSG1], P + L.b[L_MSG1], nullptr, ws + W.x2, nullptr, re, 256, 256, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.x2, P + L.w[L_MSGOUT], P + L.b[L_MSGOUT], nullptr, ws + W.msg, nullptr, re, 256, 128, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS_RELU, false, ws + W.msg, P + L.w[L_ATT0], P + L.b[L_ATT0], nullptr, ws + W.g1, nullptr, re, 128, 128, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.g1, P + L.w[L_ATT1], P + L.b[L_ATT1], nullptr, ws + W.g2, nullptr, re, 128, 128, st))) return rc;
    // 6. attention softmax + aggregation
    {
        const int grid = min((A + 7) / 8, 4 * nsm);
        attn_aggregate_kernel<<<grid, 256, 0, st>>>(A, cap, ws + W.g2, ws + W.msg, P + L.w[L_GATE], P + L.b[L_GATE],
                                                    row_start, row_deg, ws + W.att, ws + W.ag);
        count_launch();
        if ((rc = check_launch("attn_aggregate_kernel"))) return rc;
    }
    // 7-11. update MLP (agent one-hot [0,0,1] folded into the bias: row 2 of update/Dense_0) + head MLP
    if ((rc = launch_gemm_nn(EPI_BIAS_RELU, false, ws + W.ag, P + L.w[L_UPD0] + 3 * 256, P + L.b[L_UPD0], P + L.w[L_UPD0] + 2 * 256, ws + W.v1, nullptr, ra, 128, 256, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.v1, P + L.w[L_UPD1], P + L.b[L_UPD1], nullptr, ws + W.v2, nullptr, ra, 256, 256, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.v2, P + L.w[L_UPDOUT], P + L.b[L_UPDOUT], nullptr, ws + W.v3, nullptr, ra, 256, 128, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS_RELU, false, ws + W.v3, P + L.w[L_HEAD0], P + L.b[L_HEAD0], nullptr, ws + W.h1, nullptr, ra, 128, 256, st))) return rc;
    if ((rc = launch_gemm_nn(EPI_BIAS, false, ws + W.h1, P + L.w[L_HEAD1], P + L.b[L_HEAD1], nullptr, ws + W.h2, nullptr, ra, 256, 256, st))) return rc;
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
                                    const float* agent, const float* goal, const float* hits,
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
    return gnn_forward_impl(desc, out_dim, params, agent, goal, hits, row_start, row_deg, edge_recv, edge_src, counters,
                            clip_all, out, workspace, (cudaStream_t)stream);
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
