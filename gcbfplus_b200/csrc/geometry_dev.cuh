// geometry_dev.cuh -- device functions shared by geometry.cu and rollout_persist.cu: obstacle primitives
// (Rectangle / Sphere inside + ray cast), the stable 32-key warp sort of the LiDAR returns, u_ref and the Euler step.
// Both translation units are compiled with -fmad=false: every arithmetic step keeps the reference's
// one-rounding-per-op semantics (bit-exact index sets / hit ordering / masks against the CPU oracle).
//
// Replaces (reference paths): gcbfplus/env/obstacle.py:53-96 (Rectangle), :234-270 (Sphere), env/utils.py:82-131,
// env/double_integrator.py:128-143 (Euler), :332-338 (u_ref) and their SingleIntegrator / DubinsCar / LinearDrone twins.
#pragma once
#include <math.h>

#include "common.cuh"

namespace gcbf {

static constexpr int GB_WARPS = 32;  // agents (warps) per CTA in graph_build
#define NO_HIT 1e6f

// ------------------------------------------------------------------------------------
// obstacle primitives
// ------------------------------------------------------------------------------------
// Rectangle.inside (obstacle.py:53-63); ob = 16-float packed rectangle.
__device__ __forceinline__ bool rect_inside(const float* ob, float px, float py, float r) {
    float rel_x = px - ob[0];
    float rel_y = py - ob[1];
    float rel_xx = fabsf(rel_x * ob[4] + rel_y * ob[5]) - ob[2];
    float rel_yy = fabsf(rel_x * ob[5] - rel_y * ob[4]) - ob[3];
    bool is_in_down = (rel_xx < r) && (rel_yy < 0.f);
    bool is_in_up = (rel_xx < 0.f) && (rel_yy < r);
    bool is_out_corner = (rel_xx > 0.f) && (rel_yy > 0.f);
    bool is_in_circle = sqrtf(rel_xx * rel_xx + rel_yy * rel_yy) < r;
    return (is_in_down || is_in_up) || (is_out_corner && is_in_circle);
}

__device__ __forceinline__ float nanmin(float a, float b) {  // jnp.min: NaN-propagating
    return (isnan(a) || isnan(b)) ? NAN : fminf(a, b);
}

// Rectangle.raytracing (obstacle.py:65-96): min over the 4 edges.
__device__ __forceinline__ float rect_raytrace(const float* ob, float x1, float y1, float x2, float y2) {
    float best = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int kp = (k + 3) & 3;  // points[[-1,0,1,2]]
        float x3 = ob[6 + 2 * k], y3 = ob[7 + 2 * k];
        float x4 = ob[6 + 2 * kp], y4 = ob[7 + 2 * kp];
        float det = (x1 - x2) * (y4 - y3) - (y1 - y2) * (x4 - x3);
        float sgn = (det > 0.f) ? 1.f : ((det < 0.f) ? -1.f : det);  // jnp.sign (0 -> 0, NaN -> NaN)
        det = sgn * fminf(fmaxf(fabsf(det), 1e-7f), 1e7f);
        float alpha = ((y4 - y3) * (x1 - x3) - (x4 - x3) * (y1 - y3)) / det;
        float beta = (-(y1 - y2) * (x1 - x3) + (x1 - x2) * (y1 - y3)) / det;
        float v = ((alpha <= 1.f && alpha >= 0.f) && (beta <= 1.f && beta >= 0.f)) ? 1.f : 0.f;
        alpha = v * alpha + (1.f - v) * NO_HIT;
        best = (k == 0) ? alpha : nanmin(best, alpha);
    }
    return best;
}

// Sphere.inside / raytracing (obstacle.py:234-270); ob = cx,cy,cz,r.
__device__ __forceinline__ bool sphere_inside(const float* ob, float px, float py, float pz, float r) {
    float dx = px - ob[0], dy = py - ob[1], dz = pz - ob[2];
    return sqrtf(dx * dx + dy * dy + dz * dz) <= ob[3] + r;
}
__device__ __forceinline__ float sphere_raytrace(const float* ob, float x1, float y1, float z1, float x2,
                                                 float y2, float z2) {
    float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    float rmax = sqrtf(dx * dx + dy * dy + dz * dz);
    float A = rmax * rmax;
    float ex = x1 - ob[0], ey = y1 - ob[1], ez = z1 - ob[2];
    float B = 2.f * (dx * ex + dy * ey + dz * ez);
    float C = ex * ex + ey * ey + ez * ez - ob[3] * ob[3];
    float delta = B * B - 4.f * A * C;
    // the line misses the sphere (the common case: 4 small spheres, 514 rays): valid1 = 0 makes both roots
    // (-B -+ 0) / (2A) * 0 + 1 = 1 exactly (finite operands), alphas = 1 and the result 0 * 1 + 1 * 1e6 = 1e6 -- returned
    // directly, skipping the square root and the two divisions (identical bits; NaN inputs fail `delta < 0` and take the
    // full path)
    if (delta < 0.f && A > 1e-20f && fabsf(B) < 1e18f) return NO_HIT;
    float valid1 = (delta >= 0.f) ? 1.f : 0.f;
    float sq = sqrtf(delta * valid1);
    float alpha1 = (-B - sq) / (2.f * A) * valid1 + (1.f - valid1);
    float alpha2 = (-B + sq) / (2.f * A) * valid1 + (1.f - valid1);
    float a1 = ((alpha1 >= 0.f) ? 1.f : 0.f) * alpha1 + ((alpha1 < 0.f) ? 1.f : 0.f) * 1.f;
    float a2 = ((alpha2 >= 0.f) ? 1.f : 0.f) * alpha2 + ((alpha2 < 0.f) ? 1.f : 0.f) * 1.f;
    float alphas = fminf(a1, a2);
    alphas = fminf(fmaxf(alphas, 0.f), 1.f);
    return valid1 * alphas + (1.f - valid1) * NO_HIT;
}

// inside_obstacles (env/utils.py:82-107) for one point against the graph's obstacle set.
template <int PD>
__device__ __forceinline__ bool inside_any(const float* sobs, int O, const float* p, float r) {
    bool in = false;
    if (PD == 2) {
        for (int o = 0; o < O; ++o) in = in || rect_inside(sobs + 16 * o, p[0], p[1], r);
    } else {
        for (int o = 0; o < O; ++o) in = in || sphere_inside(sobs + 4 * o, p[0], p[1], p[2], r);
    }
    return in;
}

// ------------------------------------------------------------------------------------
// stable ascending sort of 32 (alpha, idx) keys held one per lane (argsort, env/utils.py:127)
// key order: (flag, alpha, idx); flag 0 = number, 1 = NaN (sorts last), 2 = padding lane.
// ------------------------------------------------------------------------------------
struct SortKey {
    int flag;
    float alpha;
    int idx;
};
__device__ __forceinline__ bool key_less(const SortKey& a, const SortKey& b) {
    if (a.flag != b.flag) return a.flag < b.flag;
    if (a.flag == 0 && a.alpha != b.alpha) return a.alpha < b.alpha;
    return a.idx < b.idx;
}
__device__ __forceinline__ SortKey warp_sort32(SortKey k, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
            SortKey o;
            o.flag = __shfl_xor_sync(0xffffffffu, k.flag, j);
            o.alpha = __shfl_xor_sync(0xffffffffu, k.alpha, j);
            o.idx = __shfl_xor_sync(0xffffffffu, k.idx, j);
            const bool up = ((lane & size) == 0);
            const bool lower = ((lane & j) == 0);
            const bool keep_min = (up == lower);
            const bool o_less = key_less(o, k);
            if (keep_min ? o_less : !o_less) k = o;
        }
    }
    return k;
}

// ------------------------------------------------------------------------------------
// u_ref (double_integrator.py:332-338; dubins_car.py:328-379)
// ------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void u_ref_dev(const gcbf_env_desc& d, const float* x, const float* gl, float* u) {
    using T = EnvTraits<KIND>;
    constexpr int SD = T::SD, NU = T::NU;
    if (KIND == GCBF_ENV_DUBINS_CAR) {
        const float PI_F = 3.14159265358979323846f;
        const float TWO_PI = 6.283185307179586f;
        const float pdx = x[0] - gl[0], pdy = x[1] - gl[1];
        const float dist = sqrtf(pdx * pdx + pdy * pdy);
        float theta_t = atan2f(-pdy, -pdx);
        theta_t = theta_t - floorf(theta_t / TWO_PI) * TWO_PI;  // python-style mod
        float theta = x[2] - floorf(x[2] / TWO_PI) * TWO_PI;
        const float theta_diff = theta_t - theta;
        const float dot = (-pdx) * cosf(theta) + (-pdy) * sinf(theta);
        const float tb = acosf(fminf(fmaxf(dot / (dist + 0.0001f), -1.f), 1.f));
        float omega = 0.f;
        const bool c1 = (theta_diff < PI_F) && (theta_diff >= 0.f);
        if (c1 && theta <= PI_F) omega = 1.0f * tb;
        if (!c1 && theta <= PI_F) omega = -1.0f * tb;
        const bool c2 = (theta_diff > -PI_F) && (theta_diff <= 0.f);
        if (c2 && theta > PI_F) omega = -1.0f * tb;
        if (!c2 && theta > PI_F) omega = 1.0f * tb;
        omega = fminf(fmaxf(omega, -5.f), 5.f);
        const float nrm = sqrtf(1e-6f + (pdx * pdx + pdy * pdy));
        const float coef = (nrm > d.comm_radius) ? d.comm_radius / fmaxf(nrm, d.comm_radius) : 1.f;
        const float qx = coef * pdx, qy = coef * pdy;
        u[0] = omega;
        u[1] = -2.5f * x[3] + 2.3f * sqrtf(qx * qx + qy * qy);
        return;
    }
    float err[SD];
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < SD; ++c) {
        err[c] = gl[c] - x[c];
        acc = (c == 0) ? err[c] * err[c] : acc + err[c] * err[c];
    }
    const float nrm = sqrtf(acc);
#pragma unroll
    for (int c = 0; c < SD; ++c) {
        const float emax = fabsf(err[c] / nrm * d.comm_radius);
        // jnp.clip(x, lo, hi) = minimum(maximum(x, lo), hi), NaN-propagating
        float e = err[c];
        e = (isnan(e) || isnan(emax)) ? NAN : fminf(fmaxf(e, -emax), emax);
        err[c] = e;
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < SD; ++c) s = (c == 0) ? err[c] * d.K[a * SD + c] : s + err[c] * d.K[a * SD + c];
        u[a] = isnan(s) ? NAN : fminf(fmaxf(s, -d.u_lim), d.u_lim);
    }
}

// agent_step_euler (double_integrator.py:128-143; SI :104-109; Dubins :104-122; LD :123-134)
template <int KIND>
__device__ __forceinline__ void euler_dev(const gcbf_env_desc& d, const float* x, const float* gl, const float* u,
                                          float* xn) {
    using T = EnvTraits<KIND>;
    constexpr int SD = T::SD, NU = T::NU;
    float xd[SD];
    if (KIND == GCBF_ENV_SINGLE_INTEGRATOR) {
        xd[0] = u[0];
        xd[1] = u[1];
    } else if (KIND == GCBF_ENV_DOUBLE_INTEGRATOR) {
        xd[0] = x[2];
        xd[1] = x[3];
        xd[2] = u[0] / d.mass;
        xd[3] = u[1] / d.mass;
    } else if (KIND == GCBF_ENV_DUBINS_CAR) {
        const float ddx = x[0] - gl[0], ddy = x[1] - gl[1];
        const float stop = (sqrtf(ddx * ddx + ddy * ddy) < d.half_r) ? 1.f : 0.f;
        const float keep = 1.f - stop;
        xd[0] = (cosf(x[2]) * x[3]) * keep;
        xd[1] = (sinf(x[2]) * x[3]) * keep;
        xd[2] = (u[0] * 20.f) * keep;
        xd[3] = u[1] * keep;
    } else {
#pragma unroll
        for (int r = 0; r < SD; ++r) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < SD; ++c) s += x[c] * d.A[r * SD + c];
            float t = 0.f;
#pragma unroll
            for (int c = 0; c < NU; ++c) t += u[c] * d.B[r * NU + c];
            xd[r] = s + t;
        }
    }
#pragma unroll
    for (int c = 0; c < SD; ++c) {
        float v = xd[c] * d.dt + x[c];
        const bool limited = (KIND == GCBF_ENV_DOUBLE_INTEGRATOR && c >= 2) || (KIND == GCBF_ENV_DUBINS_CAR && c == 3) ||
                             (KIND == GCBF_ENV_LINEAR_DRONE && c >= 3);
        if (limited) v = isnan(v) ? v : fminf(fmaxf(v, -d.v_lim), d.v_lim);
        xn[c] = v;
    }
}

}  // namespace gcbf
