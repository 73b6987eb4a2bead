// ba_kernels_f32.cuh -- optional mixed-precision PCG (options.pcg_fp32): the two sweeps of the 50-step PCG loop in fp32.
//
// Everything that decides the LM trajectory stays fp64 (linearisation, U / V / gradients, preconditioner, the reduced
// right-hand side, the camera-side PCG vectors and dot products, back-substitution, cost, gain ratio).  Only the
// application of W (V + lambda I)^-1 W^t inside the PCG loop - 95 % of the iteration's work - runs on fp32 copies of the
// per-observation and per-point data, written next to the fp64 originals by the linearisation kernels:
//   o_rec32   point-order rows   [32 x int32 camera | 32 x f32 weight]                             256 B / row
//   c_rec32   camera-order rows  [camera, segment | 32 x int32 point | 32 x f32 weight | 3 x 32 x f32 R X]   656 B / row
//   X32, v32  float4 per point;  Vinv32  8 floats per point (6 + pad);  rt32  12 floats per camera
//   ptab32    20 floats per camera [quaternion | T, 0 | p0..p3 | p4, p5, 0, 0 | pad 4]: 80-byte stride (5 x 16 B: odd, so the
//             bank group of a field is again camera mod 8), 142 KB for 1778 cameras
// The segment partials leave the by-camera sweep as fp64 sums of fp32 lane accumulators, so the camera-side step kernel
// is the same in both modes.  North star: "1e-5 in an optional fp32 mode" - tests/test_gpu_parity.py::test_fp32_pcg_mode.
#pragma once
#include "ba_kernels.cuh"

namespace b200ba {

constexpr int OREC32_BYTES = 128 + 128;
constexpr int REC32_J      = 16;
constexpr int REC32_W      = REC32_J + 128;
constexpr int REC32_R0     = REC32_W + 128;
constexpr int REC32_R1     = REC32_R0 + 128;
constexpr int REC32_R2     = REC32_R1 + 128;
constexpr int REC32_BYTES  = REC32_R2 + 128;     // 656 = 41 x 16
constexpr int TAB32_STRIDE = 20;                 // floats per camera
#ifndef B200BA_TAB32_BLOCK
#define B200BA_TAB32_BLOCK 1024
#endif
constexpr int TAB32_BLOCK  = B200BA_TAB32_BLOCK;
constexpr int PR32         = 4;                  // ring rows per warp (2 stages x 2 rows)
constexpr int TAB32_RING_BYTES = (TAB32_BLOCK / 32) * (PR32 * OREC32_BYTES + 2 * 8);
constexpr int STAGE32_BYTES  = CG * REC32_BYTES;
constexpr int CAMPASS32_SMEM = CW * CNS * STAGE32_BYTES + CW * CNS * 8;

__device__ __forceinline__ float4 ldg_f4_keep(const float4* p, uint64_t pol)
{
    float4 v; asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void st_f4_keep(float4* p, float4 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// ---- fp32 copies written once per linearisation / damping change (cheap element-wise kernels) ----
__global__ void __launch_bounds__(256) k_f32_points(Dev d, int what)      // what 0: X32 (needs compute), 1: Vinv32
{
    const LMState* st = d.st;
    if (st->done || (what == 0 && !st->compute)) return;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= d.M) return;
    if (what == 0) { double X[3]; load3p(d.X[st->cur], j, X); d.X32[j] = make_float4((float)X[0], (float)X[1], (float)X[2], 0.f); }
    else {
        double V[6]; load6(d.Vinv, j, V);
        float4* o = reinterpret_cast<float4*>(d.Vinv32 + 8 * (size_t)j);
        o[0] = make_float4((float)V[0], (float)V[1], (float)V[2], (float)V[3]); o[1] = make_float4((float)V[4], (float)V[5], 0.f, 0.f);
    }
}
__global__ void __launch_bounds__(128) k_f32_cameras(Dev d)               // rt32 and the static part of ptab32 from ptab / rt
{
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= d.N) return;
    const double* s = d.rt[st->cur] + 12 * (size_t)i;
    for (int c = 0; c < 12; ++c) d.rt32[12 * (size_t)i + c] = (float)s[c];
    const double* t = d.ptab + (size_t)i * TAB_STRIDE;
    float* o = d.ptab32 + (size_t)i * TAB32_STRIDE;
    constexpr double us = 1.0 / QUAT_SCALE;                      // ptab holds sqrt(2) x the unit quaternion, the fp32 table the unit one
    o[0] = (float)(us * t[0]); o[1] = (float)(us * t[1]); o[2] = (float)(us * t[2]); o[3] = (float)(us * t[3]); o[4] = (float)t[4]; o[5] = (float)t[5]; o[6] = (float)t[6]; o[7] = 0.f;
    for (int c = 14; c < 20; ++c) o[c] = 0.f;
}
// weights of the point-order rows: one thread per padded entry
__global__ void __launch_bounds__(256) k_f32_orows(Dev d, int n_entries)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n_entries) return;
    reinterpret_cast<float*>(d.o_rec32 + (size_t)(k >> 5) * OREC32_BYTES + 128)[k & 31] = (float)*o_w_ptr(d, k);
}
// weights and cached R X of the camera-order rows: one warp per row
__global__ void __launch_bounds__(256) k_f32_crows(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= d.n_rows) return;
    const unsigned char* s = d.c_rec + (size_t)r * REC_BYTES;
    unsigned char* o = d.c_rec32 + (size_t)r * REC32_BYTES;
    reinterpret_cast<float*>(o + REC32_W)[lane]  = (float)reinterpret_cast<const double*>(s + REC_W)[lane];
    reinterpret_cast<float*>(o + REC32_R0)[lane] = (float)reinterpret_cast<const double*>(s + REC_R0)[lane];
    reinterpret_cast<float*>(o + REC32_R1)[lane] = (float)reinterpret_cast<const double*>(s + REC_R1)[lane];
    reinterpret_cast<float*>(o + REC32_R2)[lane] = (float)reinterpret_cast<const double*>(s + REC_R2)[lane];
}
// indices of both row layouts (set_problem, after the fp64 records are complete)
__global__ void __launch_bounds__(256) k_f32_indices(Dev d, int n_orows)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r < n_orows) reinterpret_cast<int*>(d.o_rec32 + (size_t)r * OREC32_BYTES)[lane] = reinterpret_cast<const int*>(d.o_rec + (size_t)r * OREC_BYTES)[lane];
    if (r < d.n_rows) {
        const unsigned char* s = d.c_rec + (size_t)r * REC_BYTES;
        unsigned char* o = d.c_rec32 + (size_t)r * REC32_BYTES;
        if (lane < 4) reinterpret_cast<int*>(o)[lane] = reinterpret_cast<const int*>(s)[lane];
        reinterpret_cast<int*>(o + REC32_J)[lane] = reinterpret_cast<const int*>(s + REC_J)[lane];
    }
}

// ---- by-point sweep, fp32 ----
struct TabRow32 { float h0, h1, h2, qw, qx, qy, qz; };
__device__ __forceinline__ float rcp_fast(float z) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z)); return r; }   // 1 ulp, no slow path
__device__ __forceinline__ void tab32_fwd(const float4* __restrict__ tab, int cam, float w, const float X[3], float f, float ar, TabRow32& o, bool pad)
{
    const float4* t = tab + 5 * cam;
    const float4 q = t[0], T = t[1], pa4 = t[2], pb4 = t[3];
    const float tx = q.z * X[2] - q.w * X[1], ty = q.w * X[0] - q.y * X[2], tz = q.y * X[1] - q.z * X[0];     // t = qv x X, q = (w, x, y, z)
    const float r0 = fmaf(2.f, q.x * tx + (q.z * tz - q.w * ty), X[0]);
    const float r1 = fmaf(2.f, q.x * ty + (q.w * tx - q.y * tz), X[1]);
    const float r2 = fmaf(2.f, q.x * tz + (q.y * ty - q.z * tx), X[2]);
    const float x = r0 + T.x, y = r1 + T.y, z = pad ? 1.f : r2 + T.z;
    const float iz = rcp_fast(z);
    const float wf = w * f * iz;
    const float pa = wf, pb = -wf * x * iz, pc = ar * wf, pd = -ar * wf * y * iz;
    const float s0 = pa4.x + (pb4.x * r2 - pb4.y * r1);          // p = (pa4.x pa4.y pa4.z | pa4.w pb4.x pb4.y)
    const float s1 = pa4.y + (pb4.y * r0 - pa4.w * r2);
    const float s2 = pa4.z + (pa4.w * r1 - pb4.x * r0);
    const float t0 = pa * s0 + pb * s2, t1 = pc * s1 + pd * s2;
    o.h0 = pa * t0; o.h1 = pc * t1; o.h2 = pb * t0 + pd * t1;
    o.qw = q.x; o.qx = q.y; o.qy = q.z; o.qz = q.w;
}
__device__ __forceinline__ void tab32_bwd(const TabRow32& o, float& u0, float& u1, float& u2)
{
    const float gx = o.qz * o.h1 - o.qy * o.h2, gy = o.qx * o.h2 - o.qz * o.h0, gz = o.qy * o.h0 - o.qx * o.h1;
    u0 += fmaf(2.f, o.qw * gx + (o.qz * gy - o.qy * gz), o.h0);
    u1 += fmaf(2.f, o.qw * gy + (o.qx * gz - o.qz * gx), o.h1);
    u2 += fmaf(2.f, o.qw * gz + (o.qy * gx - o.qx * gy), o.h2);
}
__global__ void __launch_bounds__(TAB32_BLOCK, 1) k_cg_point_pass_tab32(Dev d)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const float4* tab = reinterpret_cast<const float4*>(smem_raw);
    const uint32_t total = (uint32_t)d.N * TAB32_STRIDE * 4u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.ptab32);
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float f = (float)d.in.k00, ar = (float)d.in.ar;
    const int wid = blockIdx.x * (TAB32_BLOCK / 32) + warp;
    const bool active = wid < d.tab_warps32;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0;
    const unsigned char* src_rows = d.o_rec32;
    if (active) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr32) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr32) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; src_rows += (size_t)(ha.z >> 5) * OREC32_BYTES; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (PR32 * OREC32_BYTES);
    const uint32_t rbar = smem_u32(smem_raw) + tab_bytes + (uint32_t)(TAB32_BLOCK / 32) * (PR32 * OREC32_BYTES) + (uint32_t)warp * 16u;
    auto arm = [&](int stage) {                                  // lane 0 only
        const uint32_t bytes = (uint32_t)min(2, n_rows) * OREC32_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar + 8u * stage), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring + (uint32_t)stage * (2 * OREC32_BYTES)), "l"(src_rows), "r"(bytes), "r"(rbar + 8u * stage), "l"(policy_evict_first()) : "memory");
        src_rows += 2 * OREC32_BYTES; n_rows -= 2;
    };
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar + 8u), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        if (n_rows > 0) arm(0);
    }
    float X[3] = {0, 0, 0};
    if (g < g_end && 32 * g + lane < d.M) { const float4 x4 = ldg_f4_keep(d.X32 + 32 * g + lane, policy_evict_last()); X[0] = x4.x; X[1] = x4.y; X[2] = x4.z; }
    __syncthreads();
    mbar_wait(&bar, 0);
    if (g >= g_end) return;
    float u0 = 0, u1 = 0, u2 = 0;
    auto request = [&](float4& Va, float4& Vb, float4& Xn) {
        if (32 * g + lane < d.M) { const float4* vp = reinterpret_cast<const float4*>(d.Vinv32 + 8 * (size_t)(32 * g + lane)); Va = __ldg(vp); Vb = __ldg(vp + 1); }
        if (g + 1 < g_end && 32 * (g + 1) + lane < d.M) Xn = ldg_f4_keep(d.X32 + 32 * (g + 1) + lane, policy_evict_last());
    };
    auto finish = [&](const float4& Va, const float4& Vb, const float4& Xn) -> bool {
        const int j = 32 * g + lane;
        if (j < d.M)      // Vinv = (xx xy xz yy yz zz) = (Va.x Va.y Va.z Va.w Vb.x Vb.y)
            st_f4_keep(d.v32 + j, make_float4(Va.x * u0 + Va.y * u1 + Va.z * u2, Va.y * u0 + Va.w * u1 + Vb.x * u2, Va.z * u0 + Vb.x * u1 + Vb.y * u2, 0.f), policy_evict_last());
        if (++g >= g_end) return false;
        rows_left = deg_next; X[0] = Xn.x; X[1] = Xn.y; X[2] = Xn.z; u0 = u1 = u2 = 0;
        if (g + 1 < g_end) deg_next = __ldg(d.slice_deg + g + 1);
        return true;
    };
    int slot = 0; uint32_t par = 0;
    for (;;) {
        while (rows_left == 0) {
            float4 Va = make_float4(0, 0, 0, 0), Vb = Va, Xn = Va;
            request(Va, Vb, Xn);
            if (!finish(Va, Vb, Xn)) return;
        }
        if ((slot & 1) == 0) {
            if (lane == 0 && n_rows > 0) arm((slot >> 1) ^ 1);
            uint32_t ok;
            do {
                asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }"
                             : "=r"(ok) : "r"(rbar + 4u * (uint32_t)slot), "r"(par) : "memory");
            } while (!ok);
            __syncwarp();
        }
        const uint32_t ca = ring + (uint32_t)slot * OREC32_BYTES + 4u * lane;
        const int cam = lds_s32(ca);
        float w; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(ca + 128u));
        if (++slot == PR32) { slot = 0; par ^= 1u; }
        const bool last = (--rows_left == 0);
        TabRow32 o;
        const bool pad = cam < 0;                                // padding entries: camera 0, weight 0, depth 1 - no divergent branch
        tab32_fwd(tab, pad ? 0 : cam, pad ? 0.f : w, X, f, ar, o, pad);
        float4 Va = make_float4(0, 0, 0, 0), Vb = Va, Xn = Va;
        if (last) request(Va, Vb, Xn);
        tab32_bwd(o, u0, u1, u2);
        if (last && !finish(Va, Vb, Xn)) return;
    }
}

// ---- by-camera sweep, fp32 ----
struct Cam32 { float R[9], T[3]; };
__device__ __forceinline__ void cam32_one(const Cam32& cm, float f, float ar, bool ok, float r0, float r1, float r2, const float4& v, float w, float acc[6])
{
    const float x = r0 + cm.T[0], y = r1 + cm.T[1], z = ok ? r2 + cm.T[2] : 1.f;
    const float iz = rcp_fast(z);
    const float wf = w * f * iz;
    const float pa = wf, pb = -wf * x * iz, pc = ar * wf, pd = -ar * wf * y * iz;
    const float g0 = cm.R[0] * v.x + cm.R[1] * v.y + cm.R[2] * v.z;
    const float g1 = cm.R[3] * v.x + cm.R[4] * v.y + cm.R[5] * v.z;
    const float g2 = cm.R[6] * v.x + cm.R[7] * v.y + cm.R[8] * v.z;
    const float y0 = pa * g0 + pb * g2, y1 = pc * g1 + pd * g2;
    const float h0 = pa * y0, h1 = pc * y1, h2 = pb * y0 + pd * y1;
    acc[0] += h0; acc[1] += h1; acc[2] += h2;
    acc[3] += r1 * h2 - r2 * h1;
    acc[4] += r2 * h0 - r0 * h2;
    acc[5] += r0 * h1 - r1 * h0;
}
__global__ void __launch_bounds__(CW * 32, 1) k_cg_camera_pass32(Dev d)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vw = (blockIdx.x * CW + warp) * VSPLIT;
    if (vw >= d.n_vwarps) return;
    const int rb = vrow(d, vw), n = vrow(d, min(vw + VSPLIT, d.n_vwarps)) - rb;
    if (n <= 0) return;
    const int nst = (n + CG - 1) / CG;
    unsigned char* buf = smem_raw + (size_t)warp * CNS * STAGE32_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)CW * CNS * STAGE32_BYTES) + warp * CNS;
    const unsigned char* src = d.c_rec32 + (size_t)rb * REC32_BYTES;
    const uint64_t strm = policy_evict_first(), keep = policy_evict_last();
    auto arm = [&](int t) {
        const int sl = t % CNS;
        const uint32_t bytes = (uint32_t)min(CG, n - t * CG) * REC32_BYTES;
        mbar_expect_tx(&bar[sl], bytes);
        bulk_g2s_hint(buf + sl * STAGE32_BYTES, src + (size_t)t * STAGE32_BYTES, bytes, &bar[sl], strm);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < CNS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int t = 0; t < CNS && t < nst; ++t) arm(t);
    }
    __syncwarp();
    const float* rt = d.rt32;
    const int first_free = d.first_free;
    const float f = (float)d.in.k00, ar = (float)d.in.ar;
    int seg = -1, cam = 0;
    Cam32 cm;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    auto flush = [&]() {
        double a6[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) { a6[q] = (double)acc[q]; acc[q] = 0.f; }
        seg_flush<6>(d, seg, a6, lane);
    };
    auto load_cam32 = [&](int i) {
        const float4* s = reinterpret_cast<const float4*>(rt + 12 * (size_t)i);
        const float4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
        cm.R[0] = a.x; cm.R[1] = a.y; cm.R[2] = a.z; cm.T[0] = a.w; cm.R[3] = b.x; cm.R[4] = b.y; cm.R[5] = b.z; cm.T[1] = b.w;
        cm.R[6] = c.x; cm.R[7] = c.y; cm.R[8] = c.z; cm.T[2] = c.w;
    };
    auto issue = [&](int t, int (&jj)[CG], float4 (&vv)[CG]) {
        const int sl = t % CNS;
        mbar_wait(&bar[sl], (uint32_t)(t / CNS) & 1u);
#pragma unroll
        for (int g = 0; g < CG; ++g) {
            jj[g] = (t * CG + g < n) ? reinterpret_cast<const int*>(buf + sl * STAGE32_BYTES + g * REC32_BYTES + REC32_J)[lane] : -1;
            vv[g] = make_float4(0, 0, 0, 0);
            if (jj[g] >= 0) vv[g] = ldg_f4_keep(d.v32 + jj[g], keep);
        }
    };
    auto compute = [&](int t, const int (&jj)[CG], const float4 (&vv)[CG]) {
        const int sl = t % CNS;
        const unsigned char* sb = buf + sl * STAGE32_BYTES;
        int2 h[CG];
        bool same = (t * CG + CG <= n);
#pragma unroll
        for (int g = 0; g < CG; ++g) { h[g] = *reinterpret_cast<const int2*>(sb + g * REC32_BYTES); same = same && (h[g].y == seg); }
        if (same) {
            if (cam >= first_free) {
                float w[CG], r0[CG], r1[CG], r2[CG];
#pragma unroll
                for (int g = 0; g < CG; ++g) {
                    const unsigned char* rec = sb + g * REC32_BYTES;
                    w[g] = reinterpret_cast<const float*>(rec + REC32_W)[lane]; r0[g] = reinterpret_cast<const float*>(rec + REC32_R0)[lane];
                    r1[g] = reinterpret_cast<const float*>(rec + REC32_R1)[lane]; r2[g] = reinterpret_cast<const float*>(rec + REC32_R2)[lane];
                }
#pragma unroll
                for (int g = 0; g < CG; ++g) cam32_one(cm, f, ar, jj[g] >= 0, r0[g], r1[g], r2[g], vv[g], jj[g] >= 0 ? w[g] : 0.f, acc);
            }
        } else {
#pragma unroll
            for (int g = 0; g < CG; ++g) {
                if (t * CG + g < n) {
                    if (h[g].y != seg) {
                        if (seg >= 0) flush();
                        seg = h[g].y; cam = h[g].x; load_cam32(cam);
                    }
                    if (cam >= first_free) {
                        const unsigned char* rec = sb + g * REC32_BYTES;
                        cam32_one(cm, f, ar, jj[g] >= 0, reinterpret_cast<const float*>(rec + REC32_R0)[lane], reinterpret_cast<const float*>(rec + REC32_R1)[lane],
                                  reinterpret_cast<const float*>(rec + REC32_R2)[lane], vv[g], jj[g] >= 0 ? reinterpret_cast<const float*>(rec + REC32_W)[lane] : 0.f, acc);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && t + CNS < nst) { fence_proxy_async(); arm(t + CNS); }
    };
    int jA[CG], jB[CG]; float4 vA[CG], vB[CG];
    issue(0, jA, vA);
    for (int t = 0; t < nst; t += 2) {
        if (t + 1 < nst) issue(t + 1, jB, vB);
        compute(t, jA, vA);
        if (t + 1 < nst) {
            if (t + 2 < nst) issue(t + 2, jA, vA);
            compute(t + 1, jB, vB);
        }
    }
    if (seg >= 0) flush();
}

} // namespace b200ba
