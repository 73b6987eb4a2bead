// ba_dense.cuh -- explicit reduced camera system for tracker-sized problems (see dense_host.h for why and for the pair lists;
// DESIGN.md section 4b for the measurements behind every choice).
//
//   k_dense_wobs      thread per observation (SELL entry):  W = A~^t B~ (6x3),  Y = W (V + damping)^-1,  t = Y g_p (the
//                     observation's share of the reduced right-hand side); also evaluates and stores (V + damping)^-1
//   k_dense_pairs     36 threads per chunk of <= 32 pairs of one block:  partial  sum  Y_a W_b^t   (+ partial sum of t on the diagonal)
//   k_dense_finalize  thread per (block, entry): fixed-order sum of the chunk partials ->  S = [U + damping] - sum  (both
//                     triangles, exactly symmetric), the diagonal blocks' sums for the preconditioner (ar_sd), sum W Vinv g_p (ar_q)
//   k_dense_pcg<kind> the WHOLE block-Jacobi PCG loop in one launch.  CTA c of the group owns `rows` rows of S (whole cameras);
//                     per step: q = S p on the owned rows, p.q -> every CTA, alpha, x, r, z = M^-1 r on the owned rows, z and
//                     r.z -> every CTA, beta, p = z + beta p (every CTA, all rows).  Two exchanges per step.
//                       kinds 0 / 1: the group is one thread-block cluster (8 or 16 CTAs), S in registers / shared memory; the
//                         exchange is st.async into every peer's shared memory completing bytes on the peer's mbarrier;
//                       kind 2: the group is the whole cooperative grid (two cameras per CTA), S in registers; the exchange goes
//                         through global memory and a counting barrier - for camera counts beyond one cluster (up to 296).
//                     Same recurrences, stop rule and breakdown handling as k_cg_camera_step; all sums in fixed order.
#pragma once
#include "ba_kernels.cuh"
#include "dense_host.h"
#include <cooperative_groups.h>

namespace b200ba {

constexpr int DPAIR_BLOCK = 288;            // 8 chunks x 36 entries
constexpr int DENSE_OBS_STRIDE = 24;        // doubles per observation block: 6 rows x [3 values | pad]

// one thread per SELL entry (a thread per point would walk a 49-observation track serially: 31 us at the demo sequence's size)
__global__ void __launch_bounds__(256) k_dense_wobs(Dev d, int n_entries)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done) return;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n_entries) return;
    const int i = __ldg(o_cam_ptr(d, k));
    if (i < 0) return;                                            // SELL padding
    const int code = __ldg(d.dn.entry_pt + k);                     // point of the entry; one's complement marks the point's first slot
    const bool first = code < 0;
    const int j = first ? ~code : code;
    const int cur = st->cur;
    double X[3], g[3], V[6], Vi[6];
    load3p(d.X[cur], j, X); load3p(d.gp, j, g); load6(d.V, j, V);
    // (V + damping)^-1 here instead of a k_point_vinv launch: every entry of the point evaluates it, the first slot stores it
    const bool pba = st->variant == 1;
    const bool vok = point_vinv_eval(V, st->lambda, pba, Vi);
    if (first) {
        if (pba && st->first && st->compute) { double2* o = reinterpret_cast<double2*>(d.pt_D0 + 4 * (size_t)j); o[0] = make_double2(V[0], V[3]); o[1] = make_double2(V[5], 0.0); }
        double2* o = reinterpret_cast<double2*>(d.Vinv + 6 * (size_t)j);
        o[0] = make_double2(Vi[0], Vi[1]); o[1] = make_double2(Vi[2], Vi[3]); o[2] = make_double2(Vi[4], Vi[5]);
        if (!vok) d.st->solve_ok = 0;
    }
    const double vg0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2];
    const double vg1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2];
    const double vg2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
    Cam cm; load_cam(d.rt[cur], i, cm);
    Obs o; obs_jac(cm, X, __ldg(o_w_ptr(d, k)), d.in, o);
    double A0[6], A1[6], B0[3], B1[3];
    jac_rows_A(o, A0, A1); jac_rows_B(o, cm, B0, B1);
    const double live = i < d.first_free ? 0.0 : 1.0;           // constant camera: its W block is zero
    double* Wo = d.dn.Wobs + DENSE_OBS_STRIDE * (size_t)k;         // 6 rows of [3 values | pad]: one 32-byte sector per row
    double* Yo = d.dn.Yobs + DENSE_OBS_STRIDE * (size_t)k;
    double2* To = reinterpret_cast<double2*>(d.dn.Tobs + 6 * (size_t)k);
    double w[18], y[18], t[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const double w0 = live * (A0[a] * B0[0] + A1[a] * B1[0]), w1 = live * (A0[a] * B0[1] + A1[a] * B1[1]), w2 = live * (A0[a] * B0[2] + A1[a] * B1[2]);
        w[3 * a] = w0; w[3 * a + 1] = w1; w[3 * a + 2] = w2;
        y[3 * a]     = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
        y[3 * a + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
        y[3 * a + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
        t[a] = w0 * vg0 + w1 * vg1 + w2 * vg2;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(Wo + 4 * a), "d"(w[3 * a]), "d"(w[3 * a + 1]), "d"(w[3 * a + 2]), "d"(0.0) : "memory");
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(Yo + 4 * a), "d"(y[3 * a]), "d"(y[3 * a + 1]), "d"(y[3 * a + 2]), "d"(0.0) : "memory");
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) To[q] = make_double2(t[2 * q], t[2 * q + 1]);
}

// 36 threads per chunk, 8 chunks per block.  The chunk's pairs are staged in shared memory by one coalesced load; the operand
// loads of 8 pairs are issued together (L2 / L1 hits: independent loads hide the latency) and summed by a fixed tree, the groups
// of 8 in pair order.  (Measured alternatives at Trafalgar size, 552 K pairs: 16 pairs in flight 110 us, operand blocks staged
// through shared memory by coalesced loads 120 us, this form 81 us.)
__global__ void __launch_bounds__(DPAIR_BLOCK) k_dense_pairs(Dev d)
{
    pdl_chain_enter();
    constexpr int CPB = DPAIR_BLOCK / 36;
    __shared__ int2 spair[CPB][DENSE_CHUNK];
    __shared__ int4 schunk[CPB];
    if (d.st->done) return;
    if (threadIdx.x < CPB) { const int c = blockIdx.x * CPB + threadIdx.x; schunk[threadIdx.x] = c < d.dn.n_chunks ? __ldg(d.dn.chunks + c) : make_int4(0, 0, 0, 0); }
    __syncthreads();
    if (threadIdx.x < CPB * DENSE_CHUNK) {
        const int cq = threadIdx.x / DENSE_CHUNK, l = threadIdx.x % DENSE_CHUNK;
        const int4 ch = schunk[cq];
        if (ch.y + l < ch.z) spair[cq][l] = __ldg(d.dn.pairs + ch.y + l);
    }
    __syncthreads();
    const int cl = threadIdx.x / 36, e = threadIdx.x % 36, c = blockIdx.x * CPB + cl;
    if (c >= d.dn.n_chunks) return;
    const int4 ch = schunk[cl];                                   // block, first pair, end pair, diagonal flag
    const int a = e / 6, b = e - 6 * a, cnt = ch.z - ch.y;
    const bool diag = ch.w != 0;
    if (diag && b < a) return;                                    // diagonal blocks: upper triangle, mirrored by the finalising kernel
    const double* Y = d.dn.Yobs + 4 * a; const double* W = d.dn.Wobs + 4 * b;
    const bool want_rhs = diag && a == 0;
    double acc = 0, rhs = 0;
    for (int q0 = 0; q0 < cnt; q0 += 8) {
        double pr_[8], tt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool on = q0 + u < cnt;
            const int2 pr = on ? spair[cl][q0 + u] : make_int2(0, 0);
            double yv[3], wv[3];                                  // one 256-bit load per operand row (rows are 32-byte records)
            load3p_256(Y + DENSE_OBS_STRIDE * (size_t)pr.x, 0, yv); load3p_256(W + DENSE_OBS_STRIDE * (size_t)pr.y, 0, wv);
            pr_[u] = on ? (yv[0] * wv[0] + yv[1] * wv[1]) + yv[2] * wv[2] : 0.0;
            tt[u] = (want_rhs && on && pr.x == pr.y) ? d.dn.Tobs[6 * (size_t)pr.x + b] : 0.0;
        }
#pragma unroll
        for (int w = 1; w < 8; w <<= 1)
#pragma unroll
            for (int u = 0; u < 8; u += 2 * w) { pr_[u] += pr_[u + w]; tt[u] += tt[u + w]; }
        acc += pr_[0]; rhs += tt[0];
    }
    d.dn.chunk_part[36 * (size_t)c + e] = acc;
    if (want_rhs) d.dn.chunk_rhs[6 * (size_t)c + b] = rhs;
}

__global__ void __launch_bounds__(256) k_dense_finalize(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done) return;
    const int idx = blockIdx.x * 256 + threadIdx.x, blk = idx / 36, e = idx - 36 * blk;
    if (idx < 32) d.dn.bar_counter[32 * idx] = 0u;                // counting barrier (top + group counters) of the grid-wide PCG kernel that follows
    if (blk >= d.dn.n_blocks) return;
    const int i = __ldg(d.dn.blk_ik + 2 * blk), k = __ldg(d.dn.blk_ik + 2 * blk + 1);
    const int a = e / 6, b = e - 6 * a;
    const bool diag = i == k;
    if (diag && b < a) return;
    const int c0 = __ldg(d.dn.blk_chunk_beg + blk), c1 = __ldg(d.dn.blk_chunk_beg + blk + 1);
    double s = 0;
    for (int c = c0; c < c1; ++c) s += d.dn.chunk_part[36 * (size_t)c + e];
    const int n = 6 * d.N;
    double* S = d.dn.S;
    if (diag) {
        if (a == 0) {
            double rs = 0;
            for (int c = c0; c < c1; ++c) rs += d.dn.chunk_rhs[6 * (size_t)c + b];
            d.ar_q[6 * (size_t)i + b] = rs;
        }
        d.ar_sd[21 * (size_t)i + (6 * a - a * (a - 1) / 2 + (b - a))] = s;
        double u = d.U[36 * (size_t)i + 6 * a + b];
        if (a == b) { const double lam = st->lambda; u = st->variant == 1 ? u * (1.0 + lam) : u + lam; }
        const double v = u - s;
        S[(size_t)(6 * i + a) * n + 6 * i + b] = v;
        S[(size_t)(6 * i + b) * n + 6 * i + a] = v;
    } else {
        const double v = -s;
        S[(size_t)(6 * i + a) * n + 6 * k + b] = v;
        S[(size_t)(6 * k + b) * n + 6 * i + a] = v;
    }
}

// ---- the PCG loop ---------------------------------------------------------------------------------------
// shared-memory carve-up (doubles): [Ssm[rows4 x ld] (kind 1)] | p[ld] | z[ld] | q[4 rows4] | r[rows4] | pq_slot | rz_slot | red[64] incl. 2 mbarriers
struct DensePcgSmem {
    double *S, *p, *z, *q, *r, *pq, *rz, *red; uint64_t* bar;
    __device__ DensePcgSmem(double* base, int rows, int ld, bool s_in_smem, int slots)
    {
        const int rows4 = (rows + 3) & ~3;
        S = base; p = S + (s_in_smem ? (size_t)rows4 * ld : 0); z = p + ld; q = z + ld; r = q + 4 * rows4; pq = r + rows4; rz = pq + slots; red = rz + slots;
        bar = reinterpret_cast<uint64_t*>(red + 60);
    }
};

// ---- exchange inside a thread-block cluster: st.async into a peer's shared memory, completing bytes on the PEER's mbarrier ----
// A value and the signal that it has arrived travel in one DSMEM transaction; the receiver waits on its OWN mbarrier for the
// byte count of the phase.  No cluster-wide barrier, no fence: measured 330 cycles from the sends to the completed wait.
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_smem_addr, int cta)
{
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- exchange across the whole (cooperative, co-resident) grid: global memory + a counting barrier ----
// One atomic per CTA and phase on a counter that only grows (zeroed by k_dense_finalize); the waiters poll that ONE word with
// acquire loads, then warp 0 sums the slots.  Measured alternatives at 257 cameras / 148 CTAs: cooperative_groups' grid.sync()
// 12.8 us per PCG step; this form 6.7 us; slots tagged with a "not written yet" sentinel and polled directly (barrier and data in
// one round trip on paper) 9.8 us - 148 CTAs polling 148 words keep the L2 lines bouncing while the writers try to update them.
#ifndef B200BA_GRID_BAR_GROUP
#define B200BA_GRID_BAR_GROUP 0
#endif
constexpr int GRID_BAR_GROUP = B200BA_GRID_BAR_GROUP;            // > 0: two-level arrival (groups of that many CTAs), see below
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int phase, int G, int rank)
{
    __syncthreads();                                               // the CTA's global stores of this phase are issued
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int target;
        if (GRID_BAR_GROUP > 0) {
            // two levels: the last arriver of a group of CTAs (its own counter, 128 bytes apart) arrives at the top counter
            const int g = rank / GRID_BAR_GROUP, ng = (G + GRID_BAR_GROUP - 1) / GRID_BAR_GROUP;
            const unsigned int size_g = (unsigned int)min(GRID_BAR_GROUP, G - g * GRID_BAR_GROUP);
            const unsigned int old = atomicAdd(counter + 32 * (1 + g), 1u);
            if (old + 1u == phase * size_g) { __threadfence(); atomicAdd(counter, 1u); }
            target = phase * (unsigned int)ng;
        } else {
            atomicAdd(counter, 1u);
            target = phase * (unsigned int)G;
        }
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    __syncthreads();
}
__device__ __forceinline__ double grid_slot_sum(const double* slots, int count, double* red)
{
    if (threadIdx.x < 32) {
        double s = 0;
        for (int c = threadIdx.x; c < count; c += 32) s += __ldcg(slots + c);
        s = warp_sum(s);
        if (threadIdx.x == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// sum of `count` values in a FIXED order (pairwise tree over strided partials: fp64 adds have ~30 cycles of latency, a serial
// chain over 8 slots cost 280 cycles); identical in every thread of every CTA
__device__ __forceinline__ double fixed_sum_smem(const double* v, int count)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int c = 0;
    for (; c + 4 <= count; c += 4) { s0 += v[c]; s1 += v[c + 1]; s2 += v[c + 2]; s3 += v[c + 3]; }
    if (c < count) s0 += v[c];
    if (c + 1 < count) s1 += v[c + 1];
    if (c + 2 < count) s2 += v[c + 2];
    return (s0 + s1) + (s2 + s3);
}
// Work split of q = S p inside a CTA: warp (rb, cq) owns the 4 rows 4 rb .. 4 rb + 3 and every CQ-th group of 32 column pairs;
// lane l of it owns the column pairs l + 32 (cq + CQ t), t < NP.  Every p value a lane reads is used for 4 rows (the first
// version gave a row to 8 lanes: 40 p values per lane, and the sweep sat on shared-memory bandwidth: 123 KB of reads per step).
// the same for a compile-time count: all loads first, then a pairwise tree (depth log2 N instead of N / 4 + 2 chained adds)
template <int N>
__device__ __forceinline__ double fixed_sum_tree(const double* v)
{
    constexpr int P2 = N <= 8 ? 8 : 16;
    static_assert(N <= 16, "slot count");
    double a[P2];
#pragma unroll
    for (int u = 0; u < P2; ++u) a[u] = u < N ? v[u] : 0.0;
#pragma unroll
    for (int w = 1; w < P2; w <<= 1)
#pragma unroll
        for (int u = 0; u < P2; u += 2 * w) a[u] += a[u + w];
    return a[0];
}
__device__ __forceinline__ double fixed_sum_slots(const double* v, int count)      // cluster of 8 or 16 CTAs
{
    return count == 8 ? fixed_sum_tree<8>(v) : count == 16 ? fixed_sum_tree<16>(v) : fixed_sum_smem(v, count);
}

template <int KIND> struct DenseCfg {
    static constexpr bool GRID = KIND == 2;
    static constexpr int NT = DENSE_THREADS[KIND], CQ = DENSE_CQ[KIND], NP = DENSE_NP[KIND];
    static constexpr bool SREG = NP > 0;                           // S in registers (4 rows x 2 NP values per lane)
    static constexpr int NP_ = NP > 0 ? NP : 1;
};

template <int KIND>
__global__ void __launch_bounds__(DenseCfg<KIND>::NT, 1) k_dense_pcg(Dev d, int max_steps)
{
    pdl_chain_enter();
    namespace cg = cooperative_groups;
    using Cfg = DenseCfg<KIND>;
    constexpr bool GRID = Cfg::GRID;
    constexpr int NT = Cfg::NT, CQ = Cfg::CQ, NP = Cfg::NP, NWARP = NT / 32;
    constexpr bool SREG = Cfg::SREG;
    extern __shared__ __align__(16) double dsm[];
    LMState* st = d.st;
    if (st->done) return;                                          // uniform over the grid
    cg::cluster_group cl = cg::this_cluster();
    const int G = GRID ? (int)gridDim.x : (int)cl.num_blocks();
    const int rank = GRID ? (int)blockIdx.x : (int)cl.block_rank();
    const int tid = threadIdx.x, n = 6 * d.N, rows = d.dn.rows, ld = d.dn.ld;
    DensePcgSmem sm(dsm, rows, ld, !SREG, GRID ? 256 : 32);
    const int row0 = rank * rows, nrows = max(0, min(rows, n - row0));
    if (!st->solve_ok) {                                           // (V + damping) was not positive definite: the step is rejected
        if (rank == 0 && tid == 0) { st->cg_active = 0; st->cg_steps = 0; st->cg_rel = 0; }
        return;
    }
    uint32_t par_a = 0, par_b = 0; unsigned int gbar = 0;          // mbarrier phase parities / grid-barrier phase count
    if (!GRID) {
        if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); }
        cl.sync();                                                 // every CTA's mbarriers exist before the first st.async
    }
    // ---- prologue: the CTA's rows of S -> registers (kinds 0, 2) or shared memory (kind 1); r = rhs, x = 0, z = M^-1 r ----
    const int warp = tid >> 5, lane = tid & 31, rb = warp / CQ, cq = warp - rb * CQ;
    const int qrow = 4 * rb + (lane >> 3);                         // the row whose sum this lane holds after the butterfly
    double sreg[4][2 * Cfg::NP_];
    if (SREG) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = 4 * rb + rr;
            const bool on = r < nrows;
            const double* src = d.dn.S + (size_t)(row0 + (on ? r : 0)) * n;
#pragma unroll
            for (int t = 0; t < NP; ++t) {
                const int c = 2 * (lane + 32 * (cq + CQ * t));
                sreg[rr][2 * t]     = (on && c < n) ? __ldcg(src + c) : 0.0;
                sreg[rr][2 * t + 1] = (on && c + 1 < n) ? __ldcg(src + c + 1) : 0.0;
            }
        }
    } else {
        for (int r = warp; r < ((rows + 3) & ~3); r += NWARP) {
            const double* src = d.dn.S + (size_t)(row0 + r) * n; double* dst = sm.S + (size_t)r * ld;
            for (int c = lane; c < ld; c += 32) dst[c] = (r < nrows && c < n) ? __ldcg(src + c) : 0.0;
        }
    }
    for (int i = n + tid; i < ld; i += NT) { sm.p[i] = 0.0; sm.z[i] = 0.0; }   // zero padding of the vectors
    if (tid < 40) sm.red[tid] = 0.0;                               // (warps that own no row never write their p.q partial)
    const bool own = tid < nrows;
    const int e = row0 + tid, cam = e / 6, a = e - 6 * cam, lc = tid - a;
    double mrow[6] = {0, 0, 0, 0, 0, 0}, xe = 0, re = 0, ze = 0;
    if (own) {
        const double* mr = d.Minv + 36 * (size_t)cam + 6 * a;
#pragma unroll
        for (int c = 0; c < 6; ++c) mrow[c] = mr[c];
        re = d.gc[e] - d.ar_q[e];
        sm.r[tid] = re;
    }
    __syncthreads();
    const uint32_t z_addr = smem_u32(sm.z), rz_addr = smem_u32(sm.rz), pq_addr = smem_u32(sm.pq), bar_a = smem_u32(&sm.bar[0]), bar_b = smem_u32(&sm.bar[1]);
#ifdef B200BA_STAMPS
    long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define DSTAMP(i) do { const long long tn_ = clock64(); ph[i] += tn_ - tprev; tprev = tn_; } while (0)
#else
#define DSTAMP(i) do {} while (0)
#endif
    // z on the owned rows and the CTA's r.z partial go to every CTA of the group (phase B)
    auto exchange_z = [&](double zval, double partial) -> double {
        if (!GRID) {                                               // the z values travel while the r.z partial is being reduced
            if (tid == 0) mbar_expect_tx(&sm.bar[1], (uint32_t)(8 * (n + G)));
            if (own) for (int c = 0; c < G; ++c) st_async_f64(cluster_map(z_addr + 8u * (uint32_t)e, c), zval, cluster_map(bar_b, c));
        }
        // partial = this thread's r * z (0 outside the owned rows); fixed-order sum over the (<= 64) owned rows
        if (tid < 64) { const double s = warp_sum(partial); if ((tid & 31) == 0) sm.red[tid >> 5] = s; }
        __syncthreads();
        const double tot = sm.red[0] + sm.red[1];
        if (GRID) {
            if (own) d.dn.zg[e] = zval;
            if (tid == 0) d.dn.rz_slots[rank] = tot;
            grid_barrier(d.dn.bar_counter, ++gbar, G, rank);
            return grid_slot_sum(d.dn.rz_slots, G, sm.red);
        } else {
            if (tid < G) st_async_f64(cluster_map(rz_addr + 8u * (uint32_t)rank, tid), tot, cluster_map(bar_b, tid));
            mbar_wait_cluster(&sm.bar[1], par_b); par_b ^= 1u;
        }
        return fixed_sum_slots(sm.rz, G);
    };
    // the CTA's p.q (its warps' partials are in red[8 ..]) goes to every CTA of the group (phase A)
    auto exchange_pq = [&]() -> double {
        const double tot = fixed_sum_tree<NWARP>(sm.red + 8);
        DSTAMP(5);
        if (GRID) {
            if (tid == 0) d.dn.pq_slots[rank] = tot;
            grid_barrier(d.dn.bar_counter, ++gbar, G, rank);
            DSTAMP(6);
            return grid_slot_sum(d.dn.pq_slots, G, sm.red);
        } else {
            if (tid == 0) mbar_expect_tx(&sm.bar[0], (uint32_t)(8 * G));
            if (tid < G) st_async_f64(cluster_map(pq_addr + 8u * (uint32_t)rank, tid), tot, cluster_map(bar_a, tid));
            mbar_wait_cluster(&sm.bar[0], par_a); par_a ^= 1u;
        }
        DSTAMP(6);
        return fixed_sum_slots(sm.pq, G);
    };
    if (own) ze = (mrow[0] * sm.r[lc] + mrow[1] * sm.r[lc + 1]) + (mrow[2] * sm.r[lc + 2] + mrow[3] * sm.r[lc + 3]) + (mrow[4] * sm.r[lc + 4] + mrow[5] * sm.r[lc + 5]);
    double rz = exchange_z(ze, own ? re * ze : 0.0);
    const double rz0 = rz;
    for (int i = tid; i < n; i += NT) sm.p[i] = GRID ? __ldcg(d.dn.zg + i) : sm.z[i];
    __syncthreads();
    int steps = 0;
    double rel = 0;
    const double tol = st->cg_tol, guard = st->cg_guard;
    const int min_steps = st->cg_min_steps;
    const bool watch = tol > 0 || guard > 0;                       // fixed-step mode: the relative residual is only reported, not tested
    bool active = rz > 0;
    // ---- the loop ----
    while (active && steps < max_steps) {
        DSTAMP(9);
        const double inv_rz = 1.0 / rz;                            // for beta: issued here, its latency hides behind the product
        // q = S p on the CTA's rows: 4 rows per lane share every p value, two accumulators per row (fp64 latency); the four
        // row sums are reduced over the 32 lanes by a butterfly that halves the number of values per lane in its first two
        // levels (6 value-shuffles instead of 20), all in a fixed order
        {
            double a[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            if (SREG) {
#pragma unroll
                for (int t = 0; t < NP; ++t) {
                    const double2 pv = *reinterpret_cast<const double2*>(sm.p + 2 * (lane + 32 * (cq + CQ * t)));
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) { a[rr][0] = fma(sreg[rr][2 * t], pv.x, a[rr][0]); a[rr][1] = fma(sreg[rr][2 * t + 1], pv.y, a[rr][1]); }
                }
            } else {
                const int rows4 = (rows + 3) & ~3;
                const double* s0 = sm.S + (size_t)min(4 * rb, rows4 - 4) * ld;   // (warps beyond the CTA's rows read valid memory and write nothing)
                const int np = (n + 63) >> 6;
                for (int t = 0; t < np; ++t) {
                    const int c = 2 * (lane + 32 * t);
                    const double2 pv = *reinterpret_cast<const double2*>(sm.p + c);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const double2 sv = *reinterpret_cast<const double2*>(s0 + (size_t)rr * ld + c);
                        a[rr][0] = fma(sv.x, pv.x, a[rr][0]); a[rr][1] = fma(sv.y, pv.y, a[rr][1]);
                    }
                }
            }
            const double v0 = a[0][0] + a[0][1], v1 = a[1][0] + a[1][1], v2 = a[2][0] + a[2][1], v3 = a[3][0] + a[3][1];
            const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
            double k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;
            k0 += __shfl_xor_sync(0xffffffffu, hi16 ? v0 : v2, 16);
            k1 += __shfl_xor_sync(0xffffffffu, hi16 ? v1 : v3, 16);
            double acc = hi8 ? k1 : k0;
            acc += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);           // every lane of an 8-lane group: the sum of row `qrow` over this warp's columns
            const bool writer = (lane & 7) == 0 && qrow < nrows;
            if (CQ == 1) {
                if (writer) sm.q[qrow] = acc;
                double pqp = writer ? sm.p[row0 + qrow] * acc : 0.0;
                pqp += __shfl_xor_sync(0xffffffffu, pqp, 8);
                pqp += __shfl_xor_sync(0xffffffffu, pqp, 16);
                if (lane == 0) sm.red[8 + warp] = pqp;
            } else if (writer) sm.q[qrow * CQ + cq] = acc;          // partial of one column split; summed by the row's owner below
        }
        __syncthreads();
        DSTAMP(0);
        const double pe = own ? sm.p[e] : 0.0;
        double qe = 0.0;
        if (CQ == 1) { if (own) qe = sm.q[tid]; }
        else {
            if (own) { for (int c = 0; c < CQ; ++c) qe += sm.q[tid * CQ + c]; }
            if (tid < 64) { const double sp = warp_sum(pe * qe); if (lane == 0) sm.red[8 + warp] = sp; }
            __syncthreads();
        }
        const double pq = exchange_pq();
        DSTAMP(1);
        if (!(pq > 0) || !isfinite(pq)) {                          // breakdown (uniform): keep x; first-step breakdown = failed solve
            if (rank == 0 && tid == 0 && steps == 0) st->solve_ok = 0;
            break;
        }
        const double alpha = rz * fast_rcp(pq);                    // (within an ulp of rz / pq, without the IEEE division's slow path)
        if (own) { xe += alpha * pe; re -= alpha * qe; sm.r[tid] = re; }
        __syncthreads();
        if (own) ze = (mrow[0] * sm.r[lc] + mrow[1] * sm.r[lc + 1]) + (mrow[2] * sm.r[lc + 2] + mrow[3] * sm.r[lc + 3]) + (mrow[4] * sm.r[lc + 4] + mrow[5] * sm.r[lc + 5]);
        DSTAMP(2);
        const double rzn = exchange_z(ze, own ? re * ze : 0.0);
        DSTAMP(3);
        if (watch) rel = sqrt(fabs(rzn) / rz0);
        const bool stop = (tol > 0 && rel < tol && steps + 1 >= min_steps) || !(rzn > 0) || (guard > 0 && rel > guard);
        if (!stop) {
            const double beta = rzn * inv_rz;
            for (int i = tid; i < n; i += NT) sm.p[i] = (GRID ? __ldcg(d.dn.zg + i) : sm.z[i]) + beta * sm.p[i];
        }
        rz = rzn; ++steps;
        __syncthreads();
        DSTAMP(4);
        if (stop) active = false;
    }
#ifdef B200BA_STAMPS
    if (rank == 0 && tid == 0) { for (int i = 0; i < 10; ++i) d.stamps[i] = (unsigned long long)ph[i]; d.stamps[10] = (unsigned long long)steps; }
#endif
    if (!watch && steps > 0) rel = sqrt(fabs(rz) / rz0);
    if (own) d.x[e] = xe;
    if (rank == 0 && tid == 0) { st->cg_steps = steps; st->total_cg += steps; st->cg_rel = rel; st->rz = rz; st->rz0 = rz0; st->cg_active = 0; }
    // a CTA must not exit while a peer may still store into its shared memory; every st.async of a phase has landed when the
    // receiver passes the phase's wait, so nothing is in flight here - the barrier only keeps the exit order tidy
    if (!GRID) cl.sync();
}

} // namespace b200ba
