// ba_dense.cuh -- explicit reduced camera system for small problems (see dense_host.h for why and for the pair lists).
//
//   k_dense_wobs      per observation (thread per point over its SELL slots):  W = A~^t B~ (6x3),  Y = W (V + damping)^-1,
//                     t = Y g_p  (the observation's share of the reduced right-hand side)
//   k_dense_pairs     36 threads per chunk of <= 32 pairs of one block:  partial  sum  Y_a W_b^t   (+ partial sum of t on the diagonal)
//   k_dense_finalize  thread per (block, entry): fixed-order sum of the chunk partials ->  S = [U + damping] - sum  (both
//                     triangles, exactly symmetric), the diagonal blocks' sums for the preconditioner (ar_sd), sum W Vinv g_p (ar_q)
//   k_dense_pcg       the WHOLE block-Jacobi PCG loop in one launch.  CTA c of the group owns `rows` rows of S in shared memory
//                     (whole cameras); per step: q = S p on the owned rows, p.q -> every CTA, alpha, x, r, z = M^-1 r on the
//                     owned rows, z and r.z -> every CTA, beta, p = z + beta p (every CTA, all rows).  Two barriers per step.
//                       CLUSTER variant: the group is one thread-block cluster (8 or 16 CTAs); the exchange is
//                         st.shared::cluster into every peer's shared memory + barrier.cluster (~0.2 us);
//                       GRID variant: the group is the whole (cooperative) grid, one CTA per SM; the exchange goes through
//                         global memory + grid.sync() - for camera counts whose S does not fit one cluster (up to ~300 cameras).
//                     Same recurrences, stop rule and breakdown handling as k_cg_camera_step; all sums in fixed order.
#pragma once
#include "ba_kernels.cuh"
#include <cooperative_groups.h>

namespace b200ba {

constexpr int DPCG_THREADS = 512;
constexpr int DPCG_LPR = 8;                 // lanes per row of the S p product
constexpr int DPCG_MAX_ROWS = 64;           // rows per CTA (the per-row vector state lives in the first 64 threads)
constexpr int DPAIR_BLOCK = 288;            // 8 chunks x 36 entries

__global__ void __launch_bounds__(PT_BLOCK) k_dense_wobs(Dev d)
{
    const LMState* st = d.st;
    if (st->done) return;
    const int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    if (j >= d.M) return;
    const int cur = st->cur;
    double X[3], g[3], Vi[6];
    load3p(d.X[cur], j, X); load3p(d.gp, j, g); load6(d.Vinv, j, Vi);
    const double vg0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2];
    const double vg1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2];
    const double vg2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
    int k0, deg; slice_of(d, j, k0, deg);
    const double* rt = d.rt[cur];
    for (int s = 0, k = k0; s < deg; ++s, k += 32) {
        const int i = __ldg(o_cam_ptr(d, k));
        if (i < 0) continue;
        Cam cm; load_cam(rt, i, cm);
        Obs o; obs_jac(cm, X, __ldg(o_w_ptr(d, k)), d.in, o);
        double A0[6], A1[6], B0[3], B1[3];
        jac_rows_A(o, A0, A1); jac_rows_B(o, cm, B0, B1);
        const double live = i < d.first_free ? 0.0 : 1.0;       // constant camera: its W block is zero
        double2* Wo = reinterpret_cast<double2*>(d.dn.Wobs + 18 * (size_t)k);
        double2* Yo = reinterpret_cast<double2*>(d.dn.Yobs + 18 * (size_t)k);
        double* To = d.dn.Tobs + 6 * (size_t)k;
        double w[18], y[18];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double w0 = live * (A0[a] * B0[0] + A1[a] * B1[0]), w1 = live * (A0[a] * B0[1] + A1[a] * B1[1]), w2 = live * (A0[a] * B0[2] + A1[a] * B1[2]);
            w[3 * a] = w0; w[3 * a + 1] = w1; w[3 * a + 2] = w2;
            y[3 * a]     = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
            y[3 * a + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
            y[3 * a + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
            To[a] = w0 * vg0 + w1 * vg1 + w2 * vg2;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) { Wo[q] = make_double2(w[2 * q], w[2 * q + 1]); Yo[q] = make_double2(y[2 * q], y[2 * q + 1]); }
    }
}

__global__ void __launch_bounds__(DPAIR_BLOCK) k_dense_pairs(Dev d)
{
    if (d.st->done) return;
    const int c = blockIdx.x * (DPAIR_BLOCK / 36) + threadIdx.x / 36, e = threadIdx.x % 36;
    if (c >= d.dn.n_chunks) return;
    const int4 ch = __ldg(d.dn.chunks + c);                       // block, first pair, end pair, diagonal flag
    const int a = e / 6, b = e - 6 * a;
    const bool diag = ch.w != 0;
    if (diag && b < a) return;                                    // diagonal blocks: upper triangle, mirrored by the finalising kernel
    const double* Y = d.dn.Yobs + 3 * a; const double* W = d.dn.Wobs + 3 * b;
    double acc = 0, rhs = 0;
#pragma unroll 4
    for (int q = ch.y; q < ch.z; ++q) {
        const int2 pr = __ldg(d.dn.pairs + q);
        const double* y = Y + 18 * (size_t)pr.x; const double* w = W + 18 * (size_t)pr.y;
        acc += y[0] * w[0] + y[1] * w[1] + y[2] * w[2];
        if (diag && a == 0 && pr.x == pr.y) rhs += d.dn.Tobs[6 * (size_t)pr.x + b];
    }
    d.dn.chunk_part[36 * (size_t)c + e] = acc;
    if (diag && a == 0) d.dn.chunk_rhs[6 * (size_t)c + b] = rhs;
}

__global__ void __launch_bounds__(256) k_dense_finalize(Dev d)
{
    const LMState* st = d.st;
    if (st->done) return;
    const int idx = blockIdx.x * 256 + threadIdx.x, blk = idx / 36, e = idx - 36 * blk;
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
// shared-memory carve-up (doubles): Ssm[rows x ld] | p[ld] | z[ld] | qrow[rows] | rrow[rows] | pq_slot[256] | rz_slot[256] | red[64]
struct DensePcgSmem {
    double *S, *p, *z, *q, *r, *pq, *rz, *red;
    __device__ DensePcgSmem(double* base, int rows, int ld)
    {
        S = base; p = S + (size_t)rows * ld; z = p + ld; q = z + ld; r = q + rows; pq = r + rows; rz = pq + 256; red = rz + 256;
    }
};

// sum of the first `count` values of a (shared or global) array in a FIXED order, identical in every thread of every CTA:
// warp 0 adds strided partials and reduces them by the fixed butterfly, the result goes through shared memory
template <bool GRID>
__device__ __forceinline__ double dense_slot_sum(const double* slots, int count, double* red)
{
    if (!GRID) {                                                   // <= 16 slots in shared memory: every thread adds them itself
        double s = 0;
        for (int c = 0; c < count; ++c) s += slots[c];
        return s;
    }
    if (threadIdx.x < 32) {
        double s = 0;
        for (int c = threadIdx.x; c < count; c += 32) s += __ldcg(slots + c);
        s = warp_sum(s);
        if (threadIdx.x == 0) red[32] = s;
    }
    __syncthreads();
    const double s = red[32];
    __syncthreads();
    return s;
}

template <bool GRID>
__global__ void __launch_bounds__(DPCG_THREADS, 1) k_dense_pcg(Dev d, int max_steps)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double dsm[];
    LMState* st = d.st;
    if (st->done) return;                                          // uniform over the grid
    cg::cluster_group cl = cg::this_cluster();
    cg::grid_group grid = cg::this_grid();
    const int G = GRID ? (int)gridDim.x : (int)cl.num_blocks();
    const int rank = GRID ? (int)blockIdx.x : (int)cl.block_rank();
    const int tid = threadIdx.x, n = 6 * d.N, rows = d.dn.rows, ld = d.dn.ld;
    DensePcgSmem sm(dsm, rows, ld);
    const int row0 = rank * rows, nrows = max(0, min(rows, n - row0));
    auto group_sync = [&]() { if (GRID) grid.sync(); else cl.sync(); };
    if (!st->solve_ok) {                                           // (V + damping) was not positive definite: the step is rejected
        if (rank == 0 && tid == 0) { st->cg_active = 0; st->cg_steps = 0; st->cg_rel = 0; }
        return;
    }
    // ---- prologue: owned rows of S -> shared memory; r = rhs, x = 0, z = M^-1 r ----
    for (int r = tid >> 5; r < nrows; r += DPCG_THREADS / 32) {
        const double* src = d.dn.S + (size_t)(row0 + r) * n; double* dst = sm.S + (size_t)r * ld;
        for (int c = tid & 31; c < n; c += 32) dst[c] = __ldcg(src + c);
    }
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
    // z on the owned rows, r.z partial, both to every CTA of the group
    auto publish_z = [&](double zval, double partial) {
        // partial = this thread's r * z (0 outside the owned rows); fixed-order sum over the (<= 64) owned rows
        if (tid < 64) { const double s = warp_sum(partial); if ((tid & 31) == 0) sm.red[tid >> 5] = s; }
        __syncthreads();
        const double tot = sm.red[0] + sm.red[1];
        if (GRID) {
            if (own) d.dn.zg[e] = zval;
            if (tid == 0) d.dn.rz_slots[rank] = tot;
        } else {
            if (own) for (int c = 0; c < G; ++c) cl.map_shared_rank(sm.z, c)[e] = zval;
            if (tid < G) cl.map_shared_rank(sm.rz, tid)[rank] = tot;
        }
    };
    auto publish_pq = [&](double partial) {
        if (tid < 64) { const double s = warp_sum(partial); if ((tid & 31) == 0) sm.red[2 + (tid >> 5)] = s; }
        __syncthreads();
        const double tot = sm.red[2] + sm.red[3];
        if (GRID) { if (tid == 0) d.dn.pq_slots[rank] = tot; }
        else if (tid < G) cl.map_shared_rank(sm.pq, tid)[rank] = tot;
    };
    if (own) ze = mrow[0] * sm.r[lc] + mrow[1] * sm.r[lc + 1] + mrow[2] * sm.r[lc + 2] + mrow[3] * sm.r[lc + 3] + mrow[4] * sm.r[lc + 4] + mrow[5] * sm.r[lc + 5];
    publish_z(ze, own ? re * ze : 0.0);
    group_sync();
    double rz = dense_slot_sum<GRID>(GRID ? d.dn.rz_slots : sm.rz, G, sm.red);
    const double rz0 = rz;
    for (int i = tid; i < n; i += DPCG_THREADS) sm.p[i] = GRID ? __ldcg(d.dn.zg + i) : sm.z[i];
    __syncthreads();
    int steps = 0;
    double rel = 0;
    const double tol = st->cg_tol, guard = st->cg_guard;
    const int min_steps = st->cg_min_steps;
    bool active = rz > 0;
    // ---- the loop ----
    const int grp = tid / DPCG_LPR, gl = tid % DPCG_LPR;            // 64 row groups of 8 lanes
    while (active && steps < max_steps) {
        // q = S p on the owned rows: 8 lanes per row, two accumulators per lane, fixed butterfly over the 8 lanes
        {
            const bool on = grp < nrows;
            const double* srow = sm.S + (size_t)(on ? grp : 0) * ld;
            double a0 = 0, a1 = 0;
            int c = gl;
            for (; c + DPCG_LPR < n; c += 2 * DPCG_LPR) { a0 = fma(srow[c], sm.p[c], a0); a1 = fma(srow[c + DPCG_LPR], sm.p[c + DPCG_LPR], a1); }
            if (c < n) a0 = fma(srow[c], sm.p[c], a0);
            double acc = a0 + a1;
#pragma unroll
            for (int o = DPCG_LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (on && gl == 0) sm.q[grp] = acc;
        }
        __syncthreads();
        const double pe = own ? sm.p[e] : 0.0, qe = own ? sm.q[tid] : 0.0;
        publish_pq(pe * qe);
        group_sync();
        const double pq = dense_slot_sum<GRID>(GRID ? d.dn.pq_slots : sm.pq, G, sm.red);
        if (!(pq > 0) || !isfinite(pq)) {                          // breakdown (uniform): keep x; first-step breakdown = failed solve
            if (rank == 0 && tid == 0 && steps == 0) st->solve_ok = 0;
            break;
        }
        const double alpha = rz / pq;
        if (own) { xe += alpha * pe; re -= alpha * qe; sm.r[tid] = re; }
        __syncthreads();
        if (own) ze = mrow[0] * sm.r[lc] + mrow[1] * sm.r[lc + 1] + mrow[2] * sm.r[lc + 2] + mrow[3] * sm.r[lc + 3] + mrow[4] * sm.r[lc + 4] + mrow[5] * sm.r[lc + 5];
        publish_z(ze, own ? re * ze : 0.0);
        group_sync();
        const double rzn = dense_slot_sum<GRID>(GRID ? d.dn.rz_slots : sm.rz, G, sm.red);
        rel = sqrt(fabs(rzn) / rz0);
        const bool stop = (tol > 0 && rel < tol && steps + 1 >= min_steps) || !(rzn > 0) || (guard > 0 && rel > guard);
        if (!stop) {
            const double beta = rzn / rz;
            for (int i = tid; i < n; i += DPCG_THREADS) sm.p[i] = (GRID ? __ldcg(d.dn.zg + i) : sm.z[i]) + beta * sm.p[i];
        }
        rz = rzn; ++steps;
        __syncthreads();
        if (stop) active = false;
    }
    if (own) d.x[e] = xe;
    if (rank == 0 && tid == 0) { st->cg_steps = steps; st->total_cg += steps; st->cg_rel = rel; st->rz = rz; st->rz0 = rz0; st->cg_active = 0; }
    // a CTA must not exit while a peer may still store into its shared memory: every remote store precedes a barrier that
    // all CTAs pass, except after a breakdown (between the two barriers of a step) - one more barrier covers that
    if (!GRID) cl.sync();
}

} // namespace b200ba
