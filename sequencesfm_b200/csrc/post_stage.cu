// post_stage.cu -- post-BA ground-plane stage of libb200ba.so (include/b200ba_post.h), sm_100a only.
//
// Reference (paths relative to the reference tree): PointsPlaneFit_RANSAC src/sfm_utils.cpp:807-977 (FindSigmaSquared
// :792-805), CalcCameraGround_axes :1098-1146, TransCloudPoint :1148-1179, TransCamera :1215-1259, strung together by
// SfM_Tracker::bundleAdjustFrames src/sfm_tracker.cpp:876-908.
//
// The reference scores each of its n_ransac hypotheses by sorting the M point-plane distances (std::sort) to read the
// median, then sums the distances truncated at sigma^2 - 100 sorts of M doubles per keyframe.  Here the points are
// uploaded once (24 B / point, L2-resident for every later sweep up to ~5 M points) and all hypotheses are processed
// together by sweeps over the points, every CTA handling HG hypotheses x a grid-stride slice of the points:
//
//   k_post_hist    radix histogram of the distance bit patterns (non-negative doubles order like their uint64 bits),
//                  11 bits per pass, per-CTA histograms in shared memory, warp-aggregated (match_any) shared atomics
//   k_post_scan    one CTA per hypothesis: locate the digit holding the rank-M/2 element, narrow prefix and rank
//   k_post_gather  after two passes (exponent + 11 mantissa bits) the surviving candidates of each hypothesis are
//   k_post_select  appended to a small buffer and the exact order statistic is read off by rank counting in shared
//                  memory; a hypothesis that overflows the buffer sends the call down the remaining four radix passes
//   k_post_sigma   sigma^2 from the median in the reference's operation order (IEEE div / sqrt, no FMA)
//   k_post_score   truncated sums, fixed-order tree reduction (bit-reproducible run to run), k_post_pick = argmin
//   k_post_mask    distances to the winning plane: cpFilter, inlier count and sum; k_post_scatter = inlier scatter
//   k_post_reaxis_* re-axis of points (with order-preserving compaction through an exclusive scan) and cameras
//
// Distances are evaluated exactly like the reference's Eigen expression |(p - mean) . n| = |d0 n0 + (d1 n1 + d2 n2)|
// with __dsub_rn / __dmul_rn / __dadd_rn so that no FMA is contracted: medians, sigma^2 and the mask are bit-exact.
#include "../../include/b200ba_post.h"
#include "pinned_pool.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <utility>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// Compile-time knobs of the sweeps, A/B-timed on a B200 at 993 923 points x 100 hypotheses (tools/build_post_variants.sh;
// kernels of one call, ms): HG/PPT/MINB = 8/1/2 1.02, 8/2/2 1.10, 8/2/1 1.11, 8/1/1 1.36, 8/4/1 0.98, 4/4/2 0.96, 4/2/3 0.92;
// with the grid filling exactly MINB CTAs per SM: 4/2/3 0.72, 4/1/4 0.83; + register prefetch of the next points
// (POST_PREFETCH): 4/2/2 0.68, 4/2/3 0.95 (spills), 4/1/3 0.74, 4/1/4 0.96.
#ifndef POST_HG
#define POST_HG 4
#endif
#ifndef POST_PPT
#define POST_PPT 2
#endif
#ifndef POST_MINB
#define POST_MINB 2
#endif
#ifndef POST_PREFETCH
#define POST_PREFETCH 1
#endif
constexpr int kHG = POST_HG;           // hypotheses per CTA
constexpr int kPPT = POST_PPT;         // points per thread and loop iteration (all loads issued before the FP64 chains)
constexpr int kBits = 11;
constexpr int kBins = 1 << kBits;
constexpr int kSweepThreads = 256;
constexpr int kHypStride = 8;          // doubles per hypothesis: mean[3], normal[3], sigma2, valid
constexpr int kResDoubles = 32;        // device result block, see enum below
enum { R_BEST = 0, R_SCORE = 1, R_SIGMA2 = 2, R_MEAN = 3, R_NORMAL = 6, R_CNT = 9, R_KEPT = 10, R_IMEAN = 11,
       R_SCATTER = 14, R_FLAGS = 20 /* bit 0 non-finite input, bit 1 candidate overflow */ };

struct SelState { unsigned long long prefix; unsigned int rank; unsigned int count; };
struct Se3 { double a[12]; };

__device__ __forceinline__ double plane_dist(double x, double y, double z, const double* hp)
{
    const double d0 = __dsub_rn(x, hp[0]), d1 = __dsub_rn(y, hp[1]), d2 = __dsub_rn(z, hp[2]);
    return fabs(__dadd_rn(__dmul_rn(d0, hp[3]), __dadd_rn(__dmul_rn(d1, hp[4]), __dmul_rn(d2, hp[5]))));
}

// kPPT points of one thread and loop iteration: point u is element first + u * kSweepThreads of the cloud.
struct PtBatch { double x[kPPT], y[kPPT], z[kPPT]; bool in[kPPT]; };
__device__ __forceinline__ void load_batch(const double* __restrict__ pts, int M, int first, PtBatch& b)
{
#pragma unroll
    for (int u = 0; u < kPPT; ++u) {
        const int i = first + u * kSweepThreads;
        b.in[u] = i < M;
        b.x[u] = b.in[u] ? pts[3 * (size_t)i] : 0.0; b.y[u] = b.in[u] ? pts[3 * (size_t)i + 1] : 0.0; b.z[u] = b.in[u] ? pts[3 * (size_t)i + 2] : 0.0;
    }
}
// POST_PREFETCH: the next iteration's points are requested before the current ones are used (register double buffer).
#if POST_PREFETCH
#define SWEEP_LOOP(first_expr, cond_expr)                                                                     \
    const int sweep_stride = gridDim.x * kPPT * kSweepThreads;                                                \
    int sweep_first = (first_expr);                                                                           \
    PtBatch cur; load_batch(pts, M, sweep_first, cur);                                                        \
    for (int base = blockIdx.x * kPPT * kSweepThreads; (cond_expr); base += sweep_stride, sweep_first += sweep_stride)
#define SWEEP_NEXT() PtBatch nxt; load_batch(pts, M, sweep_first + sweep_stride, nxt)
#define SWEEP_ADVANCE() cur = nxt
#else
#define SWEEP_LOOP(first_expr, cond_expr)                                                                     \
    const int sweep_stride = gridDim.x * kPPT * kSweepThreads;                                                \
    int sweep_first = (first_expr);                                                                           \
    for (int base = blockIdx.x * kPPT * kSweepThreads; (cond_expr); base += sweep_stride, sweep_first += sweep_stride)
#define SWEEP_NEXT() PtBatch cur; load_batch(pts, M, sweep_first, cur)
#define SWEEP_ADVANCE() ((void)0)
#endif

__global__ void k_post_init(SelState* st, unsigned* cand_n, const double* hyp, int H, unsigned rank, unsigned M, double* res)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) { st[h].prefix = 0ull; st[h].rank = rank; st[h].count = hyp[h * kHypStride + 7] != 0.0 ? M : 0u; cand_n[h] = 0u; }
    if (h < kResDoubles) res[h] = 0.0;
}

// 3 points -> (mean, unit normal) per hypothesis, src/sfm_utils.cpp:847-854, in the reference's operation order with IEEE
// division / square root and no FMA: ((A + B) + C) / 3, n = (C - A) x (B - A), skipped when |n.n| < 1e-9 (valid = 0),
// n / sqrt(nx^2 + (ny^2 + nz^2)).  One thread per hypothesis; the points are gathered from the resident cloud.
__global__ void k_post_hypotheses(const double* __restrict__ pts, const int* __restrict__ triples, int H, double* __restrict__ hyp,
                                  unsigned* __restrict__ n_degenerate)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const double* A = pts + 3 * (size_t)triples[3 * h];
    const double* B = pts + 3 * (size_t)triples[3 * h + 1];
    const double* C = pts + 3 * (size_t)triples[3 * h + 2];
    double mean[3], ca[3], ba[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mean[k] = __ddiv_rn(__dadd_rn(__dadd_rn(A[k], B[k]), C[k]), 3.0);
        ca[k] = __dsub_rn(C[k], A[k]);
        ba[k] = __dsub_rn(B[k], A[k]);
    }
    double n[3];
    n[0] = __dsub_rn(__dmul_rn(ca[1], ba[2]), __dmul_rn(ca[2], ba[1]));
    n[1] = __dsub_rn(__dmul_rn(ca[2], ba[0]), __dmul_rn(ca[0], ba[2]));
    n[2] = __dsub_rn(__dmul_rn(ca[0], ba[1]), __dmul_rn(ca[1], ba[0]));
    const double nn = __dadd_rn(__dmul_rn(n[0], n[0]), __dadd_rn(__dmul_rn(n[1], n[1]), __dmul_rn(n[2], n[2])));
    double* q = hyp + (size_t)h * kHypStride;
    const bool ok = fabs(nn) >= 1e-9;                      // NaN counts as degenerate
    const double len = __dsqrt_rn(nn);
#pragma unroll
    for (int k = 0; k < 3; ++k) { q[k] = mean[k]; q[3 + k] = ok ? __ddiv_rn(n[k], len) : 0.0; }
    q[6] = 0.0;
    q[7] = ok ? 1.0 : 0.0;
    if (!ok) atomicAdd(n_degenerate, 1u);
}

// One radix pass for all hypotheses.  grid = (point slices, hypothesis groups); dynamic smem = kHG * kBins counters.
// AGG: warp-aggregate equal bins before the shared atomic (the exponent pass puts a whole warp into 2-3 bins); the later
// passes scatter over 2048 mantissa bins and use plain shared atomics (MATCH.ANY costs one round per distinct value).
// Two points per thread and iteration: both loads are in flight before the 2 x kHG dependent FP64 chains start.
template <bool AGG>
__global__ void __launch_bounds__(kSweepThreads, POST_MINB)
k_post_hist(const double* __restrict__ pts, int M, int H, const double* __restrict__ hyp, const SelState* __restrict__ st,
            int shift, int hi_shift, unsigned mask, unsigned* __restrict__ hist, double* res, int check_finite)
{
    extern __shared__ unsigned s_hist[];
    __shared__ double s_hyp[kHG][6];
    __shared__ unsigned long long s_prefix[kHG];
    __shared__ int s_valid[kHG];
    const int tid = threadIdx.x, lane = tid & 31, h0 = blockIdx.y * kHG;
    for (int i = tid; i < kHG * kBins; i += kSweepThreads) s_hist[i] = 0u;
    if (tid < kHG) {
        const int h = h0 + tid;
        const bool v = h < H && hyp[h * kHypStride + 7] != 0.0;
        s_valid[tid] = v;
        s_prefix[tid] = v ? st[h].prefix : 0ull;
        for (int k = 0; k < 6; ++k) s_hyp[tid][k] = v ? hyp[h * kHypStride + k] : 0.0;
    }
    __syncthreads();
    bool bad = false;
    SWEEP_LOOP(blockIdx.x * kPPT * kSweepThreads + tid, base < M) {                                   // trip count uniform per CTA
        SWEEP_NEXT();
        const double (&x)[kPPT] = cur.x; const double (&y)[kPPT] = cur.y; const double (&z)[kPPT] = cur.z; const bool (&in)[kPPT] = cur.in;
        if (check_finite) {
#pragma unroll
            for (int u = 0; u < kPPT; ++u) if (in[u] && !(isfinite(x[u]) && isfinite(y[u]) && isfinite(z[u]))) bad = true;
        }
#pragma unroll
        for (int g = 0; g < kHG; ++g) {
            if (!s_valid[g]) continue;                                    // uniform over the CTA
#pragma unroll
            for (int u = 0; u < kPPT; ++u) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(plane_dist(x[u], y[u], z[u], s_hyp[g]));
                const bool pass = in[u] && (key >> hi_shift) == s_prefix[g];
                const unsigned bin = (unsigned)(key >> shift) & mask;
                if (AGG) {
                    const unsigned peers = __match_any_sync(0xffffffffu, pass ? bin : 0xffffffffu);
                    if (pass && lane == __ffs(peers) - 1) atomicAdd(&s_hist[g * kBins + bin], (unsigned)__popc(peers));
                } else if (pass) {
                    atomicAdd(&s_hist[g * kBins + bin], 1u);
                }
            }
        }
        SWEEP_ADVANCE();
    }
    __syncthreads();
    for (int i = tid; i < kHG * kBins; i += kSweepThreads) {
        const unsigned v = s_hist[i];
        if (v) atomicAdd(&hist[(size_t)(h0 + i / kBins) * kBins + (i % kBins)], v);
    }
    if (bad) atomicOr((unsigned long long*)&res[R_FLAGS], 1ull);
}

// One CTA (1024 threads, 2 bins each) per hypothesis: find the digit that holds rank `rank`, clear the histogram.
__global__ void __launch_bounds__(1024)
k_post_scan(unsigned* __restrict__ hist, SelState* __restrict__ st, int bits)
{
    __shared__ unsigned s_warp[32];
    const int h = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (st[h].count == 0u) return;                                       // invalid hypothesis (uniform)
    const int nb = 1 << bits;
    unsigned* hh = hist + (size_t)h * kBins;
    unsigned a = 0, b = 0;
    if (2 * t < nb) { a = hh[2 * t]; b = hh[2 * t + 1]; hh[2 * t] = 0u; hh[2 * t + 1] = 0u; }
    const unsigned s = a + b;
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned v = s_warp[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
        s_warp[lane] = iv - v;                                           // exclusive warp offsets
    }
    __syncthreads();
    const unsigned excl = s_warp[w] + incl - s;
    const unsigned r = st[h].rank;
    __syncthreads();                                                     // everyone has read rank before it changes
    if (excl <= r && r < excl + s) {
        const bool first = r < excl + a;
        st[h].prefix = (st[h].prefix << bits) | (unsigned long long)(first ? 2 * t : 2 * t + 1);
        st[h].rank = first ? r - excl : r - excl - a;
        st[h].count = first ? a : b;
    }
}

// Append the keys that survived the first passes (key >> hi_shift == prefix) to each hypothesis' candidate buffer.
__global__ void __launch_bounds__(kSweepThreads, POST_MINB)
k_post_gather(const double* __restrict__ pts, int M, int H, const double* __restrict__ hyp, const SelState* __restrict__ st,
              int hi_shift, unsigned long long* __restrict__ cand, unsigned* __restrict__ cand_n, int cap)
{
    __shared__ double s_hyp[kHG][6];
    __shared__ unsigned long long s_prefix[kHG];
    __shared__ int s_valid[kHG];
    const int tid = threadIdx.x, h0 = blockIdx.y * kHG;
    if (tid < kHG) {
        const int h = h0 + tid;
        const bool v = h < H && hyp[h * kHypStride + 7] != 0.0;
        s_valid[tid] = v;
        s_prefix[tid] = v ? st[h].prefix : 0ull;
        for (int k = 0; k < 6; ++k) s_hyp[tid][k] = v ? hyp[h * kHypStride + k] : 0.0;
    }
    __syncthreads();
    SWEEP_LOOP(blockIdx.x * kPPT * kSweepThreads + tid, base + tid < M) {
        SWEEP_NEXT();
        const double (&x)[kPPT] = cur.x; const double (&y)[kPPT] = cur.y; const double (&z)[kPPT] = cur.z; const bool (&in)[kPPT] = cur.in;
#pragma unroll
        for (int g = 0; g < kHG; ++g) {
            if (!s_valid[g]) continue;
#pragma unroll
            for (int u = 0; u < kPPT; ++u) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(plane_dist(x[u], y[u], z[u], s_hyp[g]));
                if (in[u] && (key >> hi_shift) == s_prefix[g]) {
                    const unsigned slot = atomicAdd(&cand_n[h0 + g], 1u);
                    if (slot < (unsigned)cap) cand[(size_t)(h0 + g) * cap + slot] = key;
                }
            }
        }
        SWEEP_ADVANCE();
    }
}

// One CTA per hypothesis: the rank-th smallest candidate by counting (any order of the buffer gives the same value).
__global__ void __launch_bounds__(1024)
k_post_select(const unsigned long long* __restrict__ cand, const unsigned* __restrict__ cand_n, const SelState* __restrict__ st,
              int cap, unsigned long long* __restrict__ med, double* res)
{
    extern __shared__ unsigned long long s_key[];
    const int h = blockIdx.x, t = threadIdx.x;
    if (st[h].count == 0u) return;
    const unsigned c = cand_n[h];
    if (c > (unsigned)cap || c != st[h].count) { if (t == 0) atomicOr((unsigned long long*)&res[R_FLAGS], 2ull); return; }
    for (unsigned i = t; i < c; i += 1024) s_key[i] = cand[(size_t)h * cap + i];
    __syncthreads();
    const unsigned r = st[h].rank;
    for (unsigned j = t; j < c; j += 1024) {
        const unsigned long long kj = s_key[j];
        unsigned less = 0, eq = 0;
        for (unsigned k = 0; k < c; ++k) { const unsigned long long kk = s_key[k]; less += kk < kj; eq += kk == kj; }
        if (less <= r && r < less + eq) med[h] = kj;
    }
}

// sigma^2 = FindSigmaSquared(err) / 20, src/sfm_utils.cpp:792-805,865, same operation order, IEEE-rounded, no FMA.
__global__ void k_post_sigma(double* hyp, const SelState* st, const unsigned long long* med, int from_state, int H, double den)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H || hyp[h * kHypStride + 7] == 0.0) return;
    const double m = __longlong_as_double((long long)(from_state ? st[h].prefix : med[h]));
    double s = __dmul_rn(__dmul_rn(1.4826, __dadd_rn(1.0, __ddiv_rn(5.0, den))), __dsqrt_rn(m));
    s = __dmul_rn(4.6851, s);
    hyp[h * kHypStride + 6] = __ddiv_rn(__dmul_rn(s, s), 20.0);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Truncated sums: partial[(h, slice)] = sum over this CTA's points of min(err, sigma2).
__global__ void __launch_bounds__(kSweepThreads, POST_MINB)
k_post_score(const double* __restrict__ pts, int M, int H, const double* __restrict__ hyp, double* __restrict__ partial)
{
    __shared__ double s_hyp[kHG][7];
    __shared__ int s_valid[kHG];
    __shared__ double s_red[kSweepThreads / 32][kHG];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, h0 = blockIdx.y * kHG;
    if (tid < kHG) {
        const int h = h0 + tid;
        const bool v = h < H && hyp[h * kHypStride + 7] != 0.0;
        s_valid[tid] = v;
        for (int k = 0; k < 7; ++k) s_hyp[tid][k] = v ? hyp[h * kHypStride + k] : 0.0;
    }
    __syncthreads();
    double acc[kHG];
#pragma unroll
    for (int g = 0; g < kHG; ++g) acc[g] = 0.0;
    SWEEP_LOOP(blockIdx.x * kPPT * kSweepThreads + tid, base + tid < M) {
        SWEEP_NEXT();
        const double (&x)[kPPT] = cur.x; const double (&y)[kPPT] = cur.y; const double (&z)[kPPT] = cur.z; const bool (&in)[kPPT] = cur.in;
#pragma unroll
        for (int g = 0; g < kHG; ++g) {
            const double s2 = s_hyp[g][6];
#pragma unroll
            for (int u = 0; u < kPPT; ++u) {        // fixed order u = 0, 1, ...: run-to-run reproducible
                const double e = plane_dist(x[u], y[u], z[u], s_hyp[g]);
                acc[g] += in[u] ? (e > s2 ? s2 : e) : 0.0;
            }
        }
        SWEEP_ADVANCE();
    }
#pragma unroll
    for (int g = 0; g < kHG; ++g) { const double v = warp_sum(acc[g]); if (lane == 0) s_red[w][g] = v; }
    __syncthreads();
    if (tid < kHG && h0 + tid < H) {
        double v = 0.0;
        for (int k = 0; k < kSweepThreads / 32; ++k) v += s_red[k][tid];
        partial[(size_t)(h0 + tid) * gridDim.x + blockIdx.x] = s_valid[tid] ? v : 0.0;
    }
}

// Sum the slices in order, then the reference's `if (dSumError < dBestDistSquared)` scan from 9e99: first minimum wins.
// dynamic smem: H doubles.
__global__ void __launch_bounds__(256)
k_post_pick(const double* __restrict__ partial, int nx, int H, const double* __restrict__ hyp, double* __restrict__ score, double* res)
{
    extern __shared__ double s_score[];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        double v = 0.0;
        for (int k = 0; k < nx; ++k) v += partial[(size_t)h * nx + k];
        score[h] = v;
        s_score[h] = hyp[h * kHypStride + 7] != 0.0 ? v : INFINITY;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double best = 9e99; int bi = -1;
        for (int h = 0; h < H; ++h)
            if (s_score[h] < best) { best = s_score[h]; bi = h; }
        res[R_BEST] = (double)bi; res[R_SCORE] = best;
        if (bi >= 0) {
            res[R_SIGMA2] = hyp[bi * kHypStride + 6];
            for (int k = 0; k < 6; ++k) res[R_MEAN + k] = hyp[bi * kHypStride + k];
        }
    }
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* out /* NV per CTA */)
{
    __shared__ double s_red[kSweepThreads / 32][NV];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) { const double r = warp_sum(v[k]); if (lane == 0) s_red[w][k] = r; }
    __syncthreads();
    if (threadIdx.x < NV) {
        double r = 0.0;
        for (int k = 0; k < kSweepThreads / 32; ++k) r += s_red[k][threadIdx.x];
        out[(size_t)blockIdx.x * NV + threadIdx.x] = r;
    }
}

// Distances to the winning plane: cpFilter (err < k sigma2), inliers (err < sigma2): count and coordinate sums.
__global__ void __launch_bounds__(kSweepThreads, POST_MINB)
k_post_mask(const double* __restrict__ pts, int M, const double* __restrict__ res, double k_threshold,
            int* __restrict__ keep, double* __restrict__ partial /* 5 per CTA: cnt, kept, sx, sy, sz */)
{
    __shared__ double s_b[7];
    if (threadIdx.x < 6) s_b[threadIdx.x] = res[R_MEAN + threadIdx.x];
    if (threadIdx.x == 6) s_b[6] = res[R_SIGMA2];
    __syncthreads();
    const double t_in = s_b[6], t_flt = __dmul_rn(s_b[6], k_threshold);
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = blockIdx.x * kSweepThreads + threadIdx.x; i < M; i += gridDim.x * kSweepThreads) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        const double e = plane_dist(x, y, z, s_b);
        const bool kp = e < t_flt;
        keep[i] = kp ? 1 : 0;
        v[1] += kp ? 1.0 : 0.0;
        if (e < t_in) { v[0] += 1.0; v[2] += x; v[3] += y; v[4] += z; }
    }
    block_reduce_store<5>(v, partial);
}

// Second-level sums: warp k adds the nx per-CTA partials of quantity k (lane-strided, then a fixed shuffle tree).
template <int NV>
__device__ __forceinline__ double second_level_sum(const double* __restrict__ partial, int nx)
{
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    if (k < NV) for (int b = lane; b < nx; b += 32) v += partial[(size_t)b * NV + k];
    return warp_sum(v);
}

__global__ void __launch_bounds__(32 * 5) k_post_mean(const double* __restrict__ partial, int nx, double* res)
{
    __shared__ double s[5];
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double v = second_level_sum<5>(partial, nx);
    if (lane == 0) s[k] = v;
    __syncthreads();
    if (threadIdx.x == 0) res[R_CNT] = s[0];
    if (threadIdx.x == 1) res[R_KEPT] = s[1];
    if (threadIdx.x >= 2 && threadIdx.x < 5) res[R_IMEAN + threadIdx.x - 2] = __ddiv_rn(s[threadIdx.x], s[0]);   // v3MeanOfInliers / size
}

// Scatter matrix of the inliers about their mean (6 unique entries), src/sfm_utils.cpp:911-916.
__global__ void __launch_bounds__(kSweepThreads, POST_MINB)
k_post_scatter(const double* __restrict__ pts, int M, const double* __restrict__ res, double* __restrict__ partial /* 6 per CTA */)
{
    __shared__ double s_b[10];
    if (threadIdx.x < 6) s_b[threadIdx.x] = res[R_MEAN + threadIdx.x];
    if (threadIdx.x == 6) s_b[6] = res[R_SIGMA2];
    if (threadIdx.x >= 7 && threadIdx.x < 10) s_b[threadIdx.x] = res[R_IMEAN + threadIdx.x - 7];
    __syncthreads();
    const double t_in = s_b[6];
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * kSweepThreads + threadIdx.x; i < M; i += gridDim.x * kSweepThreads) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        if (plane_dist(x, y, z, s_b) < t_in) {
            const double dx = x - s_b[7], dy = y - s_b[8], dz = z - s_b[9];
            v[0] += dx * dx; v[1] += dx * dy; v[2] += dx * dz; v[3] += dy * dy; v[4] += dy * dz; v[5] += dz * dz;
        }
    }
    block_reduce_store<6>(v, partial);
}

__global__ void __launch_bounds__(32 * 6) k_post_scatter_final(const double* __restrict__ partial, int nx, double* res)
{
    const double v = second_level_sum<6>(partial, nx);
    if ((threadIdx.x & 31) == 0) res[R_SCATTER + (threadIdx.x >> 5)] = v;
}

// TransCloudPoint (Matrix4d * [p; 1], accumulated left to right, src/sfm_utils.cpp:1148-1179); pos == NULL: in place,
// otherwise point i goes to slot pos[i] of out when keep[i] (order-preserving compaction).
__global__ void __launch_bounds__(256)
k_post_reaxis_points(const double* __restrict__ pts, int M, Se3 T, const int* __restrict__ keep, const int* __restrict__ pos,
                     double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    if (keep && !keep[i]) return;
    const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    const size_t o = 3 * (size_t)(pos ? pos[i] : i);
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out[o + r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.a[4 * r], x), __dmul_rn(T.a[4 * r + 1], y)),
                                         __dmul_rn(T.a[4 * r + 2], z)), __dmul_rn(T.a[4 * r + 3], 1.0));
}

// TransCamera (src/sfm_utils.cpp:1215-1259): R' = Rc RtI, t' = Rc ttI + tc, fixed-size products left to right.
__global__ void __launch_bounds__(128)
k_post_reaxis_cameras(double* __restrict__ cams, int N, Se3 T)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double RtI[9], ttI[3], c[12], o[12];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) RtI[3 * i + j] = T.a[4 * j + i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        ttI[r] = __dadd_rn(__dadd_rn(__dmul_rn(-RtI[3 * r], T.a[3]), __dmul_rn(-RtI[3 * r + 1], T.a[7])),
                           __dmul_rn(-RtI[3 * r + 2], T.a[11]));
#pragma unroll
    for (int k = 0; k < 12; ++k) c[k] = cams[12 * (size_t)n + k];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            o[4 * r + k] = __dadd_rn(__dadd_rn(__dmul_rn(c[4 * r], RtI[k]), __dmul_rn(c[4 * r + 1], RtI[3 + k])),
                                     __dmul_rn(c[4 * r + 2], RtI[6 + k]));
        o[4 * r + 3] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[4 * r], ttI[0]), __dmul_rn(c[4 * r + 1], ttI[1])),
                                           __dmul_rn(c[4 * r + 2], ttI[2])), c[4 * r + 3]);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) cams[12 * (size_t)n + k] = o[k];
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side: the O(n_ransac) and O(1) scalar work, in the reference's operation order (this TU is built without FMA
// contraction on the host: x86-64 baseline has none).
// ---------------------------------------------------------------------------------------------------------------------
inline double dot3_pairwise(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline void unit3(double* v) { const double n = std::sqrt(dot3_pairwise(v, v)); v[0] /= n; v[1] /= n; v[2] /= n; }
inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

// Two-sided Jacobi SVD of a 3x3 in single precision, the algorithm the reference calls through Eigen 3.1.3
// (JacobiSVD<MatrixXf>, Thirdparty/eigen3/Eigen/src/SVD/JacobiSVD.h:724-818): sweep order (1,0),(2,0),(2,1); a 2x2 block
// is first symmetrised by a rotation, then diagonalised by a Jacobi rotation; U columns get the sign of the diagonal
// and are sorted by descending singular value.  Only U is needed.
struct PlaneRot { float c, s; };
struct Svd3f {
    float w[3][3], u[3][3], sv[3];
    void rows(int p, int q, PlaneRot j) { for (int i = 0; i < 3; ++i) { const float x = w[p][i], y = w[q][i]; w[p][i] = j.c * x + j.s * y; w[q][i] = -j.s * x + j.c * y; } }
    static void cols(float (&m)[3][3], int p, int q, PlaneRot j) { for (int i = 0; i < 3; ++i) { const float x = m[i][p], y = m[i][q]; m[i][p] = j.c * x + j.s * y; m[i][q] = -j.s * x + j.c * y; } }
    static PlaneRot symmetriser(float m00, float m01, float m10, float m11)
    {
        const float t = m00 + m11, d = m10 - m01;
        if (t == 0.f) return PlaneRot{0.f, d > 0.f ? 1.f : -1.f};
        const float u = d / t, c = 1.f / std::sqrt(1.f + u * u);
        return PlaneRot{c, c * u};
    }
    static PlaneRot jacobi(float x, float y, float z)
    {
        if (y == 0.f) return PlaneRot{1.f, 0.f};
        const float tau = (x - z) / (2.f * std::fabs(y)), w = std::sqrt(tau * tau + 1.f);
        const float t = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
        const float sgn = t > 0.f ? 1.f : -1.f, n = 1.f / std::sqrt(t * t + 1.f);
        return PlaneRot{n, -sgn * (y / std::fabs(y)) * std::fabs(t) * n};
    }
    void run(const float (&a)[3][3])
    {
        std::memcpy(w, a, sizeof(w));
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) u[i][j] = i == j ? 1.f : 0.f;
        const float precision = 2.f * FLT_EPSILON, tiny = 2.f * std::numeric_limits<float>::denorm_min();
        for (bool done = false; !done;) {
            done = true;
            for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {
                const float thr = std::max(tiny, precision * std::max(std::fabs(w[p][p]), std::fabs(w[q][q])));
                if (!(std::max(std::fabs(w[p][q]), std::fabs(w[q][p])) > thr)) continue;
                done = false;
                const PlaneRot r1 = symmetriser(w[p][p], w[p][q], w[q][p], w[q][q]);
                const float m00 = r1.c * w[p][p] + r1.s * w[q][p], m01 = r1.c * w[p][q] + r1.s * w[q][q];
                const float m11 = -r1.s * w[p][q] + r1.c * w[q][q];
                const PlaneRot jr = jacobi(m00, m01, m11);
                const PlaneRot jl{r1.c * jr.c - r1.s * (-jr.s), r1.c * (-jr.s) + r1.s * jr.c};   // rot1 * j_right^t
                rows(p, q, jl);
                cols(u, p, q, jl);
                cols(w, p, q, PlaneRot{jr.c, -jr.s});
            }
        }
        for (int i = 0; i < 3; ++i) {
            sv[i] = std::fabs(w[i][i]);
            if (sv[i] != 0.f) { const float z = w[i][i] / sv[i]; for (int r = 0; r < 3; ++r) u[r][i] *= z; }
        }
        for (int i = 0; i < 3; ++i) {
            int pos = i;
            for (int k = i + 1; k < 3; ++k) if (sv[k] > sv[pos]) pos = k;
            if (sv[pos] == 0.f) break;
            if (pos != i) { std::swap(sv[i], sv[pos]); for (int r = 0; r < 3; ++r) std::swap(u[r][i], u[r][pos]); }
        }
    }
};

// src/sfm_utils.cpp:918-974: principal axes of the inlier scatter -> plane (normal = least-variance axis) and plane_se3.
void plane_from_moments(const double* mean, const double* scatter9, double* plane4, double* se3)
{
    float a[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = (float)scatter9[3 * i + j];
    Svd3f svd; svd.run(a);
    float c0[3], c2[3];
    const float n0 = std::sqrt((svd.u[0][0] * svd.u[0][0] + svd.u[1][0] * svd.u[1][0]) + svd.u[2][0] * svd.u[2][0]);
    const float n2 = std::sqrt((svd.u[0][2] * svd.u[0][2] + svd.u[1][2] * svd.u[1][2]) + svd.u[2][2] * svd.u[2][2]);
    for (int r = 0; r < 3; ++r) { c0[r] = svd.u[r][0] / n0; c2[r] = svd.u[r][2] / n2; }
    double ax[3] = {c0[0], c0[1], c0[2]}, az[3] = {c2[0], c2[1], c2[2]}, ay[3];
    ax[2] = (-az[0] * ax[0] - az[1] * ax[1]) / az[2];          // force ax into the plane
    unit3(ax);
    cross3(az, ax, ay);
    unit3(ay);
    plane4[0] = az[0]; plane4[1] = az[1]; plane4[2] = az[2]; plane4[3] = -dot3_pairwise(mean, az);
    const double* R[3] = {ax, ay, az};
    for (int r = 0; r < 3; ++r) {
        double t = 0.0;                                         // t = -R * mean (dynamic-size gemv: column by column)
        for (int k = 0; k < 3; ++k) t += R[r][k] * (-1.0 * mean[k]);
        for (int k = 0; k < 3; ++k) se3[4 * r + k] = R[r][k];
        se3[4 * r + 3] = t;
    }
}

// cv::Matx helpers of src/sfm_utils.cpp:994-1096
void camera_centre(const double* m, double* c)                // LocalCoord2G: -R^t t, accumulated from 0
{
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += (-m[4 * k + i]) * m[4 * k + 3]; c[i] = s; }
}
double plane_residual(const double* pt, const double* p) { return std::fabs(p[0] * pt[0] + p[1] * pt[1] + p[2] * pt[2] + p[3]); }
void foot_on_plane(const double* p, const double* pt, double* out)   // point2Plane :1070-1096
{
    const double d = (p[0] * pt[0] + p[1] * pt[1] + p[2] * pt[2] + p[3]) / std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = pt[k] - d * p[k]; hi[k] = pt[k] + d * p[k]; }
    const double* c = plane_residual(lo, p) < plane_residual(hi, p) ? lo : hi;
    out[0] = c[0]; out[2] = c[2];
    out[1] = (-p[0] * out[0] - p[2] * out[2] - p[3]) / p[1];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct b200post_handle {
    int device = 0;
    int sms = 148;
    int cap = 4096;                     // candidate buffer per hypothesis; <= 0: radix passes only
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    int launches = 0;
    // grow-only device buffers
    double* d_pts = nullptr; double* d_out = nullptr; int* d_keep = nullptr; int* d_pos = nullptr; size_t cap_pts = 0;
    double* d_hyp = nullptr; unsigned* d_hist = nullptr; SelState* d_st = nullptr; SelState* d_st_b = nullptr;
    unsigned long long* d_cand = nullptr; unsigned* d_cand_n = nullptr; unsigned long long* d_med = nullptr;
    double* d_score = nullptr; int* d_tri = nullptr; unsigned* d_ndeg = nullptr; size_t cap_hyp = 0;
    double* d_partial = nullptr; size_t cap_partial = 0;
    double* d_res = nullptr;
    double* d_cams = nullptr; size_t cap_cams = 0;
    void* d_scan_tmp = nullptr; size_t cap_scan_tmp = 0;
    // the resident problem (for b200post_debug_time_kernel)
    int M = 0, H = 0;
    double k_threshold = 2.0;
};

namespace {
thread_local std::string g_post_create_error;

#define POST_CUDA(call)                                                                                         \
    do { cudaError_t e_ = (call);                                                                               \
         if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return B200BA_E_CUDA; } } while (0)

template <class T> int grow(b200post_handle* h, T*& p, size_t n)
{
    if (p) cudaFree(p);
    p = nullptr;
    POST_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    return 0;
}

int reserve_points(b200post_handle* h, size_t M)
{
    if (M <= h->cap_pts) return 0;
    const size_t n = M + M / 8 + 1024;
    h->cap_pts = 0;                                    // a failed growth must not leave a stale capacity behind
    h->M = 0;
    int rc;
    if ((rc = grow(h, h->d_pts, 3 * n)) || (rc = grow(h, h->d_out, 3 * n)) || (rc = grow(h, h->d_keep, n)) || (rc = grow(h, h->d_pos, n))) return rc;
    h->cap_pts = n;
    return 0;
}

int reserve_hyp(b200post_handle* h, size_t H)
{
    if (H <= h->cap_hyp) return 0;
    const size_t n = ((H + kHG - 1) / kHG) * kHG + 64;
    h->cap_hyp = 0;
    int rc;
    if ((rc = grow(h, h->d_hyp, n * kHypStride)) || (rc = grow(h, h->d_hist, n * kBins)) || (rc = grow(h, h->d_st, n)) ||
        (rc = grow(h, h->d_st_b, n)) || (rc = grow(h, h->d_cand_n, n)) || (rc = grow(h, h->d_med, n)) || (rc = grow(h, h->d_score, n)) || (rc = grow(h, h->d_tri, 3 * n)) ||
        (rc = grow(h, h->d_ndeg, (size_t)1))) return rc;
    if (h->cap > 0 && (rc = grow(h, h->d_cand, n * (size_t)h->cap))) return rc;
    POST_CUDA(cudaMemsetAsync(h->d_hist, 0, n * kBins * sizeof(unsigned), h->stream));
    POST_CUDA(cudaMemsetAsync(h->d_hyp, 0, n * kHypStride * sizeof(double), h->stream));
    h->cap_hyp = n;
    return 0;
}

int reserve_partial(b200post_handle* h, size_t n)
{
    if (n <= h->cap_partial) return 0;
    h->cap_partial = 0;
    int rc = grow(h, h->d_partial, n);
    if (rc) return rc;
    h->cap_partial = n;
    return 0;
}

struct Grid { int groups, nx, nx1; };
Grid grid_for(const b200post_handle* h, int H)
{
    Grid g;
    g.groups = (H + kHG - 1) / kHG;
    // one wave that fills (not overfills) the POST_MINB resident CTAs per SM: the sweeps are grid-stride, a few CTAs too
    // many would cost a whole second wave.  (ncu, 2 CTAs / SM: 31-41 % of the stall samples waited for the point loads.)
    g.nx = std::max(1, (POST_MINB * h->sms) / g.groups);
    g.nx1 = 2 * h->sms;                                            // single-hypothesis sweeps
    return g;
}

int launch_check(b200post_handle* h, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string(what) + ": " + cudaGetErrorString(e); return B200BA_E_CUDA; }
    ++h->launches;
    return 0;
}

void launch_hist(b200post_handle* h, const Grid& g, int pass, int check_finite)
{
    static const int shifts[6] = {52, 41, 30, 19, 8, 0};
    const int shift = shifts[pass], bits = pass == 5 ? 8 : kBits;
    const dim3 grid(g.nx, g.groups);
    const size_t smem = kHG * kBins * sizeof(unsigned);
    if (pass == 0)
        k_post_hist<true><<<grid, kSweepThreads, smem, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_st, shift, shift + bits,
                                                                   (1u << bits) - 1u, h->d_hist, h->d_res, check_finite);
    else
        k_post_hist<false><<<grid, kSweepThreads, smem, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_st, shift, shift + bits,
                                                                    (1u << bits) - 1u, h->d_hist, h->d_res, check_finite);
}

// Median of every hypothesis' distances -> sigma2 in d_hyp.  radix_only: all six passes.
int run_select(b200post_handle* h, const Grid& g, bool radix_only)
{
    int rc;
    const unsigned M = (unsigned)h->M;
    k_post_init<<<(std::max(h->H, kResDoubles) + 127) / 128, 128, 0, h->stream>>>(h->d_st, h->d_cand_n, h->d_hyp, h->H, M / 2u, M, h->d_res);
    if ((rc = launch_check(h, "k_post_init"))) return rc;
    const int n_pass = radix_only ? 6 : 2;
    for (int p = 0; p < n_pass; ++p) {
        launch_hist(h, g, p, p == 0);
        if ((rc = launch_check(h, "k_post_hist"))) return rc;
        k_post_scan<<<h->H, 1024, 0, h->stream>>>(h->d_hist, h->d_st, p == 5 ? 8 : kBits);
        if ((rc = launch_check(h, "k_post_scan"))) return rc;
    }
    if (!radix_only) {
        POST_CUDA(cudaMemcpyAsync(h->d_st_b, h->d_st, h->H * sizeof(SelState), cudaMemcpyDeviceToDevice, h->stream));
        k_post_gather<<<dim3(g.nx, g.groups), kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_st, 41, h->d_cand, h->d_cand_n, h->cap);
        if ((rc = launch_check(h, "k_post_gather"))) return rc;
        k_post_select<<<h->H, 1024, (size_t)h->cap * sizeof(unsigned long long), h->stream>>>(h->d_cand, h->d_cand_n, h->d_st, h->cap, h->d_med, h->d_res);
        if ((rc = launch_check(h, "k_post_select"))) return rc;
    }
    const double den = (double)((size_t)h->M * 2 - 6);
    k_post_sigma<<<(h->H + 127) / 128, 128, 0, h->stream>>>(h->d_hyp, h->d_st, h->d_med, radix_only ? 1 : 0, h->H, den);
    return launch_check(h, "k_post_sigma");
}

int run_score_and_moments(b200post_handle* h, const Grid& g)
{
    int rc;
    k_post_score<<<dim3(g.nx, g.groups), kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_partial);
    if ((rc = launch_check(h, "k_post_score"))) return rc;
    k_post_pick<<<1, 256, (size_t)h->H * sizeof(double), h->stream>>>(h->d_partial, g.nx, h->H, h->d_hyp, h->d_score, h->d_res);
    if ((rc = launch_check(h, "k_post_pick"))) return rc;
    k_post_mask<<<g.nx1, kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->d_res, h->k_threshold, h->d_keep, h->d_partial);
    if ((rc = launch_check(h, "k_post_mask"))) return rc;
    k_post_mean<<<1, 32 * 5, 0, h->stream>>>(h->d_partial, g.nx1, h->d_res);
    if ((rc = launch_check(h, "k_post_mean"))) return rc;
    k_post_scatter<<<g.nx1, kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->d_res, h->d_partial);
    if ((rc = launch_check(h, "k_post_scatter"))) return rc;
    k_post_scatter_final<<<1, 32 * 6, 0, h->stream>>>(h->d_partial, g.nx1, h->d_res);
    return launch_check(h, "k_post_scatter_final");
}

using b200ba::threaded_copy;                              // pinned_pool.h

// Fit on the points in d_pts (uploaded from `pts` first unless pts == NULL: already resident, b200post_ground_stage_ba);
// points and mask stay resident.  keep_host may be NULL.
int fit_resident(b200post_handle* h, int32_t M, const double* pts, int32_t n_ransac, const int32_t* triples,
                 double k_threshold, b200post_plane* out, int32_t* keep_host, bool resident = false)
{
    if (!h) return B200BA_E_INVALID;
    if ((!pts && !resident) || !triples || !out || n_ransac <= 0) { h->err = "b200post_plane_fit: null argument or n_ransac <= 0"; return B200BA_E_INVALID; }
    if (M < 10) { h->err = "b200post_plane_fit: fewer than 10 points (the reference returns -1, src/sfm_utils.cpp:826-829)"; return B200BA_E_INVALID; }
    if (n_ransac > 4096) { h->err = "b200post_plane_fit: at most 4096 hypotheses per call (the reference's default is 100)"; return B200BA_E_INVALID; }
    POST_CUDA(cudaSetDevice(h->device));
    if (!resident) h->launches = 0;
    for (int i = 0; i < 3 * n_ransac; ++i)
        if (triples[i] < 0 || triples[i] >= M) { h->err = "b200post_plane_fit: hypothesis index out of range"; return B200BA_E_INVALID; }

    int rc;
    if ((rc = reserve_points(h, (size_t)M)) || (rc = reserve_hyp(h, (size_t)n_ransac))) return rc;
    h->M = M; h->H = n_ransac; h->k_threshold = k_threshold;
    const Grid g = grid_for(h, n_ransac);
    if ((rc = reserve_partial(h, std::max((size_t)g.nx * (size_t)(g.groups * kHG), (size_t)g.nx1 * 6)))) return rc;

    const size_t pts_bytes = (size_t)M * 3 * sizeof(double), keep_bytes = (size_t)M * sizeof(int);
    b200ba::PinnedPool::Lease stage;                           // held until this call returns (all copies synchronised by then)
    b200ba::g_pinned.lease(stage, (resident ? 0 : pts_bytes) + keep_bytes);
    char* stage_keep = stage.p + (resident ? 0 : pts_bytes);
    POST_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if (!resident) {
        threaded_copy(stage.p, pts, pts_bytes);
        POST_CUDA(cudaMemcpyAsync(h->d_pts, stage.p, pts_bytes, cudaMemcpyHostToDevice, h->stream));
    }
    POST_CUDA(cudaMemcpyAsync(h->d_tri, triples, (size_t)n_ransac * 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    POST_CUDA(cudaMemsetAsync(h->d_ndeg, 0, sizeof(unsigned), h->stream));
    POST_CUDA(cudaEventRecord(h->ev[1], h->stream));
    k_post_hypotheses<<<(n_ransac + 127) / 128, 128, 0, h->stream>>>(h->d_pts, h->d_tri, n_ransac, h->d_hyp, h->d_ndeg);
    if ((rc = launch_check(h, "k_post_hypotheses"))) return rc;
    unsigned degenerate_u = 0;
    POST_CUDA(cudaMemcpyAsync(&degenerate_u, h->d_ndeg, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    const bool radix_only = h->cap <= 0;
    if ((rc = run_select(h, g, radix_only)) || (rc = run_score_and_moments(h, g))) return rc;
    POST_CUDA(cudaEventRecord(h->ev[2], h->stream));
    double res[kResDoubles];
    POST_CUDA(cudaMemcpyAsync(res, h->d_res, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
    POST_CUDA(cudaStreamSynchronize(h->stream));
    unsigned long long flags; std::memcpy(&flags, &res[R_FLAGS], sizeof(flags));
    int fallback = 0;
    if (!radix_only && (flags & 2ull)) {            // a candidate buffer overflowed: finish by radix passes (same result)
        fallback = 1;
        if ((rc = run_select(h, g, true)) || (rc = run_score_and_moments(h, g))) return rc;
        POST_CUDA(cudaEventRecord(h->ev[2], h->stream));
        POST_CUDA(cudaMemcpyAsync(res, h->d_res, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
        POST_CUDA(cudaStreamSynchronize(h->stream));
        std::memcpy(&flags, &res[R_FLAGS], sizeof(flags));
    }
    if (keep_host) POST_CUDA(cudaMemcpyAsync(stage_keep, h->d_keep, keep_bytes, cudaMemcpyDeviceToHost, h->stream));
    POST_CUDA(cudaEventRecord(h->ev[3], h->stream));
    POST_CUDA(cudaStreamSynchronize(h->stream));
    if (keep_host) threaded_copy(keep_host, stage_keep, keep_bytes);
    const int degenerate = (int)degenerate_u;
    if (flags & 1ull) { h->err = "b200post_plane_fit: non-finite point coordinate"; return B200BA_E_INVALID; }
    if (degenerate == n_ransac) { h->err = "b200post_plane_fit: every hypothesis is degenerate"; return B200BA_E_INVALID; }
    if (res[R_BEST] < 0) { h->err = "b200post_plane_fit: no hypothesis scored below 9e99"; return B200BA_E_INVALID; }
    if (res[R_CNT] < 1) { h->err = "b200post_plane_fit: no inliers"; return B200BA_E_INVALID; }

    std::memset(out, 0, sizeof(*out));
    out->best_hypothesis = (int32_t)res[R_BEST];
    out->score = res[R_SCORE];
    out->sigma2 = res[R_SIGMA2];
    out->n_inliers = (int32_t)res[R_CNT];
    out->n_kept = (int32_t)res[R_KEPT];
    out->n_degenerate = degenerate;
    out->select_fallback = fallback;
    for (int k = 0; k < 3; ++k) out->inlier_mean[k] = res[R_IMEAN + k];
    const double* s = &res[R_SCATTER];
    const double full[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    std::memcpy(out->inlier_scatter, full, sizeof(full));
    plane_from_moments(out->inlier_mean, out->inlier_scatter, out->plane, out->plane_se3);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); out->ms_device = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); out->ms_kernels = ms;
    out->gpu_launches = h->launches;
    return B200BA_OK;
}

Se3 to_se3(const double* a) { Se3 t; std::memcpy(t.a, a, sizeof(t.a)); return t; }

int reaxis_cameras(b200post_handle* h, const double* axes, int32_t N, double* cam_RT)
{
    if (N <= 0 || !cam_RT) return 0;
    if ((size_t)N > h->cap_cams) { h->cap_cams = 0; int rc = grow(h, h->d_cams, (size_t)N * 12); if (rc) return rc; h->cap_cams = (size_t)N; }
    POST_CUDA(cudaMemcpyAsync(h->d_cams, cam_RT, (size_t)N * 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_post_reaxis_cameras<<<(N + 127) / 128, 128, 0, h->stream>>>(h->d_cams, N, to_se3(axes));
    int rc = launch_check(h, "k_post_reaxis_cameras");
    if (rc) return rc;
    POST_CUDA(cudaMemcpyAsync(cam_RT, h->d_cams, (size_t)N * 12 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

}  // namespace

extern "C" {

int b200post_create(int32_t device, int32_t select_cap, b200post_handle** out)
{
    if (!out) return B200BA_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_post_create_error = std::string("b200post_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                              "); the post-BA stage has no CPU fallback";
        return B200BA_E_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { g_post_create_error = "b200post_create: device ordinal out of range"; return B200BA_E_INVALID; }
    b200post_handle* h = new b200post_handle();
    h->device = device;
    h->cap = select_cap == 0 ? 4096 : (select_cap < 0 ? 0 : std::min(select_cap, 6000));   // 6000 keys = 48 KB of shared memory
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_res, kResDoubles * sizeof(double)) == cudaSuccess;
    ok = ok && cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_post_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kHG * kBins * sizeof(unsigned))) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_post_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kHG * kBins * sizeof(unsigned))) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_post_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 6000 * (int)sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) {
        g_post_create_error = std::string("b200post_create: ") + cudaGetErrorString(cudaGetLastError());
        b200post_destroy(h);
        return B200BA_E_CUDA;
    }
    *out = h;
    return B200BA_OK;
}

void b200post_destroy(b200post_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* bufs[] = {h->d_pts, h->d_out, h->d_keep, h->d_pos, h->d_hyp, h->d_hist, h->d_st, h->d_st_b, h->d_cand, h->d_cand_n,
                    h->d_med, h->d_score, h->d_tri, h->d_ndeg, h->d_partial, h->d_res, h->d_cams, h->d_scan_tmp};
    for (void* p : bufs) if (p) cudaFree(p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* b200post_last_error(const b200post_handle* h) { return h ? h->err.c_str() : g_post_create_error.c_str(); }

int b200post_draw_triples(int32_t M, int32_t n_ransac, int32_t* triples)
{
    if (M < 3 || n_ransac < 0 || !triples) return B200BA_E_INVALID;
    for (int i = 0; i < n_ransac; ++i) {
        const int a = rand() % M;
        int b = a, c = a;
        while (b == a) b = rand() % M;
        while (c == a || c == b) c = rand() % M;
        triples[3 * i] = a; triples[3 * i + 1] = b; triples[3 * i + 2] = c;
    }
    return B200BA_OK;
}

int b200post_plane_fit(b200post_handle* h, int32_t M, const double* pts, int32_t n_ransac, const int32_t* triples,
                       double k_threshold, b200post_plane* out, int32_t* keep)
{
    return fit_resident(h, M, pts, n_ransac, triples, k_threshold, out, keep);
}

int b200post_ground_axes(const double* cam, const double* plane4, const double* plane_se3, double* axes)
{
    if (!cam || !plane4 || !axes) return B200BA_E_INVALID;
    (void)plane_se3;                                     // the reference computes p_mean from it and never uses it (:1108)
    double centre[3], foot[3], tip[3], tip_foot[3], along[3];
    camera_centre(cam, centre);
    foot_on_plane(plane4, centre, foot);
    for (int k = 0; k < 3; ++k) tip[k] = centre[k] + cam[k];          // centre + camera x axis (first row of R)
    foot_on_plane(plane4, tip, tip_foot);
    for (int k = 0; k < 3; ++k) along[k] = tip_foot[k] - foot[k];
    double s_along = 0, s_norm = 0;
    for (int k = 0; k < 3; ++k) { s_along += along[k] * along[k]; s_norm += plane4[k] * plane4[k]; }
    const double inv_along = 1. / std::sqrt(s_along), inv_norm = 1. / std::sqrt(s_norm);   // cv::Vec / double scales by 1/x
    double nx[3], ny[3], nz[3];
    for (int k = 0; k < 3; ++k) { nx[k] = along[k] * inv_along; nz[k] = plane4[k] * inv_norm; }
    cross3(nz, nx, ny);
    const double* R[3] = {nx, ny, nz};
    for (int r = 0; r < 3; ++r) {
        double t = 0;
        for (int k = 0; k < 3; ++k) { axes[4 * r + k] = R[r][k]; t += (-R[r][k]) * foot[k]; }
        axes[4 * r + 3] = t;
    }
    return B200BA_OK;
}

int b200post_transform(b200post_handle* h, const double* axes, int32_t M, double* pts, int32_t N, double* cam_RT)
{
    if (!h) return B200BA_E_INVALID;
    if (!axes || M < 0 || N < 0) { h->err = "b200post_transform: bad argument"; return B200BA_E_INVALID; }
    POST_CUDA(cudaSetDevice(h->device));
    h->launches = 0;
    int rc;
    if (M > 0 && pts) {
        if ((rc = reserve_points(h, (size_t)M))) return rc;
        const size_t bytes = (size_t)M * 3 * sizeof(double);
        b200ba::PinnedPool::Lease stage;
        b200ba::g_pinned.lease(stage, bytes);
        threaded_copy(stage.p, pts, bytes);
        POST_CUDA(cudaMemcpyAsync(h->d_pts, stage.p, bytes, cudaMemcpyHostToDevice, h->stream));
        k_post_reaxis_points<<<(M + 255) / 256, 256, 0, h->stream>>>(h->d_pts, M, to_se3(axes), nullptr, nullptr, h->d_out);
        if ((rc = launch_check(h, "k_post_reaxis_points"))) return rc;
        POST_CUDA(cudaMemcpyAsync(stage.p, h->d_out, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->M = 0;                                           // d_pts no longer holds a fitted problem
        POST_CUDA(cudaStreamSynchronize(h->stream));
        threaded_copy(pts, stage.p, bytes);
    }
    if ((rc = reaxis_cameras(h, axes, N, cam_RT))) return rc;
    POST_CUDA(cudaStreamSynchronize(h->stream));
    return B200BA_OK;
}

// Tail shared by the two ground-stage entry points: the fit is done, points and mask are resident.
static int ground_stage_tail(b200post_handle* h, int32_t M, double* pts_out, int32_t N, double* cam_RT,
                             b200post_plane* out, int32_t* m_out, double* axes_out)
{
    int rc;
    double axes[12];
    if ((rc = b200post_ground_axes(cam_RT, out->plane, out->plane_se3, axes))) return rc;
    for (double v : axes) if (!std::isfinite(v)) { h->err = "b200post_ground_stage: ground axes are not finite (plane through the camera axis?)"; return B200BA_E_INVALID; }
    if (axes_out) std::memcpy(axes_out, axes, sizeof(axes));
    POST_CUDA(cudaEventRecord(h->ev[0], h->stream));
    size_t need = 0;
    POST_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, h->d_keep, h->d_pos, M, h->stream));
    if (need > h->cap_scan_tmp) {
        if (h->d_scan_tmp) cudaFree(h->d_scan_tmp);
        h->d_scan_tmp = nullptr;
        POST_CUDA(cudaMalloc(&h->d_scan_tmp, need));
        h->cap_scan_tmp = need;
    }
    POST_CUDA(cub::DeviceScan::ExclusiveSum(h->d_scan_tmp, need, h->d_keep, h->d_pos, M, h->stream));
    h->launches += 2;                                        // cub: DeviceScanInitKernel + DeviceScanKernel
    k_post_reaxis_points<<<(M + 255) / 256, 256, 0, h->stream>>>(h->d_pts, M, to_se3(axes), h->d_keep, h->d_pos, h->d_out);
    if ((rc = launch_check(h, "k_post_reaxis_points"))) return rc;
    const size_t out_bytes = (size_t)out->n_kept * 3 * sizeof(double);
    b200ba::PinnedPool::Lease stage;
    b200ba::g_pinned.lease(stage, out_bytes);
    POST_CUDA(cudaMemcpyAsync(stage.p, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = reaxis_cameras(h, axes, N, cam_RT))) return rc;
    POST_CUDA(cudaEventRecord(h->ev[1], h->stream));
    POST_CUDA(cudaStreamSynchronize(h->stream));
    threaded_copy(pts_out, stage.p, out_bytes);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    out->ms_device += ms;
    out->gpu_launches = h->launches;
    *m_out = out->n_kept;
    return B200BA_OK;
}

int b200post_ground_stage(b200post_handle* h, int32_t M, double* pts, int32_t N, double* cam_RT,
                          int32_t n_ransac, const int32_t* triples, double k_threshold,
                          b200post_plane* out, int32_t* keep, int32_t* m_out, double* axes_out)
{
    if (!h) return B200BA_E_INVALID;
    if (!cam_RT || N < 1 || !m_out) { h->err = "b200post_ground_stage: needs at least one camera and m_out"; return B200BA_E_INVALID; }
    int rc = fit_resident(h, M, pts, n_ransac, triples, k_threshold, out, keep);
    if (rc) return rc;
    return ground_stage_tail(h, M, pts, N, cam_RT, out, m_out, axes_out);
}

int b200post_ground_stage_ba(b200post_handle* h, b200ba_handle* ba, int32_t n_ransac, const int32_t* triples, double k_threshold,
                             b200post_plane* out, int32_t* keep, int32_t* m_out, double* pts_out, double* cam_RT,
                             double* axes_out)
{
    if (!h) return B200BA_E_INVALID;
    if (!ba || !pts_out || !cam_RT || !m_out) { h->err = "b200post_ground_stage_ba: null argument"; return B200BA_E_INVALID; }
    int32_t N = 0, M = 0;
    if (b200ba_problem_size(ba, &N, &M, nullptr) != B200BA_OK || N < 1) { h->err = "b200post_ground_stage_ba: the BA handle holds no problem"; return B200BA_E_STATE; }
    if (M < 10) { h->err = "b200post_plane_fit: fewer than 10 points (the reference returns -1, src/sfm_utils.cpp:826-829)"; return B200BA_E_INVALID; }
    POST_CUDA(cudaSetDevice(h->device));
    h->launches = 0;
    int rc;
    if ((rc = reserve_points(h, (size_t)M))) return rc;          // before the export: growing frees the old buffer
    POST_CUDA(cudaStreamSynchronize(h->stream));
    if (b200ba_get_result_device(ba, cam_RT, h->d_pts) != B200BA_OK) {   // must live on the same device as this handle
        h->err = std::string("b200post_ground_stage_ba: ") + b200ba_last_error(ba);
        return B200BA_E_STATE;
    }
    ++h->launches;                                               // k_export_points
    if ((rc = fit_resident(h, M, nullptr, n_ransac, triples, k_threshold, out, keep, true))) return rc;
    return ground_stage_tail(h, M, pts_out, N, cam_RT, out, m_out, axes_out);
}

int b200post_debug_time_kernel(b200post_handle* h, int32_t which, int32_t reps, double* ms_per_launch)
{
    if (!h || !ms_per_launch || reps < 1) return B200BA_E_INVALID;
    if (h->M < 10 || h->H < 1) { h->err = "b200post_debug_time_kernel: no fitted problem resident"; return B200BA_E_STATE; }
    if (which < 0 || which > 3 || (which == 1 && h->cap <= 0)) { h->err = "b200post_debug_time_kernel: no such kernel"; return B200BA_E_INVALID; }
    POST_CUDA(cudaSetDevice(h->device));
    const Grid g = grid_for(h, h->H);
    double total = 0.0;
    for (int r = 0; r < reps + 1; ++r) {                 // first repetition is a warm-up
        if (which == 0) {                                // exponent pass from the initial state (highest bin contention)
            k_post_init<<<(std::max(h->H, kResDoubles) + 127) / 128, 128, 0, h->stream>>>(h->d_st, h->d_cand_n, h->d_hyp, h->H, (unsigned)h->M / 2u, (unsigned)h->M, h->d_partial);
            POST_CUDA(cudaMemsetAsync(h->d_hist, 0, (size_t)h->H * kBins * sizeof(unsigned), h->stream));
        }
        if (which == 1) POST_CUDA(cudaMemsetAsync(h->d_cand_n, 0, (size_t)h->H * sizeof(unsigned), h->stream));
        POST_CUDA(cudaEventRecord(h->ev[0], h->stream));
        if (which == 0) launch_hist(h, g, 0, 0);
        if (which == 1) k_post_gather<<<dim3(g.nx, g.groups), kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_st_b, 41, h->d_cand, h->d_cand_n, h->cap);
        if (which == 2) k_post_score<<<dim3(g.nx, g.groups), kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->H, h->d_hyp, h->d_partial);
        if (which == 3) k_post_mask<<<g.nx1, kSweepThreads, 0, h->stream>>>(h->d_pts, h->M, h->d_res, h->k_threshold, h->d_keep, h->d_partial);
        POST_CUDA(cudaEventRecord(h->ev[1], h->stream));
        POST_CUDA(cudaStreamSynchronize(h->stream));
        int rc = launch_check(h, "b200post_debug_time_kernel");
        if (rc) return rc;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
        if (r > 0) total += ms;
    }
    if (which == 0) POST_CUDA(cudaMemsetAsync(h->d_hist, 0, (size_t)h->H * kBins * sizeof(unsigned), h->stream));
    POST_CUDA(cudaStreamSynchronize(h->stream));
    *ms_per_launch = total / reps;
    return B200BA_OK;
}

}  // extern "C"
