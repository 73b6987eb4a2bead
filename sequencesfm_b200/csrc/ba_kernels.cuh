// ba_kernels.cuh -- sm_100a device code of the B200 bundle-adjustment engine (fp64).
//
// Cost model and LM rules are those of the reference's live path (SSBA-3.0 as driven by
// src/sfm_ba_ssba.cpp); every kernel cites the reference lines whose arithmetic it carries out.  The
// organisation is NOT the reference's: nothing per-observation is stored except one weight; Jacobian
// blocks are recomputed in registers in every pass; the damped normal equations are never assembled --
// the reduced camera system S = (U + lambda I) - W (V + lambda I)^-1 W^t is applied matrix-free in two
// sweeps (by point, then by camera) and solved by block-Jacobi PCG; all reductions are segmented
// (thread-per-point over point-sorted observations, warp-per-row-range over camera-sorted observations with a
// fixed-order second level) so there is no atomic on any accumulator and results are bit-reproducible.
//
// Memory layout (HBM, all fp64 unless noted):
//   cameras   rt[2]   N x 12   row-major [R|T] (96 B = 3 sectors), current / trial copy selected by st->cur
//             U, Minv N x 36   full symmetric 6x6 blocks;  gc, x, r, z, p  N x 6
//   points    X[2]    M x 4    xyz + pad (32 B = 1 sector so a by-camera gather costs one sector)
//             V, Vinv M x 6    symmetric 3x3 (xx xy xz yy yz zz);  gp, v  M x 4
//   obs (point order)   SELL-32: points are sorted by track length (descending) and grouped in slices of 32; slot s of
//                       lane l of slice g is entry slice_base[g] + 32 s + l.  One thread owns one point, every lane of a
//                       warp runs the same trip count.  32 consecutive entries form one 384-byte ROW RECORD
//                       [32 x camera index (int32, -1 = padding) | 32 x IRLS weight (f64)] (o_rec) so that the CG by-point
//                       sweep stages a row with a single TMA bulk copy; o_meas (double2, normalised) is a flat array.
//   obs (camera order)  ROW RECORDS: every camera's observations are padded to a multiple of 32 and cut into rows of 32;
//                       one row is ONE contiguous 1168-byte record  [cam, segment | 32 x point index (int32, -1 = padding) |
//                       32 x weight | 32 x (R X).x | .y | .z]  so that a single TMA bulk copy stages everything the CG
//                       by-camera sweep streams for 32 observations, and the linearisation writes weight / cached R X back
//                       with coalesced 256-byte stores.  c_meas (double2, same row order) is only read by the linearisation.
//                       The rows are split evenly over n_vwarps "virtual warps"; a SEGMENT is a maximal run of rows of one
//                       camera inside one virtual warp's range and owns one partial-sum slot (seg_part), so the second level
//                       of every by-camera reduction is a fixed-order sum over cam_seg_beg[i] .. cam_seg_beg[i+1].
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b200ba {

constexpr int PT_BLOCK     = 128;   // threads per block in thread-per-point kernels
constexpr int CH_BLOCK     = 128;   // threads per block in the by-camera row sweeps of the linearisation (4 virtual warps)
constexpr int PART_STRIDE  = 28;    // doubles per segment partial: 21 (sym 6x6) + 6 (+1 pad)
// camera-order row record (bytes)
constexpr int REC_HDR      = 0;                 // int32 cam, int32 segment, 2 x pad
constexpr int REC_J        = 16;                // 32 x int32 point index (new order), -1 = padding
constexpr int REC_W        = REC_J + 128;       // 32 x f64 IRLS weight
constexpr int REC_R0       = REC_W + 256;       // 32 x f64 (R X).x, .y, .z cached at linearisation
constexpr int REC_R1       = REC_R0 + 256;
constexpr int REC_R2       = REC_R1 + 256;
constexpr int REC_BYTES    = REC_R2 + 256;      // 1168 = 73 x 16
constexpr int OREC_BYTES   = 128 + 256;         // point-order row record: 32 x int32 camera index | 32 x f64 weight
#ifndef B200BA_CW
#define B200BA_CW 16
#endif
#ifndef B200BA_CG
#define B200BA_CG 2
#endif
#ifndef B200BA_CNS
#define B200BA_CNS 4
#endif
constexpr int CW  = B200BA_CW;      // warps per CTA of the persistent CG by-camera sweep
#ifndef B200BA_VSPLIT
#define B200BA_VSPLIT 1
#endif
constexpr int VSPLIT = B200BA_VSPLIT;   // virtual warps per sweep warp.  2 (finer units for the once-per-iteration by-camera kernels: 2.67 waves of
                                        // fat warps instead of 1.33) took linearize_cameras from 156 to 136 us but the extra segments cost the
                                        // camera-side step as much over 50 PCG steps: 1 stays
constexpr int CG  = B200BA_CG;      // rows per stage = rows whose arithmetic is interleaved (ILP) = gathers per register set
constexpr int CNS = B200BA_CNS;     // stages staged in shared memory per warp (TMA ring, one mbarrier each)
static_assert(CNS >= 3, "two stages are in use (gather front, arithmetic) while a third is being refilled");
constexpr int STAGE_BYTES  = CG * REC_BYTES;
constexpr int CAMPASS_SMEM = CW * CNS * STAGE_BYTES + CW * CNS * 8;
constexpr int ONE_CTA      = 1024;  // threads of the single-CTA camera-side kernels
constexpr int TRACE_COLS   = 8;
#ifndef B200BA_WHATIF
#define B200BA_WHATIF 0      // bit mask (1, 2, 4, 8, 16, 32): timing experiments on the by-point PCG sweep (tools/gpu_r2d.sh whatif), results are WRONG with them
#endif
constexpr double QUAT_SCALE = 1.4142135623730951;   // the camera tables hold sqrt(2) x the unit quaternion: the sandwich's factor 2 comes with the components
constexpr int TAB_STRIDE   = 14;    // doubles per camera in the packed table: quaternion 4, T 3, p 6, pad 1 (112 B = 7 x 16 B)
#ifndef B200BA_TAB_BLOCK
#define B200BA_TAB_BLOCK 608
#ifndef B200BA_TAB_MAXNREG
#define B200BA_TAB_MAXNREG 104
#endif
#endif
constexpr int TAB_BLOCK    = B200BA_TAB_BLOCK;   // threads of the persistent shared-memory-table CG point sweep (1 CTA / SM; 640 x 96 registers measured faster than 768 x 80, which spills)
constexpr int STEP_BLOCK   = 192;   // threads of the cooperative camera-side CG kernel: 32 cameras x 6 components
constexpr int MAX_SMEM     = 227 * 1024;

struct Intrinsics {
    double k00, k01, k02, k11, k12;   // unit-focal-length intrinsics (src/sfm_ba_ssba.cpp:98-102)
    double huber;                     // 2 / |f0| (src/sfm_ba_ssba.cpp:193); <= 0: robustifier off
    double f0;
    double ar;                        // k11 / k00 (aspect ratio), precomputed so that kernels read it from the constant bank
};

// Device-resident LM / CG control block: all accept/reject and stop decisions are taken on the device
// (reference keeps them in host scalars, v3d_optimization.cpp:879-978).
struct LMState {
    double lambda, err, new_err, rho;
    double rz, rz0, cg_rel;
    double d2c, denomc;
    double param_length;
    double tau, grad_thr, upd_thr, cg_tol;
    int iter, status, done, compute, cur, first;
    int cg_active, cg_steps, solve_ok, total_cg;
    int max_iter, min_iter, n_trace, trace_cap;
    int n_accept;                                 // accepted LM steps since b200ba_begin / b200ba_restart
    // PBA's LM variant (options.lm_variant = 1; SparseBundleCPU.cpp:3772-3949): multiplicative damping, its own step control
    int variant, cg_min_steps;
    double damp_adjust;                           // growth factor of the damping after a rejected step (2, 4, 8, ...)
    double cg_guard;                              // PCG aborts when sqrt(r.z / r0.z0) exceeds it (__cg_norm_guard)
    double mse_scale, mse_thr;                    // cost * mse_scale = mean squared error in px^2; stop at <= mse_thr (__lm_mse_threshold)
    double delta_thr, norm_scale;                 // stop when norm_scale * sqrt(sum delta^2 D0) <= delta_thr (__lm_delta_threshold)
    double njx;                                   // |J delta|^2 (camera + point parts arrive through ar_dec[3])
};

// NVLink peer-memory exchange window (multi-rank), PUSH design.  Every rank owns one window
//   [ flag[2][world] (uint64, 64 bytes apart) | data[2][world][cap] (f64) ]
// allocated with cudaMalloc and mapped into every peer through CUDA IPC.  Collective number c (the same on every rank: all
// ranks run the same kernel sequence and take the same device-side decisions) uses half c & 1.  A rank STORES its
// contribution into slot [c & 1][own rank] of EVERY window (its own included) - posted NVLink writes - and then raises
// flag [c & 1][own rank] = c + 1 in every window (release, system scope).  A receiver polls and reads only ITS OWN memory:
// when all world flags of the half have arrived it sums the world slots in rank order - identical bits on every rank, one
// one-way NVLink trip, no host call.  (The first version pulled: remote flag polls + remote loads = two round trips,
// measured 16 us per PCG step at 2 GPUs.)  Half c & 1 is rewritten at collective c + 2, which a rank only starts after it
// has seen every flag of c + 1, i.e. after every peer has finished reading c.
constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_MAX_BLOCKS = 511;                              // per-block flags of the fused PCG-step exchange (6 N <= 511 x 192)
constexpr int P2P_HDR_BYTES = 2 * P2P_MAX_WORLD * (P2P_MAX_BLOCKS + 1) * 8;
struct P2P {
    int on, world, rank, pad;
    unsigned long long cap;                  // doubles per slot
    unsigned char* win[P2P_MAX_WORLD];       // win[rank] = own window, others = peer mappings
    unsigned long long* seq;                 // device counter of collectives executed (own memory)
    unsigned int* done_blocks;               // last-block-done counter of k_p2p_reduce
    int* err;                                // set when a peer's flag did not arrive (spin limit)
};
// explicit reduced camera system of small problems (ba_dense.cuh, dense_host.h); all null / 0 when the plan is matrix-free
struct DenseDev {
    int on, rows, ld, n_chunks, n_blocks, pad;
    const int2* pairs; const int4* chunks; const int* blk_chunk_beg; const int* blk_ik;
    const int* entry_pt;                  // SELL entry -> point (new order)
    double *Wobs, *Yobs, *Tobs;           // per SELL entry: W, W (V + damping)^-1 (6 rows of [3 | pad] = 24 doubles each), W (V + damping)^-1 g_p (6)
    double *chunk_part, *chunk_rhs;       // per chunk of pairs: 36 + 6 partial sums
    double* S;                            // 6N x 6N, both triangles
    double *zg, *pq_slots, *rz_slots;     // exchange buffers of the grid-wide variant of k_dense_pcg
    unsigned int* bar_counter;            // its counting barrier (zeroed by k_dense_finalize)
};
struct Dev {
    P2P p2p;
    DenseDev dn;
    int N, M, K, n_rows, n_vwarps, n_segs, first_free, n_pt_blocks, world, rank, n_slices, tab_ok;
    Intrinsics in;
    LMState* st;
    double* rt[2];
    double *U, *gc, *Minv, *x, *r, *z, *p, *q;
    double* ar_lin;     // [N x 28 | cost | ginf slots(world) | vmax slots(world)]  all-reduced after linearisation
    double* ar_q;       // N x 6  reduced sum_k W_k v_j, all-reduced every CG step
    double* ar_sd;      // N x 21 Schur-diagonal correction, all-reduced
    double* ar_dec;     // [d2p, denomp, new_cost, pad]
    double* X[2];
    double *V, *gp, *Vinv, *v;
    unsigned char* o_rec;         // (padded entries / 32) x OREC_BYTES point-order row records [32 x camera index | 32 x weight]
    double2* o_meas; int *slice_base, *slice_deg;
    double* ptab;       // N x 14 packed camera table [quat 4 | T 3 | p 6 | pad] staged into shared memory by the CG point sweep
    double *cam_D0, *pt_D0;   // PBA variant: diag(J^t J) at the first linearisation = the Jacobian column scaling (6 N, 4 M)
    double* fin_part;   // 4 x blocks of k_cg_finish: camera part of |delta|^2, delta . g, non-finite flag
    double* obs_B;      // large-camera-count plan: per SELL row of 32 entries, the weighted point Jacobian B~ = w dp/dXX R (2 x 3) as 6 x 32 doubles
    double* wtab;       // N x 8: the PCG direction in WORLD axes [R^t p_t | R^t p_w | pad] for k_cg_point_pass_wB
    double* ctab;       // (N + 1) x 8 compact table [quat 4 | T 3 | pad] of the CURRENT cameras (layout of qtab) + the padding entries' dummy camera:
                        // the static half of the PCG by-point sweep's shared-memory table (k_cg_point_pass_tabc), written once per linearisation
    double* qtab;       // N x 8 compact table [quat 4 | T 3 | pad] (swizzled, see qtab_unit) of the trial cameras for k_cost_tab
    unsigned char* c_rec;         // n_rows x REC_BYTES camera-order row records (see REC_*)
    double2* c_meas;              // n_rows x 32 normalised measurements in row order
    // optional fp32 copies for the mixed-precision PCG (ba_kernels_f32.cuh); all null when options.pcg_fp32 = 0
    unsigned char *o_rec32, *c_rec32; float4 *X32, *v32; float *Vinv32, *ptab32, *rt32; int* tab_wr32; int tab_warps32;
    int* tab_wr;                  // tab_warps 32-byte headers: contiguous slice ranges of the persistent CG point sweep (one per warp, or per chunk)
    int tab_warps;

    int* cam_seg_beg;     // N + 1: segments of camera i
    double* seg_part;     // n_segs x PART_STRIDE
    double* pt_part;      // n_pt_blocks x 4
    double* trace;
    unsigned long long* step_bar;  // counting barrier of k_cg_camera_step: [arrivals, launches completed]
    unsigned long long* stamps;   // profiling builds only (-DB200BA_STAMPS): per-CTA / per-warp clock stamps of the two PCG sweeps
};

// ------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// deterministic block reductions (fixed butterfly inside a warp, then the same butterfly over the warps' partials: a serial
// sum over 32 partials is 32 dependent fp64 adds, ~1500 cycles on this chip); result valid in all threads
template <int NT> __device__ __forceinline__ double block_sum(double v, double* sm)
{
    static_assert(NT % 32 == 0 && NT <= 1024, "block size");
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    return warp_sum(l < NT / 32 ? sm[l] : 0.0);
}
template <int NT> __device__ __forceinline__ double block_max(double v, double* sm)
{
    static_assert(NT % 32 == 0 && NT <= 1024, "block size");
    v = warp_max(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    return warp_max(sm[l < NT / 32 ? l : 0]);
}

// 1/z to within an ulp without the IEEE-division slow path (~20 instructions with a branch): MUFU.RCP64H seed
// (rcp.approx.ftz.f64, 2^-23) + two Newton steps.  z is a camera-space depth: finite, far from 0 and from overflow.
__device__ __forceinline__ double fast_rcp(double z)
{
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
    double e = fma(-z, r, 1.0); r = fma(r, e, r);
    e = fma(-z, r, 1.0); r = fma(r, e, r);
    return r;
}
struct Cam { double R[9], T[3]; };
__device__ __forceinline__ void load_cam(const double* __restrict__ rt, int i, Cam& c)
{
    const double2* s = reinterpret_cast<const double2*>(rt + 12 * (size_t)i);
    double2 a = __ldg(s), b = __ldg(s + 1), d = __ldg(s + 2), e = __ldg(s + 3), f = __ldg(s + 4), g = __ldg(s + 5);
    c.R[0] = a.x; c.R[1] = a.y; c.R[2] = b.x; c.T[0] = b.y;
    c.R[3] = d.x; c.R[4] = d.y; c.R[5] = e.x; c.T[1] = e.y;
    c.R[6] = f.x; c.R[7] = f.y; c.R[8] = g.x; c.T[2] = g.y;
}
__device__ __forceinline__ void load3p(const double* __restrict__ base, int j, double X[3])   // padded 4-double record
{
    const double2* s = reinterpret_cast<const double2*>(base + 4 * (size_t)j);
    double2 a = __ldg(s), b = __ldg(s + 1);
    X[0] = a.x; X[1] = a.y; X[2] = b.x;
}
__device__ __forceinline__ void load6(const double* __restrict__ base, int i, double v[6])
{
    const double2* s = reinterpret_cast<const double2*>(base + 6 * (size_t)i);
    double2 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}

// 256-bit gather of one padded 4-double record (LDG.E.256 on sm_100a: one L1 wavefront per lane instead of two)
__device__ __forceinline__ void load3p_256(const double* __restrict__ base, int j, double X[3])
{
    double pad;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(X[0]), "=d"(X[1]), "=d"(X[2]), "=d"(pad) : "l"(base + 4 * (size_t)j));
    (void)pad;
}
// L2 residency control.  Per PCG step ~260 MB stream through a 126 MB L2 while the point-sized vector that is GATHERED
// (v: 32 MB) must stay resident, otherwise every gather becomes a random 32-byte DRAM access.  Streams are loaded
// evict_first / no L1 allocation; gathered arrays are loaded and stored evict_last.
__device__ __forceinline__ uint64_t policy_evict_last()  { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void load3p_256_keep(const double* __restrict__ base, int j, double X[3], uint64_t pol)
{
    double pad;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(X[0]), "=d"(X[1]), "=d"(X[2]), "=d"(pad) : "l"(base + 4 * (size_t)j), "l"(pol));
    (void)pad;
}
__device__ __forceinline__ void store4_keep(double* p, double a, double b, double c, double e, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(e), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
// ---- mbarrier / TMA bulk-copy helpers (sm_90+ PTX; the only async-copy engine used here is cp.async.bulk) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// The spin loop is written in C++ (not as branches inside the asm block) so that the compiler sees it and places a
// reconvergence point behind it: lanes can leave a try_wait loop in different trips, and with asm-internal branches the
// warp stayed split for the rest of the kernel (ncu: every instruction of the sweep executed twice).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
    __syncwarp();
}

#ifdef B200BA_STAMPS
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int STAMP_STRIDE = 40;   // per CTA: [0] entry (globaltimer) [1] table ready [2] entry clock64 [3] ready clock64 [8 + w] end clock64 of warp w
#endif
// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor is still running; everything it does before pdl_wait() must not depend on the predecessor's output.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// First statement of every kernel of the small-problem LM iteration (options.reduced_system explicit plan): ~20 tiny kernels
// that depend on each other, i.e. ~20 launch gaps per iteration.  Launched with the programmatic-stream-serialization
// attribute, a kernel lets its successor be scheduled at once (the successor's blocks become resident and wait here) and then
// waits until its predecessor has COMPLETED and flushed - the same ordering as plain stream order, minus the launch latency.
// Without the attribute both instructions are no-ops.  Nothing may be read before this call (LMState included).
__device__ __forceinline__ void pdl_chain_enter() { pdl_trigger(); pdl_wait(); }
// SELL-32 addressing: first entry and (warp-uniform) trip count of point j
__device__ __forceinline__ void slice_of(const Dev& d, int j, int& k0, int& deg)
{
    int g = j >> 5;
    if (g < d.n_slices) { k0 = __ldg(d.slice_base + g) + (j & 31); deg = __ldg(d.slice_deg + g); }
    else { k0 = 0; deg = 0; }
}

// point-order row records: camera index / weight of entry k
__device__ __forceinline__ const int* o_cam_ptr(const Dev& d, int k) { return reinterpret_cast<const int*>(d.o_rec + (size_t)(k >> 5) * OREC_BYTES) + (k & 31); }
__device__ __forceinline__ double* o_w_ptr(const Dev& d, int k) { return reinterpret_cast<double*>(d.o_rec + (size_t)(k >> 5) * OREC_BYTES + 128) + (k & 31); }

// Per-observation geometry: camera-space point, residual, Huber weight, weighted projection Jacobian.
//   residual  e = K^ (XX / z) - m                         v3d_metricbundle.h:185-209
//   weight    w = 1 | sqrt(thr / |e|)                     v3d_metricbundle.h:49-57
//   dp/dXX    [[f/z, 0, -f x/z^2], [0, ar f/z, -ar f y/z^2]]   v3d_metricbundle.cpp:126-141 (skew ignored, as there)
// The weighted 2x3 matrix is kept as (a, b, c, d):  [[a, 0, b], [0, c, d]].
struct Obs {
    double e0, e1, w;          // unweighted residual, weight
    double a, b, c, d;         // w * dp/dXX
    double r0, r1, r2;         // R X  (= XX - T), the lever arm of the rotation derivative
};
__device__ __forceinline__ void obs_residual(const Cam& cm, const double X[3], double2 m, const Intrinsics& in,
                                             double& e0, double& e1, double& x, double& y, double& iz,
                                             double& r0, double& r1, double& r2)
{
    r0 = cm.R[0] * X[0] + cm.R[1] * X[1] + cm.R[2] * X[2];
    r1 = cm.R[3] * X[0] + cm.R[4] * X[1] + cm.R[5] * X[2];
    r2 = cm.R[6] * X[0] + cm.R[7] * X[1] + cm.R[8] * X[2];
    x = r0 + cm.T[0]; y = r1 + cm.T[1];
    double z = r2 + cm.T[2];
    iz = fast_rcp(z);
    double u = x * iz, v = y * iz;
    e0 = in.k00 * u + in.k01 * v + in.k02 - m.x;
    e1 = in.k11 * v + in.k12 - m.y;
}
__device__ __forceinline__ double huber_weight(double e0, double e1, double thr)
{
    if (thr <= 0.0) return 1.0;
    double n = sqrt(e0 * e0 + e1 * e1);
    return (n < thr) ? 1.0 : sqrt(thr / n);
}
// full geometry with the weight recomputed from the residual
__device__ __forceinline__ void obs_full(const Cam& cm, const double X[3], double2 m, const Intrinsics& in, Obs& o)
{
    double x, y, iz;
    obs_residual(cm, X, m, in, o.e0, o.e1, x, y, iz, o.r0, o.r1, o.r2);
    o.w = huber_weight(o.e0, o.e1, in.huber);
    double f = in.k00, ar = in.ar;
    double wf = o.w * f * iz;
    o.a = wf; o.b = -wf * x * iz; o.c = ar * wf; o.d = -ar * wf * y * iz;
}
// geometry for the CG sweeps: weight given (stored at linearisation), no measurement needed
__device__ __forceinline__ void obs_jac(const Cam& cm, const double X[3], double w, const Intrinsics& in, Obs& o)
{
    o.r0 = cm.R[0] * X[0] + cm.R[1] * X[1] + cm.R[2] * X[2];
    o.r1 = cm.R[3] * X[0] + cm.R[4] * X[1] + cm.R[5] * X[2];
    o.r2 = cm.R[6] * X[0] + cm.R[7] * X[1] + cm.R[8] * X[2];
    double x = o.r0 + cm.T[0], y = o.r1 + cm.T[1], z = o.r2 + cm.T[2];
    double iz = fast_rcp(z);
    double f = in.k00, ar = in.ar;
    double wf = w * f * iz;
    o.w = w;
    o.a = wf; o.b = -wf * x * iz; o.c = ar * wf; o.d = -ar * wf * y * iz;
}
// rows of the weighted camera Jacobian A~ = w dp/dXX [I | -[R X]x]   (poseDerivatives, v3d_metricbundle.cpp:52-71;
// parameter order t then omega)
__device__ __forceinline__ void jac_rows_A(const Obs& o, double A0[6], double A1[6])
{
    A0[0] = o.a; A0[1] = 0.0; A0[2] = o.b; A0[3] = o.b * o.r1; A0[4] = o.a * o.r2 - o.b * o.r0; A0[5] = -o.a * o.r1;
    A1[0] = 0.0; A1[1] = o.c; A1[2] = o.d; A1[3] = o.d * o.r1 - o.c * o.r2; A1[4] = -o.d * o.r0; A1[5] = o.c * o.r0;
}
// rows of the weighted point Jacobian B~ = w dp/dXX R
__device__ __forceinline__ void jac_rows_B(const Obs& o, const Cam& cm, double B0[3], double B1[3])
{
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        B0[c] = o.a * cm.R[c] + o.b * cm.R[6 + c];
        B1[c] = o.c * cm.R[3 + c] + o.d * cm.R[6 + c];
    }
}
// t = A~ p   (2-vector);  p = (pt, pw):  A~ p = P~ (pt + pw x r)
__device__ __forceinline__ void apply_A(const Obs& o, const double p[6], double& t0, double& t1)
{
    double s0 = p[0] + (p[4] * o.r2 - p[5] * o.r1);
    double s1 = p[1] + (p[5] * o.r0 - p[3] * o.r2);
    double s2 = p[2] + (p[3] * o.r1 - p[4] * o.r0);
    t0 = o.a * s0 + o.b * s2;
    t1 = o.c * s1 + o.d * s2;
}
// h = P~^t y  (camera-space 3-vector)
__device__ __forceinline__ void apply_Pt(const Obs& o, double y0, double y1, double h[3])
{
    h[0] = o.a * y0; h[1] = o.c * y1; h[2] = o.b * y0 + o.d * y1;
}

// ------------------------------------------------------------------------------------------------------
// K1  linearize_points: one thread per point over its (contiguous) observations.
//     residual + Huber weight (stored), cost, V_j = sum B~^t B~, g_p,j = sum B~^t (-w e)
//     (v3d_optimization.cpp:512-524, 543-558, 616-630).  Block partials: cost, max|g_p|, max diag V.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_BLOCK) k_linearize_points(Dev d)
{
    pdl_chain_enter();
    __shared__ double sm[PT_BLOCK / 32];
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    double cost = 0, gmax = 0, vmax = -1e30;
    int k0, deg; slice_of(d, j, k0, deg);
    {
        const double* Xb = d.X[st->cur]; const double* rt = d.rt[st->cur];
        double X[3] = {0, 0, 0}; if (j < d.M) load3p(Xb, j, X);
        double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int s = 0, k = k0; s < deg; ++s, k += 32) {
            int ci = __ldg(o_cam_ptr(d, k));
            if (ci < 0) continue;
            Cam cm; load_cam(rt, ci, cm);
            Obs o; obs_full(cm, X, __ldg(d.o_meas + k), d.in, o);
            *o_w_ptr(d, k) = o.w;
            double we0 = o.w * o.e0, we1 = o.w * o.e1;
            cost += we0 * we0 + we1 * we1;
            double B0[3], B1[3]; jac_rows_B(o, cm, B0, B1);
            if (d.obs_B) {                                        // kept for the PCG by-point sweep of the large-camera-count plan (k_cg_point_pass_wB)
                double* ob = d.obs_B + (size_t)(k >> 5) * 192 + (k & 31);
                const double live = ci < d.first_free ? 0.0 : 1.0;    // constant camera: no contribution to S
                ob[0] = live * B0[0]; ob[32] = live * B0[1]; ob[64] = live * B0[2]; ob[96] = live * B1[0]; ob[128] = live * B1[1]; ob[160] = live * B1[2];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] += B0[c] * (-we0) + B1[c] * (-we1);
            V[0] += B0[0] * B0[0] + B1[0] * B1[0]; V[1] += B0[0] * B0[1] + B1[0] * B1[1]; V[2] += B0[0] * B0[2] + B1[0] * B1[2];
            V[3] += B0[1] * B0[1] + B1[1] * B1[1]; V[4] += B0[1] * B0[2] + B1[1] * B1[2]; V[5] += B0[2] * B0[2] + B1[2] * B1[2];
        }
        if (j < d.M) {
            double2* Vo = reinterpret_cast<double2*>(d.V + 6 * (size_t)j);
            Vo[0] = make_double2(V[0], V[1]); Vo[1] = make_double2(V[2], V[3]); Vo[2] = make_double2(V[4], V[5]);
            double2* go = reinterpret_cast<double2*>(d.gp + 4 * (size_t)j);
            go[0] = make_double2(g[0], g[1]); go[1] = make_double2(g[2], 0.0);
            gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
            vmax = fmax(V[0], fmax(V[3], V[5]));
        }
    }
    cost = block_sum<PT_BLOCK>(cost, sm);
    gmax = block_max<PT_BLOCK>(gmax, sm);
    vmax = block_max<PT_BLOCK>(vmax, sm);
    if (threadIdx.x == 0) {
        double* o = d.pt_part + 4 * (size_t)blockIdx.x;
        o[0] = cost; o[1] = gmax; o[2] = vmax; o[3] = 0;
    }
}

// ------------------------------------------------------------------------------------------------------
// By-camera row sweeps.  Virtual warp v owns rows [vrow(v), vrow(v+1)); a row holds 32 observations of ONE camera; the
// record header names the camera and the segment (= partial-sum slot).  A segment change flushes the warp's
// accumulators (fixed butterfly) and loads the next camera.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int vrow(const Dev& d, int v) { return (int)((long long)d.n_rows * v / d.n_vwarps); }
template <int NACC> __device__ __forceinline__ void seg_flush(const Dev& d, int seg, double (&acc)[NACC], int lane)
{
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) {
        double* o = d.seg_part + PART_STRIDE * (size_t)seg;
#pragma unroll
        for (int q = 0; q < NACC; ++q) o[q] = acc[q];
    }
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
}

// K2  linearize_cameras: U_i += A~^t A~ (21 unique), g_c,i += A~^t (-w e)   (v3d_optimization.cpp:530-541, 600-614); the
//     weight is recomputed from the residual and written, with the lever arm R X, into the row record for the CG sweeps.
//     One warp per virtual warp (4 per CTA, 3 CTAs per SM at ~156 registers).  Same staging as the CG by-camera sweep: a
//     ring of LC_NS stages of CG rows per warp, each stage = the row records + the rows' measurements, filled by two TMA
//     bulk copies; the X_j gathers of stage t+1 are in flight while stage t is computed.  (The version that loaded two
//     rows at a time through registers was latency-bound at 12 warps per SM: 166 us.)
constexpr int LC_NS = 3;
constexpr int LC_HEAD_BYTES = REC_W;                             // what this kernel READS of a row record: header + point indices (144 of 1168 bytes;
                                                                 // the weight and R X fields are its OUTPUT: staging whole records read 160 MB for nothing)
constexpr int LC_STAGE_BYTES = CG * (LC_HEAD_BYTES + 512);       // CG record heads, then CG x 32 double2 measurements
constexpr int LINCAM_SMEM = (CH_BLOCK / 32) * LC_NS * (LC_STAGE_BYTES + 8);
__global__ void __launch_bounds__(CH_BLOCK) k_linearize_cameras(Dev d)
{
    pdl_chain_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vw = blockIdx.x * (CH_BLOCK / 32) + warp;
    if (vw >= d.n_vwarps) return;                                // warps are independent: no CTA-wide barrier below
    const int rb = vrow(d, vw), n = vrow(d, vw + 1) - rb;
    if (n <= 0) return;
    const int nst = (n + CG - 1) / CG;
    unsigned char* buf = smem_raw + (size_t)warp * LC_NS * LC_STAGE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(CH_BLOCK / 32) * LC_NS * LC_STAGE_BYTES) + warp * LC_NS;
    unsigned char* grec = d.c_rec + (size_t)rb * REC_BYTES;
    const double2* gmeas = d.c_meas + 32 * (size_t)rb;
    const uint64_t strm = policy_evict_first();
    auto arm = [&](int t) {
        const int sl = t % LC_NS;
        const uint32_t rows = (uint32_t)min(CG, n - t * CG);
        mbar_expect_tx(&bar[sl], rows * (LC_HEAD_BYTES + 512));
#pragma unroll
        for (uint32_t g = 0; g < (uint32_t)CG; ++g)
            if (g < rows) bulk_g2s_hint(buf + sl * LC_STAGE_BYTES + g * LC_HEAD_BYTES, grec + ((size_t)t * CG + g) * REC_BYTES, LC_HEAD_BYTES, &bar[sl], strm);
        bulk_g2s_hint(buf + sl * LC_STAGE_BYTES + CG * LC_HEAD_BYTES, gmeas + 32 * (size_t)t * CG, rows * 512, &bar[sl], strm);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < LC_NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int t = 0; t < LC_NS && t < nst; ++t) arm(t);
    }
    __syncwarp();
    const int cur = st->cur;
    const double* Xb = d.X[cur];
    const double* rt = d.rt[cur];
    int seg = -1, cam = 0;
    Cam cm;
    double acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0;
    auto issue = [&](int t, int (&jj)[CG], double (&XX)[CG][3]) {
        const int sl = t % LC_NS;
        mbar_wait(&bar[sl], (uint32_t)(t / LC_NS) & 1u);
#pragma unroll
        for (int g = 0; g < CG; ++g) {
            jj[g] = (t * CG + g < n) ? reinterpret_cast<const int*>(buf + sl * LC_STAGE_BYTES + g * LC_HEAD_BYTES + REC_J)[lane] : -1;
            XX[g][0] = 0.0; XX[g][1] = 0.0; XX[g][2] = 1.0;
            if (jj[g] >= 0) load3p_256(Xb, jj[g], XX[g]);
        }
    };
    auto compute = [&](int t, const int (&jj)[CG], const double (&XX)[CG][3]) {
        const int sl = t % LC_NS;
        const unsigned char* sb = buf + sl * LC_STAGE_BYTES;
#pragma unroll
        for (int g = 0; g < CG; ++g) {
            if (t * CG + g < n) {                                // warp-uniform
                const int2 h = *reinterpret_cast<const int2*>(sb + g * LC_HEAD_BYTES + REC_HDR);
                if (h.y != seg) {
                    if (seg >= 0) seg_flush<27>(d, seg, acc, lane);
                    seg = h.y; cam = h.x; load_cam(rt, cam, cm);
                }
                if (jj[g] >= 0 && cam >= d.first_free) {
                    const double2 m = reinterpret_cast<const double2*>(sb + CG * LC_HEAD_BYTES + g * 512)[lane];
                    Obs o; obs_full(cm, XX[g], m, d.in, o);
                    unsigned char* rec = grec + (size_t)(t * CG + g) * REC_BYTES;          // write-back goes to the record in HBM
                    reinterpret_cast<double*>(rec + REC_W)[lane] = o.w;
                    reinterpret_cast<double*>(rec + REC_R0)[lane] = o.r0;
                    reinterpret_cast<double*>(rec + REC_R1)[lane] = o.r1;
                    reinterpret_cast<double*>(rec + REC_R2)[lane] = o.r2;
                    double A0[6], A1[6]; jac_rows_A(o, A0, A1);
                    const double we0 = -o.w * o.e0, we1 = -o.w * o.e1;
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int b = a; b < 6; ++b) acc[q++] += A0[a] * A0[b] + A1[a] * A1[b];
#pragma unroll
                    for (int a = 0; a < 6; ++a) acc[21 + a] += A0[a] * we0 + A1[a] * we1;
                }
            }
        }
        __syncwarp();
        if (lane == 0 && t + LC_NS < nst) { fence_proxy_async(); arm(t + LC_NS); }
    };
    int jA[CG], jB[CG]; double XA[CG][3], XB[CG][3];
    issue(0, jA, XA);
    for (int t = 0; t < nst; t += 2) {
        if (t + 1 < nst) issue(t + 1, jB, XB);
        compute(t, jA, XA);
        if (t + 1 < nst) {
            if (t + 2 < nst) issue(t + 2, jA, XA);
            compute(t + 1, jB, XB);
        }
    }
    if (seg >= 0) seg_flush<27>(d, seg, acc, lane);
}

// second level of every by-camera reduction: fixed-order sum of a camera's segment partials.  Four lanes per (camera, value)
// take every fourth segment and are combined by a fixed butterfly (a small problem still has ~2400 virtual warps, i.e. one
// segment per row: 13-48 serial L2 round trips per camera in the one-thread version, 13 us at the demo sequence's size).
template <int NV, int OSTRIDE>
__global__ void k_camera_reduce(Dev d, double* __restrict__ out, int need_compute)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || (need_compute && !st->compute)) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, u = t >> 2, q = t & 3;
    const int i = u / NV, c = u - i * NV;
    double s0 = 0, s1 = 0;
    if (i < d.N) {
        const int b = d.cam_seg_beg[i], e = d.cam_seg_beg[i + 1];
        int sg = b + q;
        for (; sg + 4 < e; sg += 8) { s0 += d.seg_part[PART_STRIDE * (size_t)sg + c]; s1 += d.seg_part[PART_STRIDE * (size_t)(sg + 4) + c]; }
        if (sg < e) s0 += d.seg_part[PART_STRIDE * (size_t)sg + c];
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (i < d.N && q == 0) out[(size_t)i * OSTRIDE + c] = s;
}

// single-rank form of k_camera_reduce<27, PART_STRIDE> + k_lm_unpack: the reduced [21 unique | 6] record goes straight into U and g_c
__global__ void k_camera_reduce_unpack(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, u = t >> 2, q = t & 3;
    const int i = u / 27, c = u - i * 27;
    double s0 = 0, s1 = 0;
    if (i < d.N) {
        const int b = d.cam_seg_beg[i], e = d.cam_seg_beg[i + 1];
        int sg = b + q;
        for (; sg + 4 < e; sg += 8) { s0 += d.seg_part[PART_STRIDE * (size_t)sg + c]; s1 += d.seg_part[PART_STRIDE * (size_t)(sg + 4) + c]; }
        if (sg < e) s0 += d.seg_part[PART_STRIDE * (size_t)sg + c];
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (i < d.N && q == 0) {
        if (c >= 21) d.gc[6 * (size_t)i + (c - 21)] = s;
        else {
            int a = 0, rem = c;                                    // (a, b) of index c in the row-wise upper triangle
            while (rem >= 6 - a) { rem -= 6 - a; ++a; }
            const int bb = a + rem;
            d.U[36 * (size_t)i + 6 * a + bb] = s; d.U[36 * (size_t)i + 6 * bb + a] = s;
            if (st->variant == 1 && st->first && a == bb) d.cam_D0[6 * (size_t)i + a] = s;
        }
    }
}

// reduce the per-block partials of a thread-per-point kernel in a fixed order (single CTA).
// mode 0 (linearisation): out = ar_lin tail  [cost | ginf slot(rank) | vmax slot(rank)]
// mode 1 (decision):      out = ar_dec       [d2p, denomp, new_cost]
// mode 2 (statistics):    like 1 but runs regardless of st->done (mean reprojection error)
__global__ void __launch_bounds__(ONE_CTA) k_point_partials_reduce(Dev d, int mode, int count = -1)
{
    pdl_chain_enter();
    __shared__ double sm[ONE_CTA / 32];
    const LMState* st = d.st;
    if ((mode != 2 && st->done) || (mode == 0 && !st->compute)) return;
    double a = 0, b = 0.0, c = (mode == 0) ? -1e30 : 0.0, e4 = 0.0;
    const int np = count >= 0 ? count : d.n_pt_blocks;
    for (int i = threadIdx.x; i < np; i += ONE_CTA) {
        const double* p = d.pt_part + 4 * (size_t)i;
        a += p[0];
        if (mode == 0) { b = fmax(b, p[1]); c = fmax(c, p[2]); } else { b += p[1]; c += p[2]; e4 += p[3]; }
    }
    a = block_sum<ONE_CTA>(a, sm);
    if (mode == 0) { b = block_max<ONE_CTA>(b, sm); c = block_max<ONE_CTA>(c, sm); }
    else           { b = block_sum<ONE_CTA>(b, sm); c = block_sum<ONE_CTA>(c, sm); e4 = block_sum<ONE_CTA>(e4, sm); }
    if (threadIdx.x == 0) {
        if (mode == 0) {
            double* o = d.ar_lin + (size_t)d.N * PART_STRIDE;
            o[0] = a;
            for (int r = 0; r < d.world; ++r) { o[1 + r] = (r == d.rank) ? b : 0.0; o[1 + d.world + r] = (r == d.rank) ? c : 0.0; }
        } else {
            d.ar_dec[0] = a; d.ar_dec[1] = b; d.ar_dec[2] = c; d.ar_dec[3] = (mode == 1) ? e4 : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K4  lm_prepare (single CTA): unpack U / g_c, |J^t e|_inf stop (v3d_optimization.cpp:587), lambda0 on the
//     first linearisation (v3d_optimization.cpp:711-782).
// ------------------------------------------------------------------------------------------------------
// (a) one thread per (camera, element of the 6x6 block): unpack the reduced [21 unique | 6] record into U and g_c.  This used to be a
//     loop inside the single-CTA kernel below (one thread per camera, ~70 strided memory operations each): 68 us of an LM iteration.
__global__ void __launch_bounds__(256) k_lm_unpack(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= 36 * d.N) return;
    const int i = t / 36, e = t - 36 * i, a = e / 6, b = e - 6 * a;
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    const int q = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);        // index of (lo, hi) in the row-wise upper triangle
    const double* s = d.ar_lin + (size_t)i * PART_STRIDE;
    const double v = s[q];
    d.U[36 * (size_t)i + e] = v;
    if (e < 6) d.gc[6 * (size_t)i + e] = s[21 + e];
    if (st->variant == 1 && st->first && a == b) d.cam_D0[6 * (size_t)i + a] = v;   // PrepareJacobianNormalization (SparseBundleCPU.cpp:2900-2918)
}
// (b) single CTA: |J^t e|_inf, max diag U, lambda0
// fuse_count >= 0 (single rank): the block partials of the by-point linearisation (cost, max |g_p|, max diag V) are reduced
// here instead of by a k_point_partials_reduce launch
__global__ void __launch_bounds__(ONE_CTA) k_lm_prepare(Dev d, int fuse_count)
{
    pdl_chain_enter();
    __shared__ double sm[ONE_CTA / 32];
    LMState* st = d.st;
    if (st->done) return;
    if (st->compute && fuse_count >= 0) {
        double a = 0, b = 0.0, c = -1e30;
        for (int i = threadIdx.x; i < fuse_count; i += ONE_CTA) { const double* p = d.pt_part + 4 * (size_t)i; a += p[0]; b = fmax(b, p[1]); c = fmax(c, p[2]); }
        a = block_sum<ONE_CTA>(a, sm); b = block_max<ONE_CTA>(b, sm); c = block_max<ONE_CTA>(c, sm);
        if (threadIdx.x == 0) { double* o = d.ar_lin + (size_t)d.N * PART_STRIDE; o[0] = a; o[1] = b; o[2] = c; }   // world = 1: [cost | ginf | vmax]
        __syncthreads();
    }
    if (st->compute) {
        double gmax = 0, umax = -1e30;
        for (int e = threadIdx.x; e < 6 * d.N; e += ONE_CTA) {
            const int i = e / 6, a = e - 6 * i;
            if (i >= d.first_free) { gmax = fmax(gmax, fabs(d.gc[e])); umax = fmax(umax, d.U[36 * (size_t)i + 7 * a]); }
        }
        gmax = block_max<ONE_CTA>(gmax, sm);
        umax = block_max<ONE_CTA>(umax, sm);
        if (threadIdx.x == 0) {
            const double* t = d.ar_lin + (size_t)d.N * PART_STRIDE;
            double gp = 0, vm = -1e30;
            for (int r = 0; r < d.world; ++r) { gp = fmax(gp, t[1 + r]); vm = fmax(vm, t[1 + d.world + r]); }
            st->err = t[0];
            double ginf = fmax(gmax, gp);
            if (ginf < st->grad_thr) { st->status = 2; st->done = 1; }
            // (PBA variant: lambda0 = __lm_initial_damp is set by the host and `first` is cleared by k_lm_decide_pba)
            if (st->variant == 0 && st->first) { st->lambda = st->tau * fmax(umax, vm); st->first = 0; }
        }
    }
    if (threadIdx.x == 0) { st->solve_ok = 1; }
}

// K5  point_vinv: (V_j + lambda I)^-1 in registers (adjugate / determinant); not positive definite => the
//     step is rejected like an LDL failure (v3d_optimization.cpp:869-874).
// (V + damping)^-1 of one point by adjugate / determinant: Vi = (xx xy xz yy yz zz); returns false when the block is not positive definite
__device__ __forceinline__ bool point_vinv_eval(const double V[6], double l, bool pba, double Vi[6])
{
    double a = pba ? V[0] * (1.0 + l) : V[0] + l, b = V[1], c = V[2], e = pba ? V[3] * (1.0 + l) : V[3] + l, f = V[4], g = pba ? V[5] * (1.0 + l) : V[5] + l;
    double c00 = e * g - f * f, c01 = c * f - b * g, c02 = b * f - c * e;
    double det = a * c00 + b * c01 + c * c02;
    bool ok = (det > 0) && (a > 0) && (a * e - b * b > 0);
    double id = 1.0 / det;
    Vi[0] = c00 * id; Vi[1] = c01 * id; Vi[2] = c02 * id; Vi[3] = (a * g - c * c) * id; Vi[4] = (b * c - a * f) * id; Vi[5] = (a * e - b * b) * id;
    return ok;
}
// with_rhs: also v_j = (V_j + damping)^-1 g_p,j, the by-point half of the reduced right-hand side (what k_cg_point_pass<1> computes from
// the stored inverse: one launch and one pass over the inverses less per LM iteration)
__global__ void __launch_bounds__(256) k_point_vinv(Dev d, int with_rhs)
{
    pdl_chain_enter();
    LMState* st = d.st;
    if (st->done) return;
    int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= d.M) return;
    double V[6]; load6(d.V, j, V);
    const bool pba = st->variant == 1;                            // diag (1 + lambda) instead of + lambda (SparseBundleCPU.cpp:1793-1794)
    if (pba && st->first && st->compute) { double2* o = reinterpret_cast<double2*>(d.pt_D0 + 4 * (size_t)j); o[0] = make_double2(V[0], V[3]); o[1] = make_double2(V[5], 0.0); }
    double Vi[6];
    const bool ok = point_vinv_eval(V, st->lambda, pba, Vi);
    double2* o = reinterpret_cast<double2*>(d.Vinv + 6 * (size_t)j);
    o[0] = make_double2(Vi[0], Vi[1]);
    o[1] = make_double2(Vi[2], Vi[3]);
    o[2] = make_double2(Vi[4], Vi[5]);
    if (!ok) st->solve_ok = 0;
    if (with_rhs) {
        double g[3]; load3p(d.gp, j, g);
        const double v0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2];
        const double v1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2];
        const double v2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
        double2* w = reinterpret_cast<double2*>(d.v + 4 * (size_t)j);
        w[0] = make_double2(v0, v1); w[1] = make_double2(v2, 0.0);
    }
}

// ------------------------------------------------------------------------------------------------------
// K6  schur_diag_cameras: row sweep, sum_k W_k (V+lambda I)^-1 W_k^t (21 unique), W_k = A~^t B~.
//     Gives the diagonal blocks of S for the block-Jacobi preconditioner.  Streams the row records (weight, cached
//     R X) and gathers (V_j + lambda I)^-1.
// ------------------------------------------------------------------------------------------------------
#ifndef B200BA_SD_HDR
#define B200BA_SD_HDR 0
#endif
#ifndef B200BA_SD_PF
#define B200BA_SD_PF 0
#endif
constexpr int SD_PF = B200BA_SD_PF;                             // rows of records pulled into L2 ahead of a warp of k_schur_diag_cameras (0 = off: 2 / 3 / 6 rows measured 105 / 107 / 115 us against 89.5 without)
__global__ void __launch_bounds__(CH_BLOCK) k_schur_diag_cameras(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done) return;
    const int lane = threadIdx.x & 31, wpb = CH_BLOCK / 32, nw = gridDim.x * wpb;
    const double* rt = d.rt[st->cur];
    const double f = d.in.k00, ar = d.in.ar;
    for (int vw = blockIdx.x * wpb + (threadIdx.x >> 5); vw < d.n_vwarps; vw += nw) {
        const int rb = vrow(d, vw), re = vrow(d, vw + 1);
        int seg = -1, cam = 0;
        Cam cm;
        double acc[21];
#pragma unroll
        for (int q = 0; q < 21; ++q) acc[q] = 0;
        auto do_row = [&](const unsigned char* rec, int2 h, int j, const double Vi[6]) {
            if (h.y != seg) {
                if (seg >= 0) seg_flush<21>(d, seg, acc, lane);
                seg = h.y; cam = h.x; load_cam(rt, cam, cm);
            }
            if (j >= 0 && cam >= d.first_free) {
                Obs o;
                o.w = __ldg(reinterpret_cast<const double*>(rec + REC_W) + lane);
                o.r0 = __ldg(reinterpret_cast<const double*>(rec + REC_R0) + lane);
                o.r1 = __ldg(reinterpret_cast<const double*>(rec + REC_R1) + lane);
                o.r2 = __ldg(reinterpret_cast<const double*>(rec + REC_R2) + lane);
                const double x = o.r0 + cm.T[0], y = o.r1 + cm.T[1], z = o.r2 + cm.T[2];
                const double iz = fast_rcp(z), wf = o.w * f * iz;
                o.a = wf; o.b = -wf * x * iz; o.c = ar * wf; o.d = -ar * wf * y * iz;
                // camera-space covariance  C = R Vinv R^t (sym 3x3), then G = P~ C P~^t (sym 2x2), then
                // W Vinv W^t = A~^t G A~  because W = A~^t P~ R.
                double RV[9];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double r0 = cm.R[3 * q], r1 = cm.R[3 * q + 1], r2 = cm.R[3 * q + 2];
                    RV[3 * q]     = r0 * Vi[0] + r1 * Vi[1] + r2 * Vi[2];
                    RV[3 * q + 1] = r0 * Vi[1] + r1 * Vi[3] + r2 * Vi[4];
                    RV[3 * q + 2] = r0 * Vi[2] + r1 * Vi[4] + r2 * Vi[5];
                }
                const double C00 = RV[0] * cm.R[0] + RV[1] * cm.R[1] + RV[2] * cm.R[2];
                const double C02 = RV[0] * cm.R[6] + RV[1] * cm.R[7] + RV[2] * cm.R[8];
                const double C01 = RV[0] * cm.R[3] + RV[1] * cm.R[4] + RV[2] * cm.R[5];
                const double C11 = RV[3] * cm.R[3] + RV[4] * cm.R[4] + RV[5] * cm.R[5];
                const double C12 = RV[3] * cm.R[6] + RV[4] * cm.R[7] + RV[5] * cm.R[8];
                const double C22 = RV[6] * cm.R[6] + RV[7] * cm.R[7] + RV[8] * cm.R[8];
                // G = P C P^t with P = [[a,0,b],[0,c,d]]
                const double G00 = o.a * o.a * C00 + 2.0 * o.a * o.b * C02 + o.b * o.b * C22;
                const double G01 = o.a * o.c * C01 + o.a * o.d * C02 + o.b * o.c * C12 + o.b * o.d * C22;
                const double G11 = o.c * o.c * C11 + 2.0 * o.c * o.d * C12 + o.d * o.d * C22;
                double A0[6], A1[6]; jac_rows_A(o, A0, A1);
                // A^t G A : rows combined  Y0 = G00 A0 + G01 A1, Y1 = G01 A0 + G11 A1;  result[a][b] = A0[a] Y0[b] + A1[a] Y1[b]
                double Y0[6], Y1[6];
#pragma unroll
                for (int a = 0; a < 6; ++a) { Y0[a] = G00 * A0[a] + G01 * A1[a]; Y1[a] = G01 * A0[a] + G11 * A1[a]; }
                int q = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b) acc[q++] += A0[a] * Y0[b] + A1[a] * Y1[b];
            }
        };
        // (two rows of loads per trip, as in k_linearize_cameras, measured SLOWER here: 196 vs 112 us at 156 registers)
        // The warp walks its rows one at a time (16 warps per SM = exactly one wave of the 16 virtual warps per SM; more registers
        // per warp cost that): per row a DRAM round trip for the record, then the dependent gather of V^-1.  The NEXT row's point
        // index is loaded one row early, so that its gather can be issued as soon as the row starts: 112 -> 89.5 us at Venice shape.
        // (Bulk L2 prefetches of the records SD_PF rows ahead, by lane 0, made it slower again.)
        if (SD_PF > 0 && lane == 0 && rb < re) prefetch_l2_bulk(d.c_rec + (size_t)rb * REC_BYTES, (uint32_t)min(SD_PF, re - rb) * REC_BYTES);
        int j_next = rb < re ? __ldg(reinterpret_cast<const int*>(d.c_rec + (size_t)rb * REC_BYTES + REC_J) + lane) : -1;
#if B200BA_SD_HDR
        int2 h_next = rb < re ? __ldg(reinterpret_cast<const int2*>(d.c_rec + (size_t)rb * REC_BYTES + REC_HDR)) : make_int2(0, -1);
#endif
        for (int r = rb; r < re; ++r) {
            const unsigned char* rec = d.c_rec + (size_t)r * REC_BYTES;
            if (SD_PF > 0 && lane == 0 && r + SD_PF < re) prefetch_l2_bulk(rec + (size_t)SD_PF * REC_BYTES, REC_BYTES);
#if B200BA_SD_HDR
            const int2 h = h_next;
#else
            const int2 h = __ldg(reinterpret_cast<const int2*>(rec + REC_HDR));
#endif
            const int j = j_next;
            double Vi[6] = {0, 0, 0, 0, 0, 0};
            if (j >= 0) load6(d.Vinv, j, Vi);
            if (r + 1 < re) {
                j_next = __ldg(reinterpret_cast<const int*>(rec + REC_BYTES + REC_J) + lane);
#if B200BA_SD_HDR
                h_next = __ldg(reinterpret_cast<const int2*>(rec + REC_BYTES + REC_HDR));
#endif
            }
            do_row(rec, h, j, Vi);
        }
        if (seg >= 0) seg_flush<21>(d, seg, acc, lane);
    }
}

// K7  precond_invert: M_i = (U_i + damping - [precond] sum W Vinv W^t)^-1 by Cholesky.  Six threads per camera: each factors the
//     6x6 block itself (reciprocal pivots: 6 square roots and 6 divisions instead of 21 + 72 divisions in one serial thread,
//     14 us of a 213 us LM iteration at the demo sequence's size) and solves for ONE column of the inverse.
__global__ void __launch_bounds__(192) k_precond_invert(Dev d, int precond)
{
    pdl_chain_enter();
    LMState* st = d.st;
    if (st->done) return;
    const int t = blockIdx.x * 192 + threadIdx.x, i = t / 6, c = t - 6 * i;
    if (i >= d.N) return;
    double L[36];
    const double* U = d.U + 36 * (size_t)i;
    const double lam = st->lambda;
#pragma unroll
    for (int a = 0; a < 36; ++a) L[a] = U[a];
    const bool pba = st->variant == 1;
#pragma unroll
    for (int a = 0; a < 6; ++a) L[7 * a] = pba ? L[7 * a] * (1.0 + lam) : L[7 * a] + lam;
    if (pba && i < d.first_free) {                                // constant camera: all-zero rows (SparseBundleCPU.cpp:1384-1390), skipped by PBA's inverse
#pragma unroll
        for (int a = 0; a < 36; ++a) L[a] = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) L[7 * a] = 1.0;
    }
    if (precond == 1) {
        const double* s = d.ar_sd + 21 * (size_t)i;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) { const double v = s[q++]; L[6 * a + b] -= v; if (a != b) L[6 * b + a] -= v; }
    }
    bool ok = true;
    double rd[6];                                                 // reciprocal pivots
    // (all loops run over the full 0..5 range with guards: constant trip counts keep L, rd and x in registers)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double dd = L[7 * j];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < j) dd -= L[6 * j + k] * L[6 * j + k];
        if (!(dd > 0)) { ok = false; dd = 1.0; }
        dd = sqrt(dd); L[7 * j] = dd; rd[j] = 1.0 / dd;
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r > j) {
            double v = L[6 * r + j];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < j) v -= L[6 * r + k] * L[6 * j + k];
            L[6 * r + j] = v * rd[j];
        }
    }
    double x[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < r) v -= L[6 * r + k] * x[k];
        x[r] = v * rd[r];
    }
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
        const int r = 5 - rr;
        double v = x[r];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k > r) v -= L[6 * k + r] * x[k];
        x[r] = v * rd[r];
    }
    double* Mi = d.Minv + 36 * (size_t)i;
#pragma unroll
    for (int r = 0; r < 6; ++r) Mi[6 * r + c] = x[r];
    if (!ok && c == 0) st->solve_ok = 0;
}

// one observation of the by-point CG sweep against the packed camera table [quaternion 4 | T 3 | p 6 | pad]
struct TabRow { double h0, h1, h2, qw, qx, qy, qz, t0, t1; };   // t = A~ p is only read by the back-substitution (|J delta|^2)
// The table carries the rotation as sqrt(2) x the unit quaternion (QUAT_SCALE): every term of the sandwich
//   r = X + 2 (qw t + qv x t),  t = qv x X
// is quadratic in q, so the factor 2 comes with the scaled components and each component of r is a chain of three FMAs that
// starts at X (was: 2 products + 2 FMAs).  Record N of the table (tab_pad_cam) is a dummy camera for SELL padding entries -
// identity rotation, depth 1e100, zero direction; the padding weights of the row records are 0 (k_fill_orows) - so the row body
// runs without a select or a divergent branch and contributes exact zeros whatever the point is.
// The by-point sweep is bound by instruction issue (FP64 = 2 issue cycles): the row body below has 59 FP64 instructions, the
// first version had 71 (s = p_t + p_w x r as 2 FMAs per component instead of 3 operations; P~^t P~ s through x/z, y/z so that
// the two off-diagonal entries of P~ are never formed).
// first half of a row body: forward rotation, projection Jacobian, t = A~ p, h = P~^t t  (table fields die here)
// GLOBAL = false: the table lives in shared memory (k_cg_point_pass_tab); true: read-only global loads (problems whose
// camera table does not fit one SM's shared memory: 7 x LDG.128 instead of the 9 of [R|T] + p)
__device__ __forceinline__ int tab_pad_cam(int cam, int n_cams) { return (int)min((unsigned)cam, (unsigned)n_cams); }   // -1 (padding) -> record N
__device__ __forceinline__ void tab_row_core(const double2 a0, const double2 a1, const double2 a2, const double2 a3, const double2 a4, const double2 a5,
                                             const double2 a6, double w, const double X[3], const Intrinsics& in, TabRow& o)
{
    const double qw = a0.x, qx = a0.y, qy = a1.x, qz = a1.y;      // sqrt(2) x unit quaternion
    const double tx = qy * X[2] - qz * X[1], ty = qz * X[0] - qx * X[2], tz = qx * X[1] - qy * X[0];      // t = qv x X
    const double r0 = fma(qw, tx, fma(qy, tz, fma(-qz, ty, X[0])));                                        // r = X + qw t + qv x t
    const double r1 = fma(qw, ty, fma(qz, tx, fma(-qx, tz, X[1])));
    const double r2 = fma(qw, tz, fma(qx, ty, fma(-qy, tx, X[2])));
    const double x = r0 + a2.x, y = r1 + a2.y, z = r2 + a3.x;
    const double iz = fast_rcp(z);
    const double pa = w * in.k00 * iz, pc = in.ar * pa;           // P~ = [[pa, 0, -pa x/z], [0, pc, -pc y/z]]
    const double xz = x * iz, yz = y * iz;
    // s = p_t + p_w x r,  p = (a3.y a4.x a4.y | a5.x a5.y a6.x);  t = A~ p = P~ s
    const double s0 = fma(-a6.x, r1, fma(a5.y, r2, a3.y));
    const double s1 = fma(-a5.x, r2, fma(a6.x, r0, a4.x));
    const double s2 = fma(-a5.y, r0, fma(a5.x, r1, a4.y));
    const double t0 = pa * fma(-xz, s2, s0), t1 = pc * fma(-yz, s2, s1);
    o.h0 = pa * t0; o.h1 = pc * t1; o.h2 = fma(-yz, o.h1, -(xz * o.h0));
    o.qw = qw; o.qx = qx; o.qy = qy; o.qz = qz; o.t0 = t0; o.t1 = t1;
}
template <bool GLOBAL>
__device__ __forceinline__ void tab_row_fwd(const double2* __restrict__ tab, int cam, double w, const double X[3], const Intrinsics& in, TabRow& o)
{
    const double2* t = tab + 7 * cam;
    double2 a0, a1, a2, a3, a4, a5, a6;
    if (GLOBAL) { a0 = __ldg(t); a1 = __ldg(t + 1); a2 = __ldg(t + 2); a3 = __ldg(t + 3); a4 = __ldg(t + 4); a5 = __ldg(t + 5); a6 = __ldg(t + 6); }
    else { a0 = t[0]; a1 = t[1]; a2 = t[2]; a3 = t[3]; a4 = t[4]; a5 = t[5]; a6 = t[6]; }
    tab_row_core(a0, a1, a2, a3, a4, a5, a6, w, X, in, o);
}
// the same against a table addressed by its shared-memory window address (tab_smem_base): the base stays in ONE register; through the
// generic pointer the compiler re-derived it from %cluster_ctaid in every trip of the sweep's loop (S2R + MOV + VIADD + LEA)
__device__ __forceinline__ double2 lds_d2(uint32_t a) { double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t tab_smem_base(const void* smem) { uint32_t a = smem_u32(smem); asm volatile("" : "+r"(a)); return a; }
__device__ __forceinline__ void tab_row_fwd_s(uint32_t tab_s, int cam, double w, const double X[3], const Intrinsics& in, TabRow& o)
{
    const uint32_t t = tab_s + 112u * (uint32_t)cam;
#if B200BA_WHATIF & 2         // timing experiment (wrong results): ONE table load per observation instead of seven
    const double2 a0 = lds_d2(t), a1 = make_double2(a0.y, a0.x), a2 = a0, a3 = a1, a4 = a0, a5 = a1;
    double2 a6; a6.x = a0.x; a6.y = 0.0;
#else
    const double2 a0 = lds_d2(t), a1 = lds_d2(t + 16u), a2 = lds_d2(t + 32u), a3 = lds_d2(t + 48u), a4 = lds_d2(t + 64u), a5 = lds_d2(t + 80u);
    double2 a6; a6.x = lds_f64(t + 96u); a6.y = 0.0;
#endif
#if B200BA_WHATIF & 4         // timing experiment (wrong results): the loads, almost no arithmetic
    o.h0 = a0.x + a1.x + a2.x + w; o.h1 = a3.x + a4.x + a5.x + X[0]; o.h2 = a6.x + a0.y + a1.y + a2.y + a3.y + a4.y + a5.y;
    o.qw = o.qx = o.qy = o.qz = o.t0 = o.t1 = 0.0;
#else
    tab_row_core(a0, a1, a2, a3, a4, a5, a6, w, X, in, o);
#endif
}
// the PCG by-point sweep's table in two pieces: the static [quaternion | T] records (layout of qtab: 64 bytes, 16-byte units swizzled with
// (camera >> 1) & 3) and the direction p as it lies in d.p (48 bytes per camera: odd in 16-byte units) - both keep "bank group of a
// field = bijection of camera mod 8".  The static piece is staged BEFORE griddepcontrol.wait, i.e. while the camera-side step of the
// previous PCG step still runs; only the 48-byte piece waits for it.
__device__ __forceinline__ void tab_row_fwd_split(uint32_t ctab_s, uint32_t p_s, int cam, double w, const double X[3], const Intrinsics& in, TabRow& o)
{
    const uint32_t c = (uint32_t)cam, sw = (c << 3) & 0x30u;      // 16 x ((camera >> 1) & 3)
    const uint32_t tq = ctab_s + 64u * c, tp = p_s + 48u * c;
    const double2 a0 = lds_d2(tq + sw), a1 = lds_d2(tq + (16u ^ sw)), t01 = lds_d2(tq + (32u ^ sw)), t2 = lds_d2(tq + (48u ^ sw));
#if B200BA_WHATIF & 2         // timing experiment (wrong results): ONE table load per observation instead of seven
    const double2 p01 = a0, p23 = a1, p45 = a0;
#else
    const double2 p01 = lds_d2(tp), p23 = lds_d2(tp + 16u), p45 = lds_d2(tp + 32u);
#endif
    // packed-table field order of tab_row_core: a2 = (T0, T1), a3 = (T2, p0), a4 = (p1, p2), a5 = (p3, p4), a6 = (p5, -)
    tab_row_core(a0, a1, t01, make_double2(t2.x, p01.x), make_double2(p01.y, p23.x), make_double2(p23.y, p45.x), make_double2(p45.y, 0.0), w, X, in, o);
}
// second half: u += R^t h, the inverse rotation = sandwich with the conjugate quaternion (same scaled components)
__device__ __forceinline__ void tab_row_bwd(const TabRow& o, double& u0, double& u1, double& u2)
{
    const double gx = o.qz * o.h1 - o.qy * o.h2, gy = o.qx * o.h2 - o.qz * o.h0, gz = o.qy * o.h0 - o.qx * o.h1;   // g = -qv x h
    u0 = fma(o.qw, gx, fma(o.qz, gy, fma(-o.qy, gz, u0 + o.h0)));
    u1 = fma(o.qw, gy, fma(o.qx, gz, fma(-o.qz, gx, u1 + o.h1)));
    u2 = fma(o.qw, gz, fma(o.qy, gx, fma(-o.qx, gy, u2 + o.h2)));
}
// ------------------------------------------------------------------------------------------------------
// K8  cg_point_pass: v_j = (V+lambda I)^-1 u_j with
//       MODE 0: u_j = sum_k W_k^t p_i(k) = sum_k R^t P~^t P~ (pt + pw x r)     (first half of S p)
//       MODE 1: u_j = g_p,j                                                     (for the reduced rhs)
//       MODE 2: u_j = g_p,j - sum_k W_k^t x_i(k)  -> delta_p, trial point, trial cost, |delta_p|^2, delta_p.g_p
//                                                                               (back-substitution + cost re-evaluation,
//                                                                                v3d_optimization.cpp:908-921)
// ------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(PT_BLOCK) k_cg_point_pass(Dev d)
{
    pdl_chain_enter();
    __shared__ double sm[PT_BLOCK / 32];
    const LMState* st = d.st;
    if (st->done) return;
    if (MODE == 0 && !st->cg_active) return;
    if (MODE != 0 && !st->solve_ok) return;
    int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    double d2 = 0, dg = 0, cost = 0, njx = 0;
    if (j < d.M) {
        int cur = st->cur;
        double tt = 0;
        double u[3] = {0, 0, 0}, X[3], g[3];
        if (MODE != 1) load3p(d.X[cur], j, X);
        if (MODE != 0) load3p(d.gp, j, g);
        int k0, deg; slice_of(d, j, k0, deg);
        if (MODE == 0) {                                         // packed table [quat, T, p] from global memory (L2-resident)
            const double2* tabg = reinterpret_cast<const double2*>(d.ptab);
            for (int s = 0, k = k0; s < deg; ++s, k += 32) {
                const int i = __ldg(o_cam_ptr(d, k));
                if (i < 0) continue;
                TabRow o; tab_row_fwd<true>(tabg, i, __ldg(o_w_ptr(d, k)), X, d.in, o);
                tab_row_bwd(o, u[0], u[1], u[2]);
            }
        }
        if (MODE == 2) {
            const double* vec = d.x;
            const double* rt = d.rt[cur];
            for (int s = 0, k = k0; s < deg; ++s, k += 32) {
                int i = __ldg(o_cam_ptr(d, k));
                if (i < 0) continue;
                Cam cm; load_cam(rt, i, cm);
                Obs o; obs_jac(cm, X, __ldg(o_w_ptr(d, k)), d.in, o);
                double p[6]; load6(vec, i, p);
                double t0, t1; apply_A(o, p, t0, t1);
                if (i < d.first_free) { t0 = 0; t1 = 0; }
                tt += t0 * t0 + t1 * t1;
                double h[3]; apply_Pt(o, t0, t1, h);
#pragma unroll
                for (int c = 0; c < 3; ++c) u[c] += cm.R[c] * h[0] + cm.R[3 + c] * h[1] + cm.R[6 + c] * h[2];
            }
        }
        if (MODE == 1) { u[0] = g[0]; u[1] = g[1]; u[2] = g[2]; }
        if (MODE == 2) { u[0] = g[0] - u[0]; u[1] = g[1] - u[1]; u[2] = g[2] - u[2]; }
        double Vi[6]; load6(d.Vinv, j, Vi);
        double v0 = Vi[0] * u[0] + Vi[1] * u[1] + Vi[2] * u[2];
        double v1 = Vi[1] * u[0] + Vi[3] * u[1] + Vi[4] * u[2];
        double v2 = Vi[2] * u[0] + Vi[4] * u[1] + Vi[5] * u[2];
        if (MODE != 2) {
            double2* o = reinterpret_cast<double2*>(d.v + 4 * (size_t)j);
            o[0] = make_double2(v0, v1); o[1] = make_double2(v2, 0.0);
        } else {
            d2 = v0 * v0 + v1 * v1 + v2 * v2;
            dg = v0 * g[0] + v1 * g[1] + v2 * g[2];
            if (st->variant == 1) {                                     // PBA: see k_backsub_tab
                double Vu[6]; load6(d.V, j, Vu);
                const double ub0 = g[0] - u[0], ub1 = g[1] - u[1], ub2 = g[2] - u[2];      // sum B~^t A~ dc (u holds g - that sum here)
                njx = tt + 2.0 * (v0 * ub0 + v1 * ub1 + v2 * ub2)
                    + v0 * (Vu[0] * v0 + Vu[1] * v1 + Vu[2] * v2) + v1 * (Vu[1] * v0 + Vu[3] * v1 + Vu[4] * v2) + v2 * (Vu[2] * v0 + Vu[4] * v1 + Vu[5] * v2);
                const double* D0 = d.pt_D0 + 4 * (size_t)j;
                d2 = v0 * v0 * D0[0] + v1 * v1 * D0[1] + v2 * v2 * D0[2];
            }
            double Xn[3] = {X[0] + v0, X[1] + v1, X[2] + v2};           // updateParametersB, v3d_metricbundle.cpp:42-50
            double2* o = reinterpret_cast<double2*>(d.X[cur ^ 1] + 4 * (size_t)j);
            o[0] = make_double2(Xn[0], Xn[1]); o[1] = make_double2(Xn[2], 0.0);
            const double* rtn = d.rt[cur ^ 1];                          // trial cameras written by k_cg_finish
            for (int s = 0, k = k0; s < deg; ++s, k += 32) {
                int i = __ldg(o_cam_ptr(d, k));
                if (i < 0) continue;
                Cam cm; load_cam(rtn, i, cm);
                double e0, e1, x, y, iz, r0, r1, r2;
                obs_residual(cm, Xn, __ldg(d.o_meas + k), d.in, e0, e1, x, y, iz, r0, r1, r2);
                double w = huber_weight(e0, e1, d.in.huber);            // weights re-evaluated at the trial point (:917)
                double a = w * e0, b = w * e1;
                cost += a * a + b * b;
            }
        }
    }
    if (MODE == 2) {
        d2 = block_sum<PT_BLOCK>(d2, sm);
        dg = block_sum<PT_BLOCK>(dg, sm);
        cost = block_sum<PT_BLOCK>(cost, sm);
        njx = block_sum<PT_BLOCK>(njx, sm);
        if (threadIdx.x == 0) {
            double* o = d.pt_part + 4 * (size_t)blockIdx.x;
            o[0] = d2; o[1] = dg; o[2] = cost; o[3] = njx;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K8w  by-point PCG sweep of the LARGE-CAMERA-COUNT plan (camera table + row rings larger than one SM's shared memory: > ~1 797 cameras).
//      The generic sweep gathers the packed record [quaternion | T | p] - 7 x LDG.128 = 112 bytes per observation out of L2,
//      3.2 GB per sweep at Final shape (498 us: L2-gather bound) - and rebuilds R, the projection Jacobian and two rotations
//      from it (71 fp64 instructions per observation).  Here
//          W^t p = B~^t B~ (R^t p_t + (R^t p_w) x X),      B~ = w dp/dXX R  (2 x 3),
//      i.e. with the direction rotated into WORLD axes once per step and camera (k_build_wtab: 64-byte records, 2 x LDG.256
//      per observation) the rotation disappears from the sweep, and B~ - fixed for a linearisation - is STREAMED (48 bytes per
//      observation, coalesced, written by k_linearize_points) instead of recomputed: 21 fp64 instructions per observation.
//      More HBM bytes per observation (52 instead of 12), less than half the L2 gather traffic.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_build_wtab(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= d.N) return;
    const double* s = d.rt[st->cur] + 12 * (size_t)i;
    const double* p = d.ptab + (size_t)i * TAB_STRIDE + 7;
    const double p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3], p4 = p[4], p5 = p[5];
    double2* o = reinterpret_cast<double2*>(d.wtab + 8 * (size_t)i);
    // R^t v: column a of R (row-major at stride 4) dotted with v
    o[0] = make_double2(s[0] * p0 + s[4] * p1 + s[8] * p2, s[1] * p0 + s[5] * p1 + s[9] * p2);
    o[1] = make_double2(s[2] * p0 + s[6] * p1 + s[10] * p2, s[0] * p3 + s[4] * p4 + s[8] * p5);
    o[2] = make_double2(s[1] * p3 + s[5] * p4 + s[9] * p5, s[2] * p3 + s[6] * p4 + s[10] * p5);
    o[3] = make_double2(0.0, 0.0);
}
__global__ void __launch_bounds__(PT_BLOCK) k_cg_point_pass_wB(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    if (j >= d.M) return;
    double X[3]; load3p(d.X[st->cur], j, X);
    int k0, deg; slice_of(d, j, k0, deg);
    double u0 = 0, u1 = 0, u2 = 0;
    const double* wt = d.wtab;
#pragma unroll 2
    for (int s = 0, k = k0; s < deg; ++s, k += 32) {
        const int i = __ldg(o_cam_ptr(d, k));
        if (i < 0) continue;
        const double* ob = d.obs_B + (size_t)(k >> 5) * 192 + (k & 31);
        const double b00 = __ldg(ob), b01 = __ldg(ob + 32), b02 = __ldg(ob + 64), b10 = __ldg(ob + 96), b11 = __ldg(ob + 128), b12 = __ldg(ob + 160);
        double t[3], w[3];                                        // [R^t p_t | first of R^t p_w], [rest of R^t p_w | pad]
        double w0;
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(t[0]), "=d"(t[1]), "=d"(t[2]), "=d"(w0) : "l"(wt + 8 * (size_t)i));
        load3p_256(wt + 4, 2 * i, w);                             // record i's second half: (w1, w2, pad)
        const double a0 = t[0] + (w[0] * X[2] - w[1] * X[1]);     // a = t~ + w~ x X,  w~ = (w0, w[0], w[1])
        const double a1 = t[1] + (w[1] * X[0] - w0 * X[2]);
        const double a2 = t[2] + (w0 * X[1] - w[0] * X[0]);
        const double y0 = b00 * a0 + b01 * a1 + b02 * a2, y1 = b10 * a0 + b11 * a1 + b12 * a2;
        u0 += b00 * y0 + b10 * y1; u1 += b01 * y0 + b11 * y1; u2 += b02 * y0 + b12 * y1;
    }
    double Vi[6]; load6(d.Vinv, j, Vi);
    double2* o = reinterpret_cast<double2*>(d.v + 4 * (size_t)j);
    o[0] = make_double2(Vi[0] * u0 + Vi[1] * u1 + Vi[2] * u2, Vi[1] * u0 + Vi[3] * u1 + Vi[4] * u2);
    o[1] = make_double2(Vi[2] * u0 + Vi[4] * u1 + Vi[5] * u2, 0.0);
}

// K9  cg_camera_pass: sum_k W_k v_j(k) = sum_k A~^t P~ R v_j   (second half of S p), persistent row sweep.
//     Per observation the sweep STREAMS 36 bytes (point index, weight, cached R X) and GATHERS one 32-byte record v_j
//     (LDG.256, kept L2-resident with evict_last).  ncu showed the first version (warp per chunk, 3 gathers per lane in
//     flight, streams through registers, 128 registers) bound by gather latency: 43 % of its stall samples sat on the
//     first use of v_j.  Here the stream never touches a register before it is needed: each warp owns a ring of CNS
//     stages of CG row records in shared memory, each filled by ONE TMA bulk copy (cp.async.bulk + mbarrier, issued by
//     lane 0 two stages ahead of the gather front), and the registers hold two sets of CG outstanding v_j gathers per
//     lane: the gathers of stage t+1 are issued before the arithmetic of stage t, whose CG rows are computed in one
//     basic block (CG independent dependency chains per lane - a row-at-a-time version of this loop was issue-bound).
//     (Measured dead ends: a cp.async landing zone for the gathers doubles the L2 sector requests - two 16-byte LDGSTS
//     per record; a point-tile sweep with v in shared memory is barrier-bound.)
// contribution of one observation; `ok` false (padding / past the end): w must be 0 and v finite, z is replaced by 1
__device__ __forceinline__ void cam_pass_one(const Cam& cm, const Intrinsics& in, bool ok, double r0, double r1, double r2, const double v[3], double w, double acc[6])
{
    const double x = r0 + cm.T[0], y = r1 + cm.T[1], z = ok ? r2 + cm.T[2] : 1.0;
    const double iz = fast_rcp(z);
    const double pa = w * in.k00 * iz, pc = in.ar * pa;           // P~ = [[pa, 0, -pa x/z], [0, pc, -pc y/z]] (the off-diagonal entries are never formed)
    const double xz = x * iz, yz = y * iz;
    const double g0 = cm.R[0] * v[0] + cm.R[1] * v[1] + cm.R[2] * v[2];
    const double g1 = cm.R[3] * v[0] + cm.R[4] * v[1] + cm.R[5] * v[2];
    const double g2 = cm.R[6] * v[0] + cm.R[7] * v[1] + cm.R[8] * v[2];
    const double y0 = pa * fma(-xz, g2, g0), y1 = pc * fma(-yz, g2, g1);
    const double h0 = pa * y0, h1 = pc * y1, h2 = fma(-yz, h1, -(xz * h0));
    acc[0] += h0; acc[1] += h1; acc[2] += h2;
    acc[3] = fma(r1, h2, fma(-r2, h1, acc[3]));       // r x h
    acc[4] = fma(r2, h0, fma(-r0, h2, acc[4]));
    acc[5] = fma(r0, h1, fma(-r1, h0, acc[5]));
}
// the sweep of one warp over its row range (returns early for warps without rows: no CTA-wide barrier inside)
__device__ __forceinline__ void camera_pass_sweep(const Dev& d, const LMState* st, unsigned char* smem_raw)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef B200BA_STAMPS
    unsigned long long* stp = d.stamps + (size_t)(256 + blockIdx.x) * STAMP_STRIDE;
    if (threadIdx.x == 0) { stp[0] = gtime(); stp[2] = clock64(); }
    struct EndStamp { unsigned long long* p; int lane; __device__ ~EndStamp() { if (lane == 0) *p = clock64(); } } end_stamp{stp + 8 + warp, lane};
#endif
    const int vw = (blockIdx.x * CW + warp) * VSPLIT;  // this warp's VSPLIT consecutive virtual warps = one contiguous row range
    if (vw >= d.n_vwarps) return;                       // no CTA-wide barrier below: warps are independent
    const int rb = vrow(d, vw), n = vrow(d, min(vw + VSPLIT, d.n_vwarps)) - rb;
    if (n <= 0) return;
    const int nst = (n + CG - 1) / CG;                  // stages of this warp's range
    unsigned char* buf = smem_raw + (size_t)warp * CNS * STAGE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)CW * CNS * STAGE_BYTES) + warp * CNS;
    const unsigned char* src = d.c_rec + (size_t)rb * REC_BYTES;
    const uint64_t strm = policy_evict_first(), keep = policy_evict_last();
    // stage t = rows [t CG, t CG + CG) of the range, one bulk copy into ring slot t % CNS
    auto arm = [&](int t) {
        const int sl = t % CNS;
        const uint32_t bytes = (uint32_t)min(CG, n - t * CG) * REC_BYTES;
        mbar_expect_tx(&bar[sl], bytes);
        bulk_g2s_hint(buf + sl * STAGE_BYTES, src + (size_t)t * STAGE_BYTES, bytes, &bar[sl], strm);      // evict_first: 43 vs 47.5 us without the hint
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < CNS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int t = 0; t < CNS && t < nst; ++t) arm(t);
    }
    __syncwarp();
    const double* vb = d.v;
    const double* rt = d.rt[st->cur];
    const int first_free = d.first_free;
    int seg = -1, cam = 0;
    Cam cm;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    // gather front: point indices of stage t from shared memory, CG gathers per lane into one register set
    auto issue = [&](int t, int (&jj)[CG], double (&vv)[CG][3]) {
        const int sl = t % CNS;
        mbar_wait(&bar[sl], (uint32_t)(t / CNS) & 1u);
#pragma unroll
        for (int g = 0; g < CG; ++g) {
            jj[g] = (t * CG + g < n) ? reinterpret_cast<const int*>(buf + sl * STAGE_BYTES + g * REC_BYTES + REC_J)[lane] : -1;
            vv[g][0] = vv[g][1] = vv[g][2] = 0.0;
            if (jj[g] >= 0) load3p_256_keep(vb, jj[g], vv[g], keep);
        }
    };
    // arithmetic of stage t, then hand its ring slot back to the TMA
    auto compute = [&](int t, const int (&jj)[CG], const double (&vv)[CG][3]) {
        const int sl = t % CNS;
        const unsigned char* sb = buf + sl * STAGE_BYTES;
        int2 h[CG];
        bool same = (t * CG + CG <= n);
#pragma unroll
        for (int g = 0; g < CG; ++g) { h[g] = *reinterpret_cast<const int2*>(sb + g * REC_BYTES + REC_HDR); same = same && (h[g].y == seg); }
        if (same) {                                      // whole stage inside the current segment: one basic block, CG-way ILP
            if (cam >= first_free) {
                double w[CG], r0[CG], r1[CG], r2[CG];
#pragma unroll
                for (int g = 0; g < CG; ++g) {
                    const unsigned char* rec = sb + g * REC_BYTES;
                    w[g] = reinterpret_cast<const double*>(rec + REC_W)[lane];
                    r0[g] = reinterpret_cast<const double*>(rec + REC_R0)[lane];
                    r1[g] = reinterpret_cast<const double*>(rec + REC_R1)[lane];
                    r2[g] = reinterpret_cast<const double*>(rec + REC_R2)[lane];
                }
#pragma unroll
                for (int g = 0; g < CG; ++g) cam_pass_one(cm, d.in, jj[g] >= 0, r0[g], r1[g], r2[g], vv[g], w[g], acc);
            }
        } else {                                         // segment boundary or ragged end inside the stage: row by row
#pragma unroll
            for (int g = 0; g < CG; ++g) {
                if (t * CG + g < n) {
                    if (h[g].y != seg) {
                        if (seg >= 0) seg_flush<6>(d, seg, acc, lane);
                        seg = h[g].y; cam = h[g].x;
                        load_cam(rt, cam, cm);
                    }
                    if (cam >= first_free) {
                        const unsigned char* rec = sb + g * REC_BYTES;
                        cam_pass_one(cm, d.in, jj[g] >= 0, reinterpret_cast<const double*>(rec + REC_R0)[lane], reinterpret_cast<const double*>(rec + REC_R1)[lane],
                                     reinterpret_cast<const double*>(rec + REC_R2)[lane], vv[g], reinterpret_cast<const double*>(rec + REC_W)[lane], acc);
                    }
                }
            }
        }
        __syncwarp();                                    // every lane is done with this ring slot
        if (lane == 0 && t + CNS < nst) { fence_proxy_async(); arm(t + CNS); }
    };
    int jA[CG], jB[CG]; double vA[CG][3], vB[CG][3];
    issue(0, jA, vA);
    for (int t = 0; t < nst; t += 2) {
        if (t + 1 < nst) issue(t + 1, jB, vB);
        compute(t, jA, vA);
        if (t + 1 < nst) {
            if (t + 2 < nst) issue(t + 2, jA, vA);
            compute(t + 1, jB, vB);
        }
    }
    if (seg >= 0) seg_flush<6>(d, seg, acc, lane);
}
__global__ void __launch_bounds__(CW * 32, 1) k_cg_camera_pass(Dev d, int rhs_pass)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LMState* st = d.st;
    pdl_wait();                                                   // (programmatic dependent launch: v and the control words come from the by-point sweep)
    if (threadIdx.x == 0) pdl_trigger();                          // the camera-side step may be launched as a programmatic dependent of this sweep (single rank)
    if (st->done) return;
    if (rhs_pass ? !st->solve_ok : !st->cg_active) return;
    camera_pass_sweep(d, st, smem_raw);
}

// ------------------------------------------------------------------------------------------------------
// Peer-memory all-reduce (see struct P2P)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double* p2p_data(const P2P& x, int dst, unsigned long long c, int src)
{
    return reinterpret_cast<double*>(x.win[dst] + P2P_HDR_BYTES) + ((c & 1ull) * x.world + src) * x.cap;
}
// flag of (half, source rank, slot): slot 0 = "the whole contribution is there" (two-kernel collectives), slot 1 + b =
// "the elements of block b are there" (fused PCG-step exchange: block b of every rank handles the same elements, so a
// receiver block only waits for its own slot and no grid-wide sync is needed before the flags)
__device__ __forceinline__ unsigned long long* p2p_flag(const P2P& x, int dst, unsigned long long c, int src, int slot = 0)
{
    return reinterpret_cast<unsigned long long*>(x.win[dst]) + (((c & 1ull) * P2P_MAX_WORLD + src) * (P2P_MAX_BLOCKS + 1) + slot);
}
// element e of this rank's contribution -> every window
__device__ __forceinline__ void p2p_push(const P2P& x, unsigned long long c, size_t e, double v)
{
    for (int r = 0; r < x.world; ++r) p2p_data(x, r, c, x.rank)[e] = v;
}
// Threads 0 .. world-1 of ONE block, after every push of the grid has been ordered before them (grid sync / kernel
// boundary): raise this rank's flag in every window.
__device__ __forceinline__ void p2p_raise_flags(const P2P& x, unsigned long long c, int slot = 0)
{
    const int r = threadIdx.x;
    if (r < x.world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p2p_flag(x, r, c, x.rank, slot)), "l"(c + 1ull) : "memory");
    }
}
// threads 0 .. world-1 of a block each wait for one source's flag IN THE OWN WINDOW; bounded spin (about a second) so that a
// dead peer cannot hang the GPU
__device__ __forceinline__ void p2p_wait_all(const P2P& x, unsigned long long c, int slot = 0)
{
    const int r = threadIdx.x;
    if (r < x.world) {
        const unsigned long long* f = p2p_flag(x, x.rank, c, r, slot);
        unsigned long long v = 0; unsigned int spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= c + 1ull) break;
            if (++spins > (1u << 25)) { *x.err = 1; break; }
            if (spins > 4096u) __nanosleep(32);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ double p2p_sum(const P2P& x, unsigned long long c, size_t e)
{
    double s = 0;
    for (int r = 0; r < x.world; ++r) s += __ldcv(p2p_data(x, x.rank, c, r) + e);   // fixed rank order; .cv: never a stale L1 line
    return s;
}
// generic all-reduce (sum) of buf[0..count): push, then flag + reduce in place.  Two kernels: the boundary orders the
// pushes of all blocks before the flags.
__global__ void __launch_bounds__(256) k_p2p_publish(Dev d, const double* __restrict__ buf, int count, int need)
{
    const LMState* st = d.st;
    if (need == 1 && st->done) return;
    if (need == 2 && (st->done || !st->compute)) return;
    const unsigned long long c = *d.p2p.seq;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < count; e += gridDim.x * 256) p2p_push(d.p2p, c, (size_t)e, buf[e]);
}
__global__ void __launch_bounds__(256) k_p2p_reduce(Dev d, double* __restrict__ buf, int count, int need)
{
    const LMState* st = d.st;
    if (need == 1 && st->done) return;
    if (need == 2 && (st->done || !st->compute)) return;
    const unsigned long long c = *d.p2p.seq;
    if (blockIdx.x == 0) p2p_raise_flags(d.p2p, c);
    p2p_wait_all(d.p2p, c);
    for (int e = blockIdx.x * 256 + threadIdx.x; e < count; e += gridDim.x * 256) buf[e] = p2p_sum(d.p2p, c, (size_t)e);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(d.p2p.done_blocks, 1u) == gridDim.x - 1) { *d.p2p.done_blocks = 0; *d.p2p.seq = c + 1ull; __threadfence(); }
    }
}

// ------------------------------------------------------------------------------------------------------
// Camera-side PCG (single CTA; camera-sized vectors only).  Structure after PBA's implicit-Schur PCG
// (SparseBundleCPU.cpp:3323-3450) with additive damping and S-block / U-block Jacobi preconditioner.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double matvec6_row(const double* __restrict__ Mblk, const double* __restrict__ v6, int a)
{
    const double* r = Mblk + 6 * a;
    return r[0] * v6[0] + r[1] * v6[1] + r[2] * v6[2] + r[3] * v6[3] + r[4] * v6[4] + r[5] * v6[5];
}

__device__ __forceinline__ double grid_total_warp(const double* part, int n)
{
    double s = 0;
    for (int b = (threadIdx.x & 31); b < n; b += 32) s += part[b];
    return warp_sum(s);
}
// K10 cg_init: rhs = g_c - W Vinv g_p ; x = 0 ; r = rhs ; z = M r ; p = z   (one thread per (camera, component), block partials of r.z)
__global__ void __launch_bounds__(STEP_BLOCK) k_cg_init(Dev d, double* __restrict__ part)
{
    pdl_chain_enter();
    __shared__ double sv[STEP_BLOCK];
    __shared__ double sm[STEP_BLOCK / 32];
    const LMState* st = d.st;
    if (st->done) return;
    const bool ok = st->solve_ok != 0;
    const int e = blockIdx.x * STEP_BLOCK + threadIdx.x, n6 = 6 * d.N;       // a block owns 32 whole cameras
    const bool on = ok && e < n6;
    const int i = e / 6, a = e - 6 * i, lc = threadIdx.x - a;
    double re = 0;
    if (on) { re = d.gc[e] - d.ar_q[e]; d.x[e] = 0.0; d.r[e] = re; }
    sv[threadIdx.x] = re;
    __syncthreads();
    double z = 0;
    if (on) {
        const double* mr = d.Minv + 36 * (size_t)i + 6 * a;
        z = mr[0] * sv[lc] + mr[1] * sv[lc + 1] + mr[2] * sv[lc + 2] + mr[3] * sv[lc + 3] + mr[4] * sv[lc + 4] + mr[5] * sv[lc + 5];
        d.z[e] = z; d.p[e] = z; d.ptab[(size_t)i * TAB_STRIDE + 7 + a] = z;
        if (d.ptab32) d.ptab32[(size_t)i * 20 + 8 + a] = (float)z;
    }
    const double rz = block_sum<STEP_BLOCK>(re * z, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = rz;
}
// fixed-order sum of the block partials of k_cg_init -> PCG control words
__global__ void __launch_bounds__(32) k_cg_init_finish(Dev d, const double* __restrict__ part, int nblocks)
{
    pdl_chain_enter();
    LMState* st = d.st;
    if (st->done) return;
    const double rz = grid_total_warp(part, nblocks);
    if (threadIdx.x == 0) {
        st->rz = rz; st->rz0 = rz; st->cg_rel = 0; st->cg_steps = 0;
        st->cg_active = (st->solve_ok && rz > 0) ? 1 : 0;
    }
}

// K11 cg_update: q = (U + lambda I) p - sum W v ; alpha = r.z / p.q ; x += alpha p ; r -= alpha q ; z = M r ;
//     beta = r.z' / r.z ; p = z + beta p
__global__ void __launch_bounds__(ONE_CTA) k_cg_update(Dev d)
{
    pdl_chain_enter();
    __shared__ double sm[ONE_CTA / 32];
    LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    int n6 = 6 * d.N;
    double lam = st->lambda, rz = st->rz, rz0 = st->rz0, tol = st->cg_tol;
    double pq = 0;
    for (int e = threadIdx.x; e < n6; e += ONE_CTA) {
        int i = e / 6, a = e - 6 * i;
        double q = lam * (st->variant == 1 ? d.U[36 * (size_t)i + 7 * a] : 1.0) * d.p[e] + matvec6_row(d.U + 36 * (size_t)i, d.p + 6 * (size_t)i, a) - d.ar_q[e];
        d.q[e] = q; pq += d.p[e] * q;
    }
    pq = block_sum<ONE_CTA>(pq, sm);
    if (!(pq > 0) || !isfinite(pq)) {                      // breakdown: keep x, first-step breakdown = failed solve
        if (threadIdx.x == 0) { st->cg_active = 0; if (st->cg_steps == 0) st->solve_ok = 0; }
        return;
    }
    double alpha = rz / pq;
    for (int e = threadIdx.x; e < n6; e += ONE_CTA) { d.x[e] += alpha * d.p[e]; d.r[e] -= alpha * d.q[e]; }
    __syncthreads();
    double rzn = 0;
    for (int e = threadIdx.x; e < n6; e += ONE_CTA) {
        int i = e / 6, a = e - 6 * i;
        double z = matvec6_row(d.Minv + 36 * (size_t)i, d.r + 6 * (size_t)i, a);
        d.z[e] = z; rzn += d.r[e] * z;
    }
    rzn = block_sum<ONE_CTA>(rzn, sm);
    double rel = sqrt(fabs(rzn) / rz0);
    bool stop = (tol > 0 && rel < tol && st->cg_steps + 1 >= st->cg_min_steps) || !(rzn > 0) || (st->cg_guard > 0 && rel > st->cg_guard);
    if (!stop) {
        double beta = rzn / rz;
        for (int e = threadIdx.x; e < n6; e += ONE_CTA) {
            double pn = d.z[e] + beta * d.p[e];
            d.p[e] = pn; d.ptab[(size_t)(e / 6) * TAB_STRIDE + 7 + (e % 6)] = pn;
            if (d.ptab32) d.ptab32[(size_t)(e / 6) * 20 + 8 + (e % 6)] = (float)pn;
        }
    }
    if (threadIdx.x == 0) {
        st->cg_steps += 1; st->total_cg += 1; st->cg_rel = rel; st->rz = rzn;
        if (stop) st->cg_active = 0;
    }
}


// ------------------------------------------------------------------------------------------------------
// Shared-memory camera table for the by-point CG sweep.
//
// In the by-point sweep every observation needs its camera's R, T and p (18 doubles = 5 sectors of L2 traffic and,
// worse, 6 uncoalesced LDG.128 per lane = 192 L1 wavefronts per warp: ncu showed the generic sweep bound by L1
// wavefronts at 70 %).  For problems up to (MAX_SMEM - row rings) / 112 B = ~1 797 cameras (Venice: 1778) the whole camera side fits
// in one SM's shared memory once R is carried as a unit quaternion: 14 doubles per camera, staged with TMA bulk
// copies (cp.async.bulk, one mbarrier), read with 7 LDS.128 per observation.  k_build_ptab converts R -> quaternion
// once per linearisation; the CG kernels keep the p-part current.
// ------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int qtab_unit(int c, int f);
__device__ __forceinline__ void rot_to_quat(const double* s, double q[4]);
// also_qtab: the compact [quaternion | T] table of the same cameras in the same launch (k_build_qtab(d, 0); its only reader,
// k_lin_tab, runs under the same `compute` flag)
__global__ void __launch_bounds__(128) k_build_ptab(Dev d, int also_qtab)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= d.N) return;
    const double* s = d.rt[st->cur] + 12 * (size_t)i;
    double q[4]; rot_to_quat(s, q);                               // sqrt(2) x unit quaternion (QUAT_SCALE)
    double* o = d.ptab + (size_t)i * TAB_STRIDE;
    o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = s[3]; o[5] = s[7]; o[6] = s[11]; o[13] = 0.0;
    {
        double2* t = reinterpret_cast<double2*>(d.ctab);
        t[qtab_unit(i, 0)] = make_double2(q[0], q[1]); t[qtab_unit(i, 1)] = make_double2(q[2], q[3]);
        t[qtab_unit(i, 2)] = make_double2(s[3], s[7]); t[qtab_unit(i, 3)] = make_double2(s[11], 0.0);
    }
    if (also_qtab) {
        double2* t = reinterpret_cast<double2*>(d.qtab);
        t[qtab_unit(i, 0)] = make_double2(q[0], q[1]); t[qtab_unit(i, 1)] = make_double2(q[2], q[3]);
        t[qtab_unit(i, 2)] = make_double2(s[3], s[7]); t[qtab_unit(i, 3)] = make_double2(s[11], 0.0);
    }
}

// one warp per point-order row: camera indices into the row record; the weights start at 0 (the linearisation writes those of the
// real entries, the padding entries keep 0: the table sweeps run them against the dummy camera without a select).  Thread 0
// also writes that dummy camera, record N of the packed table: identity rotation, depth 1e100, zero direction (tab_row_fwd).
__global__ void __launch_bounds__(256) k_fill_orows(Dev d, const int* __restrict__ o_cam, int n_orows)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < TAB_STRIDE)
        d.ptab[(size_t)d.N * TAB_STRIDE + threadIdx.x] = threadIdx.x == 0 ? QUAT_SCALE : threadIdx.x == 6 ? 1e100 : 0.0;
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 36) {          // the same dummy camera in the compact table (fields 0 .. 3)
        const int f = threadIdx.x - 32;
        reinterpret_cast<double2*>(d.ctab)[qtab_unit(d.N, f)] = make_double2(f == 0 ? QUAT_SCALE : f == 3 ? 1e100 : 0.0, 0.0);
    }
    if (r >= n_orows) return;
    unsigned char* rec = d.o_rec + (size_t)r * OREC_BYTES;
    reinterpret_cast<int*>(rec)[lane] = o_cam[32 * (size_t)r + lane];
    reinterpret_cast<double*>(rec + 128)[lane] = 0.0;
}

// ---- device-side build of the camera-order row records (set_problem) ----
// Key of SELL entry e = (camera << 32 | new point index): a radix sort of the keys groups the observations by camera
// in point order, exactly the order the row records want; padding entries get camera = N and sort to the end.
__global__ void __launch_bounds__(PT_BLOCK) k_rows_keys(Dev d, const int* __restrict__ o_cam, unsigned long long* __restrict__ keys, int* __restrict__ vals)
{
    const int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    int k0, deg; slice_of(d, j, k0, deg);                        // warp-uniform trip count; lanes past M own only padding entries
    for (int s = 0, k = k0; s < deg; ++s, k += 32) {
        const int c = o_cam[k];
        keys[k] = ((unsigned long long)(unsigned)(c >= 0 ? c : d.N) << 32) | (unsigned)j;
        vals[k] = k;
    }
}
// one warp per row: header + "no observation" in every lane (the sorted fill below overwrites the real ones)
__global__ void __launch_bounds__(256) k_fill_rows(Dev d, const int2* __restrict__ row_hdr)
{
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= d.n_rows) return;
    unsigned char* rec = d.c_rec + (size_t)r * REC_BYTES;
    if (lane == 0) { const int2 h = row_hdr[r]; *reinterpret_cast<int4*>(rec + REC_HDR) = make_int4(h.x, h.y, 0, 0); }
    reinterpret_cast<int*>(rec + REC_J)[lane] = -1;
}
// sorted position i (< K) -> lane (i - first[cam]) of the camera's rows: point index into the record, measurement beside it
__global__ void __launch_bounds__(256) k_rows_fill(Dev d, const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                                   const int* __restrict__ cam_first, const int* __restrict__ row_beg)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.K) return;
    const unsigned long long key = keys[i];
    const int cam = (int)(key >> 32), jn = (int)(unsigned)key, e = vals[i];
    const int rank = i - __ldg(cam_first + cam);
    const int row = __ldg(row_beg + cam) + (rank >> 5), lane = rank & 31;
    reinterpret_cast<int*>(d.c_rec + (size_t)row * REC_BYTES + REC_J)[lane] = jn;
    d.c_meas[32 * (size_t)row + lane] = d.o_meas[e];
}

constexpr int PR = 4;                                            // rows of the point-order stream staged per warp: 2 stages x 2 rows
// The ring is only two stages deep (the camera table leaves 28 KB of shared memory for all the warps' rings), so a warp has
// ONE bulk copy in flight while it consumes the other stage: with the copy coming from DRAM (~1.2 us under load) every warp
// ran at "two rows per DRAM latency" whatever its instruction-level parallelism (one row or two rows in flight: the same
// 1200 cycles per row per warp, clock stamps + ncu round 2).  The stream is therefore pulled into L2 TAB_PF rows ahead of the
// ring by bulk L2 prefetches issued with each re-arm, which turns the ring's copies into L2 hits.
#ifndef B200BA_TAB_PF
#define B200BA_TAB_PF 12
#endif
constexpr int TAB_PF = B200BA_TAB_PF;
constexpr int TAB_RING_BYTES = (TAB_BLOCK / 32) * (PR * OREC_BYTES + 2 * 8);   // rings + their two mbarriers per warp
#ifdef B200BA_TAB_MAXNREG
__global__ void __maxnreg__(B200BA_TAB_MAXNREG) k_cg_point_pass_tab(Dev d)
#else
__global__ void __launch_bounds__(TAB_BLOCK, 1) k_cg_point_pass_tab(Dev d)
#endif
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const double2* tab = reinterpret_cast<const double2*>(smem_raw);
    const uint32_t total = (uint32_t)(d.N + 1) * TAB_STRIDE * 8u;    // + the padding entries' dummy camera
#ifdef B200BA_STAMPS
    unsigned long long* stp = d.stamps + (size_t)blockIdx.x * STAMP_STRIDE;
    if (threadIdx.x == 0) { stp[0] = gtime(); stp[2] = clock64(); }
#endif
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.ptab);
#ifdef B200BA_TAB_STAGGER
        // every CTA asks for the same 199 KB at the same moment: start each one at a different chunk so that the requests
        // spread over the L2 slices instead of all hitting the first lines together
        const uint32_t nch = (total + B200BA_TAB_STAGGER - 1) / B200BA_TAB_STAGGER;
        for (uint32_t c = 0; c < nch; ++c) {
            const uint32_t off = ((c + blockIdx.x) % nch) * B200BA_TAB_STAGGER;
            const uint32_t n = total - off < (uint32_t)B200BA_TAB_STAGGER ? total - off : (uint32_t)B200BA_TAB_STAGGER;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
#else
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
#endif
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Xb = d.X[st->cur];
    const int wid = blockIdx.x * (TAB_BLOCK / 32) + warp;
    const bool active = wid < d.tab_warps;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0;
    const unsigned char* src_rows = d.o_rec;
    if (active) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; src_rows += (size_t)(ha.z >> 5) * OREC_BYTES; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    // The warp's point-order row records form one contiguous stream (it owns a contiguous range of slices).  The stream
    // is staged by the TMA into a private ring of two stages of two rows (one bulk copy per stage, issued by lane 0 when
    // the previous occupant of the stage has been consumed); a row's two values are read into registers at the top of
    // its body.
    // (Prefetching the stream through REGISTERS does not survive ptxas: whatever the source says - rotation, unrolled
    // copies, a phase switch - the loop-carried value is copied, the copy waits for the load it was meant to overlap,
    // and ncu showed 12-25 % of all stall samples on that move.)  All ring addresses are 32-bit shared-window offsets
    // kept in registers; generic pointers made ptxas rebuild them every row.
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (PR * OREC_BYTES);
    const uint32_t rbar = smem_u32(smem_raw) + tab_bytes + (uint32_t)(TAB_BLOCK / 32) * (PR * OREC_BYTES) + (uint32_t)warp * 16u;
    auto arm = [&](int stage) {                                  // lane 0 only: next two unrequested rows -> ring stage
        const uint32_t bytes = (uint32_t)min(2, n_rows) * OREC_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar + 8u * stage), "r"(bytes) : "memory");
        // evict_first: the stream must not push the gathered vectors (X, v: evict_last) out of L2 (by-camera sweep: 43 vs 47.5 us)
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring + (uint32_t)stage * (2 * OREC_BYTES)), "l"(src_rows), "r"(bytes), "r"(rbar + 8u * stage), "l"(policy_evict_first()) : "memory");
        if (TAB_PF > 0 && n_rows > TAB_PF) prefetch_l2_bulk(src_rows + (size_t)TAB_PF * OREC_BYTES, (uint32_t)min(2, n_rows - TAB_PF) * OREC_BYTES);
        src_rows += 2 * OREC_BYTES; n_rows -= 2;
    };
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar + 8u), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        if (TAB_PF > 0 && n_rows > 2) prefetch_l2_bulk(src_rows + 2 * OREC_BYTES, (uint32_t)min(TAB_PF, n_rows - 2) * OREC_BYTES);
        if (n_rows > 0) arm(0);                                  // stage 1 is requested at the top of the first trip
    }
    double X[3] = {0, 0, 0};
    if (g < g_end && 32 * g + lane < d.M) load3p_256_keep(Xb, 32 * g + lane, X, policy_evict_last());
    __syncthreads();
    mbar_wait(&bar, 0);
#ifdef B200BA_STAMPS
    if (threadIdx.x == 0) { stp[1] = gtime(); stp[3] = clock64(); }
    struct EndStamp { unsigned long long* p; int lane; __device__ ~EndStamp() { if (lane == 0) *p = clock64(); } } end_stamp{stp + 8 + warp, lane};
#endif
    if (g >= g_end) return;
    double u0 = 0, u1 = 0, u2 = 0;
    // Requests what the END of slice g needs: its inverse block and the next slice's X.  Called in the middle of the
    // slice's last row body (after the table fields are dead, before the back-rotation), consumed by finish().
    auto request = [&](double Vi[6], double Xn[3]) {
        if (32 * g + lane < d.M) load6(d.Vinv, 32 * g + lane, Vi);
        if (g + 1 < g_end && 32 * (g + 1) + lane < d.M) load3p_256_keep(Xb, 32 * (g + 1) + lane, Xn, policy_evict_last());
    };
    // finish point 32 g + lane and step to the next slice; false: the warp's range is done
    auto finish = [&](const double Vi[6], const double Xn[3]) -> bool {
        const int j = 32 * g + lane;
        if (j < d.M)
            store4_keep(d.v + 4 * (size_t)j, Vi[0] * u0 + Vi[1] * u1 + Vi[2] * u2, Vi[1] * u0 + Vi[3] * u1 + Vi[4] * u2,
                        Vi[2] * u0 + Vi[4] * u1 + Vi[5] * u2, 0.0, policy_evict_last());
        if (++g >= g_end) return false;
        rows_left = deg_next; X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2]; u0 = u1 = u2 = 0;
        if (g + 1 < g_end) {
            deg_next = __ldg(d.slice_deg + g + 1);
            if (lane == 0) {                                     // L2 prefetch of what the next request() will ask for
                const int nj = min(32, d.M - 32 * (g + 1));
                if (nj > 0) { prefetch_l2_bulk(d.Vinv + 6 * (size_t)(32 * (g + 1)), (uint32_t)nj * 48u); prefetch_l2_bulk(Xb + 4 * (size_t)(32 * (g + 1)), (uint32_t)nj * 32u); }
            }
        }
        return true;
    };
    int slot = 0; uint32_t par = 0;                              // n_rows: rows not yet requested (meaningful in lane 0)
    for (;;) {
        while (rows_left == 0) {                                 // slice without observations (warp-uniform)
            double Vi[6], Xn[3];
            request(Vi, Xn);
            if (!finish(Vi, Xn)) return;
        }
        if ((slot & 1) == 0) {                                   // first row of a stage
            // the OTHER stage was consumed by the previous two trips (its reads fed their arithmetic, so the refill cannot
            // overtake them and no proxy fence is needed): request the two rows after this stage into it
            if (lane == 0 && n_rows > 0) arm((slot >> 1) ^ 1);
            uint32_t ok;                                         // wait for this stage (compiler-visible loop + reconvergence)
            do {
                asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }"
                             : "=r"(ok) : "r"(rbar + 4u * (uint32_t)slot), "r"(par) : "memory");
            } while (!ok);
            __syncwarp();
        }
        const uint32_t ca = ring + (uint32_t)slot * OREC_BYTES + 4u * lane;
        const int cam = lds_s32(ca);
        const double w = lds_f64(ca + 128u + 4u * lane);
        if (++slot == PR) { slot = 0; par ^= 1u; }
        const bool last = (--rows_left == 0);                    // warp-uniform
        TabRow o;
        tab_row_fwd<false>(tab, tab_pad_cam(cam, d.N), w, X, d.in, o);
        double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};     // (leaving these uninitialised measured 6 us SLOWER; requesting
        if (last) request(Vi, Xn);                                //  before the forward half 2 us slower)
        tab_row_bwd(o, u0, u1, u2);
        if (last && !finish(Vi, Xn)) return;
    }
}

// ------------------------------------------------------------------------------------------------------
// K8'' cg_point_pass_tab2: the same sweep with TWO rows of a slice in flight per trip (B200BA_TAB_ILP=2).
//      The clock stamps of the one-row sweep (tools/stamps.py) showed it latency-bound per warp: the scheduler with 4 of the
//      CTA's 19 warps finished 10 % ahead of the three with 5 - 25 % more work per scheduler cost 10 % more time, i.e. the
//      dependent FP64 chain of a row (LDS camera index -> LDS table -> ~25 dependent DFMA) is what a warp waits on, and
//      more independent chains per scheduler is what helps.  Registers allow no more warps (96 x 608), so the chains come
//      from instruction-level parallelism: when the current slice has >= 2 rows left, both run in one basic block (their
//      table reads, rotations and projections are independent; only the final u += differs).  Fewer, fatter warps
//      (TAB2_BLOCK threads at <= 128 registers), same ring (2 stages x 2 rows, one bulk copy per stage), same headers.
// ------------------------------------------------------------------------------------------------------
#ifndef B200BA_TAB2_BLOCK
#define B200BA_TAB2_BLOCK 512
#endif
constexpr int TAB2_BLOCK = B200BA_TAB2_BLOCK;
constexpr int TAB2_RING_BYTES = (TAB2_BLOCK / 32) * (PR * OREC_BYTES + 2 * 8);
__global__ void __launch_bounds__(TAB2_BLOCK, 1) k_cg_point_pass_tab2(Dev d)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;
    const double2* tab = reinterpret_cast<const double2*>(smem_raw);
    const uint32_t total = (uint32_t)(d.N + 1) * TAB_STRIDE * 8u;    // + the padding entries' dummy camera
#ifdef B200BA_STAMPS
    unsigned long long* stp = d.stamps + (size_t)blockIdx.x * STAMP_STRIDE;
    if (threadIdx.x == 0) { stp[0] = gtime(); stp[2] = clock64(); }
#endif
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.ptab);
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Xb = d.X[st->cur];
    const int wid = blockIdx.x * (TAB2_BLOCK / 32) + warp;
    const bool active = wid < d.tab_warps;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0;
    const unsigned char* src_rows = d.o_rec;
    if (active) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; src_rows += (size_t)(ha.z >> 5) * OREC_BYTES; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (PR * OREC_BYTES);
    const uint32_t rbar = smem_u32(smem_raw) + tab_bytes + (uint32_t)(TAB2_BLOCK / 32) * (PR * OREC_BYTES) + (uint32_t)warp * 16u;
    auto arm = [&](int stage) {                                  // lane 0 only: next two unrequested rows -> ring stage
        const uint32_t bytes = (uint32_t)min(2, n_rows) * OREC_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar + 8u * stage), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring + (uint32_t)stage * (2 * OREC_BYTES)), "l"(src_rows), "r"(bytes), "r"(rbar + 8u * stage), "l"(policy_evict_first()) : "memory");
        if (TAB_PF > 0 && n_rows > TAB_PF) prefetch_l2_bulk(src_rows + (size_t)TAB_PF * OREC_BYTES, (uint32_t)min(2, n_rows - TAB_PF) * OREC_BYTES);
        src_rows += 2 * OREC_BYTES; n_rows -= 2;
    };
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rbar + 8u), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        if (TAB_PF > 0 && n_rows > 4) prefetch_l2_bulk(src_rows + 4 * OREC_BYTES, (uint32_t)min(TAB_PF, n_rows - 4) * OREC_BYTES);
        if (n_rows > 0) arm(0);
        if (n_rows > 0) arm(1);
    }
    double X[3] = {0, 0, 0};
    if (g < g_end && 32 * g + lane < d.M) load3p_256_keep(Xb, 32 * g + lane, X, policy_evict_last());
    __syncthreads();
    mbar_wait(&bar, 0);
#ifdef B200BA_STAMPS
    if (threadIdx.x == 0) { stp[1] = gtime(); stp[3] = clock64(); }
    struct EndStamp { unsigned long long* p; int lane; __device__ ~EndStamp() { if (lane == 0) *p = clock64(); } } end_stamp{stp + 8 + warp, lane};
#endif
    if (g >= g_end) return;
    double u0 = 0, u1 = 0, u2 = 0;
    auto request = [&](double Vi[6], double Xn[3]) {
        if (32 * g + lane < d.M) load6(d.Vinv, 32 * g + lane, Vi);
        if (g + 1 < g_end && 32 * (g + 1) + lane < d.M) load3p_256_keep(Xb, 32 * (g + 1) + lane, Xn, policy_evict_last());
    };
    auto finish = [&](const double Vi[6], const double Xn[3]) -> bool {
        const int j = 32 * g + lane;
        if (j < d.M)
            store4_keep(d.v + 4 * (size_t)j, Vi[0] * u0 + Vi[1] * u1 + Vi[2] * u2, Vi[1] * u0 + Vi[3] * u1 + Vi[4] * u2,
                        Vi[2] * u0 + Vi[4] * u1 + Vi[5] * u2, 0.0, policy_evict_last());
        if (++g >= g_end) return false;
        rows_left = deg_next; X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2]; u0 = u1 = u2 = 0;
        if (g + 1 < g_end) {
            deg_next = __ldg(d.slice_deg + g + 1);
            if (lane == 0) {
                const int nj = min(32, d.M - 32 * (g + 1));
                if (nj > 0) { prefetch_l2_bulk(d.Vinv + 6 * (size_t)(32 * (g + 1)), (uint32_t)nj * 48u); prefetch_l2_bulk(Xb + 4 * (size_t)(32 * (g + 1)), (uint32_t)nj * 32u); }
            }
        }
        return true;
    };
    // r: rows of the stream consumed so far; row r sits in ring slot r & 3, its pair r >> 1 in stage (r >> 1) & 1 with
    // mbarrier parity (r >> 2) & 1.  arrived: pairs known to have landed.  rearm: stage whose two rows the PREVIOUS trip finished
    // reading (it is handed back to the TMA at the top of the next trip, when those reads have long fed their arithmetic).
    int r = 0, arrived = 0, rearm = -1;
    auto need_pair = [&](int pr) {
        if (pr >= arrived) {
            uint32_t ok;
            do {
                asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }"
                             : "=r"(ok) : "r"(rbar + 8u * (uint32_t)(pr & 1)), "r"((uint32_t)(pr >> 1) & 1u) : "memory");
            } while (!ok);
            __syncwarp();
            arrived = pr + 1;
        }
    };
    for (;;) {
        while (rows_left == 0) {
            double Vi[6], Xn[3];
            request(Vi, Xn);
            if (!finish(Vi, Xn)) return;
        }
        if (rearm >= 0) { if (lane == 0 && n_rows > 0) arm(rearm); rearm = -1; }
        const bool two = rows_left >= 2;                          // warp-uniform
        need_pair(r >> 1);
        if (two) need_pair((r + 1) >> 1);
        const uint32_t ca = ring + (uint32_t)(r & 3) * OREC_BYTES + 4u * lane;
        const int camA = lds_s32(ca);
        const double wA = lds_f64(ca + 128u + 4u * lane);
        if (two) {
            const uint32_t cb = ring + (uint32_t)((r + 1) & 3) * OREC_BYTES + 4u * lane;
            const int camB = lds_s32(cb);
            const double wB = lds_f64(cb + 128u + 4u * lane);
            if (r & 1) rearm = (r >> 1) & 1; else rearm = ((r + 1) >> 1) & 1;      // exactly one pair completes: the one whose odd row is r or r + 1
            r += 2; rows_left -= 2;
            const bool last = rows_left == 0;
            TabRow oA, oB;
            tab_row_fwd<false>(tab, tab_pad_cam(camA, d.N), wA, X, d.in, oA);
            tab_row_fwd<false>(tab, tab_pad_cam(camB, d.N), wB, X, d.in, oB);
            double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};
            if (last) request(Vi, Xn);
            tab_row_bwd(oA, u0, u1, u2);
            tab_row_bwd(oB, u0, u1, u2);
            if (last && !finish(Vi, Xn)) return;
        } else {
            if (r & 1) rearm = (r >> 1) & 1;
            r += 1; rows_left = 0;
            TabRow oA;
            tab_row_fwd<false>(tab, tab_pad_cam(camA, d.N), wA, X, d.in, oA);
            double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};
            request(Vi, Xn);
            tab_row_bwd(oA, u0, u1, u2);
            if (!finish(Vi, Xn)) return;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K8c cg_point_pass_tabc: the same sweep with the row stream staged by per-lane cp.async (LDGSTS) instead of TMA bulk copies
//      (B200BA_TAB_ILP=3).  A lane only ever reads ITS OWN camera index and weight of a row, so every lane copies exactly what it
//      will read (4 + 8 bytes per row, coalesced across the warp) into a private ring of TABC_ROWS rows and waits with
//      cp.async.wait_group: no mbarrier, no elected lane, no bulk-copy issue per stage, rows in flight = TABC_ROWS - 1.
//      Measured (Venice, round 2): 51.4 us (TMA ring, one row) / 49.4 (TMA ring, two rows) -> 47.4 (this, 608 threads) ->
//      46.9 (640 threads = 5 warps on every scheduler).  A device-wide chunk queue (guided self-scheduling, claims by atomicAdd)
//      evened out the warps' finish times (spread 16 % -> 5 %) but cost more than it gained (51-55 us): every chunk switch drains
//      and refills the ring.
//      Closing sessions of round 2: the sweep is NOT bound by instruction issue - with 17 % fewer issue
//      cycles per row (101 SASS instructions of which 59 FP64) it ran 2 % faster, with (almost) no arithmetic at all as long
//      (-DB200BA_WHATIF builds); it is a sum of latencies per warp at 5 warps per scheduler.  What moved it: the slice ranges dealt
//      round-robin over the CTAs and a slice end weighted 4/3 of a row on the host (46.6 -> 41.1 us), and the launch as a
//      programmatic dependent of the camera-side step with the static [quaternion | T] piece of the table staged before
//      griddepcontrol.wait (-> 37.7 us).
// ------------------------------------------------------------------------------------------------------
#ifndef B200BA_TABC_BLOCK
#define B200BA_TABC_BLOCK 640
#endif
#ifndef B200BA_TABC_ROWS
#define B200BA_TABC_ROWS 4
#endif
#ifndef B200BA_TABC_PF
#define B200BA_TABC_PF 0      // rows of the stream kept in L2 ahead of the ring (0 = off; 8 .. 32 measured 0.3 .. 2.4 us SLOWER)
#endif
#ifndef B200BA_TABC_EARLY
#define B200BA_TABC_EARLY 0   // 1: the slice-end loads are requested before the last row body instead of in its middle
#endif
#ifndef B200BA_TABC_HINT
#define B200BA_TABC_HINT 0    // bit 1: row stream copied with an L2 evict_first hint, bit 2: (V + damping)^-1 read with evict_last
#endif
#ifndef B200BA_TABC_PFN
#define B200BA_TABC_PFN 4     // rows per bulk L2 prefetch (power of two)
#endif
constexpr int TABC_BLOCK = B200BA_TABC_BLOCK, TABC_ROWS = B200BA_TABC_ROWS, TABC_PF = B200BA_TABC_PF, TABC_PFN = B200BA_TABC_PFN;
constexpr int TABC_RING_BYTES = (TABC_BLOCK / 32) * (TABC_ROWS * OREC_BYTES);
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async4_hint(uint32_t dst, const void* src, uint64_t pol) { asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory"); }
__device__ __forceinline__ void cp_async8_hint(uint32_t dst, const void* src, uint64_t pol) { asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory"); }
__device__ __forceinline__ void load6_hint(const double* __restrict__ base, int i, double v[6], uint64_t pol)
{
    const double* s = base + 6 * (size_t)i;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[0]), "=d"(v[1]) : "l"(s), "l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[2]), "=d"(v[3]) : "l"(s + 2), "l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[4]), "=d"(v[5]) : "l"(s + 4), "l"(pol));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__global__ void __launch_bounds__(TABC_BLOCK, 1) k_cg_point_pass_tabc(Dev d)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    const uint32_t total = (uint32_t)(d.N + 1) * TAB_STRIDE * 8u;    // + the padding entries' dummy camera
    const uint32_t q_bytes = (uint32_t)(d.N + 1) * 64u, p_bytes = (uint32_t)d.N * 48u;     // static piece (with the dummy record) | direction (d.p)
#ifdef B200BA_STAMPS
    unsigned long long* stp = d.stamps + (size_t)blockIdx.x * STAMP_STRIDE;
    if (threadIdx.x == 0) { stp[0] = gtime(); stp[2] = clock64(); }
#endif
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Xb = d.X[st->cur];
    const int wid = blockIdx.x * (TABC_BLOCK / 32) + warp;
    const bool active = wid < d.tab_warps;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0;
    const unsigned char* src_rows = d.o_rec;
    if (active) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; src_rows += (size_t)(ha.z >> 5) * OREC_BYTES; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (TABC_ROWS * OREC_BYTES);
    // every lane: its two words of the next unrequested row -> ring slot; one commit group per row (empty past the end, so that
    // the group arithmetic of wait_group stays uniform)
    const unsigned char* src_lane = src_rows + 4 * lane;          // camera index of this lane in the row; weight at + 128 + 4 lane more
    int fetched = 0;                                              // rows requested so far
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    auto fetch = [&]() {
        if (fetched < n_rows && !(B200BA_WHATIF & 16)) {          // (16 = timing experiment: no row stream at all)
            const uint32_t dst = ring + (uint32_t)(fetched % TABC_ROWS) * OREC_BYTES + 4u * lane;
#if B200BA_TABC_HINT & 1
            cp_async4_hint(dst, src_lane, pol_stream);
            cp_async8_hint(dst + 128u + 4u * lane, src_lane + 128 + 4 * lane, pol_stream);
#else
            cp_async4(dst, src_lane);
            cp_async8(dst + 128u + 4u * lane, src_lane + 128 + 4 * lane);
#endif
            src_lane += OREC_BYTES;
        }
        ++fetched;
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < TABC_ROWS - 1; ++i) fetch();
    // Experiment (TABC_PF > 0, off): the row stream pulled into L2 TABC_PF rows ahead of the ring, TABC_PFN rows per bulk prefetch.
    // The ring holds TABC_ROWS - 1 rows in flight per warp, and a warp needs ~1200 cycles per row, so a row's copy has landed
    // long before its turn (never waiting for the ring changes nothing: -DB200BA_WHATIF=8); the prefetches only cost, 0.3 - 2.4 us.
    if (TABC_PF > 0 && lane == 0 && n_rows > TABC_ROWS - 1)
        prefetch_l2_bulk(src_rows + (size_t)(TABC_ROWS - 1) * OREC_BYTES, (uint32_t)min(TABC_PF, n_rows - (TABC_ROWS - 1)) * OREC_BYTES);
    double X[3] = {0, 0, 0};
    if (g < g_end && 32 * g + lane < d.M) load3p_256_keep(Xb, 32 * g + lane, X, policy_evict_last());
    // Everything above reads what no PCG kernel writes (slice headers, row records, X): with programmatic dependent launch it runs
    // while the camera-side step of the previous PCG step is still in flight.  The control words and the camera table are its output.
    // static piece of the table: in flight while the predecessor (the camera-side step of the previous PCG step) still runs
    if (threadIdx.x == 0) {
        fence_proxy_async();                                      // (mbarrier init -> async proxy)
        mbar_expect_tx(&bar, q_bytes + p_bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.ctab);
        for (uint32_t off = 0; off < q_bytes; off += 32768u) bulk_g2s(smem_raw + off, src + off, min(32768u, q_bytes - off), &bar);
    }
    if (threadIdx.x < 6) reinterpret_cast<double*>(smem_raw + q_bytes)[6 * d.N + threadIdx.x] = 0.0;      // direction of the dummy camera
    pdl_wait();
    if (threadIdx.x == 0) {                                       // (issued even when the sweep is switched off: the barrier must complete)
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.p);
        for (uint32_t off = 0; off < p_bytes; off += 32768u) bulk_g2s(smem_raw + q_bytes + off, src + off, min(32768u, p_bytes - off), &bar);
        pdl_trigger();                                            // the by-camera sweep may take this SM as soon as this CTA leaves it
    }
    if (st->done || !st->cg_active) { cp_async_wait<0>(); __syncthreads(); mbar_wait(&bar, 0); return; }        // uniform over the grid
    __syncthreads();
    mbar_wait(&bar, 0);
#ifdef B200BA_STAMPS
    if (threadIdx.x == 0) { stp[1] = gtime(); stp[3] = clock64(); }
    struct EndStamp { unsigned long long* p; int lane; __device__ ~EndStamp() { if (lane == 0) *p = clock64(); } } end_stamp{stp + 8 + warp, lane};
#endif
    if (g >= g_end) return;
    const uint32_t tab_s = tab_smem_base(smem_raw), p_s = tab_s + q_bytes;
    double u0 = 0, u1 = 0, u2 = 0;
    auto request = [&](double Vi[6], double Xn[3]) {
#if B200BA_WHATIF & 1         // timing experiment (wrong results): no global loads at the end of a slice
        Vi[0] = Vi[1] = Vi[2] = Vi[3] = Vi[4] = Vi[5] = 1.0; Xn[0] = Xn[1] = Xn[2] = 0.5; return;
#endif
#if B200BA_TABC_HINT & 2
        if (32 * g + lane < d.M) load6_hint(d.Vinv, 32 * g + lane, Vi, pol_keep);
#else
        if (32 * g + lane < d.M) load6(d.Vinv, 32 * g + lane, Vi);
#endif
        if (g + 1 < g_end && 32 * (g + 1) + lane < d.M) load3p_256_keep(Xb, 32 * (g + 1) + lane, Xn, pol_keep);
    };
    auto finish = [&](const double Vi[6], const double Xn[3]) -> bool {
        const int j = 32 * g + lane;
        if (j < d.M && (!(B200BA_WHATIF & 32) || u0 == 123.456))   // (32 = timing experiment: v is never written)
            store4_keep(d.v + 4 * (size_t)j, Vi[0] * u0 + Vi[1] * u1 + Vi[2] * u2, Vi[1] * u0 + Vi[3] * u1 + Vi[4] * u2,
                        Vi[2] * u0 + Vi[4] * u1 + Vi[5] * u2, 0.0, policy_evict_last());
        if (++g >= g_end) return false;
        rows_left = deg_next; X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2]; u0 = u1 = u2 = 0;
        if (g + 1 < g_end) {
            deg_next = (B200BA_WHATIF & 128) ? 5 : __ldg(d.slice_deg + g + 1);
            if (lane == 0 && !(B200BA_WHATIF & 64)) {
                const int nj = min(32, d.M - 32 * (g + 1));
                if (nj > 0) { prefetch_l2_bulk(d.Vinv + 6 * (size_t)(32 * (g + 1)), (uint32_t)nj * 48u); prefetch_l2_bulk(Xb + 4 * (size_t)(32 * (g + 1)), (uint32_t)nj * 32u); }
            }
        }
        return true;
    };
    int r = 0;                                                    // rows consumed
    while (rows_left == 0) {                                      // leading slices without observations
        double Vi[6], Xn[3];
        request(Vi, Xn);
        if (!finish(Vi, Xn)) return;
    }
#if B200BA_TABC_EARLY == 2
    // one row: next ring slot, this lane's camera and weight, forward half of the row body
    auto next_row = [&](TabRow& o) {
        fetch();                                                  // row r + TABC_ROWS - 1 into the slot row r - 1 has left
        cp_async_wait<TABC_ROWS - 1>();                           // all but the newest TABC_ROWS - 1 groups have landed: row r is there
        const uint32_t ca = ring + (uint32_t)(r % TABC_ROWS) * OREC_BYTES + 4u * lane;
        const int cam = lds_s32(ca);
        const double w = lds_f64(ca + 128u + 4u * lane);
        ++r;
        tab_row_fwd_split(tab_s, p_s, tab_pad_cam(cam, d.N), w, X, d.in, o);
    };
    for (;;) {                                                    // invariant: rows_left > 0
        TabRow o;
        if (rows_left > 2) {                                      // warp-uniform; the common path carries no slice-end state
            next_row(o);
            tab_row_bwd(o, u0, u1, u2);
            --rows_left;
            continue;
        }
        // the slice ends within the next two rows: what its END needs ((V + damping)^-1 of its points, X of the next slice) is
        // requested now and stays in flight during both row bodies (requested in the middle of the last row it was waited for:
        // 4 us of a 41 us sweep, tools/gpu_r2d.sh whatif)
        double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};
        request(Vi, Xn);
        if (rows_left == 2) { next_row(o); tab_row_bwd(o, u0, u1, u2); }
        next_row(o);
        tab_row_bwd(o, u0, u1, u2);
        if (!finish(Vi, Xn)) return;
        while (rows_left == 0) {                                  // slices without observations
            double Vj[6], Xm[3];
            request(Vj, Xm);
            if (!finish(Vj, Xm)) return;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
#else
    for (;;) {                                                    // invariant: rows_left > 0 (the common path tests it once per row, below)
        fetch();                                                  // row r + TABC_ROWS - 1 into the slot row r - 1 has left
#if !(B200BA_WHATIF & 8)     // 8 = timing experiment (wrong results): never wait for the row stream
        cp_async_wait<TABC_ROWS - 1>();                           // all but the newest TABC_ROWS - 1 groups have landed: row r is there
#endif
        const uint32_t ca = ring + (uint32_t)(r % TABC_ROWS) * OREC_BYTES + 4u * lane;
        const int cam = lds_s32(ca);
        const double w = lds_f64(ca + 128u + 4u * lane);
        ++r;
        if (TABC_PF > 0 && (r & (TABC_PFN - 1)) == 0 && lane == 0) {        // rows [r + PF + ROWS - 1, + PFN) -> L2
            const int first = r + TABC_PF + TABC_ROWS - 1;
            if (first < n_rows) prefetch_l2_bulk(src_rows + (size_t)first * OREC_BYTES, (uint32_t)min(TABC_PFN, n_rows - first) * OREC_BYTES);
        }
        TabRow o;
#if B200BA_TABC_EARLY
        if (--rows_left != 0) {                                   // warp-uniform; the common path carries no slice-end state
            tab_row_fwd_split(tab_s, p_s, tab_pad_cam(cam, d.N), w, X, d.in, o);
            tab_row_bwd(o, u0, u1, u2);
            continue;
        }
        {
            double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};
            request(Vi, Xn);                                      // what the END of the slice needs, in flight during the whole last row body
            tab_row_fwd_split(tab_s, p_s, tab_pad_cam(cam, d.N), w, X, d.in, o);
            tab_row_bwd(o, u0, u1, u2);
            if (!finish(Vi, Xn)) return;
        }
#else
        tab_row_fwd_split(tab_s, p_s, tab_pad_cam(cam, d.N), w, X, d.in, o);
        if (--rows_left != 0) {                                   // warp-uniform; the common path carries no slice-end state
            tab_row_bwd(o, u0, u1, u2);
            continue;
        }
        {
            double Vi[6] = {0, 0, 0, 0, 0, 0}, Xn[3] = {0, 0, 0};
            request(Vi, Xn);                                      // what the END of the slice needs, in flight during the back-rotation
            tab_row_bwd(o, u0, u1, u2);
            if (!finish(Vi, Xn)) return;
        }
#endif
        while (rows_left == 0) {                                  // slices without observations
            double Vi[6], Xn[3];
            request(Vi, Xn);
            if (!finish(Vi, Xn)) return;
        }
    }
}
#endif

// ------------------------------------------------------------------------------------------------------
// K9a/K9b  back-substitution and trial cost on the shared-memory camera table (replaces k_cg_point_pass<2> when the table plan
//      is active).  The generic kernel gathered [R|T] and x per observation from global memory twice (6 + 3 + 6 uncoalesced
//      LDG.128 per observation): 186 us at Venice shape, L1TEX-bound at 89 %.  Here both halves are sweeps of the same structure
//      as k_cg_point_pass_tabc (persistent, one warp = a contiguous slice range, rows staged by per-lane cp.async):
//        k_backsub_tab   table [quaternion | T | x]: u_j = g_p,j - sum_k W_k^t x_i(k), delta_p = (V + lambda I)^-1 u_j, trial
//                        point X + delta_p (updateParametersB, v3d_metricbundle.cpp:42-50), |delta_p|^2, delta_p . g_p
//        k_cost_tab      compact trial table [quaternion | T] of rt[cur ^ 1] (64-byte records, XOR-swizzled so that the bank
//                        group of a field is again a bijection of camera mod 8), rows = camera index + measurement:
//                        residual at the trial point with the weights re-evaluated there (v3d_optimization.cpp:916-921)
//      Per-warp partial sums go to pt_part[4 (block x warps + warp) + {0, 1, 2}] and are reduced in a fixed order.
// ------------------------------------------------------------------------------------------------------
constexpr int QTAB_STRIDE = 8;                                   // doubles per camera of the compact table: quaternion 4, T 3, pad
// physical 16-byte unit of logical field f (0: qw qx, 1: qy qz, 2: T0 T1, 3: T2 pad) of camera c
__host__ __device__ __forceinline__ int qtab_unit(int c, int f) { return 4 * c + (f ^ ((c >> 1) & 3)); }   // (declared above k_build_ptab)
// R (row-major 3x3 at stride 4 of an [R|T] record) -> QUAT_SCALE x unit quaternion (what every camera table carries, see tab_row_fwd)
__device__ __forceinline__ void rot_to_quat(const double* s, double q[4])
{
    double R0 = s[0], R1 = s[1], R2 = s[2], R3 = s[4], R4 = s[5], R5 = s[6], R6 = s[8], R7 = s[9], R8 = s[10];
    double w, x, y, z, tr = R0 + R4 + R8;
    if (tr > 0) { double t = sqrt(tr + 1.0) * 2.0; w = 0.25 * t; x = (R7 - R5) / t; y = (R2 - R6) / t; z = (R3 - R1) / t; }
    else if (R0 > R4 && R0 > R8) { double t = sqrt(1.0 + R0 - R4 - R8) * 2.0; w = (R7 - R5) / t; x = 0.25 * t; y = (R1 + R3) / t; z = (R2 + R6) / t; }
    else if (R4 > R8) { double t = sqrt(1.0 + R4 - R0 - R8) * 2.0; w = (R2 - R6) / t; x = (R1 + R3) / t; y = 0.25 * t; z = (R5 + R7) / t; }
    else { double t = sqrt(1.0 + R8 - R0 - R4) * 2.0; w = (R3 - R1) / t; x = (R2 + R6) / t; y = (R5 + R7) / t; z = 0.25 * t; }
    double n = QUAT_SCALE / sqrt(w * w + x * x + y * y + z * z);
    q[0] = w * n; q[1] = x * n; q[2] = y * n; q[3] = z * n;
}
// compact [quaternion | T] table of camera set `which` (0 / 1 = rt[which]); one thread per camera
__global__ void __launch_bounds__(128) k_build_qtab(Dev d, int trial)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || (trial && !st->solve_ok)) return;
    int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= d.N) return;
    const double* s = d.rt[trial ? (st->cur ^ 1) : st->cur] + 12 * (size_t)i;
    double q[4]; rot_to_quat(s, q);
    double2* o = reinterpret_cast<double2*>(d.qtab);
    o[qtab_unit(i, 0)] = make_double2(q[0], q[1]); o[qtab_unit(i, 1)] = make_double2(q[2], q[3]);
    o[qtab_unit(i, 2)] = make_double2(s[3], s[7]); o[qtab_unit(i, 3)] = make_double2(s[11], 0.0);
}
// delta_c into the p-slots of the packed table (zero for constant cameras): the back-substitution sweep reads it there
__global__ void __launch_bounds__(256) k_ptab_set_x(Dev d)
{
    pdl_chain_enter();
    const LMState* st = d.st;
    if (st->done || !st->solve_ok) return;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= 6 * d.N) return;
    const int i = e / 6, a = e - 6 * i;
    d.ptab[(size_t)i * TAB_STRIDE + 7 + a] = i < d.first_free ? 0.0 : d.x[e];
}

__global__ void __launch_bounds__(TABC_BLOCK, 1) k_backsub_tab(Dev d)
{
    pdl_chain_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->solve_ok) return;
    const uint32_t total = (uint32_t)(d.N + 1) * TAB_STRIDE * 8u;    // + the padding entries' dummy camera
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.ptab);
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cur = st->cur;
    const double* Xb = d.X[cur];
    double* Xt = d.X[cur ^ 1];
    const int wid = blockIdx.x * (TABC_BLOCK / 32) + warp;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0;
    const unsigned char* src_rows = d.o_rec;
    if (wid < d.tab_warps) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; src_rows += (size_t)(ha.z >> 5) * OREC_BYTES; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (TABC_ROWS * OREC_BYTES);
    const unsigned char* src_lane = src_rows + 4 * lane;
    int fetched = 0;
    auto fetch = [&]() {
        if (fetched < n_rows) {
            const uint32_t dst = ring + (uint32_t)(fetched % TABC_ROWS) * OREC_BYTES + 4u * lane;
            cp_async4(dst, src_lane);
            cp_async8(dst + 128u + 4u * lane, src_lane + 128 + 4 * lane);
            src_lane += OREC_BYTES;
        }
        ++fetched;
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < TABC_ROWS - 1; ++i) fetch();
    double X[3] = {0, 0, 0};
    if (g < g_end && 32 * g + lane < d.M) load3p(Xb, 32 * g + lane, X);
    __syncthreads();
    mbar_wait(&bar, 0);
    const uint32_t tab_s = tab_smem_base(smem_raw);
    double d2 = 0, dg = 0, njx = 0;                               // this lane's share of |delta_p|^2, delta_p . g_p, |J delta|^2
    double u0 = 0, u1 = 0, u2 = 0, tt = 0;                        // tt: sum over the point's observations of |A~ delta_c|^2
    const bool pba = st->variant == 1;
    // end of slice g: delta_p, trial point, partial sums; then the next slice's X
    // what the END of slice g needs (requested in the middle of its last row body, like k_cg_point_pass_tabc)
    auto request = [&](double Vi[6], double gp[3], double Xn[3]) {
        const int j = 32 * g + lane;
        if (j < d.M) { load6(d.Vinv, j, Vi); load3p(d.gp, j, gp); }
        if (g + 1 < g_end && j + 32 < d.M) load3p(Xb, j + 32, Xn);
    };
    auto finish = [&](const double Vi[6], const double gp[3], const double Xn[3]) -> bool {
        const int j = 32 * g + lane;
        if (j < d.M) {
            const double a0 = gp[0] - u0, a1 = gp[1] - u1, a2 = gp[2] - u2;
            const double v0 = Vi[0] * a0 + Vi[1] * a1 + Vi[2] * a2;
            const double v1 = Vi[1] * a0 + Vi[3] * a1 + Vi[4] * a2;
            const double v2 = Vi[2] * a0 + Vi[4] * a1 + Vi[5] * a2;
            dg += v0 * gp[0] + v1 * gp[1] + v2 * gp[2];
            if (pba) {
                // |J delta|^2 of this point's observations = sum |A~ dc|^2 + 2 dp . (sum B~^t A~ dc) + dp^t V dp (V undamped), for
                // PBA's gain ratio (SaveUpdatedSystem, SparseBundleCPU.cpp:3681-3739); |dp|^2 in column-scaled coordinates
                double Vu[6]; load6(d.V, j, Vu);
                const double2* D0 = reinterpret_cast<const double2*>(d.pt_D0 + 4 * (size_t)j);
                const double2 da = D0[0], db = D0[1];
                njx += tt + 2.0 * (v0 * u0 + v1 * u1 + v2 * u2)
                     + v0 * (Vu[0] * v0 + Vu[1] * v1 + Vu[2] * v2) + v1 * (Vu[1] * v0 + Vu[3] * v1 + Vu[4] * v2) + v2 * (Vu[2] * v0 + Vu[4] * v1 + Vu[5] * v2);
                d2 += v0 * v0 * da.x + v1 * v1 * da.y + v2 * v2 * db.x;
            } else d2 += v0 * v0 + v1 * v1 + v2 * v2;
            double2* o = reinterpret_cast<double2*>(Xt + 4 * (size_t)j);
            o[0] = make_double2(X[0] + v0, X[1] + v1); o[1] = make_double2(X[2] + v2, 0.0);
        }
        if (++g >= g_end) return false;
        rows_left = deg_next; u0 = u1 = u2 = 0; tt = 0;
        X[0] = Xn[0]; X[1] = Xn[1]; X[2] = Xn[2];
        if (g + 1 < g_end) deg_next = __ldg(d.slice_deg + g + 1);
        return true;
    };
    int r = 0;
    bool more = g < g_end;
    while (more) {
        while (more && rows_left == 0) { double Vi[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0}, Xn[3] = {0, 0, 0}; request(Vi, gp, Xn); more = finish(Vi, gp, Xn); }
        if (!more) break;
        fetch();
        cp_async_wait<TABC_ROWS - 1>();
        const uint32_t ca = ring + (uint32_t)(r % TABC_ROWS) * OREC_BYTES + 4u * lane;
        const int cam = lds_s32(ca);
        const double w = lds_f64(ca + 128u + 4u * lane);
        ++r;
        TabRow o;
        tab_row_fwd_s(tab_s, tab_pad_cam(cam, d.N), w, X, d.in, o);
        if (pba) tt += o.t0 * o.t0 + o.t1 * o.t1;
        if (--rows_left != 0) tab_row_bwd(o, u0, u1, u2);
        else {
            double Vi[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0}, Xn[3] = {0, 0, 0};
            request(Vi, gp, Xn);
            tab_row_bwd(o, u0, u1, u2);
            more = finish(Vi, gp, Xn);
        }
    }
    d2 = warp_sum(d2); dg = warp_sum(dg); njx = warp_sum(njx);
    if (lane == 0) { double* o = d.pt_part + 4 * (size_t)(blockIdx.x * (TABC_BLOCK / 32) + warp); o[0] = d2; o[1] = dg; o[3] = njx; }
}

// trial cost: rows of [camera index | measurement] against the compact trial table
constexpr int COST_ROW_BYTES = 128 + 512;                        // staged per row and warp: 32 x int32 camera | 32 x double2 measurement
constexpr int COST_ROWS = 4;
constexpr int COST_RING_BYTES = (TABC_BLOCK / 32) * (COST_ROWS * COST_ROW_BYTES);
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__global__ void __launch_bounds__(TABC_BLOCK, 1) k_cost_tab(Dev d)
{
    pdl_chain_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->solve_ok) return;
    const double2* tab = reinterpret_cast<const double2*>(smem_raw);
    const uint32_t total = (uint32_t)d.N * QTAB_STRIDE * 8u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.qtab);
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Xt = d.X[st->cur ^ 1];
    const int wid = blockIdx.x * (TABC_BLOCK / 32) + warp;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0, e0 = 0;
    if (wid < d.tab_warps) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; e0 = ha.z; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (COST_ROWS * COST_ROW_BYTES);
    const unsigned char* src_cam = d.o_rec + (size_t)(e0 >> 5) * OREC_BYTES + 4 * lane;
    const double2* src_meas = d.o_meas + e0 + lane;
    int fetched = 0;
    auto fetch = [&]() {
        if (fetched < n_rows) {
            const uint32_t dst = ring + (uint32_t)(fetched % COST_ROWS) * COST_ROW_BYTES;
            cp_async4(dst + 4u * lane, src_cam);
            cp_async16(dst + 128u + 16u * lane, src_meas);
            src_cam += OREC_BYTES; src_meas += 32;
        }
        ++fetched;
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < COST_ROWS - 1; ++i) fetch();
    double X[3] = {0, 0, 1};
    if (g < g_end && 32 * g + lane < d.M) load3p(Xt, 32 * g + lane, X);
    __syncthreads();
    mbar_wait(&bar, 0);
    double cost = 0;
    const Intrinsics& in = d.in;
    int r = 0;
    auto next_slice = [&]() -> bool {
        if (++g >= g_end) return false;
        rows_left = deg_next;
        X[0] = X[1] = 0.0; X[2] = 1.0;
        if (32 * g + lane < d.M) load3p(Xt, 32 * g + lane, X);
        if (g + 1 < g_end) deg_next = __ldg(d.slice_deg + g + 1);
        return true;
    };
    bool more = g < g_end;
    while (more) {
        while (more && rows_left == 0) more = next_slice();
        if (!more) break;
        fetch();
        cp_async_wait<COST_ROWS - 1>();
        const uint32_t ca = ring + (uint32_t)(r % COST_ROWS) * COST_ROW_BYTES;
        const int cam = lds_s32(ca + 4u * lane);
        const double2 m = *reinterpret_cast<const double2*>(smem_raw + (ca - smem_u32(smem_raw)) + 128u + 16u * lane);
        ++r;
        const bool pad = cam < 0;
        const int c = pad ? 0 : cam;
        const double2 a0 = tab[qtab_unit(c, 0)], a1 = tab[qtab_unit(c, 1)], a2 = tab[qtab_unit(c, 2)], a3 = tab[qtab_unit(c, 3)];
        const double qw = a0.x, qx = a0.y, qy = a1.x, qz = a1.y;
        const double tx = qy * X[2] - qz * X[1], ty = qz * X[0] - qx * X[2], tz = qx * X[1] - qy * X[0];
        const double r0 = fma(qw, tx, fma(qy, tz, fma(-qz, ty, X[0])));      // scaled quaternion, see tab_row_fwd
        const double r1 = fma(qw, ty, fma(qz, tx, fma(-qx, tz, X[1])));
        const double r2 = fma(qw, tz, fma(qx, ty, fma(-qy, tx, X[2])));
        const double x = r0 + a2.x, y = r1 + a2.y, z = pad ? 1.0 : r2 + a3.x;
        const double iz = fast_rcp(z);
        const double uu = x * iz, vv = y * iz;
        const double ex = in.k00 * uu + in.k01 * vv + in.k02 - m.x;
        const double ey = in.k11 * vv + in.k12 - m.y;
        const double w = huber_weight(ex, ey, in.huber);        // weights re-evaluated at the trial point (:917)
        const double wa = w * ex, wb = w * ey;
        if (!pad) cost += wa * wa + wb * wb;
        if (--rows_left == 0) more = next_slice();
    }
    cost = warp_sum(cost);
    if (lane == 0) d.pt_part[4 * (size_t)(blockIdx.x * (TABC_BLOCK / 32) + warp) + 2] = cost;
}

// K1t linearize_points on the compact shared-memory table (replaces k_linearize_points when the table plan is active): the same
//     sweep structure as k_cost_tab - persistent, rows of [camera index | measurement] staged by per-lane cp.async, [quaternion | T]
//     of the CURRENT cameras in shared memory - instead of 6 uncoalesced LDG.128 of [R|T] per observation (111 us at Venice shape,
//     L1TEX-bound at 65 %).  Residual + Huber weight (written into the point-order row record), cost, V_j = sum B~^t B~,
//     g_p,j = sum B~^t (-w e)  (v3d_optimization.cpp:512-524, 543-558, 616-630).  R is rebuilt from the unit quaternion per
//     observation (9 products); per-warp partials (cost, max |g_p|, max diag V) are reduced in a fixed order.
__global__ void __launch_bounds__(TABC_BLOCK, 1) k_lin_tab(Dev d)
{
    pdl_chain_enter();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const LMState* st = d.st;
    if (st->done || !st->compute) return;
    const double2* tab = reinterpret_cast<const double2*>(smem_raw);
    const uint32_t total = (uint32_t)d.N * QTAB_STRIDE * 8u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.qtab);
        for (uint32_t off = 0; off < total; off += 32768u) {
            uint32_t n = total - off < 32768u ? total - off : 32768u;
            bulk_g2s(smem_raw + off, src + off, n, &bar);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* Xb = d.X[st->cur];
    const int wid = blockIdx.x * (TABC_BLOCK / 32) + warp;
    int g = 0, g_end = 0, n_rows = 0, rows_left = 0, deg_next = 0, e0 = 0;
    if (wid < d.tab_warps) {
        const int4 ha = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid), hb = __ldg(reinterpret_cast<const int4*>(d.tab_wr) + 2 * wid + 1);
        g = ha.x; g_end = ha.y; e0 = ha.z; n_rows = (ha.w - ha.z) >> 5; rows_left = hb.x; deg_next = hb.y;
    }
    const uint32_t tab_bytes = (total + 127u) & ~127u;
    const uint32_t ring = smem_u32(smem_raw) + tab_bytes + (uint32_t)warp * (COST_ROWS * COST_ROW_BYTES);
    const unsigned char* src_cam = d.o_rec + (size_t)(e0 >> 5) * OREC_BYTES + 4 * lane;
    const double2* src_meas = d.o_meas + e0 + lane;
    double* w_out = reinterpret_cast<double*>(d.o_rec + (size_t)(e0 >> 5) * OREC_BYTES + 128) + lane;    // weight slot of this lane in row 0 of the range
    int fetched = 0;
    auto fetch = [&]() {
        if (fetched < n_rows) {
            const uint32_t dst = ring + (uint32_t)(fetched % COST_ROWS) * COST_ROW_BYTES;
            cp_async4(dst + 4u * lane, src_cam);
            cp_async16(dst + 128u + 16u * lane, src_meas);
            src_cam += OREC_BYTES; src_meas += 32;
        }
        ++fetched;
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < COST_ROWS - 1; ++i) fetch();
    double X[3] = {0, 0, 1};
    if (g < g_end && 32 * g + lane < d.M) load3p(Xb, 32 * g + lane, X);
    __syncthreads();
    mbar_wait(&bar, 0);
    double cost = 0, gmax = 0, vmax = -1e30;
    double V[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
    const Intrinsics& in = d.in;
    int r = 0;
    auto next_slice = [&]() -> bool {
        const int j = 32 * g + lane;
        if (j < d.M) {
            double2* Vo = reinterpret_cast<double2*>(d.V + 6 * (size_t)j);
            Vo[0] = make_double2(V[0], V[1]); Vo[1] = make_double2(V[2], V[3]); Vo[2] = make_double2(V[4], V[5]);
            double2* go = reinterpret_cast<double2*>(d.gp + 4 * (size_t)j);
            go[0] = make_double2(gp[0], gp[1]); go[1] = make_double2(gp[2], 0.0);
            gmax = fmax(gmax, fmax(fabs(gp[0]), fmax(fabs(gp[1]), fabs(gp[2]))));
            vmax = fmax(vmax, fmax(V[0], fmax(V[3], V[5])));
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) V[q] = 0.0;
        gp[0] = gp[1] = gp[2] = 0.0;
        if (++g >= g_end) return false;
        rows_left = deg_next;
        X[0] = X[1] = 0.0; X[2] = 1.0;
        if (32 * g + lane < d.M) load3p(Xb, 32 * g + lane, X);
        if (g + 1 < g_end) deg_next = __ldg(d.slice_deg + g + 1);
        return true;
    };
    bool more = g < g_end;
    while (more) {
        while (more && rows_left == 0) more = next_slice();
        if (!more) break;
        fetch();
        cp_async_wait<COST_ROWS - 1>();
        const uint32_t ca = ring + (uint32_t)(r % COST_ROWS) * COST_ROW_BYTES;
        const int cam = lds_s32(ca + 4u * lane);
        const double2 m = *reinterpret_cast<const double2*>(smem_raw + (ca - smem_u32(smem_raw)) + 128u + 16u * lane);
        double* wp = w_out + (size_t)r * (OREC_BYTES / 8);
        ++r;
        if (cam >= 0) {                                           // padding entries only occur where track lengths differ inside a slice
            const double2 a0 = tab[qtab_unit(cam, 0)], a1 = tab[qtab_unit(cam, 1)], a2 = tab[qtab_unit(cam, 2)], a3 = tab[qtab_unit(cam, 3)];
            const double qw = a0.x, qx = a0.y, qy = a1.x, qz = a1.y;
            Cam cm;                                               // R from the sqrt(2)-scaled quaternion: the factors 2 come with the components
            const double wx = qw * qx, wy = qw * qy, wz = qw * qz;
            cm.R[0] = fma(-qz, qz, fma(-qy, qy, 1.0)); cm.R[1] = fma(qx, qy, -wz); cm.R[2] = fma(qx, qz, wy);
            cm.R[3] = fma(qx, qy, wz); cm.R[4] = fma(-qz, qz, fma(-qx, qx, 1.0)); cm.R[5] = fma(qy, qz, -wx);
            cm.R[6] = fma(qx, qz, -wy); cm.R[7] = fma(qy, qz, wx); cm.R[8] = fma(-qy, qy, fma(-qx, qx, 1.0));
            cm.T[0] = a2.x; cm.T[1] = a2.y; cm.T[2] = a3.x;
            Obs o; obs_full(cm, X, m, in, o);
            *wp = o.w;
            const double we0 = o.w * o.e0, we1 = o.w * o.e1;
            cost += we0 * we0 + we1 * we1;
            double B0[3], B1[3]; jac_rows_B(o, cm, B0, B1);
#pragma unroll
            for (int c = 0; c < 3; ++c) gp[c] += B0[c] * (-we0) + B1[c] * (-we1);
            V[0] += B0[0] * B0[0] + B1[0] * B1[0]; V[1] += B0[0] * B0[1] + B1[0] * B1[1]; V[2] += B0[0] * B0[2] + B1[0] * B1[2];
            V[3] += B0[1] * B0[1] + B1[1] * B1[1]; V[4] += B0[1] * B0[2] + B1[1] * B1[2]; V[5] += B0[2] * B0[2] + B1[2] * B1[2];
        }
        if (--rows_left == 0) more = next_slice();
    }
    cost = warp_sum(cost); gmax = warp_max(gmax); vmax = warp_max(vmax);
    if (lane == 0) { double* o = d.pt_part + 4 * (size_t)(blockIdx.x * (TABC_BLOCK / 32) + warp); o[0] = cost; o[1] = gmax; o[2] = vmax; o[3] = 0.0; }
}

// ------------------------------------------------------------------------------------------------------
// K11' cg_camera_step: the whole camera side of one PCG step in ONE cooperative kernel (grid-wide syncs instead of
//      three kernel boundaries): second-level segment sums, q = (U + lambda I) p - sum W v, alpha, x, r, z = M r, beta, p.
//      One thread per (camera, component); a block owns 32 cameras.  Dot products: fixed-order block partials,
//      summed in the same fixed order by every warp, so every thread sees identical scalars.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double grid_total(const double* part, int n)
{
    double s = 0;
    for (int b = (threadIdx.x & 31); b < n; b += 32) s += __ldcg(part + b);
    return warp_sum(s);
}
// Grid-wide barrier of the (cooperative, co-resident) camera-side step: one atomic per block on a counter that only grows, the
// waiters poll it with acquire loads.  cooperative_groups' grid.sync() measured ~3.9 us per call on 148 blocks against ~1.3 us
// for this form (tools/dense_stamps.py); the step kernel has two of them in each of the 50 PCG steps of an LM iteration.
// bar[0] = arrivals so far (every launch that reaches the barriers adds exactly 2 x gridDim.x), bar[1] = launches completed.
__device__ __forceinline__ void step_barrier_arrive(unsigned long long* ctr)
{
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(ctr, 1ull); }
}
__device__ __forceinline__ void step_barrier(unsigned long long* ctr, unsigned long long target)
{
    step_barrier_arrive(ctr);
    if (threadIdx.x == 0) {
        unsigned long long v;
        do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
}
// phases of the camera-side step (after the by-camera partial sums exist); every thread of the grid must call it
// NT = threads per block; FUSED: called at the end of the by-camera sweep (k_cg_camera_pass_step, NT = 512, only the first
// STEP_BLOCK threads of the first step_blocks blocks own elements): one more barrier up front ("every segment partial of the
// sweep is visible") replaces the kernel boundary, and the loads that do not depend on the sweep are in flight while it waits.
#ifndef B200BA_SEG_BATCH
#define B200BA_SEG_BATCH 8
#endif
constexpr int SEG_BATCH = B200BA_SEG_BATCH;
template <int NT, bool FUSED>
__device__ __forceinline__ void camera_step_phases(const Dev& d, int sums_ready,
                                                   double* part_pq, double* part_rz, int step_blocks, double* sv, double* sm,
                                                   double lam, double rz, double rz0, double tol, int steps)
{
    LMState* st = d.st;
    unsigned long long* const bar = d.step_bar + (FUSED ? 2 : 0);                                // [arrivals, launches completed] of this kernel
    constexpr unsigned long long NBAR = FUSED ? 3ull : 2ull;
    const unsigned long long gen = __ldcg(bar + 1), bar_base = NBAR * gen * gridDim.x;           // read by every block before it arrives at the first barrier
    const int e = blockIdx.x * STEP_BLOCK + threadIdx.x, n6 = 6 * d.N;
    const bool on = threadIdx.x < STEP_BLOCK && (int)blockIdx.x < step_blocks && e < n6;
    const int i = e / 6, a = e - 6 * i, lc = threadIdx.x - a;    // lc: first thread of this camera inside the block
    // Every load whose address does not depend on this step's scalars is issued up front (U row, M^-1 row, x, r, segment
    // range): the kernel is a chain of dependent L2 round trips with almost no parallelism, so overlapping them is what counts.
    double pe = 0.0, xe0 = 0.0, re0 = 0.0, urow[6] = {0, 0, 0, 0, 0, 0}, mrow[6] = {0, 0, 0, 0, 0, 0}, dmul = 1.0;
    int sg0 = 0, sg1 = 0;
    if (on) {
        pe = d.p[e]; xe0 = d.x[e]; re0 = d.r[e];
        if (sums_ready != 1) { sg0 = d.cam_seg_beg[i]; sg1 = d.cam_seg_beg[i + 1]; }
        const double* ur = d.U + 36 * (size_t)i + 6 * a; const double* mr = d.Minv + 36 * (size_t)i + 6 * a;
#pragma unroll
        for (int c = 0; c < 6; ++c) { urow[c] = ur[c]; mrow[c] = mr[c]; }
        if (st->variant == 1) dmul = ur[a];                      // PBA: lambda diag(U) p instead of lambda p
    }
    if (threadIdx.x < STEP_BLOCK) sv[threadIdx.x] = pe;
    __syncthreads();
    if (FUSED) step_barrier(bar, bar_base + gridDim.x);          // the sweep's segment partials of every block are visible
    if (!FUSED) pdl_wait();                                      // launched as a programmatic dependent of the by-camera sweep (single rank): everything
                                                                 // above reads what that sweep does not write, the segment partials below are its output
    double q = 0, s = 0;
    if (on) {
        if (sums_ready == 1) s = d.ar_q[e];
        else {
            // the first SEG_BATCH partials are loaded together (a camera has 2.3 x VSPLIT segments on average at Venice shape; the plain loop waited
            // for one L2 round trip per segment), summed in the same order as before
            auto ld = [&](int sg) { return FUSED ? __ldcg(d.seg_part + PART_STRIDE * (size_t)sg + a) : d.seg_part[PART_STRIDE * (size_t)sg + a]; };
            double v[SEG_BATCH];
#pragma unroll
            for (int k = 0; k < SEG_BATCH; ++k) v[k] = (sg0 + k < sg1) ? ld(sg0 + k) : 0.0;
            s = v[0];
#pragma unroll
            for (int k = 1; k < SEG_BATCH; ++k) s += v[k];
            for (int sg = sg0 + SEG_BATCH; sg < sg1; ++sg) s += ld(sg);
        }
    }
    unsigned long long coll = 0;
    if (sums_ready == 2) {
        // multi-rank: the all-reduce of the 6N camera sums happens HERE, over NVLink peer memory, between the second-level
        // sum and the q = S p assembly - push into every window, raise this block's flags, poll the own window, sum its slots
        coll = *d.p2p.seq;
        if (on) p2p_push(d.p2p, coll, (size_t)e, s);
        __syncthreads();                                         // this block's pushes are ordered before its flags (CTA barrier + release)
        p2p_raise_flags(d.p2p, coll, 1 + (int)blockIdx.x);
        p2p_wait_all(d.p2p, coll, 1 + (int)blockIdx.x);           // block b of every rank pushes exactly the elements block b reads
        if (on) s = p2p_sum(d.p2p, coll, (size_t)e);
    }
    if (on) q = lam * dmul * pe + (urow[0] * sv[lc] + urow[1] * sv[lc + 1] + urow[2] * sv[lc + 2] + urow[3] * sv[lc + 3] + urow[4] * sv[lc + 4] + urow[5] * sv[lc + 5]) - s;
    double bp = block_sum<NT>(pe * q, sm);
    if (threadIdx.x == 0 && (int)blockIdx.x < step_blocks) part_pq[blockIdx.x] = bp;
    step_barrier(bar, bar_base + (NBAR - 1ull) * gridDim.x);
    const double pq = grid_total(part_pq, step_blocks);
    if (!(pq > 0) || !isfinite(pq)) {                            // breakdown (uniform): keep x; first-step breakdown = failed solve
        step_barrier_arrive(bar);                                // (keeps the arrival count at NBAR per block and launch)
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->cg_active = 0; if (steps == 0) st->solve_ok = 0; if (sums_ready == 2) *d.p2p.seq = coll + 1ull; *(bar + 1) = gen + 1ull; }
        return;
    }
    const double alpha = rz / pq;
    double re = 0, xe = 0;
    if (on) { xe = xe0 + alpha * pe; re = re0 - alpha * q; d.x[e] = xe; d.r[e] = re; }
    __syncthreads();
    if (threadIdx.x < STEP_BLOCK) sv[threadIdx.x] = re;
    __syncthreads();
    double z = on ? (mrow[0] * sv[lc] + mrow[1] * sv[lc + 1] + mrow[2] * sv[lc + 2] + mrow[3] * sv[lc + 3] + mrow[4] * sv[lc + 4] + mrow[5] * sv[lc + 5]) : 0.0;
    double br = block_sum<NT>(re * z, sm);
    if (threadIdx.x == 0 && (int)blockIdx.x < step_blocks) part_rz[blockIdx.x] = br;
    step_barrier(bar, bar_base + NBAR * gridDim.x);
    const double rzn = grid_total(part_rz, step_blocks);
    const double rel = sqrt(fabs(rzn) / rz0);
    const bool stop = (tol > 0 && rel < tol && steps + 1 >= st->cg_min_steps) || !(rzn > 0) || (st->cg_guard > 0 && rel > st->cg_guard);
    if (!stop && on) {
        double pn = z + (rzn / rz) * pe;
        d.p[e] = pn; d.ptab[(size_t)i * TAB_STRIDE + 7 + a] = pn;
        if (d.ptab32) d.ptab32[(size_t)i * 20 + 8 + a] = (float)pn;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->cg_steps = steps + 1; st->total_cg += 1; st->cg_rel = rel; st->rz = rzn;
        if (stop) st->cg_active = 0;
        if (sums_ready == 2) *d.p2p.seq = coll + 1ull;           // every thread read the counter before the first grid sync
        *(bar + 1) = gen + 1ull;                                 // every block read the launch count before the first barrier
    }
}
__global__ void __launch_bounds__(STEP_BLOCK) k_cg_camera_step(Dev d, int sums_ready, double* part_pq, double* part_rz)
{
    __shared__ double sv[STEP_BLOCK];
    __shared__ double sm[STEP_BLOCK / 32];
    const LMState* st = d.st;
    if (threadIdx.x == 0) pdl_trigger();                          // the by-point sweep of the next PCG step may start its prologue
    if (st->done || !st->cg_active) return;                      // uniform over the grid
    camera_step_phases<STEP_BLOCK, false>(d, sums_ready, part_pq, part_rz, gridDim.x, sv, sm, st->lambda, st->rz, st->rz0, st->cg_tol, st->cg_steps);
}
// K9 + K11' in one launch (single rank, cooperative: one CTA per SM): the by-camera sweep, then - after a grid-wide counting
// barrier instead of a kernel boundary - the camera-side step in the first step_blocks blocks.  Two launches per PCG step
// instead of three.  (Round 1 tried this with cooperative_groups' grid.sync(), ~3.9 us per call on 148 CTAs: no gain.)
__global__ void __launch_bounds__(CW * 32, 1) k_cg_camera_pass_step(Dev d, double* part_pq, double* part_rz, int step_blocks)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sv[STEP_BLOCK];
    __shared__ double sm[CW];
    const LMState* st = d.st;
    if (st->done || !st->cg_active) return;                      // uniform over the grid
    const double lam = st->lambda, rz = st->rz, rz0 = st->rz0, tol = st->cg_tol; const int steps = st->cg_steps;
    camera_pass_sweep(d, st, smem_raw);
    camera_step_phases<CW * 32, true>(d, 0, part_pq, part_rz, step_blocks, sv, sm, lam, rz, rz0, tol, steps);
}

// K12 cg_finish: delta_c = x; trial cameras T += dt, R <- Rodrigues(dw) R (v3d_metricbundle.cpp:17-39,
//     v3d_mathutilities.h:173-194); camera part of |delta|^2 and delta . J^t e.
constexpr int FIN_BLOCK = 128;
// fuse_tabs: also delta_c into the p-slots of the packed table (k_ptab_set_x) and the compact table of the TRIAL cameras
// (k_build_qtab(d, 1)) - three launches in one for the explicit plan
__device__ __forceinline__ void rot_to_quat(const double* s, double q[4]);
__global__ void __launch_bounds__(FIN_BLOCK) k_cg_finish(Dev d, double* __restrict__ part, int fuse_tabs)
{
    pdl_chain_enter();
    __shared__ double sm[FIN_BLOCK / 32];
    const LMState* st = d.st;
    if (st->done) return;
    const bool ok = st->solve_ok != 0;
    const int cur = st->cur;
    double d2 = 0, dg = 0;
    int bad = 0;
    const int i = blockIdx.x * FIN_BLOCK + threadIdx.x;
    if (ok && i < d.N) {
        double x[6];
        for (int a = 0; a < 6; ++a) {
            x[a] = d.x[6 * (size_t)i + a]; if (i < d.first_free) x[a] = 0.0;
            d2 += x[a] * x[a] * (st->variant == 1 ? d.cam_D0[6 * (size_t)i + a] : 1.0);      // PBA measures the update in column-scaled coordinates
            dg += x[a] * d.gc[6 * (size_t)i + a]; if (!isfinite(x[a])) bad = 1;
        }
        const double* s = d.rt[cur] + 12 * (size_t)i; double* o = d.rt[cur ^ 1] + 12 * (size_t)i;
        double R[9] = {s[0], s[1], s[2], s[4], s[5], s[6], s[8], s[9], s[10]};
        double T[3] = {s[3] + x[0], s[7] + x[1], s[11] + x[2]};
        double w0 = x[3], w1 = x[4], w2 = x[5];
        double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
        double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (fabs(th) > 1e-6) {
            double J[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0}, J2[9];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J2[3 * r + c] = J[3 * r] * J[c] + J[3 * r + 1] * J[3 + c] + J[3 * r + 2] * J[6 + c];
            double c1 = sin(th) / th, c2 = (1.0 - cos(th)) / (th * th);
            for (int a = 0; a < 9; ++a) dR[a] += c1 * J[a] + c2 * J2[a];
        }
        double Rt[12];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rt[4 * r + c] = dR[3 * r] * R[c] + dR[3 * r + 1] * R[3 + c] + dR[3 * r + 2] * R[6 + c];
            Rt[4 * r + 3] = T[r];
        }
        for (int a = 0; a < 12; ++a) o[a] = Rt[a];
        if (fuse_tabs) {
            for (int a = 0; a < 6; ++a) d.ptab[(size_t)i * TAB_STRIDE + 7 + a] = x[a];          // (already zero for constant cameras)
            double q[4]; rot_to_quat(Rt, q);
            double2* t = reinterpret_cast<double2*>(d.qtab);
            t[qtab_unit(i, 0)] = make_double2(q[0], q[1]); t[qtab_unit(i, 1)] = make_double2(q[2], q[3]);
            t[qtab_unit(i, 2)] = make_double2(Rt[3], Rt[7]); t[qtab_unit(i, 3)] = make_double2(Rt[11], 0.0);
        }
    }
    d2 = block_sum<FIN_BLOCK>(d2, sm);
    dg = block_sum<FIN_BLOCK>(dg, sm);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) { double* o = part + 4 * (size_t)blockIdx.x; o[0] = d2; o[1] = dg; o[2] = bad ? 1.0 : 0.0; o[3] = 0.0; }
}
// camera part of |delta|^2 and delta . J^t e: fixed-order sum of k_cg_finish's block partials (called by the decide kernels)
__device__ __forceinline__ void camera_update_sums(const Dev& d, LMState* st)
{
    double d2 = 0, dg = 0; bool bad = false;
    const int nb = (d.N + FIN_BLOCK - 1) / FIN_BLOCK;
    for (int b = 0; b < nb; ++b) { d2 += d.fin_part[4 * b]; dg += d.fin_part[4 * b + 1]; bad = bad || d.fin_part[4 * b + 2] != 0.0; }
    st->d2c = d2; st->denomc = dg; if (bad) st->solve_ok = 0;
}

// K13 lm_decide (one thread): update stop test, gain ratio, accept/reject, damping update
//     (v3d_optimization.cpp:879-978, v3d_optimization.h:49-50,60-64).
// fuse_count >= 0 (single rank, launched with ONE_CTA threads): the per-warp partials of the back-substitution / trial-cost
// sweeps (|delta_p|^2, delta_p . g_p, trial cost, |J delta|^2) are reduced here instead of by a k_point_partials_reduce launch
__device__ __forceinline__ void decide_reduce(const Dev& d, int fuse_count, double* sm)
{
    double a = 0, b = 0, c = 0, e4 = 0;
    for (int i = threadIdx.x; i < fuse_count; i += ONE_CTA) { const double* p = d.pt_part + 4 * (size_t)i; a += p[0]; b += p[1]; c += p[2]; e4 += p[3]; }
    a = block_sum<ONE_CTA>(a, sm); b = block_sum<ONE_CTA>(b, sm); c = block_sum<ONE_CTA>(c, sm); e4 = block_sum<ONE_CTA>(e4, sm);
    if (threadIdx.x == 0) { d.ar_dec[0] = a; d.ar_dec[1] = b; d.ar_dec[2] = c; d.ar_dec[3] = e4; }
}
__global__ void k_lm_decide(Dev d, int fuse_count)
{
    pdl_chain_enter();
    __shared__ double sm[ONE_CTA / 32];
    LMState* st = d.st;
    if (st->done) return;
    if (fuse_count >= 0) { decide_reduce(d, fuse_count, sm); if (threadIdx.x != 0) return; }
    if (st->solve_ok) camera_update_sums(d, st);
    int it = st->iter;
    double* tr = (d.trace && st->n_trace < st->trace_cap) ? d.trace + TRACE_COLS * (size_t)st->n_trace : nullptr;
    double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (tr) { tr[0] = it; tr[1] = st->err; tr[2] = st->lambda; tr[3] = nan; tr[4] = nan; tr[5] = 0; tr[6] = st->cg_steps; tr[7] = st->cg_rel; st->n_trace += 1; }
    bool accept = false;
    if (st->solve_ok) {
        double d2 = st->d2c + d.ar_dec[0], denom2 = st->denomc + d.ar_dec[1], new_err = d.ar_dec[2];
        if (it >= st->min_iter && sqrt(d2) < st->upd_thr * (st->param_length + st->upd_thr)) {
            st->status = 1; st->done = 1;
            if (tr) tr[5] = -1;
            return;
        }
        double rho = (st->err - new_err) / (st->lambda * d2 + denom2);
        st->new_err = new_err; st->rho = rho;
        accept = rho > 0;
        if (tr) { tr[3] = new_err; tr[4] = rho; tr[5] = accept ? 1 : 0; }
    }
    if (accept) { st->lambda = fmax(1e-10, st->lambda * 0.1); st->compute = 1; st->cur ^= 1; st->n_accept += 1; }
    else        { st->lambda *= 10.0; st->compute = 0; }
    st->iter = it + 1;
    if (it + 1 >= st->max_iter) st->done = 1;
}

// K13' lm_decide_pba (one thread): PBA's step control (NonlinearOptimizeLM, SparseBundleCPU.cpp:3772-3949; options.lm_variant = 1).
//      quit when the update is small (|delta| in the column-scaled coordinates of PrepareJacobianNormalization, <= 1e-6: 'S');
//      accept iff the squared error decreases; gain ratio rho = reduction / (2 delta.g - |J delta|^2) (SaveUpdatedSystem :3681-3739,
//      __accurate_gain_ratio; a non-positive denominator is replaced by 0.001 x reduction); quit when the mean squared error
//      is <= 0.25 px^2 ('M'); damping *= max(1/3, 1 - (2 rho - 1)^3) after an accepted step and *= 2, 4, 8, ... after rejected
//      ones, clamped to [1e-10, 1e5].  PBA keeps these scalars in fp32; here they are fp64 (tolerance in the parity tests).
__global__ void k_lm_decide_pba(Dev d, int fuse_count)
{
    pdl_chain_enter();
    __shared__ double sm[ONE_CTA / 32];
    LMState* st = d.st;
    if (st->done) return;
    if (fuse_count >= 0) { decide_reduce(d, fuse_count, sm); if (threadIdx.x != 0) return; }
    if (st->solve_ok) camera_update_sums(d, st);
    const int it = st->iter;
    double* tr = (d.trace && st->n_trace < st->trace_cap) ? d.trace + TRACE_COLS * (size_t)st->n_trace : nullptr;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (tr) { tr[0] = it; tr[1] = st->err; tr[2] = st->lambda; tr[3] = nan; tr[4] = nan; tr[5] = 0; tr[6] = st->cg_steps; tr[7] = st->cg_rel; st->n_trace += 1; }
    st->first = 0;
    bool accept = false;
    if (st->solve_ok) {
        const double d2 = st->d2c + d.ar_dec[0], dxtg = st->denomc + d.ar_dec[1], new_err = d.ar_dec[2], njx = d.ar_dec[3];
        if (st->norm_scale * sqrt(d2) <= st->delta_thr) {          // "quit on too small change"
            st->status = 1; st->done = 1;
            if (tr) tr[5] = -1;
            return;
        }
        // PBA keeps the squared error in fp32 (float _projection_sse, new_residual: SparseBundleCPU.cpp:3855-3859), so an improvement
        // below fp32 resolution counts as "no better solution"; mirrored here, otherwise the run would keep accepting at the noise floor
        const double reduction = (double)((float)st->err - (float)new_err);
        st->new_err = new_err;
        if (tr) tr[3] = new_err;
        if (isfinite(new_err) && reduction > 0) {
            double expected = 2.0 * dxtg - njx;
            if (expected <= 0) expected = 0.001 * reduction;
            const double rho = reduction / expected;
            st->rho = rho; accept = true;
            if (tr) { tr[4] = rho; tr[5] = 1; }
            st->compute = 1; st->cur ^= 1; st->n_accept += 1;
            if (new_err * st->mse_scale <= st->mse_thr) { st->status = 3; st->done = 1; st->iter = it + 1; st->err = new_err; return; }   // "satisfies MSE threshold"
            const double t = 2.0 * rho - 1.0;
            st->lambda = fmin(1e5, fmax(1e-10, st->lambda * fmax(1.0 / 3.0, 1.0 - t * t * t)));
            st->damp_adjust = 2.0;
        }
    }
    if (!accept) { st->lambda *= st->damp_adjust; st->damp_adjust *= 2.0; st->compute = 0; }
    st->iter = it + 1;
    if (it + 1 >= st->max_iter) st->done = 1;
}

// mean reprojection error in pixels, the reference's own self-check (src/sfm_ba_ssba.cpp:26-51):
// block partial sums of f0 * |e_k| (unweighted), thread per point.
__global__ void __launch_bounds__(PT_BLOCK) k_reproj_sum(Dev d, int which)
{
    pdl_chain_enter();
    __shared__ double sm[PT_BLOCK / 32];
    int j = blockIdx.x * PT_BLOCK + threadIdx.x;
    double s = 0;
    if (j < d.M) {
        double X[3]; load3p(d.X[which], j, X);
        int k0, deg; slice_of(d, j, k0, deg);
        for (int sl = 0, k = k0; sl < deg; ++sl, k += 32) {
            int i = __ldg(o_cam_ptr(d, k));
            if (i < 0) continue;
            Cam cm; load_cam(d.rt[which], i, cm);
            double e0, e1, x, y, iz, r0, r1, r2;
            obs_residual(cm, X, __ldg(d.o_meas + k), d.in, e0, e1, x, y, iz, r0, r1, r2);
            s += d.in.f0 * sqrt(e0 * e0 + e1 * e1);
        }
    }
    s = block_sum<PT_BLOCK>(s, sm);
    if (threadIdx.x == 0) d.pt_part[4 * (size_t)blockIdx.x] = s;
}

} // namespace b200ba
