// ba_engine.cu -- host side of libb200ba.so: problem upload, LM driver, C ABI (include/b200ba.h).
//
// One handle = one CUDA device (+ optionally one NCCL rank).  The host never touches an observation after
// upload: it only enqueues kernels; accept/reject, damping and every stop decision live in the device-side
// LMState (ba_kernels.cuh), polled every `poll_every` LM iterations.
//
// Reference lines replaced by each entry point are cited in include/b200ba.h.
#include "../../include/b200ba.h"
#include "ba_kernels.cuh"
#include "ba_kernels_f32.cuh"
#include "ba_dense.cuh"
#include "dense_host.h"
#include "pinned_pool.h"
#include "layout_host.h"

#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b200ba;

namespace b200ba { PinnedPool g_pinned; }

#ifndef B200BA_TAB_ILP_DEFAULT
#define B200BA_TAB_ILP_DEFAULT 3
#endif
#ifndef B200BA_PDL_DEFAULT
#define B200BA_PDL_DEFAULT 2
#endif
#ifndef B200BA_SLICE_W_NUM
#define B200BA_SLICE_W_NUM 4
#endif
constexpr int SLICE_W_NUM = B200BA_SLICE_W_NUM, SLICE_W_DEN = 3;   // epilogue of a slice, in row bodies

namespace {

thread_local std::string g_create_error;

// ---- NCCL through dlopen: the library has no link-time dependency on NCCL (single-GPU use needs none) ----
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (so) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
        if (!so) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(so, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(so, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(so, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(so, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(so, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
NcclApi g_nccl;

// One device arena per problem: every buffer is carved out of a single cudaMalloc (40 separate allocations cost
// 40-100 ms per set_problem) and zeroed with a single memset.
struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0, allocated = 0;   // cap: bytes of the current problem; allocated: grow-only size
    void* take(size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return base ? (void*)(base + o) : nullptr; }
};
template <class T> struct DBuf {
    T* p = nullptr; size_t n = 0; bool owned = false;
    cudaError_t alloc(size_t count)
    {
        release(); n = count; owned = true;
        if (!count) return cudaSuccess;
        return cudaMalloc((void**)&p, sizeof(T) * count);
    }
    cudaError_t take(Arena& a, size_t count) { release(); n = count; owned = false; p = (T*)a.take(sizeof(T) * count); return cudaSuccess; }
    void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = false; }
};



// fork-join helper for the host-side layout build (std::thread: no OpenMP runtime dependency)
template <class F> void parallel_for(int nthreads, F&& fn)
{
    if (nthreads <= 1) { fn(0, 1); return; }
    std::vector<std::thread> th; th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back([&fn, t, nthreads] { fn(t, nthreads); });
    fn(0, nthreads);
    for (auto& x : th) x.join();
}

} // namespace

struct b200ba_handle {
    b200ba_options opt{};
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // communicator
    ncclComm_t comm = nullptr; int rank = 0, world = 1;
    // NVLink peer-memory exchange (struct P2P in ba_kernels.cuh)
    unsigned char* win_own = nullptr; void* win_peer[P2P_MAX_WORLD] = {nullptr}; size_t win_cap = 0; bool p2p_ready = false;
    unsigned long long* p2p_ctl = nullptr;           // [seq | done_blocks | err] device words
    // host copies of the caller's problem
    int N = 0, M = 0, K = 0;
    std::vector<double> h_cam, h_pts;        // as given (un-normalised)
    std::vector<double> s_xy; std::vector<int32_t> s_cam, s_pt; int s_K = -1;   // options.keep_problem: the marshalled observations (b200ba_append)
    double K5[5] = {0, 0, 0, 0, 0};
    double mean[3] = {0, 0, 0}, scale = 1.0;
    bool have_problem = false, begun = false, finished = false;
    // device
    Dev d{};
    DBuf<double> rt0, rt1, U, gc, Minv, x, r, z, p, q, ar_lin, ar_q, ar_sd, ar_dec, X0, X1, V, gp, Vinv, v, seg_part, pt_part, trace;
    DBuf<double2> o_meas, c_meas;
    DBuf<int> o_cam, slice_base, slice_deg, cam_seg_beg, sort_vals, row_tab;
    DBuf<unsigned long long> sort_keys;
    DBuf<unsigned char> sort_tmp;
    DBuf<int2> row_hdr;
    DBuf<unsigned char> c_rec, o_rec;
    DBuf<double> fin_part, ptab, qtab, ctab, part_pq, part_rz, X_init, rt_init, cam_D0, pt_D0;
    double pba_depth_scale = 1.0;                     // normalize = 2: PBA's depth scaling of T and X
    bool use_tab_backsub = false; size_t cost_smem = 0;          // linearisation / back-substitution / trial cost as table sweeps
    int use_pdl = 0;                                 // programmatic dependent launch of the PCG sweeps (B200BA_PDL = 1: both, 2: by-point only, 3: by-camera only)
    bool quat_ok = true;                             // the caller's rotations are orthonormal (checked by b200ba_begin): quaternion tables may stand in for R
    DBuf<unsigned long long> stamps, step_bar;
    DBuf<int> tab_wr, tab_wr32;
    DBuf<unsigned char> o_rec32, c_rec32; DBuf<float4> X32, v32; DBuf<float> Vinv32, ptab32, rt32;
    bool use_f32 = false; int tab32_grid = 0; size_t tab32_smem = 0;
    // explicit reduced camera system (small problems): pair lists on the device, S, the per-observation blocks
    bool pdl_chain = false;                          // programmatic dependent launch of every kernel of the iteration (explicit plan)
    bool step_pdl = false;                           // camera-side step as a programmatic dependent of the by-camera sweep (single rank)
    bool fuse_small = false;                         // small kernels of the LM iteration fused (single rank: no all-reduce between them)
    bool use_dense = false; int dense_ctas = 0, dense_kind = -1, dense_threads = 0; bool dense_grid = false; size_t dense_smem = 0; long long dense_pairs = 0;
    DBuf<int2> dn_pairs; DBuf<int4> dn_chunks; DBuf<int> dn_blk_chunk_beg, dn_blk_ik, dn_entry_pt;
    DBuf<double> dn_Wobs, dn_Yobs, dn_Tobs, dn_chunk_part, dn_chunk_rhs, dn_S, dn_zg, dn_pq, dn_rz; DBuf<unsigned int> dn_bar;
    std::vector<int> orig_of_new;                    // SELL permutation (new point index -> caller's)
    std::unique_ptr<int[]> entry_of_obs;             // obs -> padded SELL entry
    int Kp = 0;                                      // padded SELL entries
    int num_sms = 148, tab_grid = 0, step_grid = 0, cam_grid = 0;
    size_t tab_smem = 0;
    bool use_tab = false, use_coop = false, use_graph = false, fuse_step = false;
    bool use_wb = false;                             // large-camera-count plan: B~ records + world-axes direction table (k_cg_point_pass_wB)
    DBuf<double> obs_B, wtab;
    int tab_ilp = 1, tab_block = TAB_BLOCK;          // rows of a slice in flight per trip of the shared-memory-table sweep (1 or 2)
    cudaGraphExec_t graph_exec = nullptr;
    long long launches_per_graph = 0;
    DBuf<LMState> st;
    DBuf<double> norm_buf;
    Arena arena;
    LMState h_st{}, h_st_init{};
    long long launches = 0;
    double ms_h2d = 0, ms_d2h = 0, ms_solve = 0, ms_last_iterate = 0, ms_last_restart = 0;
    long long iters_before_restart = 0, accepts_before_restart = 0;   // totals of the runs ended by b200ba_restart
    double mean_px[2] = {0, 0};
    double initial_cost = 0;
    int trace_cap = 0;
};

namespace {

int fail(b200ba_handle* h, int code, const char* fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, B200BA_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(h, B200BA_E_COMM, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"); } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Sum over ranks, in place.  `need`: 0 = unconditional, 1 = skipped on device once the solve is done, 2 = skipped as well
// when the iteration does not re-linearise (the same device-side flags on every rank).  With the peer windows mapped this
// is two small kernels (no host-side collective call, graph-capturable); otherwise NCCL.
int allreduce(b200ba_handle* h, double* buf, size_t count, int need = 0)
{
    if (h->world <= 1) return 0;
    if (h->p2p_ready) {
        if (count > h->win_cap) return fail(h, B200BA_E_COMM, "peer window too small (%zu > %zu doubles)", count, h->win_cap);
        const int grid = (int)std::min<size_t>(64, (count + 255) / 256);
        k_p2p_publish<<<grid, 256, 0, h->stream>>>(h->d, buf, (int)count, need);
        k_p2p_reduce<<<grid, 256, 0, h->stream>>>(h->d, buf, (int)count, need);
        h->launches += 2;
        return 0;
    }
    NC(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream));
    return 0;
}

// ---- kernel launch sequences ---------------------------------------------------------------------------
// Every kernel of an LM iteration goes through here.  h->pdl_chain (explicit small-problem plan): programmatic dependent launch,
// see pdl_chain_enter() in ba_kernels.cuh.  EVERY kernel that can be launched through this helper starts with pdl_chain_enter()
// (or pdl_wait()): one that did not could start before its predecessor has finished.
template <class... KA, class... A>
void launch_chain(b200ba_handle* h, void (*kern)(KA...), int grid, int block, size_t smem, A&&... a)
{
    if (h->pdl_chain) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(a)...);
    } else kern<<<grid, block, smem, h->stream>>>(static_cast<KA>(a)...);
    h->launches++;
}
#define LAUNCH(kern, grid, block, ...) launch_chain(h, kern, (grid), (block), 0, __VA_ARGS__)

// by-point PCG sweep against the shared-memory camera table (one or two rows in flight per trip)
void launch_point_tab(b200ba_handle* h)
{
    if (h->tab_ilp == 3 && (h->use_pdl == 1 || h->use_pdl == 2)) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        cfg.gridDim = dim3(h->tab_grid); cfg.blockDim = dim3(TABC_BLOCK); cfg.dynamicSmemBytes = h->tab_smem; cfg.stream = h->stream;
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_cg_point_pass_tabc, h->d);
    }
    else if (h->tab_ilp == 3) k_cg_point_pass_tabc<<<h->tab_grid, TABC_BLOCK, h->tab_smem, h->stream>>>(h->d);
    else if (h->tab_ilp == 2) k_cg_point_pass_tab2<<<h->tab_grid, TAB2_BLOCK, h->tab_smem, h->stream>>>(h->d);
    else k_cg_point_pass_tab<<<h->tab_grid, TAB_BLOCK, h->tab_smem, h->stream>>>(h->d);
    h->launches++;
}

// by-point PCG sweep: shared-memory camera table (<= ~1 797 cameras), else B~ records + world-axes direction table
// (k_cg_point_pass_wB; needs orthonormal input rotations like every table plan), else the generic packed-table gathers
void launch_point_sweep(b200ba_handle* h, bool allow_tab = true)
{
    Dev& d = h->d;
    if (h->use_tab && allow_tab) { launch_point_tab(h); return; }
    if (h->use_wb && h->quat_ok) {
        LAUNCH(k_build_wtab, cdiv(d.N, 128), 128, d);
        LAUNCH(k_cg_point_pass_wB, d.n_pt_blocks, PT_BLOCK, d);
        return;
    }
    LAUNCH(k_cg_point_pass<0>, d.n_pt_blocks, PT_BLOCK, d);
}

int enqueue_linearize(b200ba_handle* h)
{
    Dev& d = h->d;
    const bool tabs = h->use_tab_backsub && h->quat_ok;
    const bool fuse = h->fuse_small;                              // single rank: fused small kernels (explicit plan 20 -> 14, matrix-free plan 174 -> 168 launches per iteration)
    const int lin_parts = tabs ? h->tab_grid * (TABC_BLOCK / 32) : d.n_pt_blocks;
    LAUNCH(k_build_ptab, cdiv(d.N, 128), 128, d, (fuse && tabs) ? 1 : 0);
    if (tabs) {                                                  // table plan: by-point linearisation as a shared-memory-table sweep
        if (!fuse) LAUNCH(k_build_qtab, cdiv(d.N, 128), 128, d, 0);
        launch_chain(h, k_lin_tab, h->tab_grid, TABC_BLOCK, h->cost_smem, d);
    } else LAUNCH(k_linearize_points, d.n_pt_blocks, PT_BLOCK, d);
    launch_chain(h, k_linearize_cameras, cdiv(d.n_vwarps, CH_BLOCK / 32), CH_BLOCK, LINCAM_SMEM, d);
    if (fuse) {
        LAUNCH(k_camera_reduce_unpack, cdiv((long long)d.N * 27 * 4, 256), 256, d);
        LAUNCH(k_lm_prepare, 1, ONE_CTA, d, lin_parts);
    } else {
        LAUNCH((k_camera_reduce<27, PART_STRIDE>), cdiv((long long)d.N * 27 * 4, 256), 256, d, d.ar_lin, 1);
        LAUNCH(k_point_partials_reduce, 1, ONE_CTA, d, 0, tabs ? lin_parts : -1);
        if (int rc = allreduce(h, d.ar_lin, (size_t)d.N * PART_STRIDE + 1 + 2 * h->world, 2)) return rc;
        LAUNCH(k_lm_unpack, cdiv(36 * (long long)d.N, 256), 256, d);
        LAUNCH(k_lm_prepare, 1, ONE_CTA, d, -1);
    }
    if (h->use_f32) {                                            // fp32 copies for the mixed-precision PCG sweeps
        LAUNCH(k_f32_points, cdiv(d.M, 256), 256, d, 0);
        LAUNCH(k_f32_cameras, cdiv(d.N, 128), 128, d);
        LAUNCH(k_f32_orows, cdiv(h->Kp, 256), 256, d, h->Kp);
        LAUNCH(k_f32_crows, cdiv(d.n_rows, 8), 256, d);
    }
    return 0;
}

int enqueue_precond(b200ba_handle* h)
{
    Dev& d = h->d;
    LAUNCH(k_point_vinv, cdiv(d.M, 256), 256, d, 1);           // + v = (V + damping)^-1 g_p for enqueue_cg_init
    if (h->use_f32) LAUNCH(k_f32_points, cdiv(d.M, 256), 256, d, 1);
    if (h->opt.precond == 1) {
        LAUNCH(k_schur_diag_cameras, cdiv(d.n_vwarps, CH_BLOCK / 32), CH_BLOCK, d);
        LAUNCH((k_camera_reduce<21, 21>), cdiv((long long)d.N * 21 * 4, 256), 256, d, d.ar_sd, 0);
        if (int rc = allreduce(h, d.ar_sd, (size_t)d.N * 21, 1)) return rc;
    }
    LAUNCH(k_precond_invert, cdiv(6ll * d.N, 192), 192, d, h->opt.precond);
    return 0;
}

int enqueue_camera_half(b200ba_handle* h, int rhs_pass);
int enqueue_camera_sums(b200ba_handle* h);
const void* dense_pcg_kernel(int kind);

int enqueue_cg_init(b200ba_handle* h)
{
    Dev& d = h->d;
    // (v = (V + damping)^-1 g_p was written by k_point_vinv in enqueue_precond)
    if (int rc = enqueue_camera_half(h, 1)) return rc;
    if (int rc = enqueue_camera_sums(h)) return rc;
    LAUNCH(k_cg_init, h->step_grid, STEP_BLOCK, d, h->part_rz.p);
    LAUNCH(k_cg_init_finish, 1, 32, d, h->part_rz.p, h->step_grid);
    return 0;
}

// sum_k W_k v_j by camera (persistent row sweep, TMA-staged row records)
int enqueue_camera_half(b200ba_handle* h, int rhs_pass)
{
    Dev& d = h->d;
    if ((h->use_pdl == 1 || h->use_pdl == 3) && !rhs_pass) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
        cfg.gridDim = dim3(h->cam_grid); cfg.blockDim = dim3(CW * 32); cfg.dynamicSmemBytes = CAMPASS_SMEM; cfg.stream = h->stream;
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_cg_camera_pass, d, rhs_pass);
    } else k_cg_camera_pass<<<h->cam_grid, CW * 32, CAMPASS_SMEM, h->stream>>>(d, rhs_pass);
    h->launches++;
    return 0;
}
// second level -> ar_q (+ all-reduce over ranks)
int enqueue_camera_sums(b200ba_handle* h)
{
    Dev& d = h->d;
    LAUNCH((k_camera_reduce<6, 6>), cdiv((long long)d.N * 6 * 4, 256), 256, d, d.ar_q, 0);
    return allreduce(h, d.ar_q, (size_t)d.N * 6, 1);
}

int launch_camera_step(b200ba_handle* h, int sums_ready)
{
    Dev& d = h->d;
    double* ppq = h->part_pq.p; double* prz = h->part_rz.p;
    void* args[] = {(void*)&d, (void*)&sums_ready, (void*)&ppq, (void*)&prz};
    // Single rank: the step is launched as a programmatic dependent of the by-camera sweep (which triggers at its top): its blocks
    // take the SMs the sweep's CTAs leave and have their loads of U, M^-1, x, r, p in flight when the sweep ends (LM iteration
    // 5.010 -> 4.984 ms at Venice shape).  B200BA_NO_STEP_PDL=1 or more than one rank: plain cooperative launch.
    if (h->step_pdl) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[2];
        cfg.gridDim = dim3(h->step_grid); cfg.blockDim = dim3(STEP_BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        CU(cudaLaunchKernelExC(&cfg, (const void*)k_cg_camera_step, args));
    } else CU(cudaLaunchCooperativeKernel((const void*)k_cg_camera_step, dim3(h->step_grid), dim3(STEP_BLOCK), args, 0, h->stream));
    h->launches++;
    return 0;
}

int enqueue_cg_step(b200ba_handle* h)
{
    Dev& d = h->d;
    if (h->use_f32) {                                            // mixed-precision PCG: both sweeps in fp32, the step kernel as usual
        k_cg_point_pass_tab32<<<h->tab32_grid, TAB32_BLOCK, h->tab32_smem, h->stream>>>(d);
        k_cg_camera_pass32<<<h->cam_grid, CW * 32, CAMPASS32_SMEM, h->stream>>>(d);
        h->launches += 2;
        if (!h->use_coop) {                                      // no cooperative launch: second-level sums + single-CTA step
            if (int rc = enqueue_camera_sums(h)) return rc;
            LAUNCH(k_cg_update, 1, ONE_CTA, d);
            return 0;
        }
        if (h->world <= 1) return launch_camera_step(h, 0);
        if (h->p2p_ready && h->step_grid <= P2P_MAX_BLOCKS) return launch_camera_step(h, 2);
        if (int rc = enqueue_camera_sums(h)) return rc;
        return launch_camera_step(h, 1);
    }
    launch_point_sweep(h);
    if (h->fuse_step) {                                          // single rank: by-camera sweep + camera-side step in one cooperative launch
        double* ppq = h->part_pq.p; double* prz = h->part_rz.p; int sb = h->step_grid;
        void* args[] = {(void*)&d, (void*)&ppq, (void*)&prz, (void*)&sb};
        CU(cudaLaunchCooperativeKernel((const void*)k_cg_camera_pass_step, dim3(h->cam_grid), dim3(CW * 32), args, CAMPASS_SMEM, h->stream));
        h->launches++;
        return 0;
    }
    if (int rc = enqueue_camera_half(h, 0)) return rc;
    if (h->use_coop && h->world <= 1) return launch_camera_step(h, 0);
    if (h->use_coop && h->p2p_ready && h->step_grid <= P2P_MAX_BLOCKS) return launch_camera_step(h, 2);   // all-reduce fused into the step kernel (NVLink peer memory)
    if (int rc = enqueue_camera_sums(h)) return rc;
    if (h->use_coop) return launch_camera_step(h, 1);
    LAUNCH(k_cg_update, 1, ONE_CTA, d);
    return 0;
}

// back-substitution, trial point, trial cost -> ar_dec [d2p, denomp, new_cost]
void enqueue_backsub(b200ba_handle* h)
{
    Dev& d = h->d;
    if (h->use_tab_backsub && h->quat_ok) {
        if (!h->fuse_small) {                                    // (fused: both written by k_cg_finish)
            LAUNCH(k_ptab_set_x, cdiv(6 * (long long)d.N, 256), 256, d);
            LAUNCH(k_build_qtab, cdiv(d.N, 128), 128, d, 1);
        }
        launch_chain(h, k_backsub_tab, h->tab_grid, TABC_BLOCK, h->tab_smem, d);
        launch_chain(h, k_cost_tab, h->tab_grid, TABC_BLOCK, h->cost_smem, d);
        if (!h->fuse_small) LAUNCH(k_point_partials_reduce, 1, ONE_CTA, d, 1, h->tab_grid * (TABC_BLOCK / 32));
    } else {
        LAUNCH(k_cg_point_pass<2>, d.n_pt_blocks, PT_BLOCK, d);
        if (!h->fuse_small) LAUNCH(k_point_partials_reduce, 1, ONE_CTA, d, 1, -1);
    }
}

int enqueue_step_and_decide(b200ba_handle* h)
{
    Dev& d = h->d;
    const bool tabs = h->use_tab_backsub && h->quat_ok;
    LAUNCH(k_cg_finish, cdiv(d.N, FIN_BLOCK), FIN_BLOCK, d, h->fin_part.p, (h->fuse_small && tabs) ? 1 : 0);
    enqueue_backsub(h);
    if (int rc = allreduce(h, d.ar_dec, 4, 1)) return rc;
    const int dec_parts = h->fuse_small ? (tabs ? h->tab_grid * (TABC_BLOCK / 32) : d.n_pt_blocks) : -1;
    const int dec_threads = h->fuse_small ? ONE_CTA : 1;
    if (h->opt.lm_variant == 1) LAUNCH(k_lm_decide_pba, 1, dec_threads, d, dec_parts); else LAUNCH(k_lm_decide, 1, dec_threads, d, dec_parts);
    return 0;
}

int enqueue_lm_iteration_raw(b200ba_handle* h);

// One LM iteration = a fixed kernel sequence (all data-dependent decisions are device-side flags), so in the
// fixed-step mode it is captured once into a CUDA graph and replayed.
int enqueue_lm_iteration(b200ba_handle* h)
{
    // (the convergence-controlled PCG of the matrix-free plan polls a device flag between batches of steps: not capturable;
    //  the explicit plan runs its whole PCG loop, stop rule included, inside one kernel)
    if (!h->use_graph || (h->opt.cg_rel_tol > 0 && !h->use_dense)) return enqueue_lm_iteration_raw(h);
    if (!h->graph_exec) {
        cudaGraph_t g = nullptr;
        long long before = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); h->use_graph = false; return enqueue_lm_iteration_raw(h); }
        int rc = enqueue_lm_iteration_raw(h);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        h->launches_per_graph = h->launches - before;
        h->launches = before;
        if (rc || e != cudaSuccess || !g || cudaGraphInstantiate(&h->graph_exec, g, 0) != cudaSuccess) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            h->graph_exec = nullptr; h->use_graph = false;
            return enqueue_lm_iteration_raw(h);
        }
        cudaGraphDestroy(g);
    }
    CU(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->launches_per_graph;
    return 0;
}

const void* dense_pcg_kernel(int kind)
{
    return kind == 0 ? (const void*)k_dense_pcg<0> : kind == 1 ? (const void*)k_dense_pcg<1> : (const void*)k_dense_pcg<2>;
}

// Explicit plan (small problems): S assembled from the pair lists, preconditioner from its diagonal blocks, the whole PCG
// loop in one cluster / cooperative kernel (ba_dense.cuh).
int enqueue_dense_solve(b200ba_handle* h)
{
    Dev& d = h->d;
    LAUNCH(k_dense_wobs, cdiv(h->Kp, 256), 256, d, h->Kp);      // (V + damping)^-1 included
    if (d.dn.n_chunks) LAUNCH(k_dense_pairs, cdiv(d.dn.n_chunks, DPAIR_BLOCK / 36), DPAIR_BLOCK, d);
    LAUNCH(k_dense_finalize, cdiv(36ll * d.dn.n_blocks, 256), 256, d);
    LAUNCH(k_precond_invert, cdiv(6ll * d.N, 192), 192, d, h->opt.precond);
    int max_steps = h->opt.cg_max_steps;
    const void* fn = dense_pcg_kernel(h->dense_kind);
    void* args[] = {(void*)&d, (void*)&max_steps};
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[2];
    cfg.gridDim = dim3(h->dense_ctas); cfg.blockDim = dim3(h->dense_threads); cfg.dynamicSmemBytes = h->dense_smem; cfg.stream = h->stream;
    if (h->dense_grid) { at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1; }
    else { at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = h->dense_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; }
    cfg.attrs = at; cfg.numAttrs = 1;
    if (h->pdl_chain) { at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1; cfg.numAttrs = 2; }
    CU(cudaLaunchKernelExC(&cfg, fn, args));
    h->launches++;
    return 0;
}

int enqueue_lm_iteration_raw(b200ba_handle* h)
{
    if (int rc = enqueue_linearize(h)) return rc;
    if (h->use_dense) {
        if (int rc = enqueue_dense_solve(h)) return rc;
        return enqueue_step_and_decide(h);
    }
    if (int rc = enqueue_precond(h)) return rc;
    if (int rc = enqueue_cg_init(h)) return rc;
    const int steps = h->opt.cg_max_steps;
    if (h->opt.cg_rel_tol > 0) {
        // convergence-controlled PCG: poll the device flag between batches of steps
        int done_steps = 0;
        while (done_steps < steps) {
            int batch = std::min(25, steps - done_steps);
            for (int s = 0; s < batch; ++s) if (int rc = enqueue_cg_step(h)) return rc;
            done_steps += batch;
            int active = 0;
            CU(cudaMemcpyAsync(&active, &h->d.st->cg_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (!active) break;
        }
    } else {
        for (int s = 0; s < steps; ++s) if (int rc = enqueue_cg_step(h)) return rc;
    }
    return enqueue_step_and_decide(h);
}

int mean_reproj(b200ba_handle* h, int which, double* out)
{
    Dev& d = h->d;
    if (d.K == 0 && h->world <= 1) { *out = 0; return 0; }
    LAUNCH(k_reproj_sum, d.n_pt_blocks, PT_BLOCK, d, which);
    LAUNCH(k_point_partials_reduce, 1, ONE_CTA, d, 2, -1);
    double k_local = (double)d.K;
    CU(cudaMemcpyAsync(d.ar_dec + 1, &k_local, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (int rc = allreduce(h, d.ar_dec, 4)) return rc;
    double s[2];
    CU(cudaMemcpyAsync(s, d.ar_dec, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = s[1] > 0 ? s[0] / s[1] : 0.0;
    return 0;
}

// host-side global sums for the normaliser: local sums go through one tiny all-reduce when world > 1
int global_sum(b200ba_handle* h, double* vals, int n)
{
    if (h->world <= 1) return 0;
    CU(cudaMemcpyAsync(h->norm_buf.p, vals, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (int rc = allreduce(h, h->norm_buf.p, n)) return rc;
    CU(cudaMemcpyAsync(vals, h->norm_buf.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

void camera_centre(const double* P, double* C)   // -R^t T
{
    for (int c = 0; c < 3; ++c) C[c] = -(P[c] * P[3] + P[4 + c] * P[7] + P[8 + c] * P[11]);
}
void set_camera_centre(double* P, const double* C)   // T = -R C   (v3d_cameramatrix.h:132)
{
    for (int r = 0; r < 3; ++r) P[4 * r + 3] = -1.0 * (P[4 * r] * C[0] + P[4 * r + 1] * C[1] + P[4 * r + 2] * C[2]);
}

} // namespace

// =========================================================================================================
extern "C" {

int b200ba_version(void) { return B200BA_VERSION; }

int b200ba_default_options(b200ba_options* o)
{
    if (!o) return B200BA_E_INVALID;
    memset(o, 0, sizeof *o);
    o->struct_size = (int32_t)sizeof *o;
    o->device = -1;
    o->max_iterations = 50; o->min_iterations = 10; o->tau = 1e-3; o->inlier_px = 2.0;
    o->gradient_threshold = 1e-8; o->update_threshold = 1e-8;
    o->round_pixels = 1; o->normalize = 1; o->fix_first_camera = 0; o->discard_converged = 1;
    o->cg_max_steps = 50; o->cg_rel_tol = 0.0; o->precond = 1; o->poll_every = 5; o->lane_window = 64;
    return 0;
}

int b200ba_pba_preset(b200ba_options* o)
{
    if (int rc = b200ba_default_options(o)) return rc;
    o->fix_first_camera = 1;      // camera_data[0].SetConstantCamera()   src/sfm_ba_pba.cpp:94
    o->inlier_px = 0.0;           // PBA has no robustifier                SparseBundleCPU.cpp:1176-1180
    o->round_pixels = 0;          // float sub-pixel measurements          src/sfm_ba_pba.cpp:122-123
    o->discard_converged = 0;     // adjustBundle_pba always returns 0     src/sfm_ba_pba.cpp:173
    return 0;
}

int b200ba_pba_lm_preset(b200ba_options* o)
{
    if (int rc = b200ba_pba_preset(o)) return rc;
    o->lm_variant = 1;            // NonlinearOptimizeLM                   SparseBundleCPU.cpp:3772-3949
    o->normalize = 2;             // NormalizeDataD (depth scaling)         SparseBundleCPU.cpp:3189-3282
    o->precond = 0;               // block-Jacobi of the damped U blocks    SparseBundleCPU.cpp:1641-1783
    o->cg_max_steps = 100; o->cg_min_steps = 10; o->cg_rel_tol = 0.1;      // __cg_max/min_iteration, __cg_norm_threshold (ConfigBA.cpp:58-62)
    o->gradient_threshold = 0.0;  // __lm_check_gradient = false            ConfigBA.cpp:50
    o->min_iterations = 0;
    return 0;
}

const char* b200ba_last_error(const b200ba_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int b200ba_create(const b200ba_options* opt, b200ba_handle** out)
{
    if (!out) return B200BA_E_INVALID;
    *out = nullptr;
    b200ba_handle* h = nullptr;
    b200ba_options o;
    if (opt) { if (opt->struct_size != (int32_t)sizeof o) return fail(nullptr, B200BA_E_INVALID, "options.struct_size mismatch"); o = *opt; }
    else b200ba_default_options(&o);
    if (o.max_iterations < 0 || o.cg_max_steps < 0) return fail(nullptr, B200BA_E_INVALID, "negative iteration count");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, B200BA_E_CUDA, "no usable CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    int dev = o.device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) return fail(nullptr, B200BA_E_INVALID, "device %d out of range (%d devices)", dev, ndev);
    h = new b200ba_handle();
    h->opt = o; h->device = dev;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        fail(nullptr, B200BA_E_CUDA, "cannot initialise device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
        delete h; return B200BA_E_CUDA;
    }
    {
        // experiment knob: L2 set-aside for persisting (evict_last) lines, in MB (-1 = the device maximum)
        const char* e = getenv("B200BA_L2_PERSIST");
        if (e) {
            int maxp = 0; cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
            long long want = atoll(e) < 0 ? (long long)maxp : atoll(e) << 20;
            if (want > maxp) want = maxp;
            cudaError_t pe = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want);
            if (getenv("B200BA_VERBOSE")) fprintf(stderr, "[b200ba] persisting L2: max %d bytes, set %lld bytes (%s)\n", maxp, want, cudaGetErrorString(pe));
            cudaGetLastError();
        }
    }
    *out = h;
    return 0;
}

void b200ba_destroy(b200ba_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < P2P_MAX_WORLD; ++r) if (h->win_peer[r]) cudaIpcCloseMemHandle(h->win_peer[r]);
    if (h->win_own) cudaFree(h->win_own);
    if (h->p2p_ctl) cudaFree(h->p2p_ctl);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    h->rt0.release(); h->rt1.release(); h->U.release(); h->gc.release(); h->Minv.release(); h->x.release(); h->r.release();
    h->z.release(); h->p.release(); h->q.release(); h->ar_lin.release(); h->ar_q.release(); h->ar_sd.release(); h->ar_dec.release();
    h->X0.release(); h->X1.release(); h->V.release(); h->gp.release(); h->Vinv.release(); h->v.release(); h->o_rec.release();
    h->seg_part.release(); h->pt_part.release(); h->trace.release(); h->o_meas.release(); h->c_meas.release(); h->o_cam.release();
    h->sort_keys.release(); h->sort_vals.release(); h->row_tab.release(); h->sort_tmp.release(); h->row_hdr.release(); h->c_rec.release(); h->cam_seg_beg.release();
    h->st.release(); h->norm_buf.release(); h->slice_base.release(); h->slice_deg.release(); h->ptab.release(); h->ctab.release(); h->part_pq.release(); h->part_rz.release();
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->arena.base) cudaFree(h->arena.base);
    h->tab_wr.release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int b200ba_comm_unique_id(void* id_bytes)
{
    if (!id_bytes) return B200BA_E_INVALID;
    if (!g_nccl.load()) return fail(nullptr, B200BA_E_COMM, "libnccl.so.2 not loadable: %s", dlerror());
    static_assert(sizeof(ncclUniqueId) == B200BA_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return fail(nullptr, B200BA_E_COMM, "ncclGetUniqueId failed");
    memcpy(id_bytes, &id, sizeof id);
    return 0;
}

int b200ba_comm_init(b200ba_handle* h, const void* id_bytes, int32_t rank, int32_t world)
{
    if (!h || !id_bytes || world < 1 || rank < 0 || rank >= world) return B200BA_E_INVALID;
    if (h->have_problem) return fail(h, B200BA_E_STATE, "b200ba_comm_init must precede b200ba_set_problem");
    if (world == 1) { h->rank = 0; h->world = 1; return 0; }
    if (!g_nccl.load()) return fail(h, B200BA_E_COMM, "libnccl.so.2 not loadable: %s", dlerror());
    CU(cudaSetDevice(h->device));
    ncclUniqueId id; memcpy(&id, id_bytes, sizeof id);
    NC(g_nccl.CommInitRank(&h->comm, world, id, rank));
    h->rank = rank; h->world = world;
    return 0;
}

int b200ba_comm_peer_handle(b200ba_handle* h, int32_t max_cameras, void* handle_bytes)
{
    if (!h || !handle_bytes || max_cameras <= 0) return B200BA_E_INVALID;
    if (h->world <= 1) return fail(h, B200BA_E_STATE, "b200ba_comm_peer_handle needs a multi-rank handle (b200ba_comm_init first)");
    if (h->world > P2P_MAX_WORLD) return fail(h, B200BA_E_INVALID, "peer windows support at most %d ranks", P2P_MAX_WORLD);
    if (h->have_problem) return fail(h, B200BA_E_STATE, "b200ba_comm_peer_handle must precede b200ba_set_problem");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200BA_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    CU(cudaSetDevice(h->device));
    if (h->win_own) { cudaFree(h->win_own); h->win_own = nullptr; }
    h->win_cap = (size_t)max_cameras * PART_STRIDE + 1 + 2 * (size_t)h->world + 16;
    const size_t bytes = P2P_HDR_BYTES + 2 * (size_t)h->world * h->win_cap * sizeof(double);
    CU(cudaMalloc((void**)&h->win_own, bytes));
    CU(cudaMemset(h->win_own, 0, bytes));
    if (!h->p2p_ctl) { CU(cudaMalloc((void**)&h->p2p_ctl, 64)); CU(cudaMemset(h->p2p_ctl, 0, 64)); }
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, h->win_own));
    memcpy(handle_bytes, &hd, sizeof hd);
    return 0;
}

int b200ba_comm_open_peers(b200ba_handle* h, const void* all_handles)
{
    if (!h || !all_handles) return B200BA_E_INVALID;
    if (!h->win_own) return fail(h, B200BA_E_STATE, "b200ba_comm_open_peers before b200ba_comm_peer_handle");
    CU(cudaSetDevice(h->device));
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) continue;
        cudaIpcMemHandle_t hd; memcpy(&hd, (const char*)all_handles + (size_t)r * B200BA_PEER_HANDLE_BYTES, sizeof hd);
        cudaError_t e = cudaIpcOpenMemHandle(&h->win_peer[r], hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (h->win_peer[q]) { cudaIpcCloseMemHandle(h->win_peer[q]); h->win_peer[q] = nullptr; }
            return fail(h, B200BA_E_COMM, "cudaIpcOpenMemHandle(rank %d): %s (collectives stay on NCCL)", r, cudaGetErrorString(e));
        }
    }
    h->p2p_ready = true;
    return 0;
}

static int set_problem_impl(b200ba_handle* h, int32_t N, int32_t M, int32_t K,
                            const double* cam_RT, const double* K5, const double* pts,
                            const double* obs_xy, const int32_t* obs_cam, const int32_t* obs_pt)
{
    if (!h) return B200BA_E_INVALID;
    if (N <= 0 || M < 0 || K < 0 || !cam_RT || !K5 || (M && !pts) || (K && (!obs_xy || !obs_cam || !obs_pt)))
        return fail(h, B200BA_E_INVALID, "null or empty problem arrays (N=%d M=%d K=%d)", N, M, K);
    if (h->world <= 1 && (M == 0 || K == 0)) return fail(h, B200BA_E_INVALID, "empty problem (M=%d K=%d)", M, K);
    if (!(K5[0] != 0.0) || !std::isfinite(K5[0])) return fail(h, B200BA_E_INVALID, "focal length fx must be finite and non-zero");
    const int T = (K > 200000) ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
    {
        std::atomic<long long> bad(-1);
        parallel_for(T, [&](int t, int nt) {
            const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
            for (long long k = k0; k < k1; ++k)
                if (obs_cam[k] < 0 || obs_cam[k] >= N || obs_pt[k] < 0 || obs_pt[k] >= M || (k && obs_pt[k] < obs_pt[k - 1])) {
                    long long cur = bad.load();
                    while ((cur < 0 || k < cur) && !bad.compare_exchange_weak(cur, k)) {}
                    break;
                }
        });
        const long long k = bad.load();
        if (k >= 0) {
            if (obs_cam[k] < 0 || obs_cam[k] >= N || obs_pt[k] < 0 || obs_pt[k] >= M) return fail(h, B200BA_E_INVALID, "observation %lld: index out of range", k);
            return fail(h, B200BA_E_INVALID, "observation %lld: obs_pt must be non-decreasing", k);
        }
    }
    double t0 = now_ms();
    const bool verbose = getenv("B200BA_VERBOSE") != nullptr;
    double tm = t0;
    auto lap = [&](const char* what) { if (verbose) { double t = now_ms(); fprintf(stderr, "[b200ba] set_problem %-28s %8.2f ms\n", what, t - tm); tm = t; } };
    CU(cudaSetDevice(h->device));
    // A failed call must not leave a half-replaced problem behind: from here on the handle has NO problem until the very
    // end of this function (begin / solve / get_result then fail with B200BA_E_STATE instead of touching freed buffers).
    h->have_problem = false; h->begun = false; h->finished = false;
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    h->N = N; h->M = M; h->K = K;
    h->h_cam.assign(cam_RT, cam_RT + 12 * (size_t)N);
    h->h_pts.assign(pts, pts + 3 * (size_t)M);
    memcpy(h->K5, K5, sizeof h->K5);

    // ---- marshalling (src/sfm_ba_ssba.cpp:88-102,171-178; intrinsic normaliser v3d_metricbundle.h:395-413) ----
    Dev& d = h->d;
    const b200ba_options& o = h->opt;
    double f0 = K5[0], rf = 1.0 / f0;
    Intrinsics in;
    in.k00 = K5[0] * rf; in.k01 = K5[1] * rf; in.k02 = K5[2] * rf; in.k11 = K5[3] * rf; in.k12 = K5[4] * rf;
    double r2 = 1.0 / in.k00;
    in.k00 *= r2; in.k01 *= r2; in.k02 *= r2; in.k11 *= r2; in.k12 *= r2;
    in.huber = o.inlier_px > 0 ? o.inlier_px / std::fabs(f0) : 0.0;
    in.f0 = f0;
    in.ar = in.k11 / in.k00;

    std::unique_ptr<double2[]> meas(new double2[(size_t)std::max(K, 1)]);
    parallel_for(T, [&](int t, int nt) {
        const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
        for (long long k = k0; k < k1; ++k) {
            double x = obs_xy[2 * (size_t)k], y = obs_xy[2 * (size_t)k + 1];
            if (o.round_pixels) { x = (double)lrintf((float)x); y = (double)lrintf((float)y); }
            meas[k] = make_double2((x * rf) * r2, (y * rf) * r2);
        }
    });
    h->pba_depth_scale = 1.0;
    if (o.normalize == 2) {
        // NormalizeDataD (SparseBundleCPU.cpp:3189-3282): the median depth of all observations becomes 1 / __data_normalize_median
        // (= 2).  (Its degeneracy fix - shifting cameras that sit inside the point cloud - is not reproduced.)
        if (h->world > 1) return fail(h, B200BA_E_INVALID, "normalize = 2 (PBA's data normalisation) is single-rank only");
        std::vector<float> oz((size_t)std::max(K, 1));
        parallel_for(T, [&](int t, int nt) {
            const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
            for (long long k = k0; k < k1; ++k) {
                const double* P = cam_RT + 12 * (size_t)obs_cam[k]; const double* X = pts + 3 * (size_t)obs_pt[k];
                oz[k] = (float)(P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11]);
            }
        });
        std::nth_element(oz.begin(), oz.begin() + K / 2, oz.begin() + K);
        const double med = oz[K / 2];
        if (!(med > 0) || !std::isfinite(med)) return fail(h, B200BA_E_INVALID, "normalize = 2: median depth %g is not positive", med);
        h->pba_depth_scale = (1.0 / med) / 0.5;
    }
    lap("validate + measurements");
    h->pdl_chain = false;
    h->use_dense = false; h->dense_ctas = 0; h->dense_grid = false; h->dense_smem = 0; h->dense_pairs = 0;
    DenseLists dl; DenseShape dsh; std::vector<int> dense_entry_pt;
    int smem_optin = 0, n_sm = 0;
    CU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    h->num_sms = n_sm;
    {
        int mode = o.reduced_system;
        if (const char* e = getenv("B200BA_DENSE")) mode = atoi(e) ? 2 : 1;
        int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
        // (pair count bounded here, before the layout is built for one plan or the other: sum over points of d^2)
        long long pair_cap = DENSE_MAX_PAIRS;
        if (const char* e = getenv("B200BA_DENSE_MAX_PAIRS")) pair_cap = std::min<long long>(DENSE_MAX_PAIRS, std::max(0ll, atoll(e)));   // (tests: force the fallback)
        if (mode != 1 && h->world <= 1 && !o.pcg_fp32 && K > 0 && dense_pair_bound(K, obs_pt) <= pair_cap) {
            const int cand_kind[4] = {0, 1, 1, 2}, cand_ctas[4] = {8, 8, 16, coop ? n_sm : 0};
            for (int c = 0; c < 4 && !dsh.ctas; ++c) {
                if (cand_ctas[c] <= 0) continue;
                const DenseShape t = dense_shape(N, cand_kind[c], cand_ctas[c]);
                if (t.kind < 0 || t.smem + 1024 > (size_t)smem_optin) continue;
                const void* fn = dense_pcg_kernel(t.kind);
                // (function attributes are per kernel and process, not per handle: always the device's limit, so that handles with
                //  different camera counts on different threads cannot lower each other's)
                if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) != cudaSuccess) { cudaGetLastError(); continue; }
                if (t.kind != 2) {                                      // can one such cluster be resident?
                    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
                    cfg.gridDim = dim3(t.ctas); cfg.blockDim = dim3(t.threads); cfg.dynamicSmemBytes = t.smem;
                    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = t.ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    int nclusters = 0;
                    if ((t.ctas > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
                        cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
                } else {
                    int per_sm = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, t.threads, t.smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); continue; }
                }
                dsh = t; h->dense_grid = t.kind == 2;
            }
        }
    }
    // ---- by-point layout: SELL-32, points sorted by track length (descending, stable), quarter-warps grouped and slots
    //      assigned for conflict-free reads of the shared-memory camera table (layout_host.h) ----
    PointLayout lay;
    // (the conflict-free grouping serves the shared-memory-table PCG sweep of the matrix-free plan; a problem that will run the
    //  explicit plan skips it: 0.9 of the 1.8 ms of b200ba_set_problem at the demo sequence's size)
    build_point_layout(M, K, obs_cam, obs_pt, (getenv("B200BA_NO_LANE_OPT") || dsh.ctas) ? 0 : o.lane_window, T, lay);
    h->orig_of_new.swap(lay.orig_of_new);
    const std::vector<int>& pstart = lay.pstart;
    const std::vector<int>& deg_new = lay.deg_new;
    const std::vector<int>& slot_obs = lay.slot_obs;
    lap("point order + lane grouping + slots");
    const int n_slices = cdiv(std::max(M, 1), 32);
    std::vector<int> slice_base((size_t)n_slices + 1, 0), slice_deg((size_t)n_slices, 0);
    for (int g = 0; g < n_slices; ++g) {
        int j0 = 32 * g;
        int dg = j0 < M ? deg_new[j0] : 0;       // (the grouping permutes only points of equal track length)
        slice_deg[g] = dg; slice_base[g + 1] = slice_base[g] + 32 * dg;
    }
    const int Kp = slice_base[n_slices];
    h->Kp = Kp;
    // staging of the two point-order arrays in the pinned pool (held until the uploads have completed at the end of this call)
    PinnedPool::Lease stage;
    const size_t stage_meas = sizeof(double2) * (size_t)std::max(Kp, 1);
    g_pinned.lease(stage, stage_meas + sizeof(int) * (size_t)std::max(Kp, 1));
    // Declared after the lease, so it runs BEFORE the lease is released on every exit path (the CU() early returns
    // included): no asynchronous upload may still be reading the staging memory when another handle gets it.
    struct SyncOnExit { cudaStream_t s; ~SyncOnExit() { cudaStreamSynchronize(s); } } sync_on_exit{h->stream};
    double2* o_meas_p = reinterpret_cast<double2*>(stage.p);
    int* o_cam_p = reinterpret_cast<int*>(stage.p + stage_meas);
    h->entry_of_obs.reset(new int[(size_t)std::max(K, 1)]);
    // camera order.  The host only COUNTS (per-camera histogram -> rows, first sorted position, segments); the row records
    // themselves are built on the device from the uploaded point-order arrays (radix sort by (camera, point), see k_rows_*):
    // every camera's run is padded to whole ROWS of 32 observations (point index -1 = padding); see REC_* in ba_kernels.cuh.
    std::vector<int> row_beg((size_t)N + 1, 0), cam_first((size_t)N + 1, 0);
    int n_rows = 0;
    {
        std::vector<std::vector<int>> cnt((size_t)T, std::vector<int>((size_t)N, 0));
        parallel_for(T, [&](int t, int nt) {
            const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
            std::vector<int>& c = cnt[t];
            for (long long k = k0; k < k1; ++k) c[obs_cam[k]]++;
        });
        for (int i = 0; i < N; ++i) {
            int c = 0; for (int t = 0; t < T; ++t) c += cnt[t][i];
            cam_first[i + 1] = cam_first[i] + c;
            row_beg[i + 1] = row_beg[i] + (c + 31) / 32;
        }
        n_rows = row_beg[N];
        auto slice_range = [&](int t, int nt, int& g0, int& g1) { g0 = (int)((long long)n_slices * t / nt); g1 = (int)((long long)n_slices * (t + 1) / nt); };
        parallel_for(T, [&](int t, int nt) {
            int g0, g1; slice_range(t, nt, g0, g1);
            for (int e = slice_base[g0]; e < slice_base[g1]; ++e) { o_cam_p[e] = -1; o_meas_p[e] = make_double2(0.0, 0.0); }
            for (int jn = 32 * g0; jn < std::min(M, 32 * g1); ++jn) {
                const int j = h->orig_of_new[jn], g = jn >> 5, l = jn & 31;
                for (int q = pstart[j], sl = 0; q < pstart[j + 1]; ++q, ++sl) {
                    const int k = slot_obs[q], kk = slice_base[g] + 32 * sl + l;      // observation k sits in slot sl of its point
                    o_cam_p[kk] = obs_cam[k]; o_meas_p[kk] = meas[k]; h->entry_of_obs[k] = kk;
                }
            }
        });
    }
    // virtual warps (equal row ranges) and segments (runs of one camera inside one virtual warp's range)
    const int n_vwarps = std::max(1, n_sm) * CW * VSPLIT;
    std::vector<int2> row_hdr((size_t)std::max(n_rows, 1), make_int2(0, 0));
    std::vector<int> cam_seg_beg((size_t)N + 1, 0);
    int n_segs = 0;
    {
        for (int i = 0; i < N; ++i) for (int r = row_beg[i]; r < row_beg[i + 1]; ++r) row_hdr[r].x = i;
        int seg = -1;
        for (int v = 0; v < n_vwarps; ++v) {
            const int rb = (int)((long long)n_rows * v / n_vwarps), re = (int)((long long)n_rows * (v + 1) / n_vwarps);
            for (int r = rb; r < re; ++r) { if (r == rb || row_hdr[r].x != row_hdr[r - 1].x) ++seg; row_hdr[r].y = seg; }
        }
        n_segs = seg + 1;
        cam_seg_beg[N] = n_segs;
        for (int i = N - 1; i >= 0; --i) cam_seg_beg[i] = row_beg[i + 1] > row_beg[i] ? row_hdr[row_beg[i]].y : cam_seg_beg[i + 1];
    }
    int n_pt_blocks = std::max(1, cdiv(M, PT_BLOCK));

    lap("SELL placement + camera counts");

    // ---- explicit reduced camera system for small problems (dense_host.h): S in the shared memory of one thread-block
    //      cluster (8 or 16 CTAs) or, for a few hundred cameras, of the whole cooperative grid ----
    //      (the kernel shape was chosen above, before the layout; here the pair lists, which need the SELL entry of every observation)
    if (dsh.ctas && build_dense_lists(N, M, K, obs_cam, obs_pt, h->entry_of_obs.get(), dl)) {
        h->use_dense = true; h->dense_ctas = dsh.ctas; h->dense_smem = dsh.smem; h->dense_pairs = (long long)dl.pairs.size();
        h->dense_kind = dsh.kind; h->dense_threads = dsh.threads;
        { const char* e = getenv("B200BA_PDL_CHAIN"); h->pdl_chain = e ? atoi(e) != 0 : true; }
        dense_entry_pt.assign((size_t)std::max(Kp, 1), 0);
        for (int jn = 0; jn < M; ++jn)
            for (int sl = 0, e = slice_base[jn >> 5] + (jn & 31); sl < deg_new[jn]; ++sl, e += 32) dense_entry_pt[e] = sl == 0 ? ~jn : jn;
    }
    lap("explicit reduced system: pair lists");

    // ---- device allocation ----
    size_t n6 = 6 * (size_t)N;
    int world = h->world;
    h->trace_cap = o.max_iterations + 2;
    const size_t Mp = (size_t)std::max(M, 1), Kpp = (size_t)std::max(Kp, 1), Rp = (size_t)std::max(n_rows, 1);
    char* const arena_old = h->arena.base;                        // grow-only: kept when the new problem fits (persistent handles)
    h->arena.base = nullptr;
    int tab_warps_cap = (n_sm > 0 ? n_sm : 148) * (std::max(std::max(TAB_BLOCK, TAB2_BLOCK), TABC_BLOCK) / 32);
    const int step_grid = cdiv((long long)N * 6, STEP_BLOCK);
    // large-camera-count plan: the packed camera table will not fit one SM's shared memory (same test as `use_tab` below)
    const bool want_wb = (((size_t)(N + 1) * TAB_STRIDE * sizeof(double) + 127) & ~(size_t)127) + TABC_RING_BYTES + 256 > (size_t)smem_optin || getenv("B200BA_NO_TAB") != nullptr;
    h->use_wb = want_wb && !h->use_dense && getenv("B200BA_NO_WB") == nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Arena& A = h->arena;
        A.off = 0;
        if (h->use_wb) { h->obs_B.take(A, (Kpp / 32 + 1) * 192); h->wtab.take(A, 8 * (size_t)N); } else { h->obs_B.p = nullptr; h->wtab.p = nullptr; }
        h->rt0.take(A, 12 * (size_t)N); h->rt1.take(A, 12 * (size_t)N);
        h->U.take(A, 36 * (size_t)N); h->Minv.take(A, 36 * (size_t)N);
        h->gc.take(A, n6); h->x.take(A, n6); h->r.take(A, n6); h->z.take(A, n6); h->p.take(A, n6); h->q.take(A, n6);
        h->ar_lin.take(A, (size_t)N * PART_STRIDE + 1 + 2 * (size_t)world); h->ar_q.take(A, n6); h->ar_sd.take(A, 21 * (size_t)N); h->ar_dec.take(A, 4);
        h->X0.take(A, 4 * Mp); h->X1.take(A, 4 * Mp);
        h->V.take(A, 6 * Mp); h->Vinv.take(A, 6 * Mp); h->gp.take(A, 4 * Mp); h->v.take(A, 4 * Mp);
        h->o_rec.take(A, (Kpp / 32 + 1) * OREC_BYTES); h->o_meas.take(A, Kpp); h->c_meas.take(A, 32 * Rp);
        h->o_cam.take(A, Kpp); h->slice_base.take(A, (size_t)n_slices + 1); h->slice_deg.take(A, (size_t)n_slices);
        h->c_rec.take(A, Rp * REC_BYTES + 16); h->row_hdr.take(A, Rp);
        h->sort_keys.take(A, 2 * Kpp); h->sort_vals.take(A, 2 * Kpp); h->row_tab.take(A, 2 * ((size_t)N + 1)); h->sort_tmp.take(A, (size_t)16 << 20);
        h->cam_seg_beg.take(A, (size_t)N + 1);
        h->seg_part.take(A, (size_t)std::max(n_segs, 1) * PART_STRIDE); h->pt_part.take(A, 4 * (size_t)std::max(n_pt_blocks, tab_warps_cap));
        h->trace.take(A, (size_t)h->trace_cap * TRACE_COLS); h->st.take(A, 1); h->norm_buf.take(A, 16);
        h->ptab.take(A, (size_t)(N + 1) * TAB_STRIDE); h->qtab.take(A, (size_t)N * QTAB_STRIDE); h->ctab.take(A, (size_t)(N + 1) * QTAB_STRIDE); h->fin_part.take(A, 4 * (size_t)cdiv(N, FIN_BLOCK) + 4);
        if (o.lm_variant == 1) { h->cam_D0.take(A, n6); h->pt_D0.take(A, 4 * Mp); }
        h->X_init.take(A, 4 * Mp); h->rt_init.take(A, 12 * (size_t)N);
        h->tab_wr.take(A, (size_t)tab_warps_cap * 8 + 8);
        h->stamps.take(A, 512 * 40); h->step_bar.take(A, 4);
        if (o.pcg_fp32) {
            h->tab_wr32.take(A, (size_t)(n_sm > 0 ? n_sm : 148) * (TAB32_BLOCK / 32) * 8 + 8);
            h->o_rec32.take(A, (Kpp / 32 + 1) * OREC32_BYTES); h->c_rec32.take(A, Rp * REC32_BYTES + 16);
            h->X32.take(A, Mp); h->v32.take(A, Mp); h->Vinv32.take(A, 8 * Mp); h->ptab32.take(A, (size_t)N * TAB32_STRIDE + 8); h->rt32.take(A, 12 * (size_t)N);
        } h->part_pq.take(A, (size_t)step_grid); h->part_rz.take(A, (size_t)step_grid);
        if (h->use_dense) {
            h->dn_pairs.take(A, dl.pairs.size() + 1); h->dn_chunks.take(A, dl.chunks.size() + 1);
            h->dn_blk_chunk_beg.take(A, dl.blk_chunk_beg.size()); h->dn_blk_ik.take(A, dl.blk_ik.size()); h->dn_entry_pt.take(A, Kpp);
            h->dn_Wobs.take(A, DENSE_OBS_STRIDE * Kpp); h->dn_Yobs.take(A, DENSE_OBS_STRIDE * Kpp); h->dn_Tobs.take(A, 6 * Kpp);
            h->dn_chunk_part.take(A, 36 * (dl.chunks.size() + 1)); h->dn_chunk_rhs.take(A, 6 * (dl.chunks.size() + 1));
            h->dn_S.take(A, n6 * n6); h->dn_zg.take(A, n6); h->dn_pq.take(A, 512); h->dn_rz.take(A, 512); h->dn_bar.take(A, 32 * 32);
        }
        if (pass == 0) {
            A.cap = A.off;
            if (arena_old && A.cap <= A.allocated) A.base = arena_old;
            else {
                if (arena_old) cudaFree(arena_old);
                const size_t want = arena_old ? A.cap + A.cap / 4 : A.cap;     // a handle that is reused grows with headroom
                A.allocated = 0;
                CU(cudaMalloc((void**)&A.base, want));
                A.allocated = want;
            }
        }
    }
    lap("cudaMalloc");
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(h->arena.base, 0, h->arena.cap, s));
    CU(cudaMemcpyAsync(h->o_meas.p, o_meas_p, sizeof(double2) * Kpp, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->o_cam.p, o_cam_p, sizeof(int) * Kpp, cudaMemcpyHostToDevice, s));
    if (n_rows) CU(cudaMemcpyAsync(h->row_hdr.p, row_hdr.data(), sizeof(int2) * (size_t)n_rows, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->slice_base.p, slice_base.data(), sizeof(int) * ((size_t)n_slices + 1), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->slice_deg.p, slice_deg.data(), sizeof(int) * (size_t)n_slices, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->cam_seg_beg.p, cam_seg_beg.data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, s));
    if (h->use_dense) {
        static_assert(sizeof(DensePair) == sizeof(int2) && sizeof(DenseChunk) == sizeof(int4), "device views of the pair lists");
        if (!dl.pairs.empty()) CU(cudaMemcpyAsync(h->dn_pairs.p, dl.pairs.data(), sizeof(int2) * dl.pairs.size(), cudaMemcpyHostToDevice, s));
        if (!dl.chunks.empty()) CU(cudaMemcpyAsync(h->dn_chunks.p, dl.chunks.data(), sizeof(int4) * dl.chunks.size(), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->dn_blk_chunk_beg.p, dl.blk_chunk_beg.data(), sizeof(int) * dl.blk_chunk_beg.size(), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->dn_blk_ik.p, dl.blk_ik.data(), sizeof(int) * dl.blk_ik.size(), cudaMemcpyHostToDevice, s));
        if (Kp) CU(cudaMemcpyAsync(h->dn_entry_pt.p, dense_entry_pt.data(), sizeof(int) * (size_t)Kp, cudaMemcpyHostToDevice, s));
    }
    if (Kp) {
        Dev dr{}; dr.o_rec = h->o_rec.p; dr.ptab = h->ptab.p; dr.ctab = h->ctab.p; dr.N = N;
        k_fill_orows<<<cdiv(Kp / 32, 8), 256, 0, s>>>(dr, h->o_cam.p, Kp / 32);
    }
    if (n_rows) {
        // camera-order row records on the device: keys, radix sort by (camera, point), scatter into the rows
        Dev dr{}; dr.N = N; dr.M = M; dr.K = K; dr.n_rows = n_rows; dr.n_slices = n_slices; dr.c_rec = h->c_rec.p; dr.c_meas = h->c_meas.p;
        dr.o_meas = h->o_meas.p; dr.slice_base = h->slice_base.p; dr.slice_deg = h->slice_deg.p;
        CU(cudaMemcpyAsync(h->row_tab.p, cam_first.data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->row_tab.p + N + 1, row_beg.data(), sizeof(int) * ((size_t)N + 1), cudaMemcpyHostToDevice, s));
        k_fill_rows<<<cdiv(n_rows, 8), 256, 0, s>>>(dr, h->row_hdr.p);
        k_rows_keys<<<cdiv(32 * n_slices, PT_BLOCK), PT_BLOCK, 0, s>>>(dr, h->o_cam.p, h->sort_keys.p, h->sort_vals.p);
        cub::DoubleBuffer<unsigned long long> kb(h->sort_keys.p, h->sort_keys.p + Kpp);
        cub::DoubleBuffer<int> vb(h->sort_vals.p, h->sort_vals.p + Kpp);
        int end_bit = 33; while (end_bit < 64 && ((unsigned long long)N >> (end_bit - 32)) != 0) ++end_bit;
        size_t tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, Kp, 0, end_bit, s));
        void* tmp = h->sort_tmp.p; bool own_tmp = false;
        if (tmp_bytes > ((size_t)16 << 20)) { CU(cudaMalloc(&tmp, tmp_bytes)); own_tmp = true; }
        CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, Kp, 0, end_bit, s));
        k_rows_fill<<<cdiv(K, 256), 256, 0, s>>>(dr, kb.Current(), vb.Current(), h->row_tab.p, h->row_tab.p + N + 1);
        if (own_tmp) { CU(cudaStreamSynchronize(s)); cudaFree(tmp); }
    }
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());

    lap("upload");
    // ---- execution plan: shared-memory camera table, cooperative camera-side step, CUDA graph ----
    const size_t tab_bytes = (size_t)(N + 1) * TAB_STRIDE * sizeof(double);   // + the padding entries' dummy camera (record N)
    { const char* e = getenv("B200BA_TAB_ILP"); h->tab_ilp = (e && atoi(e) >= 1 && atoi(e) <= 3) ? atoi(e) : B200BA_TAB_ILP_DEFAULT; }
    h->tab_block = h->tab_ilp == 3 ? TABC_BLOCK : h->tab_ilp == 2 ? TAB2_BLOCK : TAB_BLOCK;
    h->tab_smem = ((tab_bytes + 127) & ~(size_t)127) + (h->tab_ilp == 3 ? TABC_RING_BYTES : h->tab_ilp == 2 ? TAB2_RING_BYTES : TAB_RING_BYTES);   // camera table + per-warp stream rings
    h->use_tab = h->tab_smem + 256 <= (size_t)smem_optin && getenv("B200BA_NO_TAB") == nullptr;
    if (h->use_tab) {
        cudaError_t ea = h->tab_ilp == 3 ? cudaFuncSetAttribute(k_cg_point_pass_tabc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tab_smem)
                       : h->tab_ilp == 2 ? cudaFuncSetAttribute(k_cg_point_pass_tab2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tab_smem)
                                         : cudaFuncSetAttribute(k_cg_point_pass_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tab_smem);
        if (ea != cudaSuccess) { cudaGetLastError(); h->use_tab = false; }
        h->tab_grid = std::max(1, std::min(h->num_sms, cdiv(n_slices, h->tab_block / 32)));
    }
    // contiguous slice ranges per warp of the persistent CG point sweep, balanced by rows (+1 per slice for the epilogue);
    // one 32-byte header per warp: {first slice, end slice, first entry, end entry, deg(first), deg(first+1), 0, 0}
    { const char* e = getenv("B200BA_PDL"); h->use_pdl = e ? atoi(e) : B200BA_PDL_DEFAULT; }
    h->use_tab_backsub = false;
    if (h->use_tab && h->tab_ilp == 3 && getenv("B200BA_NO_TAB_BACKSUB") == nullptr) {
        h->cost_smem = ((((size_t)N * QTAB_STRIDE * sizeof(double)) + 127) & ~(size_t)127) + COST_RING_BYTES;
        if (h->cost_smem + 256 <= (size_t)smem_optin &&
            cudaFuncSetAttribute(k_backsub_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tab_smem) == cudaSuccess &&
            cudaFuncSetAttribute(k_cost_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cost_smem) == cudaSuccess &&
            cudaFuncSetAttribute(k_lin_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cost_smem) == cudaSuccess) h->use_tab_backsub = true;
        else cudaGetLastError();
    }
    int tab_warps = std::max(1, h->tab_grid) * (h->tab_block / 32);
    auto warp_headers = [&](int nwarps, int warps_per_cta, int* dst) -> int {
        std::vector<int> wr((size_t)nwarps + 1, n_slices);
        // cost of a slice = SLICE_W_NUM / SLICE_W_DEN row bodies for its epilogue + one per row (clock stamps of the sweep's
        // warps at Venice shape: a slice end costs 0.6-0.7 of a row body; weighting it as a whole row left the CTAs that own
        // the long tracks 10 % behind those with the two-observation points)
        const long long SLICE_W_NUM = getenv("B200BA_SLICE_W") ? atoll(getenv("B200BA_SLICE_W")) : ::SLICE_W_NUM;   // (thirds of a row body)
        long long total = 0; for (int g = 0; g < n_slices; ++g) total += (long long)SLICE_W_DEN * slice_deg[g] + SLICE_W_NUM;
        long long acc = 0; int w = 0; wr[0] = 0;
        for (int g = 0; g < n_slices && w + 1 < nwarps; ++g) {
            acc += (long long)SLICE_W_DEN * slice_deg[g] + SLICE_W_NUM;
            while (w + 1 < nwarps && acc * nwarps >= total * (long long)(w + 1)) wr[++w] = g + 1;
        }
        for (int q = w + 1; q <= nwarps; ++q) wr[q] = n_slices;
        // Range q goes to warp (q / grid) of CTA (q % grid): the points are sorted by track length, so contiguous ranges per CTA gave
        // the first CTAs only long tracks and the last ones only two-observation points, and whatever the cost model missed showed as
        // a 10 % spread between the CTAs' finish times; with the ranges dealt out round-robin every SM gets the same mix.  Which warp
        // sweeps a range changes no arithmetic.  (B200BA_XPOSE=0: contiguous ranges per CTA.)
        const int wpc = std::max(1, warps_per_cta), grid = std::max(1, nwarps / wpc);
        const bool xpose = nwarps == grid * wpc && (getenv("B200BA_XPOSE") ? atoi(getenv("B200BA_XPOSE")) != 0 : true);
        std::vector<int> hdr((size_t)nwarps * 8, 0);
        for (int q = 0; q < nwarps; ++q) {
            const int g0 = wr[q], g1 = wr[q + 1];
            int* o = hdr.data() + 8 * (size_t)(xpose ? (q % grid) * wpc + q / grid : q);
            o[0] = g0; o[1] = g1; o[2] = slice_base[g0]; o[3] = slice_base[g1];
            o[4] = g0 < n_slices ? slice_deg[g0] : 0; o[5] = g0 + 1 < g1 ? slice_deg[g0 + 1] : 0;
        }
        CU(cudaMemcpy(dst, hdr.data(), sizeof(int) * hdr.size(), cudaMemcpyHostToDevice));
        return 0;
    };
    if (int rc = warp_headers(tab_warps, h->tab_block / 32, h->tab_wr.p)) return rc;
    // ---- optional mixed-precision PCG (options.pcg_fp32): needs the fp32 camera table in shared memory ----
    h->use_f32 = false;
    int tab_warps32 = 0;
    if (o.pcg_fp32) {
        const size_t t32 = (((size_t)N * TAB32_STRIDE * sizeof(float)) + 127) & ~(size_t)127;
        h->tab32_smem = t32 + TAB32_RING_BYTES;
        if (h->tab32_smem + 256 > (size_t)smem_optin || (size_t)CAMPASS32_SMEM > (size_t)smem_optin)
            return fail(h, B200BA_E_INVALID, "options.pcg_fp32 needs the fp32 camera table (%d cameras x 80 bytes) to fit one SM's shared memory", N);
        CU(cudaFuncSetAttribute(k_cg_point_pass_tab32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tab32_smem));
        CU(cudaFuncSetAttribute(k_cg_camera_pass32, cudaFuncAttributeMaxDynamicSharedMemorySize, CAMPASS32_SMEM));
        h->tab32_grid = std::max(1, std::min(h->num_sms, cdiv(n_slices, TAB32_BLOCK / 32)));
        tab_warps32 = h->tab32_grid * (TAB32_BLOCK / 32);
        if (int rc = warp_headers(tab_warps32, TAB32_BLOCK / 32, h->tab_wr32.p)) return rc;
        h->use_f32 = true;
    }
    if ((size_t)CAMPASS_SMEM > (size_t)smem_optin) return fail(h, B200BA_E_CUDA, "device offers %d bytes of shared memory per block, the by-camera sweep needs %d", smem_optin, CAMPASS_SMEM);
    CU(cudaFuncSetAttribute(k_cg_camera_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, CAMPASS_SMEM));
    CU(cudaFuncSetAttribute(k_linearize_cameras, cudaFuncAttributeMaxDynamicSharedMemorySize, LINCAM_SMEM));
    h->cam_grid = cdiv(n_vwarps, CW * VSPLIT);
    h->step_grid = cdiv((long long)N * 6, STEP_BLOCK);
    {
        int per_sm = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_camera_step, STEP_BLOCK, 0);
        h->use_coop = coop && (long long)per_sm * h->num_sms >= h->step_grid && getenv("B200BA_NO_COOP") == nullptr;
    }
    {   // by-camera sweep + camera-side step in one cooperative launch: B200BA_FUSE_STEP=1.  Measured at Venice shape with the
        // counting barrier (1.3 us) in place of grid.sync(): 5.57 vs 5.48 ms per LM iteration - 148 CTAs x 512 threads go through
        // three barriers for a step that 56 blocks of 192 threads do in two, and the cooperative launch costs more than the
        // kernel boundary it removes.  Off by default.
        const char* e = getenv("B200BA_FUSE_STEP");
        int per_sm = 0;
        h->fuse_step = false;
        if (h->use_coop && h->world <= 1 && !h->use_f32 && h->step_grid <= h->cam_grid && (e ? atoi(e) != 0 : false) &&
            cudaFuncSetAttribute(k_cg_camera_pass_step, cudaFuncAttributeMaxDynamicSharedMemorySize, CAMPASS_SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_camera_pass_step, CW * 32, CAMPASS_SMEM) == cudaSuccess &&
            (long long)per_sm * h->num_sms >= h->cam_grid) h->fuse_step = true;
        else cudaGetLastError();
    }
    // NCCL collectives are enqueued eagerly (outside any capture); with the peer windows mapped the multi-rank iteration
    // contains no host-side collective and is captured like the single-rank one
    h->use_graph = getenv("B200BA_NO_GRAPH") == nullptr && (h->world <= 1 || h->p2p_ready);
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }

    d = Dev{};
    d.N = N; d.M = M; d.K = K; d.n_rows = n_rows; d.n_vwarps = n_vwarps; d.n_segs = n_segs; d.first_free = o.fix_first_camera ? 1 : 0; d.n_pt_blocks = n_pt_blocks;
    if (h->p2p_ready) {
        if ((size_t)N * PART_STRIDE + 1 + 2 * (size_t)world > h->win_cap)
            return fail(h, B200BA_E_COMM, "peer windows were sized for fewer cameras (capacity %zu doubles, need %zu)", h->win_cap, (size_t)N * PART_STRIDE + 1 + 2 * (size_t)world);
        d.p2p.on = 1; d.p2p.world = world; d.p2p.rank = h->rank; d.p2p.cap = h->win_cap;
        for (int r = 0; r < world; ++r) d.p2p.win[r] = (r == h->rank) ? h->win_own : (unsigned char*)h->win_peer[r];
        d.p2p.seq = h->p2p_ctl; d.p2p.done_blocks = (unsigned int*)(h->p2p_ctl + 1); d.p2p.err = (int*)(h->p2p_ctl + 2);
    }
    d.world = world; d.rank = h->rank; d.n_slices = n_slices; d.tab_ok = h->use_tab ? 1 : 0; d.in = in; d.st = h->st.p;
    d.rt[0] = h->rt0.p; d.rt[1] = h->rt1.p; d.U = h->U.p; d.gc = h->gc.p; d.Minv = h->Minv.p;
    d.x = h->x.p; d.r = h->r.p; d.z = h->z.p; d.p = h->p.p; d.q = h->q.p;
    d.ar_lin = h->ar_lin.p; d.ar_q = h->ar_q.p; d.ar_sd = h->ar_sd.p; d.ar_dec = h->ar_dec.p;
    d.X[0] = h->X0.p; d.X[1] = h->X1.p; d.V = h->V.p; d.gp = h->gp.p; d.Vinv = h->Vinv.p; d.v = h->v.p;
    d.o_rec = h->o_rec.p; d.o_meas = h->o_meas.p; d.slice_base = h->slice_base.p; d.slice_deg = h->slice_deg.p;
    d.ptab = h->ptab.p; d.qtab = h->qtab.p; d.ctab = h->ctab.p; d.obs_B = h->use_wb ? h->obs_B.p : nullptr; d.wtab = h->use_wb ? h->wtab.p : nullptr; d.fin_part = h->fin_part.p; d.cam_D0 = h->cam_D0.p; d.pt_D0 = h->pt_D0.p;
    d.tab_wr = h->tab_wr.p; d.tab_warps = tab_warps;
    d.c_rec = h->c_rec.p; d.c_meas = h->c_meas.p;
    d.cam_seg_beg = h->cam_seg_beg.p;
    d.seg_part = h->seg_part.p; d.pt_part = h->pt_part.p; d.trace = h->trace.p; d.stamps = h->stamps.p; d.step_bar = h->step_bar.p;
    if (h->use_dense) {
        DenseDev& dn = d.dn;
        dn.on = 1; dn.rows = dsh.rows; dn.ld = dsh.ld; dn.n_chunks = (int)dl.chunks.size(); dn.n_blocks = dl.n_blocks;
        dn.pairs = h->dn_pairs.p; dn.chunks = h->dn_chunks.p; dn.blk_chunk_beg = h->dn_blk_chunk_beg.p; dn.blk_ik = h->dn_blk_ik.p; dn.entry_pt = h->dn_entry_pt.p;
        dn.Wobs = h->dn_Wobs.p; dn.Yobs = h->dn_Yobs.p; dn.Tobs = h->dn_Tobs.p; dn.chunk_part = h->dn_chunk_part.p; dn.chunk_rhs = h->dn_chunk_rhs.p;
        dn.S = h->dn_S.p; dn.zg = h->dn_zg.p; dn.pq_slots = h->dn_pq.p; dn.rz_slots = h->dn_rz.p;
        dn.bar_counter = h->dn_bar.p;
        // (the kernel's function attributes were set when the shape was chosen: shared memory = the device limit, non-portable cluster size)
    }
    if (h->use_f32) {
        d.o_rec32 = h->o_rec32.p; d.c_rec32 = h->c_rec32.p; d.X32 = h->X32.p; d.v32 = h->v32.p; d.Vinv32 = h->Vinv32.p;
        d.ptab32 = h->ptab32.p; d.rt32 = h->rt32.p; d.tab_wr32 = h->tab_wr32.p; d.tab_warps32 = tab_warps32;
        k_f32_indices<<<cdiv(std::max(Kp / 32, n_rows), 8), 256, 0, h->stream>>>(d, Kp / 32);
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaGetLastError());
    }
    lap("plan");
    h->have_problem = true; h->begun = false; h->finished = false;
    h->ms_h2d = now_ms() - t0;
    return 0;
}

int b200ba_set_problem(b200ba_handle* h, int32_t N, int32_t M, int32_t K,
                       const double* cam_RT, const double* K5, const double* pts,
                       const double* obs_xy, const int32_t* obs_cam, const int32_t* obs_pt)
{
    if (!h) return B200BA_E_INVALID;
    h->s_K = -1;                                                  // the stored observations (if any) describe the previous problem
    if (int rc = set_problem_impl(h, N, M, K, cam_RT, K5, pts, obs_xy, obs_cam, obs_pt)) return rc;
    if (h->opt.keep_problem) {                                    // persistent handles: what b200ba_append extends
        h->s_xy.assign(obs_xy, obs_xy + 2 * (size_t)K); h->s_cam.assign(obs_cam, obs_cam + K); h->s_pt.assign(obs_pt, obs_pt + K);
        h->s_K = K;
    }
    return 0;
}

// Persistent state across keyframes: the tracker's map only GROWS between most BA calls - one new
// camera, the points triangulated from it, and observations of old and new points in the new frame (src/sfm_tracker.cpp:
// 117-195, 389-399, 858-866) - while src/sfm_ba_ssba.cpp:124-182 re-marshals everything through std::map look-ups every time.
// b200ba_append takes only what is new; the handle keeps the marshalled problem, its device arena (grow-only) and its stream.
int b200ba_append(b200ba_handle* h, int32_t n_cam, const double* cam_RT, int32_t n_pts, const double* pts,
                  int32_t n_obs, const double* obs_xy, const int32_t* obs_cam, const int32_t* obs_pt)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->opt.keep_problem) return fail(h, B200BA_E_STATE, "b200ba_append needs options.keep_problem = 1");
    if (!h->have_problem || h->s_K < 0) return fail(h, B200BA_E_STATE, "b200ba_append before b200ba_set_problem");
    if (h->world > 1) return fail(h, B200BA_E_STATE, "b200ba_append on a multi-rank handle is not supported");
    if (n_cam < 0 || n_pts < 0 || n_obs < 0 || (n_cam && !cam_RT) || (n_pts && !pts) || (n_obs && (!obs_xy || !obs_cam || !obs_pt)))
        return fail(h, B200BA_E_INVALID, "b200ba_append: null or negative-sized arrays");
    const int N0 = h->N, M0 = h->M, K0 = h->s_K, N = N0 + n_cam, M = M0 + n_pts, K = K0 + n_obs;
    for (int k = 0; k < n_obs; ++k)
        if (obs_cam[k] < 0 || obs_cam[k] >= N || obs_pt[k] < 0 || obs_pt[k] >= M || (k && obs_pt[k] < obs_pt[k - 1]))
            return fail(h, B200BA_E_INVALID, "b200ba_append: observation %d out of range or obs_pt decreasing", k);
    std::vector<double> cam(h->h_cam), pt(h->h_pts), xy(2 * (size_t)K);
    cam.insert(cam.end(), cam_RT, cam_RT + 12 * (size_t)n_cam);
    pt.insert(pt.end(), pts, pts + 3 * (size_t)n_pts);
    std::vector<int32_t> oc((size_t)K), op((size_t)K);
    // stable merge by point: within a point the stored observations come first, then the new ones in the order given - the order
    // src/sfm_ba_ssba.cpp:161-182 produces when the new frame has the largest id (std::map<frame, keypoint> iterates by frame id)
    {
        int a = 0, b = 0, o = 0;
        while (a < K0 || b < n_obs) {
            const bool take_old = b >= n_obs || (a < K0 && h->s_pt[a] <= obs_pt[b]);
            if (take_old) { oc[o] = h->s_cam[a]; op[o] = h->s_pt[a]; xy[2 * (size_t)o] = h->s_xy[2 * (size_t)a]; xy[2 * (size_t)o + 1] = h->s_xy[2 * (size_t)a + 1]; ++a; }
            else { oc[o] = obs_cam[b]; op[o] = obs_pt[b]; xy[2 * (size_t)o] = obs_xy[2 * (size_t)b]; xy[2 * (size_t)o + 1] = obs_xy[2 * (size_t)b + 1]; ++b; }
            ++o;
        }
    }
    double K5[5]; memcpy(K5, h->K5, sizeof K5);
    h->s_K = -1;
    if (int rc = set_problem_impl(h, N, M, K, cam.data(), K5, pt.data(), xy.data(), oc.data(), op.data())) return rc;
    h->s_xy.swap(xy); h->s_cam.swap(oc); h->s_pt.swap(op); h->s_K = K;
    return 0;
}

// New parameter values for the uploaded problem (what the tracker writes into its containers between two BA calls without
// changing the map's structure: the post-BA re-axis, src/sfm_tracker.cpp:894-908); the observation layout on the device stays.
int b200ba_set_state(b200ba_handle* h, const double* cam_RT, const double* pts)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->have_problem) return fail(h, B200BA_E_STATE, "b200ba_set_state before b200ba_set_problem");
    if (cam_RT) h->h_cam.assign(cam_RT, cam_RT + 12 * (size_t)h->N);
    if (pts) h->h_pts.assign(pts, pts + 3 * (size_t)h->M);
    h->begun = false; h->finished = false;
    return 0;
}

int b200ba_begin(b200ba_handle* h)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->have_problem) return fail(h, B200BA_E_STATE, "b200ba_begin before b200ba_set_problem");
    CU(cudaSetDevice(h->device));
    const int N = h->N, M = h->M;
    const b200ba_options& o = h->opt;
    std::vector<double> cam(h->h_cam);
    {
        // The residual, the Jacobians and the trial cost use R exactly as given (the reference never re-orthonormalises,
        // v3d_metricbundle.cpp:17-39); their shared-memory-table versions carry R as a unit quaternion, which is the same thing only
        // for orthonormal input.  Rotations that are not (e.g. the float32 cameras of the PBA path, 1e-7 off) keep the generic sweeps.
        double defect = 0;
        for (int i = 0; i < N; ++i) {
            const double* P = cam.data() + 12 * (size_t)i;
            for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) {
                const double v = P[4 * a] * P[4 * b] + P[4 * a + 1] * P[4 * b + 1] + P[4 * a + 2] * P[4 * b + 2] - (a == b ? 1.0 : 0.0);
                defect = std::max(defect, std::fabs(v));
            }
        }
        const bool ok = defect < 1e-11;
        if (ok != h->quat_ok) { h->quat_ok = ok; if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; } }
    }
    const double* P0 = h->h_pts.data();
    h->mean[0] = h->mean[1] = h->mean[2] = 0; h->scale = 1.0;
    std::vector<double> C(3 * (size_t)N);
    for (int i = 0; i < N; ++i) camera_centre(cam.data() + 12 * (size_t)i, C.data() + 3 * (size_t)i);
    // Host loops over the points: sequential (the reference's own summation order) up to 200 000 points, where the parity
    // tests live; above that, contiguous blocks summed by threads and combined in block order (rounding differs at 1e-16).
    const int T = (M > 200000) ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
    auto blocked_sum = [&](int ncomp, double* out, auto&& term) {
        std::vector<double> part((size_t)T * ncomp, 0.0);
        parallel_for(T, [&](int t, int nt) {
            const int j0 = (int)((long long)M * t / nt), j1 = (int)((long long)M * (t + 1) / nt);
            double* a = part.data() + (size_t)t * ncomp;
            for (int j = j0; j < j1; ++j) term(j, a);
        });
        for (int t = 0; t < T; ++t) for (int c = 0; c < ncomp; ++c) out[c] += part[(size_t)t * ncomp + c];
    };
    double rs = 1.0;
    if (o.normalize == 2) { rs = h->pba_depth_scale; h->scale = 1.0 / rs;
        for (int i = 0; i < N; ++i) { double* P = cam.data() + 12 * (size_t)i; P[3] *= rs; P[7] *= rs; P[11] *= rs; } }
    else if (o.normalize) {
        // ScopedBundleExtrinsicNormalizer ctor (v3d_metricbundle.h:327-356); point sums are global over ranks,
        // cameras are replicated and counted once.  (The cross-rank sums run BEFORE the pinned staging pool is leased:
        // the pool is a process-wide mutex, and a rank that lives as a thread of the same process must be able to reach
        // its side of the collective.)
        double sm4[4] = {0, 0, 0, (double)M};
        blocked_sum(3, sm4, [&](int j, double* a) { for (int c = 0; c < 3; ++c) a[c] += P0[3 * (size_t)j + c]; });
        if (int rc = global_sum(h, sm4, 4)) return rc;
        double cs[3] = {0, 0, 0};
        for (int i = 0; i < N; ++i) for (int c = 0; c < 3; ++c) cs[c] += C[3 * (size_t)i + c];
        // reference adds camera centres first, then points
        double rn = 1.0 / ((double)N + sm4[3]);
        for (int c = 0; c < 3; ++c) h->mean[c] = (cs[c] + sm4[c]) * rn;
        double sq[1] = {0};
        blocked_sum(1, sq, [&](int j, double* a) { for (int c = 0; c < 3; ++c) { double v = P0[3 * (size_t)j + c] - h->mean[c]; a[0] += v * v; } });
        if (int rc = global_sum(h, sq, 1)) return rc;
        double sqc = 0;
        for (int i = 0; i < N; ++i) for (int c = 0; c < 3; ++c) { C[3 * (size_t)i + c] -= h->mean[c]; sqc += C[3 * (size_t)i + c] * C[3 * (size_t)i + c]; }
        h->scale = std::sqrt(sqc + sq[0]);
        rs = 1.0 / h->scale;
        for (int i = 0; i < N; ++i) { for (int c = 0; c < 3; ++c) C[3 * (size_t)i + c] *= rs; set_camera_centre(cam.data() + 12 * (size_t)i, C.data() + 3 * (size_t)i); }
    }
    double L[1] = {0};
    cudaStream_t s = h->stream;
    {
        PinnedPool::Lease stage;                                  // scoped: staging of the point block and its upload only
        g_pinned.lease(stage, sizeof(double) * 4 * (size_t)std::max(M, 1));
        struct SyncOnExit { cudaStream_t s; ~SyncOnExit() { cudaStreamSynchronize(s); } } sync_on_exit{s};   // before the lease goes
        double* X4 = reinterpret_cast<double*>(stage.p);         // no zero-fill of 32 MB: the pad is written with the point
        if (M == 0) X4[0] = X4[1] = X4[2] = X4[3] = 0.0;
        const bool nrm = o.normalize != 0;
        parallel_for(T, [&](int t, int nt) {
            const int a = (int)((long long)M * t / nt), b = (int)((long long)M * (t + 1) / nt);
            for (int jn = a; jn < b; ++jn) {
                const int j = h->orig_of_new[jn];
                for (int c = 0; c < 3; ++c) X4[4 * (size_t)jn + c] = nrm ? (P0[3 * (size_t)j + c] - h->mean[c]) * rs : P0[3 * (size_t)j + c];
                X4[4 * (size_t)jn + 3] = 0.0;
            }
        });
        // cached parameter length (v3d_metricbundle.h:37-45), after normalisation (summed in the caller's point order)
        if (T == 1) {
            std::vector<int> new_of_orig((size_t)M);
            for (int jn = 0; jn < M; ++jn) new_of_orig[h->orig_of_new[jn]] = jn;
            for (int j = 0; j < M; ++j) { const int jn = new_of_orig[j]; for (int c = 0; c < 3; ++c) L[0] += X4[4 * (size_t)jn + c] * X4[4 * (size_t)jn + c]; }
        } else {
            blocked_sum(1, L, [&](int jn, double* a) { for (int c = 0; c < 3; ++c) a[0] += X4[4 * (size_t)jn + c] * X4[4 * (size_t)jn + c]; });
        }
        CU(cudaMemcpyAsync(h->X0.p, X4, sizeof(double) * 4 * (size_t)std::max(M, 1), cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));
    }
    if (int rc = global_sum(h, L, 1)) return rc;
    for (int i = (o.fix_first_camera ? 1 : 0); i < N; ++i) {
        const double* P = cam.data() + 12 * (size_t)i;
        L[0] += P[3] * P[3] + P[7] * P[7] + P[11] * P[11] + 3.0;
    }
    CU(cudaMemcpyAsync(h->rt0.p, cam.data(), sizeof(double) * 12 * (size_t)N, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->rt1.p, cam.data(), sizeof(double) * 12 * (size_t)N, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->X1.p, h->X0.p, sizeof(double) * 4 * (size_t)std::max(M, 1), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(h->X_init.p, h->X0.p, sizeof(double) * 4 * (size_t)std::max(M, 1), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(h->rt_init.p, h->rt0.p, sizeof(double) * 12 * (size_t)N, cudaMemcpyDeviceToDevice, s));
    LMState st{};
    st.lambda = 1e-3; st.param_length = std::sqrt(L[0]);
    st.tau = o.tau; st.grad_thr = o.gradient_threshold; st.upd_thr = o.update_threshold; st.cg_tol = o.cg_rel_tol;
    st.iter = 0; st.status = 0; st.done = (o.max_iterations <= 0) ? 1 : 0; st.compute = 1; st.cur = 0; st.first = 1;
    st.cg_active = 0; st.solve_ok = 1; st.max_iter = o.max_iterations; st.min_iter = o.min_iterations;
    st.n_trace = 0; st.trace_cap = h->trace_cap;
    st.variant = o.lm_variant == 1 ? 1 : 0; st.cg_min_steps = o.cg_min_steps; st.damp_adjust = 2.0;
    st.cg_guard = st.variant ? 1.0 : 0.0;                        // __cg_norm_guard (ConfigBA.cpp:62)
    if (st.variant) {
        st.lambda = 1e-3; st.first = 1;                          // __lm_initial_damp (ConfigBA.cpp:45)
        double kt[1] = {(double)h->K};
        if (int rc = global_sum(h, kt, 1)) return rc;
        st.mse_scale = h->K5[0] * h->K5[0] / std::max(1.0, kt[0]); st.mse_thr = 0.25;   // cost is in (pixel / f0)^2 units
        st.delta_thr = 1e-6; st.norm_scale = 0.5;                // PBA's residuals are pixel x 0.5 / f (focal normalised to __data_normalize_median)
    }
    h->h_st = st; h->h_st_init = st;
    CU(cudaMemcpyAsync(h->st.p, &h->h_st, sizeof st, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    h->launches = 0; h->ms_solve = 0; h->iters_before_restart = 0; h->accepts_before_restart = 0;
    if (int rc = mean_reproj(h, 0, &h->mean_px[0])) return rc;
    {
        const bool fs = h->use_dense || (h->world == 1 && getenv("B200BA_NO_FUSE_SMALL") == nullptr);
        const bool sp = h->world == 1 && (h->use_pdl == 1 || h->use_pdl == 2) && getenv("B200BA_NO_STEP_PDL") == nullptr;
        if (fs != h->fuse_small || sp != h->step_pdl) { h->fuse_small = fs; h->step_pdl = sp; if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; } }
    }
    {
        // experiment knob: access policy window (persisting hits, streaming misses) on v [, Vinv] of this handle's stream
        const char* e = getenv("B200BA_L2_WINDOW");
        if (e && atoi(e) > 0) {
            cudaStreamAttrValue av = {};
            char* lo = atoi(e) == 2 ? (char*)h->Vinv.p : (char*)h->v.p;
            char* hi = (char*)h->v.p + sizeof(double) * 4 * (size_t)h->M;
            if (atoi(e) == 2 && (char*)h->Vinv.p > (char*)h->v.p) { lo = (char*)h->v.p; hi = (char*)h->Vinv.p + sizeof(double) * 6 * (size_t)h->M; }
            av.accessPolicyWindow.base_ptr = lo; av.accessPolicyWindow.num_bytes = (size_t)(hi - lo);
            av.accessPolicyWindow.hitRatio = getenv("B200BA_L2_RATIO") ? (float)atof(getenv("B200BA_L2_RATIO")) : 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t pe = cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            if (getenv("B200BA_VERBOSE")) fprintf(stderr, "[b200ba] L2 window %zu bytes: %s\n", av.accessPolicyWindow.num_bytes, cudaGetErrorString(pe));
            cudaGetLastError();
        }
    }
    h->begun = true; h->finished = false;
    return 0;
}

int b200ba_restart(b200ba_handle* h)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_restart before b200ba_begin");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t xb = sizeof(double) * 4 * (size_t)std::max(h->M, 1), rb = sizeof(double) * 12 * (size_t)h->N;
    {   // keep the totals of the run that ends here (reported by b200ba_finish)
        LMState cur; CU(cudaMemcpyAsync(&cur, h->st.p, sizeof cur, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        h->iters_before_restart += cur.iter; h->accepts_before_restart += cur.n_accept;
    }
    CU(cudaEventRecord(h->ev0, s));
    CU(cudaMemcpyAsync(h->X0.p, h->X_init.p, xb, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(h->X1.p, h->X_init.p, xb, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(h->rt0.p, h->rt_init.p, rb, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(h->rt1.p, h->rt_init.p, rb, cudaMemcpyDeviceToDevice, s));
    LMState st = h->h_st_init;
    CU(cudaMemcpyAsync(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(h->ev1, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0; CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->ms_last_restart = ms;
    h->finished = false;
    return 0;
}

int b200ba_last_restart_ms(b200ba_handle* h, double* ms)
{
    if (!h || !ms) return B200BA_E_INVALID;
    *ms = h->ms_last_restart;
    return 0;
}

int b200ba_iterate(b200ba_handle* h, int32_t n, int32_t* done)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_iterate before b200ba_begin");
    CU(cudaSetDevice(h->device));
    int poll = h->opt.poll_every > 0 ? h->opt.poll_every : n;
    int is_done = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    for (int it = 0; it < n && !is_done; ) {
        int batch = std::min(poll, n - it);
        for (int b = 0; b < batch; ++b) if (int rc = enqueue_lm_iteration(h)) return rc;
        it += batch;
        CU(cudaMemcpyAsync(&is_done, &h->d.st->done, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (h->p2p_ready) {
            int perr = 0; CU(cudaMemcpy(&perr, h->d.p2p.err, sizeof(int), cudaMemcpyDeviceToHost));
            if (perr) return fail(h, B200BA_E_COMM, "a peer rank did not publish its contribution in time (peer-memory all-reduce)");
        }
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->ms_solve += ms; h->ms_last_iterate = ms;
    CU(cudaGetLastError());
    if (done) *done = is_done;
    return 0;
}

int b200ba_last_iterate_ms(b200ba_handle* h, double* ms)
{
    if (!h || !ms) return B200BA_E_INVALID;
    *ms = h->ms_last_iterate;
    return 0;
}

int b200ba_finish(b200ba_handle* h, b200ba_report* rep)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_finish before b200ba_begin");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(&h->h_st, h->st.p, sizeof(LMState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (int rc = mean_reproj(h, h->h_st.cur, &h->mean_px[1])) return rc;
    h->finished = true;
    if (rep) {
        double* tr = rep->trace; int cap = rep->trace_cap;
        rep->status = h->h_st.status; rep->iterations = h->h_st.iter;
        rep->ret = (h->opt.discard_converged && h->h_st.status == B200BA_STATUS_CONVERGED) ? -1 : 0;
        rep->total_cg_steps = h->h_st.total_cg;
        rep->iterations_total = (int32_t)(h->iters_before_restart + h->h_st.iter);
        rep->accepted_total = (int32_t)(h->accepts_before_restart + h->h_st.n_accept);
        rep->gpu_launches = (int32_t)std::min<long long>(h->launches, 2147483647LL);
        rep->mean_px[0] = h->mean_px[0]; rep->mean_px[1] = h->mean_px[1];
        rep->ms_solve = h->ms_solve; rep->ms_h2d = h->ms_h2d; rep->ms_d2h = h->ms_d2h;
        int n = std::min(h->h_st.n_trace, h->trace_cap);
        std::vector<double> t((size_t)std::max(n, 1) * TRACE_COLS);
        if (n) { CU(cudaMemcpy(t.data(), h->trace.p, sizeof(double) * (size_t)n * TRACE_COLS, cudaMemcpyDeviceToHost)); }
        rep->initial_cost = n ? t[1] : h->h_st.err;
        rep->final_cost = h->h_st.err;
        if (n && t[(size_t)(n - 1) * TRACE_COLS + 5] == 1.0) rep->final_cost = t[(size_t)(n - 1) * TRACE_COLS + 3];
        rep->n_trace = 0;
        if (tr && cap > 0) { int m = std::min(n, cap); memcpy(tr, t.data(), sizeof(double) * (size_t)m * TRACE_COLS); rep->n_trace = m; }
    }
    return 0;
}

int b200ba_solve(b200ba_handle* h, b200ba_report* rep)
{
    if (int rc = b200ba_begin(h)) return rc;
    int done = 0;
    if (int rc = b200ba_iterate(h, h->opt.max_iterations, &done)) return rc;
    return b200ba_finish(h, rep);
}

int b200ba_get_result(b200ba_handle* h, double* cam_RT, double* pts)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->finished) return fail(h, B200BA_E_STATE, "b200ba_get_result before b200ba_solve / b200ba_finish");
    double t0 = now_ms();
    CU(cudaSetDevice(h->device));
    const int N = h->N, M = h->M;
    bool discard = h->opt.discard_converged && h->h_st.status == B200BA_STATUS_CONVERGED;   // src/sfm_ba_ssba.cpp:221,235
    if (discard) {
        if (cam_RT) memcpy(cam_RT, h->h_cam.data(), sizeof(double) * 12 * (size_t)N);
        if (pts) memcpy(pts, h->h_pts.data(), sizeof(double) * 3 * (size_t)M);
        return 0;
    }
    int cur = h->h_st.cur;
    if (cam_RT) {
        CU(cudaMemcpy(cam_RT, h->d.rt[cur], sizeof(double) * 12 * (size_t)N, cudaMemcpyDeviceToHost));
        if (h->opt.normalize)   // ScopedBundleExtrinsicNormalizer dtor (v3d_metricbundle.h:358-375)
            for (int i = 0; i < N; ++i) {
                double C[3]; camera_centre(cam_RT + 12 * (size_t)i, C);
                for (int c = 0; c < 3; ++c) C[c] = C[c] * h->scale + h->mean[c];
                set_camera_centre(cam_RT + 12 * (size_t)i, C);
            }
    }
    if (pts && M) {
        PinnedPool::Lease stage;
        g_pinned.lease(stage, sizeof(double) * 4 * (size_t)M);
        double* X4 = reinterpret_cast<double*>(stage.p);
        CU(cudaMemcpy(X4, h->d.X[cur], sizeof(double) * 4 * (size_t)M, cudaMemcpyDeviceToHost));
        const int T = (M > 200000) ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
        const bool nrm = h->opt.normalize != 0;
        parallel_for(T, [&](int t, int nt) {
            const int a = (int)((long long)M * t / nt), b = (int)((long long)M * (t + 1) / nt);
            for (int jn = a; jn < b; ++jn) {
                const int j = h->orig_of_new[jn];
                for (int c = 0; c < 3; ++c) pts[3 * (size_t)j + c] = nrm ? X4[4 * (size_t)jn + c] * h->scale + h->mean[c] : X4[4 * (size_t)jn + c];
            }
        });
    }
    h->ms_d2h = now_ms() - t0;
    return 0;
}

int b200ba_problem_size(const b200ba_handle* h, int32_t* N, int32_t* M, int32_t* K)
{
    if (!h || !h->have_problem) return B200BA_E_STATE;
    if (N) *N = h->N;
    if (M) *M = h->M;
    if (K) *K = h->K;
    return 0;
}

int b200ba_get_plan(const b200ba_handle* h, b200ba_plan* plan)
{
    if (!h || !plan) return B200BA_E_INVALID;
    if (!h->have_problem) return B200BA_E_STATE;
    memset(plan, 0, sizeof *plan);
    plan->struct_size = (int32_t)sizeof *plan;
    plan->camera_table = h->use_tab ? B200BA_TABLE_SHARED : B200BA_TABLE_GLOBAL;
    plan->cluster_size = 1;
    plan->cooperative_step = h->use_coop ? 1 : 0;
    plan->cuda_graph = (h->use_graph && (h->use_dense || !(h->opt.cg_rel_tol > 0))) ? 1 : 0;
    plan->peer_windows = h->p2p_ready ? 1 : 0;
    plan->world = h->world; plan->rank = h->rank;
    plan->launches_per_iteration = (int32_t)h->launches_per_graph;
    plan->n_slices = h->d.n_slices; plan->padded_entries = h->Kp;
    plan->n_rows = h->d.n_rows; plan->n_segments = h->d.n_segs;
    plan->fp32_pcg = h->use_f32 ? 1 : 0;
    plan->device_bytes = (int64_t)h->arena.cap;
    plan->explicit_schur = h->use_dense ? h->dense_ctas : 0;
    plan->explicit_pairs = h->use_dense ? (int32_t)h->dense_pairs : 0;
    return 0;
}

// Point half of b200ba_get_result without leaving the device: caller order restored and the normaliser undone by a kernel
// (X * scale + mean evaluated as a rounded product and a rounded sum, like the host loop above: identical bits).
__global__ void k_export_points(const double* __restrict__ X4, const int* __restrict__ orig_of_new, int M, int nrm,
                                double scale, double m0, double m1, double m2, double* __restrict__ out)
{
    const int jn = blockIdx.x * blockDim.x + threadIdx.x;
    if (jn >= M) return;
    const size_t j = (size_t)orig_of_new[jn];
    const double x = X4[4 * (size_t)jn], y = X4[4 * (size_t)jn + 1], z = X4[4 * (size_t)jn + 2];
    out[3 * j] = nrm ? __dadd_rn(__dmul_rn(x, scale), m0) : x;
    out[3 * j + 1] = nrm ? __dadd_rn(__dmul_rn(y, scale), m1) : y;
    out[3 * j + 2] = nrm ? __dadd_rn(__dmul_rn(z, scale), m2) : z;
}

int b200ba_get_result_device(b200ba_handle* h, double* cam_RT, double* d_pts)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->finished) return fail(h, B200BA_E_STATE, "b200ba_get_result_device before b200ba_solve / b200ba_finish");
    if (int rc = b200ba_get_result(h, cam_RT, nullptr)) return rc;
    const int M = h->M;
    if (!d_pts || !M) return 0;
    CU(cudaSetDevice(h->device));
    const bool discard = h->opt.discard_converged && h->h_st.status == B200BA_STATUS_CONVERGED;
    if (discard) {                                               // results discarded: the inputs, unchanged
        CU(cudaMemcpy(d_pts, h->h_pts.data(), sizeof(double) * 3 * (size_t)M, cudaMemcpyHostToDevice));
        return 0;
    }
    int* d_perm = nullptr;
    CU(cudaMalloc((void**)&d_perm, sizeof(int) * (size_t)M));
    cudaError_t e = cudaMemcpyAsync(d_perm, h->orig_of_new.data(), sizeof(int) * (size_t)M, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        k_export_points<<<cdiv(M, 256), 256, 0, h->stream>>>(h->d.X[h->h_st.cur], d_perm, M, h->opt.normalize != 0, h->scale,
                                                             h->mean[0], h->mean[1], h->mean[2], d_pts);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_perm);
    if (e != cudaSuccess) return fail(h, B200BA_E_CUDA, "b200ba_get_result_device: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// stage-level access
// ---------------------------------------------------------------------------------------------------------
int b200ba_debug_linearize(b200ba_handle* h, double* cost, double* w, double* gc, double* gp, double* U, double* V, double* lambda0)
{
    if (!h) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_debug_linearize before b200ba_begin");
    CU(cudaSetDevice(h->device));
    if (int rc = enqueue_linearize(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    LMState st; CU(cudaMemcpy(&st, h->st.p, sizeof st, cudaMemcpyDeviceToHost));
    const int N = h->N, M = h->M, K = h->K;
    if (cost) *cost = st.err;
    if (lambda0) *lambda0 = st.lambda;
    if (w && K) {
        std::vector<unsigned char> t((size_t)(h->Kp / 32) * OREC_BYTES);
        CU(cudaMemcpy(t.data(), h->o_rec.p, t.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < K; ++k) { const int e = h->entry_of_obs[k]; memcpy(&w[k], t.data() + (size_t)(e >> 5) * OREC_BYTES + 128 + 8 * (e & 31), sizeof(double)); }
    }
    if (gc) CU(cudaMemcpy(gc, h->gc.p, sizeof(double) * 6 * (size_t)N, cudaMemcpyDeviceToHost));
    if (U) CU(cudaMemcpy(U, h->U.p, sizeof(double) * 36 * (size_t)N, cudaMemcpyDeviceToHost));
    if (gp && M) {
        std::vector<double> t(4 * (size_t)M); CU(cudaMemcpy(t.data(), h->gp.p, sizeof(double) * 4 * (size_t)M, cudaMemcpyDeviceToHost));
        for (int jn = 0; jn < M; ++jn) for (int c = 0; c < 3; ++c) gp[3 * (size_t)h->orig_of_new[jn] + c] = t[4 * (size_t)jn + c];
    }
    if (V && M) {
        std::vector<double> t(6 * (size_t)M); CU(cudaMemcpy(t.data(), h->V.p, sizeof(double) * 6 * (size_t)M, cudaMemcpyDeviceToHost));
        for (int jn = 0; jn < M; ++jn) {
            const double* s = t.data() + 6 * (size_t)jn; double* o = V + 9 * (size_t)h->orig_of_new[jn];
            o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
        }
    }
    return 0;
}

int b200ba_debug_schur_matvec(b200ba_handle* h, double lambda, const double* x, double* Sx)
{
    if (!h || !x || !Sx) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_debug_schur_matvec before b200ba_begin");
    CU(cudaSetDevice(h->device));
    Dev& d = h->d;
    const int N = h->N; size_t n6 = 6 * (size_t)N;
    LMState st; CU(cudaMemcpy(&st, h->st.p, sizeof st, cudaMemcpyDeviceToHost));
    LMState sv = st;
    st.lambda = lambda; st.done = 0; st.cg_active = 1; st.solve_ok = 1;
    CU(cudaMemcpy(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d.p, x, sizeof(double) * n6, cudaMemcpyHostToDevice));
    LAUNCH(k_point_vinv, cdiv(d.M, 256), 256, d, 0);
    {
        // both by-point sweeps read the direction from the packed table: scatter x into its p-slots
        CU(cudaStreamSynchronize(h->stream));
        std::vector<double> tab((size_t)N * TAB_STRIDE);
        CU(cudaMemcpy(tab.data(), d.ptab, sizeof(double) * tab.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; ++i) for (int a = 0; a < 6; ++a) tab[(size_t)i * TAB_STRIDE + 7 + a] = x[6 * (size_t)i + a];
        CU(cudaMemcpy(d.ptab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    }
    launch_point_sweep(h, getenv("B200BA_DEBUG_GENERIC") == nullptr);
    if (int rc = enqueue_camera_half(h, 0)) return rc;
    if (int rc = enqueue_camera_sums(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    std::vector<double> U(36 * (size_t)N), wv(n6);
    CU(cudaMemcpy(U.data(), d.U, sizeof(double) * 36 * (size_t)N, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(wv.data(), d.ar_q, sizeof(double) * n6, cudaMemcpyDeviceToHost));
    for (int i = 0; i < N; ++i) for (int a = 0; a < 6; ++a) {
        double v = lambda * x[6 * (size_t)i + a];
        for (int b = 0; b < 6; ++b) v += U[36 * (size_t)i + 6 * a + b] * x[6 * (size_t)i + b];
        Sx[6 * (size_t)i + a] = v - wv[6 * (size_t)i + a];
    }
    CU(cudaMemcpy(h->st.p, &sv, sizeof sv, cudaMemcpyHostToDevice));
    return 0;
}

// Host-only (works without a CUDA device): the conflict factor of the point-order layout b200ba_set_problem would build -
// expected shared-memory wavefronts of one 16-byte camera-table read per quarter-warp phase (1.0 = conflict-free).
// out[0] = factor with the grouping / slot assignment, out[1] = factor of the plain track-length order, out[2] = milliseconds
// the layout build took, out[3] = padded SELL entries / K.
int b200ba_debug_layout_stats(int32_t N, int32_t M, int32_t K, const int32_t* obs_cam, const int32_t* obs_pt, int32_t lane_window, double* out)
{
    if (N <= 0 || M <= 0 || K <= 0 || !obs_cam || !obs_pt || !out) return B200BA_E_INVALID;
    for (int k = 0; k < K; ++k) if (obs_cam[k] < 0 || obs_cam[k] >= N || obs_pt[k] < 0 || obs_pt[k] >= M || (k && obs_pt[k] < obs_pt[k - 1])) return B200BA_E_INVALID;
    const int T = (K > 200000) ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
    PointLayout a, b;
    const double t0 = now_ms();
    build_point_layout(M, K, obs_cam, obs_pt, lane_window, T, a);
    out[2] = now_ms() - t0;
    build_point_layout(M, K, obs_cam, obs_pt, 0, T, b);
    out[0] = layout_conflict_factor(M, obs_cam, a);
    out[1] = layout_conflict_factor(M, obs_cam, b);
    long long kp = 0; for (int g = 0; 32 * g < M; ++g) kp += 32ll * a.deg_new[(size_t)32 * g];
    out[3] = (double)kp / (double)K;
    // every observation must appear in exactly one slot of its point
    std::vector<unsigned char> seen((size_t)K, 0);
    for (int j = 0; j < M; ++j) for (int q = a.pstart[j]; q < a.pstart[j + 1]; ++q) { const int k = a.slot_obs[q]; if (k < a.pstart[j] || k >= a.pstart[j + 1] || seen[k]) return B200BA_E_STATE; seen[k] = 1; }
    return 0;
}

int b200ba_debug_dense_lists(int32_t N, int32_t M, int32_t K, const int32_t* obs_cam, const int32_t* obs_pt, int64_t* n,
                             int32_t* pair_a, int32_t* pair_b, int32_t* pair_block, int64_t cap, int32_t n_sm, int64_t* shape)
{
    if (N <= 0 || M <= 0 || K <= 0 || !obs_cam || !obs_pt || !n) return B200BA_E_INVALID;
    for (int k = 0; k < K; ++k) if (obs_cam[k] < 0 || obs_cam[k] >= N || obs_pt[k] < 0 || obs_pt[k] >= M || (k && obs_pt[k] < obs_pt[k - 1])) return B200BA_E_INVALID;
    if (shape) {
        const int cand_kind[4] = {0, 1, 1, 2}, cand_ctas[4] = {8, 8, 16, n_sm};
        DenseShape pick;
        for (int c = 0; c < 4 && pick.kind < 0; ++c) {
            if (cand_ctas[c] <= 0) continue;
            const DenseShape t = dense_shape(N, cand_kind[c], cand_ctas[c]);
            if (t.kind >= 0 && t.smem + 1024 <= (size_t)MAX_SMEM) pick = t;
        }
        shape[0] = pick.kind; shape[1] = pick.ctas; shape[2] = pick.rows; shape[3] = pick.ld; shape[4] = pick.threads; shape[5] = (int64_t)pick.smem;
    }
    std::vector<int> ident((size_t)K);
    for (int k = 0; k < K; ++k) ident[k] = k;
    DenseLists dl;
    if (!build_dense_lists(N, M, K, obs_cam, obs_pt, ident.data(), dl)) return B200BA_E_INVALID;
    n[0] = (int64_t)dl.pairs.size(); n[1] = (int64_t)dl.chunks.size(); n[2] = dl.n_blocks; n[3] = 0;
    if (pair_a && pair_b && pair_block && cap >= (int64_t)dl.pairs.size()) {
        for (const DenseChunk& ch : dl.chunks) {
            if (ch.end - ch.beg > DENSE_CHUNK || ch.end <= ch.beg) return B200BA_E_STATE;
            for (int q = ch.beg; q < ch.end; ++q) { pair_a[q] = dl.pairs[q].a; pair_b[q] = dl.pairs[q].b; pair_block[q] = ch.block; }
        }
        for (int b = 0; b < dl.n_blocks; ++b)                     // chunks of a block are contiguous and in pair order
            for (int c = dl.blk_chunk_beg[b]; c < dl.blk_chunk_beg[b + 1]; ++c) {
                if (dl.chunks[c].block != b) return B200BA_E_STATE;
                if (c > dl.blk_chunk_beg[b] && dl.chunks[c].beg != dl.chunks[c - 1].end) return B200BA_E_STATE;
                if ((dl.chunks[c].diag != 0) != (dl.blk_ik[2 * b] == dl.blk_ik[2 * b + 1])) return B200BA_E_STATE;
            }
    }
    return 0;
}

#ifdef B200BA_STAMPS
// profiling builds only (not part of include/b200ba.h): the clock stamps written by the last launches of the two PCG sweeps
int b200ba_debug_stamps(b200ba_handle* h, unsigned long long* out, int32_t cap)
{
    if (!h || !out) return B200BA_E_INVALID;
    CU(cudaMemcpy(out, h->stamps.p, sizeof(unsigned long long) * (size_t)std::min<int>(cap, 512 * 40), cudaMemcpyDeviceToHost));
    return 0;
}
#endif

static const char* kKernelNames[] = {"linearize_points", "linearize_cameras", "cg_point_pass", "cg_camera_pass",
                                     "camera_reduce6", "cg_update", "schur_diag_cameras", "point_vinv",
                                     "backsub_cost_points", "precond_invert", "cg_point_pass_generic", "cg_camera_step_fused", "cg_step",
                                     "cg_camera_pass_generic", "cg_point_pass_f32", "cg_camera_pass_f32"};
const char* b200ba_kernel_name(int32_t which)
{
    return (which >= 0 && which < (int)(sizeof kKernelNames / sizeof *kKernelNames)) ? kKernelNames[which] : nullptr;
}

int b200ba_debug_time_kernel(b200ba_handle* h, int32_t which, int32_t reps, double* ms_per_launch)
{
    if (!h || !ms_per_launch || reps <= 0 || !b200ba_kernel_name(which)) return B200BA_E_INVALID;
    if (!h->begun) return fail(h, B200BA_E_STATE, "b200ba_debug_time_kernel before b200ba_begin");
    CU(cudaSetDevice(h->device));
    Dev& d = h->d;
    LMState st; CU(cudaMemcpy(&st, h->st.p, sizeof st, cudaMemcpyDeviceToHost));
    LMState sv = st;
    st.done = 0; st.compute = 1; st.cg_active = 1; st.solve_ok = 1; st.cg_tol = 0;
    CU(cudaMemcpy(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice));
    auto one = [&]() {
        switch (which) {
        case 0: if (h->use_tab_backsub && h->quat_ok) { LAUNCH(k_build_qtab, cdiv(d.N, 128), 128, d, 0); k_lin_tab<<<h->tab_grid, TABC_BLOCK, h->cost_smem, h->stream>>>(d); h->launches++; }
                else LAUNCH(k_linearize_points, d.n_pt_blocks, PT_BLOCK, d);
                break;
        case 1: k_linearize_cameras<<<cdiv(d.n_vwarps, CH_BLOCK / 32), CH_BLOCK, LINCAM_SMEM, h->stream>>>(d); h->launches++; break;
        case 2: launch_point_sweep(h); break;
        case 10: LAUNCH(k_cg_point_pass<0>, d.n_pt_blocks, PT_BLOCK, d); break;
        case 11: if (h->use_coop) { if (int rc = launch_camera_step(h, 0)) return rc; } else LAUNCH(k_cg_update, 1, ONE_CTA, d);
                 CU(cudaMemcpyAsync(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice, h->stream)); break;
        case 12: if (int rc = enqueue_cg_step(h)) return rc;
                 CU(cudaMemcpyAsync(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice, h->stream)); break;
        case 3: if (int rc = enqueue_camera_half(h, 0)) return rc; break;
        case 13: if (int rc = enqueue_camera_half(h, 0)) return rc; break;
        case 14: if (h->use_f32) { k_cg_point_pass_tab32<<<h->tab32_grid, TAB32_BLOCK, h->tab32_smem, h->stream>>>(d); h->launches++; } break;
        case 15: if (h->use_f32) { k_cg_camera_pass32<<<h->cam_grid, CW * 32, CAMPASS32_SMEM, h->stream>>>(d); h->launches++; } break;
        case 4: LAUNCH((k_camera_reduce<6, 6>), cdiv((long long)d.N * 6 * 4, 256), 256, d, d.ar_q, 0); break;
        case 5: LAUNCH(k_cg_update, 1, ONE_CTA, d); CU(cudaMemcpyAsync(h->st.p, &st, sizeof st, cudaMemcpyHostToDevice, h->stream)); break;
        case 6: LAUNCH(k_schur_diag_cameras, cdiv(d.n_vwarps, CH_BLOCK / 32), CH_BLOCK, d); break;
        case 7: LAUNCH(k_point_vinv, cdiv(d.M, 256), 256, d, 0); break;
        case 8: enqueue_backsub(h); break;
        case 9: LAUNCH(k_precond_invert, cdiv(6ll * d.N, 192), 192, d, h->opt.precond); break;
        }
        return 0;
    };
    for (int w = 0; w < 3; ++w) one();
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < reps; ++r) one();
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *ms_per_launch = ms / reps;
    CU(cudaMemcpy(h->st.p, &sv, sizeof sv, cudaMemcpyHostToDevice));
    CU(cudaGetLastError());
    return 0;
}

} // extern "C"
