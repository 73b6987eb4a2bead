// tri_stage.cu -- batched two-view triangulation of libb200ba.so (include/b200ba_tri.h), sm_100a only.
// Reference: TriangulatePoints / IterativeLinearLSTriangulation, src/sfm_triangulation.cpp:63-202; the undistortion it
// calls is OpenCV 2.4.9's cvUndistortPoints (modules/imgproc/src/undistort.cpp, an external dependency of the reference:
// 5 fixed-point iterations of x = (x0 - tangential) / radial in double, result stored as float).
//
// One thread per match, everything in registers; the matches are independent, so the kernel streams 16 B in and 44 B out
// per match and is bound by its FP64 arithmetic (up to 10 Householder QR solves of a 4x3 system).
#include "../../include/b200ba_tri.h"
#include "pinned_pool.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

namespace {

struct TriParams {
    double K[9], Kinv[9], P[12], P1[12], KP1[12], dist[8];
    int n_dist;
};

// cvUndistortPoints without R / P: pixel -> normalised, distortion removed by 5 fixed-point iterations; the reference then
// maps back to pixels with fx, cx (src/sfm_triangulation.cpp:141-155) and stores the result in a float KeyPoint.
__device__ __forceinline__ float2 undistort_pixel(float px, float py, const TriParams& t)
{
    const double fx = t.K[0], cx = t.K[2], fy = t.K[4], cy = t.K[5];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double x = ((double)px - cx) * ifx, y = ((double)py - cy) * ify;
    const double x0 = x, y0 = y;
    if (t.n_dist > 0) {
        const double* k = t.dist;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
            const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
    }
    const float ux = (float)x, uy = (float)y;                     // the CV_32FC2 destination of undistortPoints
    return make_float2((float)((double)ux * fx + cx), (float)((double)uy * fy + cy));
}

// Least squares of a 4x3 system by Householder QR; false when a pivot vanishes (rank deficient).
__device__ __forceinline__ bool solve_4x3(double (&A)[4][3], double (&b)[4], double (&x)[3])
{
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) scale = fmax(scale, fabs(A[i][j]));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double nrm = 0.0;
#pragma unroll
        for (int i = k; i < 4; ++i) nrm += A[i][k] * A[i][k];
        nrm = sqrt(nrm);
        if (!(nrm > 1e-14 * scale)) return false;
        const double alpha = A[k][k] > 0 ? -nrm : nrm;
        double v[4] = {0, 0, 0, 0};
        v[k] = A[k][k] - alpha;
#pragma unroll
        for (int i = k + 1; i < 4; ++i) v[i] = A[i][k];
        double vv = 0.0;
#pragma unroll
        for (int i = k; i < 4; ++i) vv += v[i] * v[i];
        const double beta = 2.0 / vv;
#pragma unroll
        for (int j = k + 1; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int i = k; i < 4; ++i) s += v[i] * A[i][j];
            s *= beta;
#pragma unroll
            for (int i = k; i < 4; ++i) A[i][j] -= s * v[i];
        }
        double s = 0.0;
#pragma unroll
        for (int i = k; i < 4; ++i) s += v[i] * b[i];
        s *= beta;
#pragma unroll
        for (int i = k; i < 4; ++i) b[i] -= s * v[i];
        A[k][k] = alpha;
    }
    x[2] = b[2] / A[2][2];
    x[1] = (b[1] - A[1][2] * x[2]) / A[1][1];
    x[0] = (b[0] - A[0][1] * x[1] - A[0][2] * x[2]) / A[0][0];
    return true;
}

__global__ void __launch_bounds__(128)
k_tri_points(const float2* __restrict__ pt1, const float2* __restrict__ pt2, int n, TriParams t,
             double* __restrict__ pts, double* __restrict__ err, float2* __restrict__ pt1_out,
             double* __restrict__ partial /* per CTA: sum of errors */, unsigned* __restrict__ n_degenerate)
{
    __shared__ double s_red[4];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        const float2 a = pt1[i], c = pt2[i];
        const float2 k1 = undistort_pixel(a.x, a.y, t), k2 = undistort_pixel(c.x, c.y, t);
        // u = Kinv (x, y, 1)
        double u[3], w[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            u[r] = t.Kinv[3 * r] * (double)k1.x + t.Kinv[3 * r + 1] * (double)k1.y + t.Kinv[3 * r + 2];
            w[r] = t.Kinv[3 * r] * (double)k2.x + t.Kinv[3 * r + 1] * (double)k2.y + t.Kinv[3 * r + 2];
        }
        double wi = 1.0, wi1 = 1.0, X[3] = {0, 0, 0};
        bool ok = true;
        for (int it = 0; it < 10 && ok; ++it) {                   // Hartley & Sturm: 10 iterations at most (:75)
            double A[4][3], B[4];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                A[0][j] = (u[0] * t.P[8 + j] - t.P[j]) / wi;
                A[1][j] = (u[1] * t.P[8 + j] - t.P[4 + j]) / wi;
                A[2][j] = (w[0] * t.P1[8 + j] - t.P1[j]) / wi1;
                A[3][j] = (w[1] * t.P1[8 + j] - t.P1[4 + j]) / wi1;
            }
            B[0] = -(u[0] * t.P[11] - t.P[3]) / wi;
            B[1] = -(u[1] * t.P[11] - t.P[7]) / wi;
            B[2] = -(w[0] * t.P1[11] - t.P1[3]) / wi1;
            B[3] = -(w[1] * t.P1[11] - t.P1[7]) / wi1;
            ok = solve_4x3(A, B, X);
            const double p2x = t.P[8] * X[0] + t.P[9] * X[1] + t.P[10] * X[2] + t.P[11];
            const double p2x1 = t.P1[8] * X[0] + t.P1[9] * X[1] + t.P1[10] * X[2] + t.P1[11];
            if (fabsf((float)(wi - p2x)) <= 0.0001 && fabsf((float)(wi1 - p2x1)) <= 0.0001) break;     // EPSILON, sfm_triangulation.h:25
            wi = p2x; wi1 = p2x1;
        }
        if (!ok) { X[0] = X[1] = X[2] = nan(""); atomicAdd(n_degenerate, 1u); }
        // reproject into view 2 through K P1, pixel error against the (undistorted, float) measurement (:176-182)
        double q[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) q[r] = t.KP1[4 * r] * X[0] + t.KP1[4 * r + 1] * X[1] + t.KP1[4 * r + 2] * X[2] + t.KP1[4 * r + 3];
        const float rx = (float)(q[0] / q[2]), ry = (float)(q[1] / q[2]);
        const float dx = rx - k2.x, dy = ry - k2.y;
        e = sqrt((double)dx * dx + (double)dy * dy);
        if (pts) { pts[3 * (size_t)i] = X[0]; pts[3 * (size_t)i + 1] = X[1]; pts[3 * (size_t)i + 2] = X[2]; }
        if (err) err[i] = e;
        if (pt1_out) pt1_out[i] = k1;
        if (!ok) e = 0.0;                                          // NaN points stay out of the mean
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
}

__global__ void __launch_bounds__(256) k_tri_mean(const double* __restrict__ partial, int nb, double* out)
{
    __shared__ double s[8];
    double v = 0.0;
    for (int b = threadIdx.x; b < nb; b += 256) v += partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += s[k]; out[0] = t; }
}

// TestTriangulation (src/sfm_cameraMatrices.cpp:180-204): cv::perspectiveTransform with [P; 0 0 0 1] leaves w = 1, so the
// test is the sign of the third row of P applied to the point.  counts[0] / [1]: points in front of P / P1.
__global__ void __launch_bounds__(256)
k_tri_front(const double* __restrict__ pts, int n, TriParams t, unsigned char* __restrict__ front1, unsigned* __restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool f0 = false, f1 = false;
    if (i < n) {
        const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        f0 = t.P[8] * x + t.P[9] * y + t.P[10] * z + t.P[11] > 0;
        f1 = t.P1[8] * x + t.P1[9] * y + t.P1[10] * z + t.P1[11] > 0;
        front1[i] = f1 ? 1 : 0;
    }
    const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
    if ((threadIdx.x & 31) == 0) { if (b0) atomicAdd(&counts[0], (unsigned)__popc(b0)); if (b1) atomicAdd(&counts[1], (unsigned)__popc(b1)); }
}

// ---- exact order statistic by radix selection (the tracker std::sorts all errors to read ONE of them, :738-742) ----
// Order-preserving key of a double (negative values and NaNs included): what a sort of the values would compare.
__device__ __forceinline__ unsigned long long sel_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
constexpr int SEL_BITS = 11, SEL_BINS = 1 << SEL_BITS;
// state[0] = key prefix found so far, state[1] = rank still to find among the elements that share the prefix
// one pass: histogram of the next SEL_BITS key bits over those elements (integer counters: exact, order-independent)
__global__ void __launch_bounds__(256)
k_sel_hist(const double* __restrict__ v, int n, int shift, int bits, const unsigned long long* __restrict__ state, unsigned* __restrict__ hist)
{
    __shared__ unsigned sh[SEL_BINS];
    for (int b = threadIdx.x; b < SEL_BINS; b += 256) sh[b] = 0;
    __syncthreads();
    const int top = shift + bits;
    const unsigned bin_mask = (1u << bits) - 1u;
    const unsigned long long mask = top >= 64 ? 0ull : (~0ull << top), prefix = state[0] & mask;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const unsigned long long k = sel_key(v[i]);
        if ((k & mask) == prefix) atomicAdd(&sh[(unsigned)(k >> shift) & bin_mask], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < SEL_BINS; b += 256) if (sh[b]) atomicAdd(&hist[b], sh[b]);
}
// narrow to the bin that holds the wanted rank; clears the histogram for the next pass; the last pass writes the value
__global__ void __launch_bounds__(32) k_sel_pick(int shift, unsigned long long* __restrict__ state, unsigned* __restrict__ hist, double* __restrict__ value_out)
{
    if (threadIdx.x == 0) {
        unsigned long long rank = state[1], cum = 0;
        int b = 0;
        for (; b < SEL_BINS - 1; ++b) { if (cum + hist[b] > rank) break; cum += hist[b]; }
        state[0] |= (unsigned long long)b << shift;
        state[1] = rank - cum;
        if (shift == 0) {
            const unsigned long long k = state[0], bits = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
            *value_out = __longlong_as_double((long long)bits);
        }
    }
    __syncwarp();
    for (int b = threadIdx.x; b < SEL_BINS; b += 32) hist[b] = 0;
}

// keep[i] = in front of P1 and error < 2.4 x sorted_err[4 n / 5]  (src/sfm_tracker.cpp:736-761)
__global__ void __launch_bounds__(256)
k_tri_keep(const double* __restrict__ err, const double* __restrict__ pct80, int n, const unsigned char* __restrict__ front1,
           unsigned char* __restrict__ keep, unsigned* __restrict__ counts, double* __restrict__ cutoff_out)
{
    const double cutoff = __dmul_rn(pct80[0], 2.4);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool k = false;
    if (i < n) { k = front1[i] && err[i] < cutoff; keep[i] = k ? 1 : 0; }
    const unsigned b = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counts[2], (unsigned)__popc(b));
    if (i == 0) cutoff_out[0] = cutoff;
}

thread_local std::string g_tri_create_error;

}  // namespace

struct b200tri_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    char* d_buf = nullptr; size_t cap = 0;            // one grow-only arena: pt1 | pt2 | pts | err | pt1_out | partial | scalars
    char* d_gate = nullptr; size_t cap_gate = 0;      // gating scratch: selected value | select state | histogram | front flags | keep | counters
    int n_last = 0; size_t o_pts_last = 0, o_err_last = 0; TriParams params_last;   // resident results of the last call
};

#define TRI_CUDA(call)                                                                                          \
    do { cudaError_t e_ = (call);                                                                               \
         if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return B200BA_E_CUDA; } } while (0)

extern "C" {

int b200tri_create(int32_t device, b200tri_handle** out)
{
    if (!out) return B200BA_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_tri_create_error = std::string("b200tri_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                             "); the triangulation has no CPU fallback";
        return B200BA_E_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { g_tri_create_error = "b200tri_create: device ordinal out of range"; return B200BA_E_INVALID; }
    b200tri_handle* h = new b200tri_handle();
    h->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    if (!ok) { g_tri_create_error = std::string("b200tri_create: ") + cudaGetErrorString(cudaGetLastError()); b200tri_destroy(h); return B200BA_E_CUDA; }
    *out = h;
    return B200BA_OK;
}

void b200tri_destroy(b200tri_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->d_gate) cudaFree(h->d_gate);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* b200tri_last_error(const b200tri_handle* h) { return h ? h->err.c_str() : g_tri_create_error.c_str(); }

int b200tri_triangulate(b200tri_handle* h, int32_t n, const float* pt1, const float* pt2, const double* K9,
                        const double* Kinv9, const double* dist, int32_t n_dist, const double* P, const double* P1,
                        double* pts, double* reproj_err, float* pt1_undistorted, b200tri_report* report)
{
    if (!h) return B200BA_E_INVALID;
    if (n < 0 || (n > 0 && (!pt1 || !pt2)) || !K9 || !Kinv9 || !P || !P1 || !report) { h->err = "b200tri_triangulate: null argument"; return B200BA_E_INVALID; }
    if (!(n_dist == 0 || n_dist == 4 || n_dist == 5 || n_dist == 8) || (n_dist > 0 && !dist)) { h->err = "b200tri_triangulate: n_dist must be 0, 4, 5 or 8"; return B200BA_E_INVALID; }
    std::memset(report, 0, sizeof(*report));
    h->n_last = 0;
    if (n == 0) return B200BA_OK;                                 // cv::mean of an empty vector is 0
    TRI_CUDA(cudaSetDevice(h->device));
    TriParams t;
    std::memset(&t, 0, sizeof(t));
    std::memcpy(t.K, K9, sizeof(t.K)); std::memcpy(t.Kinv, Kinv9, sizeof(t.Kinv));
    std::memcpy(t.P, P, sizeof(t.P)); std::memcpy(t.P1, P1, sizeof(t.P1));
    for (int k = 0; k < n_dist; ++k) t.dist[k] = dist[k];
    bool any = false;
    for (int k = 0; k < n_dist; ++k) any = any || dist[k] != 0.0;
    t.n_dist = any ? n_dist : 0;                                  // all-zero coefficients: the iterations are the identity
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) {     // KP1 = K * P1 (:114)
        double s = 0;
        for (int k = 0; k < 3; ++k) s += K9[3 * r + k] * P1[4 * k + c];
        t.KP1[4 * r + c] = s;
    }
    const int nb = (n + 127) / 128;
    const size_t o_pt1 = 0, o_pt2 = o_pt1 + (size_t)n * 8, o_pts = o_pt2 + (size_t)n * 8, o_err = o_pts + (size_t)n * 24,
                 o_out = o_err + (size_t)n * 8, o_part = o_out + (size_t)n * 8, o_sc = o_part + (size_t)nb * 8, total = o_sc + 16;
    if (total > h->cap) {
        if (h->d_buf) cudaFree(h->d_buf);
        h->d_buf = nullptr; h->cap = 0;
        TRI_CUDA(cudaMalloc((void**)&h->d_buf, total + total / 4));
        h->cap = total + total / 4;
    }
    char* d = h->d_buf;
    // host <-> device through the library's pinned staging pool (pinned_pool.h): [pt1 | pt2 | pts | err | pt1_out]
    b200ba::PinnedPool::Lease stage;
    b200ba::g_pinned.lease(stage, o_part);
    TRI_CUDA(cudaEventRecord(h->ev[0], h->stream));
    b200ba::threaded_copy(stage.p + o_pt1, pt1, (size_t)n * 8);
    b200ba::threaded_copy(stage.p + o_pt2, pt2, (size_t)n * 8);
    TRI_CUDA(cudaMemcpyAsync(d + o_pt1, stage.p + o_pt1, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
    TRI_CUDA(cudaMemsetAsync(d + o_sc, 0, 16, h->stream));
    TRI_CUDA(cudaEventRecord(h->ev[1], h->stream));
    k_tri_points<<<nb, 128, 0, h->stream>>>((const float2*)(d + o_pt1), (const float2*)(d + o_pt2), n, t, (double*)(d + o_pts),
                                            (double*)(d + o_err), (float2*)(d + o_out), (double*)(d + o_part), (unsigned*)(d + o_sc + 8));
    k_tri_mean<<<1, 256, 0, h->stream>>>((const double*)(d + o_part), nb, (double*)(d + o_sc));
    TRI_CUDA(cudaGetLastError());
    TRI_CUDA(cudaEventRecord(h->ev[2], h->stream));
    if (pts || reproj_err || pt1_undistorted)
        TRI_CUDA(cudaMemcpyAsync(stage.p + o_pts, d + o_pts, o_part - o_pts, cudaMemcpyDeviceToHost, h->stream));
    double sc[2];
    TRI_CUDA(cudaMemcpyAsync(sc, d + o_sc, 16, cudaMemcpyDeviceToHost, h->stream));
    TRI_CUDA(cudaStreamSynchronize(h->stream));
    if (pts) b200ba::threaded_copy(pts, stage.p + o_pts, (size_t)n * 24);
    if (reproj_err) b200ba::threaded_copy(reproj_err, stage.p + o_err, (size_t)n * 8);
    if (pt1_undistorted) b200ba::threaded_copy(pt1_undistorted, stage.p + o_out, (size_t)n * 8);
    TRI_CUDA(cudaEventRecord(h->ev[3], h->stream));
    TRI_CUDA(cudaEventSynchronize(h->ev[3]));
    h->n_last = n; h->o_pts_last = o_pts; h->o_err_last = o_err; h->params_last = t;
    unsigned ndeg; std::memcpy(&ndeg, &sc[1], sizeof(ndeg));
    report->mean_reproj_error = sc[0] / (double)n;
    report->n_degenerate = (int32_t)ndeg;
    report->gpu_launches = 2;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]); report->ms_device = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); report->ms_kernels = ms;
    return B200BA_OK;
}

int b200tri_gate(b200tri_handle* h, uint8_t* keep, b200tri_gate_report* report)
{
    if (!h) return B200BA_E_INVALID;
    if (!report) { h->err = "b200tri_gate: null report"; return B200BA_E_INVALID; }
    std::memset(report, 0, sizeof(*report));
    const int n = h->n_last;
    if (n < 10) { report->status = -1; return B200BA_OK; }        // "triangulated point is too few" (:718-721); also: nothing resident
    TRI_CUDA(cudaSetDevice(h->device));
    const double* d_pts = (const double*)(h->d_buf + h->o_pts_last);
    const double* d_err = (const double*)(h->d_buf + h->o_err_last);
    // scratch: [selected value | select state (prefix, rank) | histogram | front flags | keep | counters]
    const size_t o_sel = 0, o_state = 16, o_hist = 32, o_front = o_hist + SEL_BINS * 4, o_keep = o_front + (((size_t)n + 15) & ~(size_t)15),
                 o_cnt = o_keep + (((size_t)n + 15) & ~(size_t)15), total = o_cnt + 32;
    if (total > h->cap_gate) {
        if (h->d_gate) cudaFree(h->d_gate);
        h->d_gate = nullptr; h->cap_gate = 0;
        TRI_CUDA(cudaMalloc((void**)&h->d_gate, total + total / 4));
        h->cap_gate = total + total / 4;
    }
    char* g = h->d_gate;
    unsigned* d_cnt = (unsigned*)(g + o_cnt);
    TRI_CUDA(cudaEventRecord(h->ev[0], h->stream));
    TRI_CUDA(cudaMemsetAsync(g + o_cnt, 0, 32, h->stream));
    k_tri_front<<<(n + 255) / 256, 256, 0, h->stream>>>(d_pts, n, h->params_last, (unsigned char*)(g + o_front), d_cnt);
    // 80th percentile = element 4 n / 5 of the sorted errors (:742), found by radix selection on the key bits, most
    // significant first: 5 passes of 11 bits + one of 9 (exactly the value std::sort would put there)
    const unsigned long long st0[2] = {0ull, (unsigned long long)(4 * (size_t)n / 5)};
    TRI_CUDA(cudaMemsetAsync(g + o_hist, 0, SEL_BINS * 4, h->stream));
    TRI_CUDA(cudaMemcpyAsync(g + o_state, st0, 16, cudaMemcpyHostToDevice, h->stream));
    const int sel_grid = std::max(1, std::min(592, (n + 1023) / 1024));
    int n_sel = 0;
    for (int shift = 64 - SEL_BITS; ; shift = shift >= SEL_BITS ? shift - SEL_BITS : 0) {
        const int bits = shift == 0 ? 64 - 5 * SEL_BITS : SEL_BITS;   // 53, 42, 31, 20, 9 (11 bits each), then the last 9 bits
        k_sel_hist<<<sel_grid, 256, 0, h->stream>>>(d_err, n, shift, bits, (const unsigned long long*)(g + o_state), (unsigned*)(g + o_hist));
        k_sel_pick<<<1, 32, 0, h->stream>>>(shift, (unsigned long long*)(g + o_state), (unsigned*)(g + o_hist), (double*)(g + o_sel));
        n_sel += 2;
        if (shift == 0) break;
    }
    k_tri_keep<<<(n + 255) / 256, 256, 0, h->stream>>>(d_err, (const double*)(g + o_sel), n, (const unsigned char*)(g + o_front),
                                                      (unsigned char*)(g + o_keep), d_cnt, (double*)(g + o_cnt + 16));
    TRI_CUDA(cudaGetLastError());
    unsigned char cnt_host[32];
    TRI_CUDA(cudaMemcpyAsync(cnt_host, g + o_cnt, 32, cudaMemcpyDeviceToHost, h->stream));
    if (keep) TRI_CUDA(cudaMemcpyAsync(keep, g + o_keep, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    TRI_CUDA(cudaEventRecord(h->ev[1], h->stream));
    TRI_CUDA(cudaStreamSynchronize(h->stream));
    unsigned c[3]; std::memcpy(c, cnt_host, sizeof(c));
    std::memcpy(&report->cutoff, cnt_host + 16, sizeof(double));
    report->n_front_P = (int32_t)c[0]; report->n_front_P1 = (int32_t)c[1]; report->n_kept = (int32_t)c[2];
    report->gpu_launches = 2 + n_sel;                              // front test, keep + the selection passes
    float ms = 0.f; cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); report->ms_device = ms;
    // TestTriangulation on P, then on P1 (short-circuit ||, :725-729): less than 75 % in front of either -> -2
    if ((double)c[0] / (double)n < 0.75 || (double)c[1] / (double)n < 0.75) report->status = -2;
    else if (c[2] <= 10) report->status = -3;                      // "Filted points are too few" (:765-768)
    return B200BA_OK;
}

}  // extern "C"
