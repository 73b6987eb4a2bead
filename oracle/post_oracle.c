/* oracle/post_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's post-BA stage
 * (SURVEY.md section 8(f) rank 1): RANSAC ground-plane fit, outlier mask, ground axes, re-axis of points and cameras.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's library; the product
 * (sequencesfm_b200/, adapter/, include/) never does.
 *
 * Pinned: tests/test_post_oracle.py checks every function below against oracle/_ref/libpost_ref.so (the reference's own
 * functions compiled in place, see oracle/post_ref_harness.cpp) and against tests/golden/post_*.json made from it.
 *
 * Association order matters for bit-exactness; the reference computes with Eigen 3.1.3 and OpenCV 2.4.9 Matx:
 *   - fixed-size Vector3d reductions (dot, squaredNorm) unroll as a0 + (a1 + a2)
 *         (Thirdparty/eigen3/Eigen/src/Core/Redux.h:76-91)
 *   - fixed-size matrix products accumulate left to right          (Eigen coeff-based product)
 *   - cv::Matx products and dots start from 0 and accumulate left to right
 *   - vector / scalar is a true division                           (Eigen/src/Core/Functors.h:490-499)
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static double dot3e(const double *a, const double *b) { return a[0]*b[0] + (a[1]*b[1] + a[2]*b[2]); }  /* Eigen */
static double dot3c(const double *a, const double *b) { double s = 0; for (int i = 0; i < 3; ++i) s += a[i]*b[i]; return s; } /* cv */
static void normalize3e(double *v) { double n = sqrt(dot3e(v, v)); v[0] /= n; v[1] /= n; v[2] /= n; }
static void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1]*b[2] - a[2]*b[1]; c[1] = a[2]*b[0] - a[0]*b[2]; c[2] = a[0]*b[1] - a[1]*b[0];
}
static int cmp_double(const void *a, const void *b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

/* The hypothesis draw of src/sfm_utils.cpp:837-845: nA = rand() % n, nB, nC redrawn until distinct.  The stream is what
 * rand() would return, in order. */
int post_oracle_draw_triples(int M, int n_ransac, const int *rand_stream, int n_rand, int *triples, int *used)
{
    int pos = 0;
#define NEXT() (pos < n_rand ? rand_stream[pos++] : (pos++, 0))
    for (int i = 0; i < n_ransac; ++i) {
        int a = (int)((unsigned)NEXT() % (unsigned)M), b = a, c = a;
        while (b == a) { if (pos >= n_rand) return -2; b = (int)((unsigned)NEXT() % (unsigned)M); }
        while (c == a || c == b) { if (pos >= n_rand) return -2; c = (int)((unsigned)NEXT() % (unsigned)M); }
        triples[3*i] = a; triples[3*i+1] = b; triples[3*i+2] = c;
    }
#undef NEXT
    if (used) *used = pos;
    return pos <= n_rand ? 0 : -2;
}

/* FindSigmaSquared, src/sfm_utils.cpp:792-805 (the argument is |distance|, not its square, whatever the name says) */
static double find_sigma_squared(double *err, size_t n)
{
    qsort(err, n, sizeof(double), cmp_double);
    double med = err[n / 2];
    double s = 1.4826 * (1 + 5.0 / (double)(n * 2 - 6)) * sqrt(med);
    s = 4.6851 * s;
    return s * s;
}

/* ---- Eigen 3.1.3 JacobiSVD<MatrixXf> of a 3x3 (Eigen/src/SVD/JacobiSVD.h:724-818, Jacobi/Jacobi.h:82-110,
 *      real_2x2_jacobi_svd :395-421), U only; column-major like MatrixXf does not matter for 3x3 scalars ---- */
static void rot_rows(float W[3][3], int p, int q, float c, float s) {   /* applyOnTheLeft: x' = c x + s y, y' = -s x + c y */
    for (int i = 0; i < 3; ++i) { float x = W[p][i], y = W[q][i]; W[p][i] = c*x + s*y; W[q][i] = -s*x + c*y; }
}
static void rot_cols(float W[3][3], int p, int q, float c, float s) {   /* apply_rotation_in_the_plane(col p, col q, (c, s)) */
    for (int i = 0; i < 3; ++i) { float x = W[i][p], y = W[i][q]; W[i][p] = c*x + s*y; W[i][q] = -s*x + c*y; }
}
static void jacobi_svd3f(const float A[3][3], float U[3][3], float sv[3])
{
    float W[3][3];
    memcpy(W, A, sizeof(W));
    memset(U, 0, 9 * sizeof(float)); U[0][0] = U[1][1] = U[2][2] = 1.f;
    const float precision = 2.f * FLT_EPSILON, consider_zero = 2.f * 1.40129846e-45f;
    int finished = 0;
    while (!finished) {
        finished = 1;
        for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {
            float thr = fmaxf(consider_zero, precision * fmaxf(fabsf(W[p][p]), fabsf(W[q][q])));
            if (fmaxf(fabsf(W[p][q]), fabsf(W[q][p])) > thr) {
                finished = 0;
                float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
                float c1, s1;
                float t = m00 + m11, d = m10 - m01;
                if (t == 0.f) { c1 = 0.f; s1 = d > 0.f ? 1.f : -1.f; }
                else { float u = d / t; c1 = 1.f / sqrtf(1.f + u*u); s1 = c1 * u; }
                /* m.applyOnTheLeft(0,1,rot1) */
                { float x0 = m00, y0 = m10; m00 = c1*x0 + s1*y0; m10 = -s1*x0 + c1*y0;
                  float x1 = m01, y1 = m11; m01 = c1*x1 + s1*y1; m11 = -s1*x1 + c1*y1; }
                /* j_right.makeJacobi(m00, m01, m11) */
                float cr, sr;
                if (m01 == 0.f) { cr = 1.f; sr = 0.f; }
                else {
                    float tau = (m00 - m11) / (2.f * fabsf(m01));
                    float w = sqrtf(tau*tau + 1.f);
                    float tt = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
                    float sign_t = tt > 0.f ? 1.f : -1.f;
                    float n = 1.f / sqrtf(tt*tt + 1.f);
                    sr = -sign_t * (m01 / fabsf(m01)) * fabsf(tt) * n;
                    cr = n;
                }
                /* j_left = rot1 * j_right.transpose();  transpose = (cr, -sr) */
                float cl = c1 * cr - s1 * (-sr);
                float sl = c1 * (-sr) + s1 * cr;
                rot_rows(W, p, q, cl, sl);          /* m_workMatrix.applyOnTheLeft(p,q,j_left)                        */
                rot_cols(U, p, q, cl, sl);          /* m_matrixU.applyOnTheRight(p,q,j_left.transpose()) -> rot j_left */
                rot_cols(W, p, q, cr, -sr);         /* m_workMatrix.applyOnTheRight(p,q,j_right) -> rot j_right^T      */
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        float a = fabsf(W[i][i]); sv[i] = a;
        if (a != 0.f) { float z = W[i][i] / a; for (int r = 0; r < 3; ++r) U[r][i] *= z; }
    }
    for (int i = 0; i < 3; ++i) {           /* descending, first maximum wins */
        int pos = i; for (int k = i + 1; k < 3; ++k) if (sv[k] > sv[pos]) pos = k;
        if (sv[pos] == 0.f) break;
        if (pos != i) {
            float t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
            for (int r = 0; r < 3; ++r) { float u = U[r][i]; U[r][i] = U[r][pos]; U[r][pos] = u; }
        }
    }
}

/* From the inlier mean and scatter matrix to plane and plane_se3: src/sfm_utils.cpp:918-974. */
void post_oracle_plane_from_moments(const double mean[3], const double cov[9], double *plane4, double *se3)
{
    float A[3][3], U[3][3], sv[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = (float)cov[3*i + j];
    jacobi_svd3f(A, U, sv);
    float v1[3], v3[3];
    {   /* u.col(k).normalized() in float, dynamic-size reduction = left to right */
        float n1 = sqrtf((U[0][0]*U[0][0] + U[1][0]*U[1][0]) + U[2][0]*U[2][0]);
        float n3 = sqrtf((U[0][2]*U[0][2] + U[1][2]*U[1][2]) + U[2][2]*U[2][2]);
        for (int r = 0; r < 3; ++r) { v1[r] = U[r][0] / n1; v3[r] = U[r][2] / n3; }
    }
    double ax[3] = { v1[0], v1[1], v1[2] }, az[3] = { v3[0], v3[1], v3[2] }, ay[3];
    ax[2] = (-az[0]*ax[0] - az[1]*ax[1]) / az[2];
    normalize3e(ax);
    cross3(az, ax, ay);
    normalize3e(ay);
    double len = dot3e(mean, az);
    plane4[0] = az[0]; plane4[1] = az[1]; plane4[2] = az[2]; plane4[3] = -len;
    const double *R[3] = { ax, ay, az };
    for (int r = 0; r < 3; ++r) {
        /* t = -R * t1 with dynamic-size MatrixXd: Eigen's gemv accumulates column by column, res += R(:,k) * (-1 * t1[k]) */
        double t = 0.0;
        for (int k = 0; k < 3; ++k) t += R[r][k] * (-1.0 * mean[k]);
        se3[4*r] = R[r][0]; se3[4*r+1] = R[r][1]; se3[4*r+2] = R[r][2]; se3[4*r+3] = t;
    }
}

/* PointsPlaneFit_RANSAC, src/sfm_utils.cpp:807-977, with the hypotheses handed in (post_oracle_draw_triples).
 * returns 0, -1 (fewer than 10 points, like the reference) or -3 (every hypothesis degenerate: the reference would read
 * uninitialised memory there). */
int post_oracle_plane_fit(int M, const double *pts, int n_ransac, const int *triples, double k_threshold,
                          double *plane4, double *se3, int *keep, int *best_hyp, double *sigma2_out,
                          int *n_inliers, double *hyp_sigma2, double *hyp_score)
{
    if (M < 10) return -1;
    memset(keep, 0, (size_t)M * sizeof(int));
    double *err = (double*)malloc((size_t)M * sizeof(double)), *srt = (double*)malloc((size_t)M * sizeof(double));
    double bmean[3] = {0,0,0}, bnorm[3] = {0,0,0}, best = 9e99, bsig = 0; int bi = -1;
    for (int h = 0; h < n_ransac; ++h) {
        const double *A = pts + 3*(size_t)triples[3*h], *B = pts + 3*(size_t)triples[3*h+1], *C = pts + 3*(size_t)triples[3*h+2];
        double mean[3], ca[3], ba[3], n[3];
        for (int k = 0; k < 3; ++k) { mean[k] = ((A[k] + B[k]) + C[k]) / 3.0; ca[k] = C[k] - A[k]; ba[k] = B[k] - A[k]; }
        cross3(ca, ba, n);
        if (hyp_sigma2) hyp_sigma2[h] = -1.0;
        if (hyp_score) hyp_score[h] = -1.0;
        if (fabs(dot3e(n, n)) < 1e-9) continue;
        normalize3e(n);
        for (int i = 0; i < M; ++i) {
            double d[3] = { pts[3*(size_t)i] - mean[0], pts[3*(size_t)i+1] - mean[1], pts[3*(size_t)i+2] - mean[2] };
            err[i] = fabs(dot3e(d, n));
        }
        memcpy(srt, err, (size_t)M * sizeof(double));
        double s2 = find_sigma_squared(srt, (size_t)M) / 20.0;
        double sum = 0.0;
        for (int i = 0; i < M; ++i) sum += err[i] > s2 ? s2 : err[i];
        if (hyp_sigma2) hyp_sigma2[h] = s2;
        if (hyp_score) hyp_score[h] = sum;
        if (sum < best) { best = sum; memcpy(bmean, mean, sizeof(mean)); memcpy(bnorm, n, sizeof(n)); bsig = s2; bi = h; }
    }
    free(srt);
    if (bi < 0) { free(err); return -3; }
    double t_in = bsig, t_flt = bsig * k_threshold;
    double msum[3] = {0,0,0}; size_t cnt = 0;
    for (int i = 0; i < M; ++i) {
        double d[3] = { pts[3*(size_t)i] - bmean[0], pts[3*(size_t)i+1] - bmean[1], pts[3*(size_t)i+2] - bmean[2] };
        double e = fabs(dot3e(d, bnorm));
        err[i] = e;
        if (e < t_in) { msum[0] += pts[3*(size_t)i]; msum[1] += pts[3*(size_t)i+1]; msum[2] += pts[3*(size_t)i+2]; ++cnt; }
        if (e < t_flt) keep[i] = 1;
    }
    double mean[3] = { msum[0] / (double)cnt, msum[1] / (double)cnt, msum[2] / (double)cnt };
    double cov[9] = {0};
    for (int i = 0; i < M; ++i) if (err[i] < t_in) {
        double d[3] = { pts[3*(size_t)i] - mean[0], pts[3*(size_t)i+1] - mean[1], pts[3*(size_t)i+2] - mean[2] };
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3*r+c] += d[r] * d[c];
    }
    free(err);
    post_oracle_plane_from_moments(mean, cov, plane4, se3);
    if (best_hyp) *best_hyp = bi;
    if (sigma2_out) *sigma2_out = bsig;
    if (n_inliers) *n_inliers = (int)cnt;
    return 0;
}

/* ---- cv::Matx34d helpers of src/sfm_utils.cpp:994-1096 (Matx products start from 0, left to right) ---- */
static void local2g(const double *m, double *tg)      /* LocalCoord2G :994-1011: tg = -R^t t */
{
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += (-m[4*k + i]) * m[4*k + 3]; tg[i] = s; }
}
static double point_in_plane(const double *pt, const double *p) { return fabs(p[0]*pt[0] + p[1]*pt[1] + p[2]*pt[2] + p[3]); }
static void point2plane(const double *p, const double *pt, double *out)     /* :1070-1096 */
{
    double d = (p[0]*pt[0] + p[1]*pt[1] + p[2]*pt[2] + p[3]) / sqrt(p[0]*p[0] + p[1]*p[1] + p[2]*p[2]);
    double a[3], b[3];
    for (int k = 0; k < 3; ++k) { a[k] = pt[k] - d*p[k]; b[k] = pt[k] + d*p[k]; }
    double r1 = point_in_plane(a, p), r2 = point_in_plane(b, p);
    const double *c = r1 < r2 ? a : b;
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
    out[1] = (-p[0]*out[0] - p[2]*out[2] - p[3]) / p[1];
}

/* CalcCameraGround_axes, src/sfm_utils.cpp:1098-1146 */
int post_oracle_ground_axes(const double *cam, const double *plane4, const double *plane_se3, double *axes)
{
    (void)plane_se3;   /* p_mean = LocalCoord2G(p_se3) is computed but never used by the reference */
    double cam_p[3], p_int[3], cam_pax[3], q[3], p_ax[3];
    local2g(cam, cam_p);
    point2plane(plane4, cam_p, p_int);
    for (int k = 0; k < 3; ++k) cam_pax[k] = cam_p[k] + cam[k];
    point2plane(plane4, cam_pax, q);
    for (int k = 0; k < 3; ++k) p_ax[k] = q[k] - p_int[k];
    double nx[3], ny[3], nz[3];
    double la = sqrt(dot3c(p_ax, p_ax)), ln = sqrt(dot3c(plane4, plane4));
    /* cv::Vec / double multiplies by the reciprocal (opencv2/core/operations.hpp:1429-1432) */
    for (int k = 0; k < 3; ++k) { nx[k] = p_ax[k] * (1. / la); nz[k] = plane4[k] * (1. / ln); }
    cross3(nz, nx, ny);
    const double *R[3] = { nx, ny, nz };
    for (int r = 0; r < 3; ++r) {
        double s = 0; for (int k = 0; k < 3; ++k) s += (-R[r][k]) * p_int[k];
        axes[4*r] = R[r][0]; axes[4*r+1] = R[r][1]; axes[4*r+2] = R[r][2]; axes[4*r+3] = s;
    }
    return 0;
}

/* TransCloudPoint :1148-1179 (Matrix4d * Vector4d, left to right, the last term is t * 1.0) and
 * TransCamera :1215-1259 (Rc * RtI, Rc * ttI + tc with ttI = -RtI * tt; fixed-size products, left to right). */
int post_oracle_transform(const double *axes, int M, double *pts, int N, double *cams)
{
    for (int i = 0; i < M; ++i) {
        double *p = pts + 3*(size_t)i, x = p[0], y = p[1], z = p[2];
        for (int r = 0; r < 3; ++r) p[r] = ((axes[4*r]*x + axes[4*r+1]*y) + axes[4*r+2]*z) + axes[4*r+3]*1.0;
    }
    double RtI[9], ttI[3];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) RtI[3*i + j] = axes[4*j + i];
    for (int r = 0; r < 3; ++r) ttI[r] = ((-RtI[3*r])*axes[3] + (-RtI[3*r+1])*axes[7]) + (-RtI[3*r+2])*axes[11];
    for (int n = 0; n < N; ++n) {
        double *c = cams + 12*(size_t)n, Rn[9], tn[3];
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) Rn[3*r+k] = (c[4*r]*RtI[k] + c[4*r+1]*RtI[3+k]) + c[4*r+2]*RtI[6+k];
            tn[r] = ((c[4*r]*ttI[0] + c[4*r+1]*ttI[1]) + c[4*r+2]*ttI[2]) + c[4*r+3];
        }
        for (int r = 0; r < 3; ++r) { c[4*r] = Rn[3*r]; c[4*r+1] = Rn[3*r+1]; c[4*r+2] = Rn[3*r+2]; c[4*r+3] = tn[r]; }
    }
    return 0;
}
