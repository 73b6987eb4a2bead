// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference PBA (multicore CPU build,
// -DPBA_NO_GPU) that lives under /root/reference/Thirdparty/pba through the calls SequenceSFM makes
// when USE_PBA is defined (src/sfm_ba_pba.cpp:56-166), on raw arrays.  Built by oracle/Makefile into
// oracle/_ref/libpba_ref.so (git-ignored).  Nothing in the product path links it.
//
// What is replayed (reference file:line):
//   * f = (fx+fy)/2, measurements = pixel - (cx,cy) as float        src/sfm_ba_pba.cpp:60-62,122-125
//   * CameraT (float) from [R|T], camera 0 SetConstantCamera         src/sfm_ba_pba.cpp:75-94
//   * SetFixedIntrinsics(true); SetCameraData/PointData/Projection   src/sfm_ba_pba.cpp:136-140
//   * RunBundleAdjustment; copy back points and [R|T]                src/sfm_ba_pba.cpp:142-166
// The device is forced to a CPU one (float or double) because ProgramCU.cu cannot be built with CUDA 12.9.
#include <pba.h>

#include <cstdio>
#include <cstring>
#include <vector>

extern "C" {

// device: -2 = PBA_CPU_FLOAT, -3 = PBA_CPU_DOUBLE.  extra_args: PBA argv-style options ("-schur", "-lmi", "50", ...).
// stats[8] = {initial_mse, final_mse, lm_iterations, cg_iterations, seconds_optimization, return_code,
//             successful_lm_iterations, threads}
int pba_ref_adjust(int N, int M, int K,
                   const double* K5, const double* dist2,
                   double* cam_RT, double* pts,
                   const double* obs_xy, const int* obs_cam, const int* obs_pt,
                   int device, int nargs, const char** extra_args, int nthreads,
                   double* stats)
{
    std::vector<CameraT> camera_data(N);
    std::vector<Point3D> point_data(M);
    std::vector<Point2D> measurements; measurements.reserve(K);
    std::vector<int> camidx(obs_cam, obs_cam + K), ptidx(obs_pt, obs_pt + K);

    float f  = (float)((K5[0] + K5[3]) / 2.0);
    float cx = (float)K5[2], cy = (float)K5[4];
    float d0 = dist2 ? (float)dist2[0] : 0.0f;

    for (int i = 0; i < N; ++i) {
        const double* P = cam_RT + 12 * i;
        float q[9], c[3];
        for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) q[3*r+cc] = (float)P[4*r+cc]; c[r] = (float)P[4*r+3]; }
        camera_data[i].SetFocalLength(f);
        camera_data[i].SetMatrixRotation(q);
        camera_data[i].SetTranslation(c);
        camera_data[i].SetNormalizedMeasurementDistortion(d0);
    }
    camera_data[0].SetConstantCamera();

    for (int j = 0; j < M; ++j) {
        float pt[3] = {(float)pts[3*j], (float)pts[3*j+1], (float)pts[3*j+2]};
        point_data[j].SetPoint(pt);
    }
    for (int k = 0; k < K; ++k) {
        float imx = (float)obs_xy[2*k] - cx;      // Point2f pixel minus float principal point
        float imy = (float)obs_xy[2*k+1] - cy;
        measurements.push_back(Point2D(imx, imy));
    }

    ParallelBA pba((ParallelBA::DeviceT)device);
    std::vector<char*> argv;
    for (int a = 0; a < nargs; ++a) argv.push_back(const_cast<char*>(extra_args[a]));
    char tn[] = "-tnum"; char tv[32];
    if (nthreads > 0) { snprintf(tv, sizeof tv, "%d", nthreads); argv.push_back(tn); argv.push_back(tv); }
    if (!argv.empty()) pba.ParseParam((int)argv.size(), argv.data());
    pba.SetFixedIntrinsics(true);
    pba.SetCameraData(camera_data.size(), &camera_data[0]);
    pba.SetPointData(point_data.size(), &point_data[0]);
    pba.SetProjection(measurements.size(), &measurements[0], &ptidx[0], &camidx[0]);
    int nsuccess = pba.RunBundleAdjustment();

    ConfigBA* cfg = pba.GetInternalConfig();
    if (stats) {
        stats[0] = cfg->GetInitialMSE();
        stats[1] = cfg->GetFinalMSE();
        stats[2] = cfg->GetIterationsLM();
        stats[3] = cfg->GetIterationsCG();
        stats[4] = cfg->GetBundleTiming(1 /* TIMER_OPTIMIZATION, ConfigBA.h:34 */);
        stats[5] = cfg->GetBundleReturnCode();
        stats[6] = nsuccess;
        stats[7] = nthreads;
    }

    for (int j = 0; j < M; ++j) {
        float pt[3]; point_data[j].GetPoint(pt);
        pts[3*j] = pt[0]; pts[3*j+1] = pt[1]; pts[3*j+2] = pt[2];
    }
    for (int i = 0; i < N; ++i) {
        float q[9], c[3];
        camera_data[i].GetMatrixRotation(q);
        camera_data[i].GetTranslation(c);
        double* P = cam_RT + 12 * i;
        for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) P[4*r+cc] = q[3*r+cc]; P[4*r+3] = c[r]; }
    }
    return 0;
}

} // extern "C"
