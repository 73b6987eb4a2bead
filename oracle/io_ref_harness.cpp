// oracle/io_ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the reference's OWN model-file readers and writers (Thirdparty/pba/pba/pba_util.h: LoadModelFile :392-411 ->
// LoadNVM / LoadBundlerOut / LoadBundlerModel, SaveNVM :131, SaveBundlerOut :268, SaveBundlerModel :364) on raw arrays,
// so that the product's b200io_load / b200io_save (include/b200ba_io.h) can be checked against them in both directions:
// files written by the reference must load to the same model, files written by the product must load in the reference.
// Built by oracle/Makefile into oracle/_ref/libio_ref.so from the header where it lies under /root/reference.
#include <cstring>
#include <pba.h>
#include <pba_util.h>

static vector<CameraT> g_cam; static vector<Point3D> g_pt; static vector<Point2D> g_meas;
static vector<int> g_ptidx, g_camidx, g_ptc; static vector<string> g_names;

extern "C" {

int io_ref_load(const char* path, int* N, int* M, int* K)
{
    g_cam.clear(); g_pt.clear(); g_meas.clear(); g_ptidx.clear(); g_camidx.clear(); g_ptc.clear(); g_names.clear();
    if (!LoadModelFile(path, g_cam, g_pt, g_meas, g_ptidx, g_camidx, g_names, g_ptc)) return -1;
    *N = (int)g_cam.size(); *M = (int)g_pt.size(); *K = (int)g_meas.size();
    return 0;
}

// cam_RT N x 12, focal N, radial N, dtype N, pts M x 3, rgb M x 3 (zeros when the format has none), xy K x 2, oc K, op K
void io_ref_copy(double* cam_RT, double* focal, double* radial, int* dtype, double* pts, int* rgb, double* xy, int* oc, int* op)
{
    for (size_t i = 0; i < g_cam.size(); ++i) {
        const CameraT& c = g_cam[i];
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) cam_RT[12 * i + 4 * r + k] = c.m[r][k]; cam_RT[12 * i + 4 * r + 3] = c.t[r]; }
        focal[i] = c.f; radial[i] = c.radial; dtype[i] = c.distortion_type;
    }
    for (size_t j = 0; j < g_pt.size(); ++j) for (int k = 0; k < 3; ++k) {
        pts[3 * j + k] = g_pt[j].xyz[k];
        rgb[3 * j + k] = 3 * j + k < g_ptc.size() ? g_ptc[3 * j + k] : 0;
    }
    for (size_t k = 0; k < g_meas.size(); ++k) { xy[2 * k] = g_meas[k].x; xy[2 * k + 1] = g_meas[k].y; oc[k] = g_camidx[k]; op[k] = g_ptidx[k]; }
}

// kind: 0 BAL / bundler model, 1 NVM, 2 Bundler .out
int io_ref_save(const char* path, int kind, int N, int M, int K, const double* cam_RT, const double* focal, const double* radial,
                const int* dtype, const double* pts, const int* rgb, const double* xy, const int* oc, const int* op)
{
    vector<CameraT> cam(N); vector<Point3D> pt(M); vector<Point2D> meas(K);
    vector<int> ptidx(op, op + K), camidx(oc, oc + K), ptc(rgb, rgb + 3 * (size_t)M); vector<string> names(N, string("img.jpg"));
    for (int i = 0; i < N; ++i) {
        double R[9], T[3];
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) R[3 * r + k] = cam_RT[12 * i + 4 * r + k]; T[r] = cam_RT[12 * i + 4 * r + 3]; }
        cam[i].SetFocalLength(focal[i]); cam[i].SetMatrixRotation(R); cam[i].SetTranslation(T);
        if (dtype[i] == 1) cam[i].SetProjectionDistortion(radial[i]);
        else if (dtype[i] == -1) cam[i].SetMeasumentDistortion(radial[i]);
    }
    for (int j = 0; j < M; ++j) pt[j].SetPoint(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]);
    for (int k = 0; k < K; ++k) meas[k].SetPoint2D(xy[2 * k], xy[2 * k + 1]);
    if (kind == 1) SaveNVM(path, cam, pt, meas, ptidx, camidx, names, ptc);
    else if (kind == 2) SaveBundlerOut(path, cam, pt, meas, ptidx, camidx, names, ptc);
    else SaveBundlerModel(path, cam, pt, meas, ptidx, camidx);
    return 0;
}

}  // extern "C"
