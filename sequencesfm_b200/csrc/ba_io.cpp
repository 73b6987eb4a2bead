// ba_io.cpp -- on-disk model formats of libb200ba.so (include/b200ba_io.h): BAL / bundler-model text, VisualSFM .nvm,
// Bundler .out.  Host code only.  Reference: Thirdparty/pba/pba/pba_util.h:52-411 and the camera conversions of
// Thirdparty/pba/pba/DataInterface.h:93-135,183-206,276-316 (cited per function below).
#include "../../include/b200ba_io.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <cmath>
#include <stdexcept>

namespace {

thread_local std::string g_io_error;
int io_fail(int code, const std::string& msg) { g_io_error = msg; return code; }

// Whole-file tokenizer: the formats are whitespace-separated numbers (plus image names in NVM).
struct Tokens {
    std::vector<char> buf; char* p = nullptr; char* end = nullptr;
    bool open(const char* path)
    {
        struct stat sb;
        if (stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) { errno = errno ? errno : EINVAL; return false; }   // no FIFOs, directories, devices
        FILE* f = fopen(path, "rb");
        if (!f) return false;
        if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return false; }
        const long n = ftell(f);
        if (n < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return false; }
        buf.resize((size_t)n + 1);
        const size_t got = fread(buf.data(), 1, (size_t)n, f);
        fclose(f);
        buf[got] = 0; p = buf.data(); end = p + got;
        return true;
    }
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    void skip_comment_lines() { for (skip_ws(); p < end && *p == '#'; skip_ws()) while (p < end && *p != '\n') ++p; }
    bool number(double& v)
    {
        skip_ws();
        if (p >= end) return false;
        char* q = nullptr; v = strtod(p, &q);
        if (q == p || !std::isfinite(v)) return false;                       // "nan" / "inf" tokens are not numbers of these formats
        p = q; return true;
    }
    bool integer(long long& v) { double d; if (!number(d) || std::fabs(d) > 9.0e15) return false; v = (long long)d; return true; }
    bool word(std::string& s)
    {
        skip_ws();
        if (p >= end) return false;
        char* q = p; while (q < end && !(*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r')) ++q;
        s.assign(p, q); p = q; return true;
    }
    char peek() { return p < end ? *p : 0; }
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

template <class F> void io_parallel(int T, F&& fn)
{
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back([&fn, t, T] { fn(t, T); });
    fn(0, T);
    for (auto& x : th) x.join();
}
int io_threads(size_t bytes) { return bytes < ((size_t)8 << 20) ? 1 : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())); }

// A purely numeric file (BAL text) as one flat array of doubles: the buffer is cut at whitespace into one chunk per thread,
// tokens are counted, then converted in place (strtod) - the reference reads the same file through ifstream >> double.
bool parse_all_numbers(const char* p, const char* end, std::vector<double>& out)
{
    const size_t bytes = (size_t)(end - p);
    const int T = io_threads(bytes);
    std::vector<const char*> cut((size_t)T + 1);
    cut[0] = p; cut[T] = end;
    for (int t = 1; t < T; ++t) {
        const char* c = p + bytes * t / T;
        while (c < end && !is_ws(*c)) ++c;                     // never inside a token
        cut[t] = c;
    }
    std::vector<size_t> count((size_t)T + 1, 0);
    io_parallel(T, [&](int t, int) {
        size_t n = 0;
        for (const char* c = cut[t]; c < cut[t + 1];) {
            while (c < cut[t + 1] && is_ws(*c)) ++c;
            if (c >= cut[t + 1]) break;
            ++n;
            while (c < end && !is_ws(*c)) ++c;                 // a token that starts in this chunk belongs to it
        }
        count[t + 1] = n;
    });
    for (int t = 0; t < T; ++t) count[t + 1] += count[t];
    out.resize(count[T]);
    std::vector<char> bad((size_t)T, 0);
    io_parallel(T, [&](int t, int) {
        size_t k = count[t];
        for (const char* c = cut[t]; c < cut[t + 1];) {
            while (c < cut[t + 1] && is_ws(*c)) ++c;
            if (c >= cut[t + 1]) break;
            char* q = nullptr;
            // decimal tokens only: strtod would also take "nan", "inf" and hex floats, none of which these formats hold
            if (!((*c >= '0' && *c <= '9') || *c == '-' || *c == '+' || *c == '.') || (c[0] == '0' && (c[1] == 'x' || c[1] == 'X'))) { bad[t] = 1; return; }
            const double d = strtod(c, &q);                    // the buffer is NUL-terminated
            out[k++] = d;
            if (q == c || (q < end && !is_ws(*q)) || !std::isfinite(d)) { bad[t] = 1; return; }
            c = q;
        }
    });
    for (char b : bad) if (b) return false;
    return true;
}

// Formats rows [0, n) with `row(i, buf)` (returns the characters written, at most 160) on several threads and writes the pieces in order.
template <class F> bool write_rows_parallel(FILE* f, size_t n, F&& row)
{
    const int T = n < 100000 ? 1 : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::string> piece((size_t)T);
    io_parallel(T, [&](int t, int nt) {
        const size_t a = n * t / nt, b = n * (t + 1) / nt;
        std::string& s = piece[t];
        s.reserve((b - a) * 64);
        char buf[192];
        for (size_t i = a; i < b; ++i) s.append(buf, (size_t)row(i, buf));
    });
    for (const std::string& s : piece) if (!s.empty() && fwrite(s.data(), 1, s.size(), f) != s.size()) return false;
    return true;
}

bool ends_with(const char* s, const char* suffix) { return strstr(s, suffix) != nullptr; }   // like LoadModelFile's strstr

// DataInterface.h:93-107 (SetRodriguesRotation), in double
void rodrigues_to_matrix(const double r[3], double m[9])
{
    const double a = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double ct = a == 0.0 ? 0.5 : (1.0 - std::cos(a)) / a / a;
    const double st = a == 0.0 ? 1 : std::sin(a) / a;
    m[0] = 1.0 - (r[1] * r[1] + r[2] * r[2]) * ct; m[1] = r[0] * r[1] * ct - r[2] * st; m[2] = r[2] * r[0] * ct + r[1] * st;
    m[3] = r[0] * r[1] * ct + r[2] * st; m[4] = 1.0 - (r[2] * r[2] + r[0] * r[0]) * ct; m[5] = r[1] * r[2] * ct - r[0] * st;
    m[6] = r[2] * r[0] * ct - r[1] * st; m[7] = r[1] * r[2] * ct + r[0] * st; m[8] = 1.0 - (r[0] * r[0] + r[1] * r[1]) * ct;
}

// Inverse map (what GetRodriguesRotation, DataInterface.h:108-181, computes): angle from the trace, axis from the skew part;
// near pi the axis comes from the symmetric part (largest diagonal entry), its sign from the skew part.
void matrix_to_rodrigues(const double m[9], double r[3])
{
    const double c = std::max(-1.0, std::min(1.0, (m[0] + m[4] + m[8] - 1.0) / 2.0));
    const double sx = m[7] - m[5], sy = m[2] - m[6], sz = m[3] - m[1];
    const double s = 0.5 * std::sqrt(sx * sx + sy * sy + sz * sz);
    const double a = std::atan2(s, c);
    if (s > 1e-6) { const double k = a / (2.0 * s); r[0] = k * sx; r[1] = k * sy; r[2] = k * sz; return; }
    if (c > 0) { r[0] = 0.5 * sx; r[1] = 0.5 * sy; r[2] = 0.5 * sz; return; }           // a ~ 0: R ~ I + [r]x
    const double xx = (m[0] + 1.0) / 2.0, yy = (m[4] + 1.0) / 2.0, zz = (m[8] + 1.0) / 2.0;   // a ~ pi: R ~ 2 n n^t - I
    double n[3];
    if (xx >= yy && xx >= zz) { n[0] = std::sqrt(xx); n[1] = (m[1] + m[3]) / (4.0 * n[0]); n[2] = (m[2] + m[6]) / (4.0 * n[0]); }
    else if (yy >= zz)        { n[1] = std::sqrt(yy); n[0] = (m[1] + m[3]) / (4.0 * n[1]); n[2] = (m[5] + m[7]) / (4.0 * n[1]); }
    else                      { n[2] = std::sqrt(zz); n[0] = (m[2] + m[6]) / (4.0 * n[2]); n[1] = (m[5] + m[7]) / (4.0 * n[2]); }
    if (n[0] * sx + n[1] * sy + n[2] * sz < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    r[0] = a * n[0]; r[1] = a * n[1]; r[2] = a * n[2];
}

// DataInterface.h:183-206 (SetQuaternionRotation), in double
void quaternion_to_matrix(const double q[4], double m[9])
{
    const double qq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double qw = 1, qx = 0, qy = 0, qz = 0;
    if (qq > 0) { qw = q[0] / qq; qx = q[1] / qq; qy = q[2] / qq; qz = q[3] / qq; }
    m[0] = qw * qw + qx * qx - qz * qz - qy * qy; m[1] = 2 * qx * qy - 2 * qz * qw; m[2] = 2 * qy * qw + 2 * qz * qx;
    m[3] = 2 * qx * qy + 2 * qw * qz; m[4] = qy * qy + qw * qw - qz * qz - qx * qx; m[5] = 2 * qz * qy - 2 * qx * qw;
    m[6] = 2 * qx * qz - 2 * qy * qw; m[7] = 2 * qy * qz + 2 * qw * qx; m[8] = qz * qz + qw * qw - qy * qy - qx * qx;
}

void set_camera(b200io_model* m, int i, const double R[9], const double T[3])
{
    double* c = m->cam_RT + 12 * (size_t)i;
    for (int r = 0; r < 3; ++r) { c[4 * r] = R[3 * r]; c[4 * r + 1] = R[3 * r + 1]; c[4 * r + 2] = R[3 * r + 2]; c[4 * r + 3] = T[r]; }
}
// the handedness flip of BAL / .out cameras: rows 2-3 of R and T[1..2] negated (DataInterface.h:287-316)
void set_camera_inverted(b200io_model* m, int i, const double R[9], const double T[3])
{
    double Rf[9], Tf[3] = {T[0], -T[1], -T[2]};
    for (int k = 0; k < 9; ++k) Rf[k] = k < 3 ? R[k] : -R[k];
    set_camera(m, i, Rf, Tf);
}
void get_camera_inverted(const b200io_model* m, int i, double R[9], double T[3])
{
    const double* c = m->cam_RT + 12 * (size_t)i;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R[3 * r + k] = r == 0 ? c[4 * r + k] : -c[4 * r + k];
    T[0] = c[3]; T[1] = -c[7]; T[2] = -c[11];
}

bool check_indices(const b200io_model* m, std::string& why)
{
    for (int k = 0; k < m->K; ++k) {
        if (m->obs_cam[k] < 0 || m->obs_cam[k] >= m->N) { why = "observation " + std::to_string(k) + ": camera index out of range"; return false; }
        if (m->obs_pt[k] < 0 || m->obs_pt[k] >= m->M) { why = "observation " + std::to_string(k) + ": point index out of range"; return false; }
    }
    return true;
}

// ---- BAL / bundler model: pba_util.h:309-362 ----
int load_bal(Tokens& in, b200io_model** out)
{
    std::vector<double> v;
    if (!parse_all_numbers(in.p, in.end, v)) return io_fail(B200BA_E_INVALID, "BAL: not a number where one was expected");
    if (v.size() < 3 || !(v[0] >= 0 && v[1] >= 0 && v[2] >= 0 && v[0] <= 2e9 && v[1] <= 2e9 && v[2] <= 2e9))
        return io_fail(B200BA_E_INVALID, "BAL: bad header (expected: cameras points projections)");
    const long long ncam = (long long)v[0], npt = (long long)v[1], nproj = (long long)v[2];
    const size_t o_obs = 3, o_cam = o_obs + 4 * (size_t)nproj, o_pt = o_cam + 9 * (size_t)ncam, need = o_pt + 3 * (size_t)npt;
    if (v.size() < o_cam) return io_fail(B200BA_E_INVALID, "BAL: truncated observation list");
    if (v.size() < o_pt) return io_fail(B200BA_E_INVALID, "BAL: truncated camera block");
    if (v.size() < need) return io_fail(B200BA_E_INVALID, "BAL: truncated point block");
    b200io_model* m = nullptr;
    if (int rc = b200io_alloc((int32_t)ncam, (int32_t)npt, (int32_t)nproj, 0, &m)) return rc;
    io_parallel(io_threads((size_t)nproj * 32), [&](int t, int nt) {
        for (long long i = nproj * t / nt, e = nproj * (t + 1) / nt; i < e; ++i) {
            const double* o = &v[o_obs + 4 * (size_t)i];
            // an index outside int32 must not wrap into range: -1 fails check_indices
            m->obs_cam[i] = (o[0] >= 0 && o[0] < 2147483647.0) ? (int32_t)o[0] : -1;
            m->obs_pt[i] = (o[1] >= 0 && o[1] < 2147483647.0) ? (int32_t)o[1] : -1;
            m->obs_xy[2 * i] = o[2]; m->obs_xy[2 * i + 1] = -o[3];       // measurements[i].SetPoint2D(x, -y)
        }
    });
    for (long long i = 0; i < ncam; ++i) {
        const double* p = &v[o_cam + 9 * (size_t)i];
        double R[9];
        rodrigues_to_matrix(p, R);
        set_camera_inverted(m, (int)i, R, p + 3);                       // SetInvertedRT(p, p + 3)
        m->focal[i] = p[6];
        m->radial[i] = p[7]; m->distortion_type[i] = 1;                 // SetProjectionDistortion(p[7]); p[8] is not read
    }
    std::memcpy(m->pts, &v[o_pt], 3 * (size_t)npt * sizeof(double));
    *out = m;
    return 0;
}

int save_bal(FILE* f, const b200io_model* m)                             // pba_util.h:364-390
{
    fprintf(f, "%d %d %d\n", m->N, m->M, m->K);
    bool ok = write_rows_parallel(f, (size_t)m->K, [&](size_t k, char* buf) {
        return snprintf(buf, 160, "%d %d %.17g %.17g\n", m->obs_cam[k], m->obs_pt[k], m->obs_xy[2 * k], -m->obs_xy[2 * k + 1]);
    });
    for (int i = 0; i < m->N; ++i) {
        double R[9], T[3], r[3];
        get_camera_inverted(m, i, R, T);                                 // GetInvertedRT
        matrix_to_rodrigues(R, r);
        fprintf(f, "%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0\n", r[0], r[1], r[2], T[0], T[1], T[2], m->focal[i],
                m->distortion_type[i] == 1 ? m->radial[i] : 0.0);
    }
    ok = ok && write_rows_parallel(f, (size_t)m->M, [&](size_t j, char* buf) {
        return snprintf(buf, 160, "%.17g %.17g %.17g\n", m->pts[3 * j], m->pts[3 * j + 1], m->pts[3 * j + 2]);
    });
    return ok ? 0 : io_fail(B200BA_E_INVALID, "save: write failed");
}

// Points with their observation lists, shared by NVM and .out: "x y z r g b n [cam feat x y]*n"
int load_point_lists(Tokens& in, long long npoint, double y_sign, std::vector<double>& pts, std::vector<int32_t>& rgb,
                     std::vector<double>& xy, std::vector<int32_t>& oc, std::vector<int32_t>& op, const char* fmt)
{
    pts.resize(3 * (size_t)npoint); rgb.resize(3 * (size_t)npoint);
    for (long long i = 0; i < npoint; ++i) {
        long long cc[3], npj;
        bool ok = in.number(pts[3 * i]) && in.number(pts[3 * i + 1]) && in.number(pts[3 * i + 2]) &&
                  in.integer(cc[0]) && in.integer(cc[1]) && in.integer(cc[2]) && in.integer(npj);
        if (!ok || npj < 0) return io_fail(B200BA_E_INVALID, std::string(fmt) + ": truncated point record");
        for (int c = 0; c < 3; ++c) rgb[3 * i + c] = (int32_t)std::max(-2147483647LL, std::min(2147483647LL, cc[c]));
        for (long long j = 0; j < npj; ++j) {
            long long cidx, fidx; double x, y;
            if (!in.integer(cidx) || !in.integer(fidx) || !in.number(x) || !in.number(y)) return io_fail(B200BA_E_INVALID, std::string(fmt) + ": truncated observation");
            oc.push_back((cidx >= 0 && cidx < 2147483647LL) ? (int32_t)cidx : -1);    // never wraps into range: -1 fails check_indices
            op.push_back((int32_t)i); xy.push_back(x); xy.push_back(y_sign * y);
        }
    }
    return 0;
}

int finish_model(b200io_model* m, const std::vector<double>& pts, const std::vector<int32_t>& rgb, const std::vector<double>& xy,
                 const std::vector<int32_t>& oc, const std::vector<int32_t>& op)
{
    std::memcpy(m->pts, pts.data(), pts.size() * sizeof(double));
    std::memcpy(m->pt_rgb, rgb.data(), rgb.size() * sizeof(int32_t));
    std::memcpy(m->obs_xy, xy.data(), xy.size() * sizeof(double));
    std::memcpy(m->obs_cam, oc.data(), oc.size() * sizeof(int32_t));
    std::memcpy(m->obs_pt, op.data(), op.size() * sizeof(int32_t));
    return 0;
}

// ---- NVM: pba_util.h:52-128 ----
int load_nvm(Tokens& in, b200io_model** out)
{
    bool r9t = false;
    in.skip_ws();
    if (in.peek() == 'N') { std::string hdr; in.word(hdr); r9t = hdr.find("R9T") != std::string::npos; }
    long long ncam;
    if (!in.integer(ncam) || ncam <= 1) return io_fail(B200BA_E_INVALID, "NVM: needs more than one camera (pba_util.h:71)");
    struct Cam { std::string name; double f, R[9], T[3], d; };
    std::vector<Cam> cams((size_t)ncam);
    for (auto& c : cams) {
        double q[9], cc[3], d1;
        bool ok = in.word(c.name) && in.number(c.f);
        for (int j = 0; ok && j < (r9t ? 9 : 4); ++j) ok = in.number(q[j]);
        ok = ok && in.number(cc[0]) && in.number(cc[1]) && in.number(cc[2]) && in.number(c.d) && in.number(d1);
        if (!ok) return io_fail(B200BA_E_INVALID, "NVM: truncated camera record");
        if (r9t) { std::memcpy(c.R, q, sizeof(c.R)); std::memcpy(c.T, cc, sizeof(c.T)); }          // SetMatrixRotation / SetTranslation
        else {
            quaternion_to_matrix(q, c.R);                                                           // SetQuaternionRotation
            for (int j = 0; j < 3; ++j) c.T[j] = -(c.R[3 * j] * cc[0] + c.R[3 * j + 1] * cc[1] + c.R[3 * j + 2] * cc[2]);   // SetCameraCenterAfterRotation
        }
    }
    long long npoint;
    if (!in.integer(npoint) || npoint <= 0) return io_fail(B200BA_E_INVALID, "NVM: no points (pba_util.h:104)");
    std::vector<double> pts, xy; std::vector<int32_t> rgb, oc, op;
    if (int rc = load_point_lists(in, npoint, 1.0, pts, rgb, xy, oc, op, "NVM")) return rc;
    b200io_model* m = nullptr;
    if (int rc = b200io_alloc((int32_t)ncam, (int32_t)npoint, (int32_t)oc.size(), 1, &m)) return rc;
    for (int i = 0; i < (int)ncam; ++i) {
        set_camera(m, i, cams[i].R, cams[i].T);
        m->focal[i] = cams[i].f;
        m->radial[i] = cams[i].d / (cams[i].f * cams[i].f); m->distortion_type[i] = -1;             // SetNormalizedMeasurementDistortion
        m->names[i] = strdup(cams[i].name.c_str());
    }
    finish_model(m, pts, rgb, xy, oc, op);
    *out = m;
    return 0;
}

// observations of point j as a contiguous run [a, b) - requires point order, which every loader above produces
int save_point_lists(FILE* f, const b200io_model* m, double y_sign, bool nvm)
{
    int k = 0;
    for (int j = 0; j < m->M; ++j) {
        int e = k;
        while (e < m->K && m->obs_pt[e] == j) ++e;
        const int32_t* c = m->pt_rgb + 3 * (size_t)j;
        if (nvm) {
            fprintf(f, "%.17g %.17g %.17g %d %d %d %d ", m->pts[3 * j], m->pts[3 * j + 1], m->pts[3 * j + 2], c[0], c[1], c[2], e - k);
            for (; k < e; ++k) fprintf(f, "%d  0 %.17g %.17g ", m->obs_cam[k], m->obs_xy[2 * k], y_sign * m->obs_xy[2 * k + 1]);
            fprintf(f, "\n");
        } else {
            fprintf(f, "%.17g %.17g %.17g\n%d %d %d\n%d ", m->pts[3 * j], m->pts[3 * j + 1], m->pts[3 * j + 2], c[0], c[1], c[2], e - k);
            for (; k < e; ++k) fprintf(f, "%d 0 %.17g %.17g\n", m->obs_cam[k], m->obs_xy[2 * k], y_sign * m->obs_xy[2 * k + 1]);
            fprintf(f, "\n");
        }
    }
    if (k != m->K) return io_fail(B200BA_E_INVALID, "save: observations are not grouped by point (call b200io_sort_by_point first)");
    return 0;
}

int save_nvm(FILE* f, const b200io_model* m)                             // pba_util.h:131-171 (always the R9T flavour)
{
    fprintf(f, "NVM_V3_R9T\n%d\n", m->N);
    for (int i = 0; i < m->N; ++i) {
        const double* c = m->cam_RT + 12 * (size_t)i;
        fprintf(f, "%s %.17g ", m->names && m->names[i] ? m->names[i] : "unknown", m->focal[i]);
        for (int r = 0; r < 3; ++r) fprintf(f, "%.17g %.17g %.17g ", c[4 * r], c[4 * r + 1], c[4 * r + 2]);
        fprintf(f, "%.17g %.17g %.17g %.17g 0\n", c[3], c[7], c[11],
                m->distortion_type[i] == -1 ? m->radial[i] * (m->focal[i] * m->focal[i]) : 0.0);      // GetNormalizedMeasurementDistortion
    }
    fprintf(f, "%d\n", m->M);
    return save_point_lists(f, m, 1.0, true);
}

// ---- Bundler .out: pba_util.h:173-266 (image names from the list file are not read) ----
int load_out(Tokens& in, b200io_model** out)
{
    in.skip_comment_lines();
    long long ncam, npoint;
    if (!in.integer(ncam) || !in.integer(npoint) || ncam <= 1 || npoint <= 1) return io_fail(B200BA_E_INVALID, "Bundler .out: needs more than one camera and point (pba_util.h:204)");
    struct Cam { double f, d, R[9], T[3]; };
    std::vector<Cam> cams((size_t)ncam);
    for (auto& c : cams) {
        double d1;
        bool ok = in.number(c.f) && in.number(c.d) && in.number(d1);
        for (int j = 0; ok && j < 9; ++j) ok = in.number(c.R[j]);
        for (int j = 0; ok && j < 3; ++j) ok = in.number(c.T[j]);
        if (!ok) return io_fail(B200BA_E_INVALID, "Bundler .out: truncated camera record");
    }
    std::vector<double> pts, xy; std::vector<int32_t> rgb, oc, op;
    if (int rc = load_point_lists(in, npoint, -1.0, pts, rgb, xy, oc, op, "Bundler .out")) return rc;   // Point2D(imx, -imy)
    b200io_model* m = nullptr;
    if (int rc = b200io_alloc((int32_t)ncam, (int32_t)npoint, (int32_t)oc.size(), 0, &m)) return rc;
    for (int i = 0; i < (int)ncam; ++i) {
        set_camera_inverted(m, i, cams[i].R, cams[i].T);                 // SetInvertedR9T
        m->focal[i] = cams[i].f;
        m->radial[i] = cams[i].d; m->distortion_type[i] = 1;             // SetProjectionDistortion(d[0])
    }
    finish_model(m, pts, rgb, xy, oc, op);
    *out = m;
    return 0;
}

int save_out(FILE* f, const b200io_model* m)                             // pba_util.h:268-307
{
    fprintf(f, "# Bundle file v0.3\n%d %d\n", m->N, m->M);
    for (int i = 0; i < m->N; ++i) {
        double R[9], T[3];
        get_camera_inverted(m, i, R, T);                                 // GetInvertedR9T
        fprintf(f, "%.17g %.17g 0\n", m->focal[i], m->distortion_type[i] == 1 ? m->radial[i] : 0.0);
        for (int r = 0; r < 3; ++r) fprintf(f, "%.17g %.17g %.17g\n", R[3 * r], R[3 * r + 1], R[3 * r + 2]);
        fprintf(f, "%.17g %.17g %.17g\n", T[0], T[1], T[2]);
    }
    return save_point_lists(f, m, -1.0, false);
}

}  // namespace

extern "C" {

const char* b200io_last_error(void) { return g_io_error.c_str(); }

int b200io_alloc(int32_t N, int32_t M, int32_t K, int32_t with_names, b200io_model** out)
{
    if (!out || N < 0 || M < 0 || K < 0) return io_fail(B200BA_E_INVALID, "b200io_alloc: bad sizes");
    b200io_model* m = (b200io_model*)calloc(1, sizeof(b200io_model));
    if (!m) return io_fail(B200BA_E_INVALID, "b200io_alloc: out of memory");
    m->N = N; m->M = M; m->K = K;
    auto z = [](size_t n, size_t sz) { return calloc(n ? n : 1, sz); };
    m->cam_RT = (double*)z(12 * (size_t)N, sizeof(double)); m->focal = (double*)z(N, sizeof(double)); m->radial = (double*)z(N, sizeof(double));
    m->distortion_type = (int32_t*)z(N, sizeof(int32_t)); m->pts = (double*)z(3 * (size_t)M, sizeof(double));
    m->pt_rgb = (int32_t*)z(3 * (size_t)M, sizeof(int32_t)); m->obs_xy = (double*)z(2 * (size_t)K, sizeof(double));
    m->obs_cam = (int32_t*)z(K, sizeof(int32_t)); m->obs_pt = (int32_t*)z(K, sizeof(int32_t));
    m->names = with_names ? (char**)z(N, sizeof(char*)) : nullptr;
    if (!m->cam_RT || !m->focal || !m->radial || !m->distortion_type || !m->pts || !m->pt_rgb || !m->obs_xy || !m->obs_cam || !m->obs_pt ||
        (with_names && !m->names)) { b200io_free(m); return io_fail(B200BA_E_INVALID, "b200io_alloc: out of memory"); }
    *out = m;
    return 0;
}

void b200io_free(b200io_model* m)
{
    if (!m) return;
    if (m->names) { for (int i = 0; i < m->N; ++i) free(m->names[i]); free(m->names); }
    free(m->cam_RT); free(m->focal); free(m->radial); free(m->distortion_type); free(m->pts); free(m->pt_rgb);
    free(m->obs_xy); free(m->obs_cam); free(m->obs_pt); free(m);
}

int b200io_load(const char* path, b200io_model** out)
{
    try {
        if (!path || !out) return io_fail(B200BA_E_INVALID, "b200io_load: null argument");
        *out = nullptr;
        Tokens in;
        if (!in.open(path)) return io_fail(B200BA_E_INVALID, std::string("b200io_load: cannot open ") + path + ": " + strerror(errno));
        int rc = ends_with(path, ".nvm") ? load_nvm(in, out) : ends_with(path, ".out") ? load_out(in, out) : load_bal(in, out);
        if (rc) return rc;
        std::string why;
        if (!check_indices(*out, why)) { b200io_free(*out); *out = nullptr; return io_fail(B200BA_E_INVALID, std::string("b200io_load: ") + why); }
        return 0;
    
    } catch (const std::exception& ex) { return io_fail(B200BA_E_INVALID, std::string("b200io_load: ") + ex.what()); }
      catch (...) { return io_fail(B200BA_E_INVALID, "b200io_load: unexpected exception"); }
}

int b200io_save(const char* path, const b200io_model* m)
{
    try {
        if (!path || !m) return io_fail(B200BA_E_INVALID, "b200io_save: null argument");
        std::string why;
        if (!check_indices(m, why)) return io_fail(B200BA_E_INVALID, std::string("b200io_save: ") + why);
        FILE* f = fopen(path, "w");
        if (!f) return io_fail(B200BA_E_INVALID, std::string("b200io_save: cannot open ") + path + ": " + strerror(errno));
        const int rc = ends_with(path, ".nvm") ? save_nvm(f, m) : ends_with(path, ".out") ? save_out(f, m) : save_bal(f, m);
        if (fclose(f) != 0 && rc == 0) return io_fail(B200BA_E_INVALID, std::string("b200io_save: write failed: ") + strerror(errno));
        return rc;
    
    } catch (const std::exception& ex) { return io_fail(B200BA_E_INVALID, std::string("b200io_save: ") + ex.what()); }
      catch (...) { return io_fail(B200BA_E_INVALID, "b200io_save: unexpected exception"); }
}

int b200io_sort_by_point(b200io_model* m)
{
    try {
        if (!m) return io_fail(B200BA_E_INVALID, "b200io_sort_by_point: null model");
        std::vector<int32_t> idx((size_t)m->K);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
            return m->obs_pt[a] != m->obs_pt[b] ? m->obs_pt[a] < m->obs_pt[b] : m->obs_cam[a] < m->obs_cam[b];
        });
        std::vector<double> xy(2 * (size_t)m->K); std::vector<int32_t> oc((size_t)m->K), op((size_t)m->K);
        for (int k = 0; k < m->K; ++k) { xy[2 * k] = m->obs_xy[2 * idx[k]]; xy[2 * k + 1] = m->obs_xy[2 * idx[k] + 1]; oc[k] = m->obs_cam[idx[k]]; op[k] = m->obs_pt[idx[k]]; }
        std::memcpy(m->obs_xy, xy.data(), xy.size() * sizeof(double));
        std::memcpy(m->obs_cam, oc.data(), oc.size() * sizeof(int32_t));
        std::memcpy(m->obs_pt, op.data(), op.size() * sizeof(int32_t));
        return 0;
    
    } catch (const std::exception& ex) { return io_fail(B200BA_E_INVALID, std::string("b200io_sort_by_point: ") + ex.what()); }
      catch (...) { return io_fail(B200BA_E_INVALID, "b200io_sort_by_point: unexpected exception"); }
}

int b200io_shared_focal(b200io_model* m, double* K5)
{
    try {
        if (!m || !K5 || m->N < 1) return io_fail(B200BA_E_INVALID, "b200io_shared_focal: null / empty model");
        std::vector<double> f(m->focal, m->focal + m->N);
        std::nth_element(f.begin(), f.begin() + m->N / 2, f.end());
        const double f0 = f[m->N / 2];
        for (int i = 0; i < m->N; ++i)
            if (!(m->focal[i] > 0) || !std::isfinite(m->focal[i])) return io_fail(B200BA_E_INVALID, "b200io_shared_focal: non-positive focal length");
        for (int k = 0; k < m->K; ++k) {
            const double s = f0 / m->focal[m->obs_cam[k]];
            m->obs_xy[2 * k] *= s; m->obs_xy[2 * k + 1] *= s;
        }
        for (int i = 0; i < m->N; ++i) m->focal[i] = f0;
        K5[0] = f0; K5[1] = 0; K5[2] = 0; K5[3] = f0; K5[4] = 0;
        return 0;
    
    } catch (const std::exception& ex) { return io_fail(B200BA_E_INVALID, std::string("b200io_shared_focal: ") + ex.what()); }
      catch (...) { return io_fail(B200BA_E_INVALID, "b200io_shared_focal: unexpected exception"); }
}

}  // extern "C"
