// layout_host.h -- host-side construction of the by-point (SELL-32) observation layout.  Pure C++ (no CUDA): the same code
// runs inside b200ba_set_problem and, without a device, behind b200ba_debug_layout_stats for the CPU tests.
//
// What is decided here (none of it changes the mathematics - only the order in which a point's observations are summed and
// which 32 points share a warp):
//   1. points are sorted by track length, descending and stable, and cut into slices of 32 (lane = point);
//   2. bank-conflict-free table reads for the shared-memory camera table of the PCG by-point sweep (k_cg_point_pass_tab):
//      the 8 lanes of a quarter-warp read the same 16-byte field of 8 camera records with one LDS.128 phase; the bank group of
//      a record is (camera index mod 8) for the fp64 table (stride 7 x 16 B) and for the fp32 table (5 x 16 B) alike, so the
//      phase is conflict-free exactly when the 8 cameras are distinct mod 8 (or equal).  ncu (round 1 and 2) showed the sweep
//      bound by these replays: 8.5 M shared-memory wavefronts per launch of which 3.9 M were conflicts, and the sweep's time
//      was the same for 16, 19 and 20 warps and for one or two rows in flight.  Two freedoms remove them:
//        (a) points of equal track length d are interchangeable  -> choose the 8 points of a quarter-warp such that no
//            residue class occurs more than d times among their 8 d observations (greedy balancing over a candidate window);
//        (b) the order of a point's observations over the d slots is free -> a bipartite multigraph (8 points x 8 residue
//            classes, 8 d edges) with maximum degree d has a proper edge colouring with d colours (Koenig), i.e. a slot
//            assignment in which every slot sees 8 distinct residues.  Residue classes that still exceed d are split into
//            virtual classes, so only the unavoidable excess collides.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace b200ba {

template <class F> inline void host_parallel_for(int nthreads, F&& fn)
{
    if (nthreads <= 1) { fn(0, 1); return; }
    std::vector<std::thread> th; th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back([&fn, t, nthreads] { fn(t, nthreads); });
    fn(0, nthreads);
    for (auto& x : th) x.join();
}

struct PointLayout {
    std::vector<int> pstart;        // M + 1: first observation of caller point j (obs_pt is non-decreasing)
    std::vector<int> orig_of_new;   // M: caller index of the point placed at new index jn
    std::vector<int> deg_new;       // M: track length by new index (non-increasing)
    std::vector<int> slot_obs;      // K: for caller point j, slot_obs[pstart[j] + s] = observation (caller index) placed in slot s
    int maxdeg = 0;
};

// Proper edge colouring of the (8 points x residue classes) multigraph of one quarter-warp: colour = slot.
// cams[p * d + e] = camera of the e-th observation of point p (p < 8); out_slot_edge[p * d + s] = e placed in slot s.
constexpr int COLOUR_DMAX = 48;                                  // longer tracks (a handful of points) keep their input order
inline void colour_quarter(const int* cams, int d, int* out_slot_edge)
{
    // virtual residue vertices: class rho is split into chunks of at most d edges (sum of chunks <= 16)
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, vbase[8];
    for (int i = 0; i < 8 * d; ++i) cnt[cams[i] & 7]++;
    int nv = 0;
    for (int r = 0; r < 8; ++r) { vbase[r] = nv; nv += (cnt[r] + d - 1) / d; }
    int16_t colP[8 * COLOUR_DMAX], colR[16 * COLOUR_DMAX];        // colP[p][c], colR[v][c] = edge id (valid where the used-mask has the bit)
    int8_t erv[8 * COLOUR_DMAX];                                  // erv[edge] = virtual residue vertex
    uint64_t usedP[8] = {0, 0, 0, 0, 0, 0, 0, 0}, usedR[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // colours in use
    int seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < 8; ++p)
        for (int e = 0; e < d; ++e) { const int r = cams[p * d + e] & 7; erv[p * d + e] = (int8_t)(vbase[r] + seen[r] / d); seen[r]++; }
    for (int p = 0; p < 8; ++p)
        for (int e = 0; e < d; ++e) {
            const int id = p * d + e, v = erv[id];
            const int a = __builtin_ctzll(~usedP[p]), b = __builtin_ctzll(~usedR[v]);      // first free colour on either side
            int c = a;
            if ((usedR[v] >> a) & 1) {                            // a is taken at v
                if (!((usedP[p] >> b) & 1)) c = b;                 // b is free on both sides
                else {
                    // make `a` free at v: swap a <-> b along the alternating path that starts at v with its a-edge.  P vertices
                    // are entered through a-edges only and p has none, so the path never reaches p.
                    int path[32]; int np = 0;
                    int ed = colR[v * d + a];
                    bool at_r = true;
                    while (np < 32) {
                        path[np++] = ed;
                        if (at_r) { const int pp = ed / d; if (!((usedP[pp] >> b) & 1)) break; ed = colP[pp * d + b]; }
                        else { const int vv = erv[ed]; if (!((usedR[vv] >> a) & 1)) break; ed = colR[vv * d + a]; }
                        at_r = !at_r;
                    }
                    for (int i = 0; i < np; ++i) { const int x = path[i], col = (i & 1) ? b : a; usedP[x / d] &= ~(1ull << col); usedR[erv[x]] &= ~(1ull << col); }
                    for (int i = 0; i < np; ++i) {
                        const int x = path[i], col = (i & 1) ? a : b;
                        colP[(x / d) * d + col] = (int16_t)x; colR[erv[x] * d + col] = (int16_t)x; usedP[x / d] |= 1ull << col; usedR[erv[x]] |= 1ull << col;
                    }
                }
            }
            colP[p * d + c] = (int16_t)id; colR[v * d + c] = (int16_t)id; usedP[p] |= 1ull << c; usedR[v] |= 1ull << c;
        }
    for (int p = 0; p < 8; ++p) for (int s = 0; s < d; ++s) out_slot_edge[p * d + s] = colP[p * d + s] - p * d;
}

// obs_pt non-decreasing (validated by the caller).  lane_window <= 0: steps 2a / 2b are skipped (input order kept).
inline void build_point_layout(int M, int K, const int32_t* obs_cam, const int32_t* obs_pt, int lane_window, int T, PointLayout& L)
{
    L.pstart.assign((size_t)M + 1, 0);
    L.pstart[M] = K;
    host_parallel_for(T, [&](int t, int nt) {
        const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
        for (long long k = k0; k < k1; ++k) {
            const int ja = (k == 0) ? -1 : obs_pt[k - 1], jb = obs_pt[k];
            for (int j = ja + 1; j <= jb; ++j) L.pstart[j] = (int)k;          // points ja+1 .. jb start at k (empty ones included)
        }
        if (t + 1 == nt) for (int j = (K ? obs_pt[K - 1] : -1) + 1; j < M; ++j) L.pstart[j] = K;
    });
    const std::vector<int>& pstart = L.pstart;
    int maxdeg = 0;
    {
        std::vector<int> mx((size_t)T, 0);
        host_parallel_for(T, [&](int t, int nt) {
            const int j0 = (int)((long long)M * t / nt), j1 = (int)((long long)M * (t + 1) / nt);
            int m = 0; for (int j = j0; j < j1; ++j) m = std::max(m, pstart[j + 1] - pstart[j]);
            mx[t] = m;
        });
        for (int t = 0; t < T; ++t) maxdeg = std::max(maxdeg, mx[t]);
    }
    L.maxdeg = maxdeg;
    L.orig_of_new.assign((size_t)M, 0);
    L.deg_new.assign((size_t)std::max(M, 1), 0);
    {   // stable counting sort by track length, descending (per-thread histograms over contiguous point ranges)
        const size_t nb = (size_t)maxdeg + 1;
        std::vector<int> hist((size_t)T * nb, 0);
        host_parallel_for(T, [&](int t, int nt) {
            const int j0 = (int)((long long)M * t / nt), j1 = (int)((long long)M * (t + 1) / nt);
            int* hh = hist.data() + (size_t)t * nb;
            for (int j = j0; j < j1; ++j) hh[maxdeg - (pstart[j + 1] - pstart[j])]++;
        });
        int run = 0;
        for (size_t bk = 0; bk < nb; ++bk) for (int t = 0; t < T; ++t) { int c = hist[(size_t)t * nb + bk]; hist[(size_t)t * nb + bk] = run; run += c; }
        host_parallel_for(T, [&](int t, int nt) {
            const int j0 = (int)((long long)M * t / nt), j1 = (int)((long long)M * (t + 1) / nt);
            int* hh = hist.data() + (size_t)t * nb;
            for (int j = j0; j < j1; ++j) { const int dj = pstart[j + 1] - pstart[j]; const int pos = hh[maxdeg - dj]++; L.orig_of_new[pos] = j; L.deg_new[pos] = dj; }
        });
    }
    L.slot_obs.resize((size_t)std::max(K, 1));
    host_parallel_for(T, [&](int t, int nt) {
        const long long k0 = (long long)K * t / nt, k1 = (long long)K * (t + 1) / nt;
        for (long long k = k0; k < k1; ++k) L.slot_obs[k] = (int)k;
    });
    if (lane_window <= 0 || K <= 0 || M < 16) return;

    // ---- 2a + 2b, per run of equal track length; runs are cut at thread boundaries that are multiples of 32 points ----
    const int W = std::min(std::max(lane_window, 8), 256);
    std::vector<int>& ord = L.orig_of_new;
    const std::vector<int>& deg = L.deg_new;
    // residue histogram of a point, 8 byte-counters packed in a word (built in the caller's point order: one streaming pass)
    std::vector<uint64_t> pkey((size_t)M);
    host_parallel_for(T, [&](int t, int nt) {
        const int j0 = (int)((long long)M * t / nt), j1 = (int)((long long)M * (t + 1) / nt);
        for (int j = j0; j < j1; ++j) {
            uint64_t kq = 0; const int32_t* oc = obs_cam + pstart[j]; const int ds = std::min(pstart[j + 1] - pstart[j], 255);
            for (int sl = 0; sl < ds; ++sl) kq += 1ull << (8 * (oc[sl] & 7));
            pkey[j] = kq;
        }
    });
    // thread boundaries (multiples of 32 points) at equal WORK, not equal point counts: the points are sorted by track length and the
    // grouping / colouring of a quarter-warp costs about its 8 d observations (+ a constant per point); with equal point counts the
    // thread that owned the long tracks took 35 ms where the others took 20 - 27 (Venice shape, 8 threads)
    std::vector<long long> cut((size_t)T + 1, M);
    {
        cut[0] = 0;
        const int pw = getenv("B200BA_LAYOUT_PW") ? atoi(getenv("B200BA_LAYOUT_PW")) : 3;       // cost of a point in observations
        long long total = 0; for (int j = 0; j < M; ++j) total += deg[(size_t)j] + pw;
        long long acc = 0; int t = 1;
        for (long long g = 0; g + 32 <= M && t < T; g += 32) {
            for (int q = 0; q < 32; ++q) acc += deg[(size_t)(g + q)] + pw;
            while (t < T && acc * T >= total * t) cut[t++] = g + 32;
        }
    }
    host_parallel_for(T, [&](int t, int nt) {
        const long long a = nt == T ? cut[t] : (long long)(M / 32) * t / nt * 32, b = (t + 1 == nt) ? M : (nt == T ? cut[t + 1] : (long long)(M / 32) * (t + 1) / nt * 32);
        struct Tm { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(); int t; long long a, b; ~Tm() { if (getenv("B200BA_LAYOUT_TIMES")) fprintf(stderr, "[layout] thread %d points %lld..%lld: %.2f ms\n", t, a, b, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); } } tm_; tm_.t = t; tm_.a = a; tm_.b = b;
        std::vector<int> run, out, pool, cams, slot_edge, head, nxt;
        std::vector<uint64_t> key, hkey;
        std::vector<unsigned char> taken;
        const uint64_t Hb = 0x8080808080808080ull, ONES = 0x0101010101010101ull;
        for (long long r0 = a; r0 < b;) {
            const int d = deg[(size_t)r0];
            long long r1 = r0; while (r1 < b && deg[(size_t)r1] == d) ++r1;
            const int n = (int)(r1 - r0);
            if (n >= 16 && d > 0 && d <= COLOUR_DMAX) {
                run.assign(ord.begin() + r0, ord.begin() + r1);
                key.resize(n);
                for (int q = 0; q < n; ++q) key[q] = pkey[run[q]];
                const bool swar = d <= 15;                         // byte counters of a group cannot overflow (8 d < 128)
                // exact-complement index: open-addressing table  residue histogram -> chain of the run's points that have it
                // (the LAST point of a quarter-warp is looked up, not searched: the histogram that brings every residue to
                // exactly d is unique, and a long run holds every histogram many times)
                int hbits = 4; while ((1 << hbits) < 4 * std::min(n, 8192)) ++hbits;
                const int hmask = (1 << hbits) - 1;
                auto hslot = [&](uint64_t k) { return (int)((k * 0x9E3779B97F4A7C15ull) >> (64 - hbits)); };
                if (swar) {
                    head.assign((size_t)hmask + 1, -1); hkey.assign((size_t)hmask + 1, ~0ull); nxt.assign((size_t)n, -1);
                    for (int q = n - 1; q >= 0; --q) {             // reverse: chains come out in run order
                        int hs = hslot(key[q]); int probes = 0;
                        while (hkey[hs] != ~0ull && hkey[hs] != key[q] && probes < 64) { hs = (hs + 1) & hmask; ++probes; }
                        if (hkey[hs] == ~0ull) hkey[hs] = key[q];
                        if (hkey[hs] == key[q]) { nxt[q] = head[hs]; head[hs] = q; }
                    }
                }
                taken.assign((size_t)n, 0);
                out.clear(); out.reserve(n); pool.clear();
                int next = 0;
                auto refill = [&]() { while ((int)pool.size() < W && next < n) { if (!taken[next]) pool.push_back(next); ++next; } };
                refill();
                int group = (8 - (int)(r0 & 7)) & 7;              // first (partial) group completes the quarter-warp the run starts in
                if (group == 0) group = 8;
                while ((int)out.size() < n) {
                    uint64_t used = 0;                             // residue counts of the group so far (byte counters)
                    for (int g8 = 0; g8 < group && (int)out.size() < n; ++g8) {
                        int q = -1;
                        if (swar && group == 8 && g8 == 7) {      // exact complement by lookup
                            const uint64_t x = ((uint64_t)d * ONES | Hb) - used;          // per byte: 128 + d - used
                            if ((x & Hb) == Hb) {                  // no residue above d so far
                                const uint64_t need = x & ~Hb;
                                int hs = hslot(need); int probes = 0;
                                while (hkey[hs] != ~0ull && hkey[hs] != need && probes < 64) { hs = (hs + 1) & hmask; ++probes; }
                                if (hkey[hs] == need) {
                                    int c = head[hs];
                                    while (c >= 0 && taken[c]) c = nxt[c];
                                    head[hs] = c >= 0 ? nxt[c] : -1;
                                    q = c;
                                }
                            }
                        }
                        if (q < 0) {
                            // greedy: the candidate of the window that overshoots the proportional fill tau = ceil(d (g8 + 1) / 8)
                            // least (tau = d at the last pick: exactly the colourability condition)
                            int best = -1, best_score = 1 << 30;
                            const int tau = (d * (g8 + 1 + (8 - group)) + 7) / 8;
                            const int scan = (g8 + (8 - group)) >= 5 ? W : std::min(W, 16);       // early picks have slack: short scan
                            for (int c = 0; c < std::min((int)pool.size(), scan);) {
                                if (taken[pool[c]]) { pool[c] = pool.back(); pool.pop_back(); continue; }
                                int score = 0;
                                if (g8 > 0) {
                                    const uint64_t kq = key[pool[c]];
                                    if (swar) {
                                        const uint64_t x = ((used + kq) | Hb) - (uint64_t)tau * ONES;
                                        const uint64_t over = (x & Hb) >> 7;                     // 1 in every byte with count >= tau
                                        score = (int)(((x & ~Hb & (over * 255ull)) * ONES) >> 56);   // sum of (count - tau) there
                                    } else {
                                        for (int r = 0; r < 8; ++r) score += std::max(0, (int)((used >> (8 * r)) & 255) + (int)((kq >> (8 * r)) & 255) - tau);
                                    }
                                }
                                if (score < best_score) { best_score = score; best = c; if (score == 0) break; }
                                ++c;
                            }
                            if (best < 0) { refill(); if (pool.empty()) break; continue; }      // window held only taken points
                            q = pool[best];
                            pool[best] = pool.back(); pool.pop_back();
                        }
                        taken[q] = 1;
                        used += key[q];
                        out.push_back(run[q]);
                        refill();
                    }
                    group = 8;
                }
                std::copy(out.begin(), out.end(), ord.begin() + r0);
                // 2b: slot assignment of every complete quarter-warp that lies inside this run
                cams.resize((size_t)8 * d); slot_edge.resize((size_t)8 * d);
                for (long long q0 = (r0 + 7) & ~7ll; q0 + 8 <= r1; q0 += 8) {
                    for (int p = 0; p < 8; ++p) { const int32_t* oc = obs_cam + pstart[ord[(size_t)(q0 + p)]]; for (int e = 0; e < d; ++e) cams[(size_t)p * d + e] = oc[e]; }
                    colour_quarter(cams.data(), d, slot_edge.data());
                    for (int p = 0; p < 8; ++p) { const int ps = pstart[ord[(size_t)(q0 + p)]]; for (int s = 0; s < d; ++s) L.slot_obs[(size_t)ps + s] = ps + slot_edge[(size_t)p * d + s]; }
                }
            }
            r0 = r1;
        }
    });
}

// Expected shared-memory wavefronts of ONE 16-byte table field read by a warp over the whole layout, relative to the
// conflict-free minimum (4 per row): sum over rows and quarter-warps of the largest number of DISTINCT cameras that share a
// residue class mod 8 (equal cameras are one broadcast address); padding lanes read camera 0.
inline double layout_conflict_factor(int M, const int32_t* obs_cam, const PointLayout& L)
{
    const int n_slices = (std::max(M, 1) + 31) / 32;
    long long waves = 0, phases = 0;
    for (int g = 0; g < n_slices; ++g) {
        const int dg = 32 * g < M ? L.deg_new[(size_t)32 * g] : 0;
        for (int s = 0; s < dg; ++s)
            for (int qw = 0; qw < 4; ++qw) {
                int cam[8], ncam = 0;
                for (int l = 0; l < 8; ++l) {
                    const int jn = 32 * g + 8 * qw + l;
                    int c = 0;
                    if (jn < M && s < L.deg_new[(size_t)jn]) { const int j = L.orig_of_new[(size_t)jn]; c = obs_cam[L.slot_obs[(size_t)L.pstart[j] + s]]; }
                    bool dup = false; for (int i = 0; i < ncam; ++i) if (cam[i] == c) dup = true;
                    if (!dup) cam[ncam++] = c;
                }
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
                for (int i = 0; i < ncam; ++i) mx = std::max(mx, ++cnt[cam[i] & 7]);
                waves += mx; phases++;
            }
    }
    return phases ? (double)waves / (double)phases : 1.0;
}

}  // namespace b200ba
