// dense_host.h -- host side of the EXPLICIT reduced camera system for small problems (tracker-sized maps: tens of cameras).
//
// The matrix-free PCG of the large shapes costs three kernel launches per step; at 49 cameras one step is ~21 us of launch and
// L2 latency for ~1 us of arithmetic, and a 50-step LM iteration takes 1.2 ms (profiles/: step-slope measurements).  When the
// reduced camera system S = (U + damping) - W (V + damping)^-1 W^t (6N x 6N) fits the shared memory of ONE thread-block
// cluster, it is assembled explicitly once per LM iteration and the whole PCG loop runs inside one cluster kernel
// (ba_dense.cuh).  The reference assembles and factors the full sparse system every iteration
// (v3d_optimization.cpp:822-866); this is the same idea restricted to the Schur complement, still solved by the engine's PCG.
//
// What the host prepares here (once per b200ba_set_problem, nothing of it depends on parameter values):
//   * the PAIR LIST: for every point, every ordered pair (a, b) of its observations with camera(a) <= camera(b)
//     (a == b included; both orders when a point is seen twice by one camera) contributes  W_a (V + damping)^-1 W_b^t
//     to block (camera(a), camera(b)) of the upper block triangle.  Pairs are sorted by destination block, inside a
//     block by point order - so every block is a FIXED-ORDER sum: no atomics, bit-reproducible;
//   * CHUNKS of at most DENSE_CHUNK pairs of one block: one partial sum each (36 threads per chunk), summed per block in
//     chunk order by the finalising kernel.
// Pure C++ (no CUDA) so that the CPU tests can check the lists without a device.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace b200ba {

constexpr int DENSE_CHUNK = 32;            // pairs per partial sum
constexpr long long DENSE_MAX_PAIRS = 24ll << 20;   // above this the matrix-free path is kept (memory: 8 B per pair)

struct DensePair { int32_t a, b; };        // SELL entries of the two observations (row operand W V^-1, column operand W)
struct DenseChunk { int32_t block, beg, end, diag; };
struct DenseLists {
    std::vector<DensePair> pairs;
    std::vector<DenseChunk> chunks;
    std::vector<int32_t> blk_chunk_beg;    // n_blocks + 1
    std::vector<int32_t> blk_ik;           // 2 x n_blocks: (i, k), i <= k
    int n_blocks = 0;
    long long flops = 0;
};

// upper-triangle block index of (i, k), i <= k, row-major
inline long long dense_block_index(int N, int i, int k) { return (long long)i * N - (long long)i * (i - 1) / 2 + (k - i); }

// Upper bound of the pair count (exact when no point is seen twice by one camera would be sum d (d + 1) / 2): sum over points of d^2.
inline long long dense_pair_bound(int K, const int32_t* obs_pt)
{
    long long bound = 0;
    for (int k0 = 0; k0 < K;) {
        int k1 = k0 + 1; while (k1 < K && obs_pt[k1] == obs_pt[k0]) ++k1;
        const long long dd = k1 - k0; bound += dd * dd;
        k0 = k1;
    }
    return bound;
}

// Returns false (lists untouched) when the pair count exceeds DENSE_MAX_PAIRS.
// obs_pt must be non-decreasing (the precondition of b200ba_set_problem); entry_of_obs maps an observation to its SELL entry.
inline bool build_dense_lists(int N, int M, int K, const int32_t* obs_cam, const int32_t* obs_pt, const int* entry_of_obs, DenseLists& out)
{
    (void)M;
    const long long nb = (long long)N * (N + 1) / 2;
    if (dense_pair_bound(K, obs_pt) > 2 * DENSE_MAX_PAIRS) return false;
    std::vector<int32_t> count((size_t)nb + 1, 0);
    long long total = 0;
    for (int k0 = 0; k0 < K;) {
        int k1 = k0 + 1; while (k1 < K && obs_pt[k1] == obs_pt[k0]) ++k1;
        for (int a = k0; a < k1; ++a)
            for (int b = k0; b < k1; ++b)
                if (obs_cam[a] < obs_cam[b] || (obs_cam[a] == obs_cam[b])) { count[(size_t)dense_block_index(N, obs_cam[a], obs_cam[b]) + 1]++; ++total; }
        k0 = k1;
    }
    if (total > DENSE_MAX_PAIRS) return false;
    for (long long b = 0; b < nb; ++b) count[(size_t)b + 1] += count[(size_t)b];
    out.pairs.assign((size_t)total, DensePair{0, 0});
    {
        std::vector<int32_t> pos(count.begin(), count.end() - 1);
        for (int k0 = 0; k0 < K;) {
            int k1 = k0 + 1; while (k1 < K && obs_pt[k1] == obs_pt[k0]) ++k1;
            for (int a = k0; a < k1; ++a)
                for (int b = k0; b < k1; ++b)
                    if (obs_cam[a] < obs_cam[b] || (obs_cam[a] == obs_cam[b]))
                        out.pairs[(size_t)pos[(size_t)dense_block_index(N, obs_cam[a], obs_cam[b])]++] = DensePair{entry_of_obs[a], entry_of_obs[b]};
            k0 = k1;
        }
    }
    out.n_blocks = (int)nb;
    out.blk_chunk_beg.assign((size_t)nb + 1, 0);
    out.blk_ik.assign(2 * (size_t)nb, 0);
    out.chunks.clear();
    long long blk = 0;
    for (int i = 0; i < N; ++i)
        for (int k = i; k < N; ++k, ++blk) {
            out.blk_ik[2 * (size_t)blk] = i; out.blk_ik[2 * (size_t)blk + 1] = k;
            out.blk_chunk_beg[(size_t)blk] = (int32_t)out.chunks.size();
            for (int32_t q = count[(size_t)blk]; q < count[(size_t)blk + 1]; q += DENSE_CHUNK)
                out.chunks.push_back(DenseChunk{(int32_t)blk, q, std::min(q + DENSE_CHUNK, count[(size_t)blk + 1]), i == k ? 1 : 0});
        }
    out.blk_chunk_beg[(size_t)nb] = (int32_t)out.chunks.size();
    out.flops = total * 36 * 6;
    return true;
}

// Shapes of the PCG kernel (ba_dense.cuh).  A CTA owns `rows` rows of S (whole cameras, so that the block-Jacobi preconditioner
// is CTA-local); a warp owns 4 of them (and, when cq > 1, every cq-th group of 32 column pairs).
//   kind 0  one 8-CTA thread-block cluster, S in REGISTERS (4 rows x 10 columns per lane): up to 53 cameras
//   kind 1  one 8- or 16-CTA cluster, S in shared memory: up to ~105 cameras
//   kind 2  the whole cooperative grid (one CTA per SM), S in registers (4 rows x 14 columns per lane, 4 column splits): up to 2 cameras per SM
struct DenseShape { int kind = -1, ctas = 0, rows = 0, ld = 0, threads = 0; size_t smem = 0; };
constexpr int DENSE_KINDS = 3;
constexpr int DENSE_THREADS[DENSE_KINDS] = {384, 512, 384};
constexpr int DENSE_CQ[DENSE_KINDS]      = {1, 1, 4};      // column splits (warps sharing a block of 4 rows)
constexpr int DENSE_NP[DENSE_KINDS]      = {5, 0, 7};      // column pairs per lane kept in registers (0: S in shared memory)
inline DenseShape dense_shape(int N, int kind, int ctas)
{
    DenseShape s; s.kind = kind; s.ctas = ctas; s.threads = DENSE_THREADS[kind];
    const int n = 6 * N, cq = DENSE_CQ[kind], np = DENSE_NP[kind];
    s.rows = 6 * ((N + ctas - 1) / ctas);
    if (kind == 2) {                                              // two cameras per CTA (12 rows = 3 row blocks x 4 column splits = its 12 warps) and
        if (N > 2 * ctas) { s.kind = -1; return s; }              // no CTA without rows: every CTA of the grid is a participant of the two barriers per step
        s.rows = 12; s.ctas = (N + 1) / 2;
    }
    const int rows4 = (s.rows + 3) & ~3;
    // padded vector length: every column pair a lane touches exists (zeros beyond n), so the product runs without bounds checks
    const int plen = np ? 64 * cq * np : ((n + 63) / 64) * 64;
    s.ld = plen + 8;
    const bool fits_threads = rows4 / 4 * cq <= s.threads / 32 && s.rows <= 64;
    const bool fits_regs = !np || n <= plen;
    if (!fits_threads || !fits_regs) { s.kind = -1; return s; }
    // [S rows (kind 1)] | p | z | q (4 x rows4) | r (rows4) | pq slots | rz slots | scratch + mbarriers
    const int slots = kind == 2 ? 256 : 32;
    s.smem = sizeof(double) * ((kind == 1 ? (size_t)rows4 * s.ld : 0) + 2 * (size_t)s.ld + 5 * (size_t)rows4 + 2 * (size_t)slots + 64);
    return s;
}

} // namespace b200ba
