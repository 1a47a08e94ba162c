// pm_bfs.cu — all-pairs maze distances for L boards at once (sm_100a).
//
// Replaces distanceCalculator.computeDistances (distanceCalculator.py:111-150):
// the reference runs one uniform-cost search per open cell through a Python
// priority queue (9-335 ms per board, rebuilt per bot per reset).  With unit
// edge costs UCS == BFS, and a BFS level over a bit-packed board is
//     next = (F<<1 | F>>1 | F<<W | F>>W) & open & ~visited
// (boards have closed borders, so +-1 never wraps a row onto an open cell).
// A board has at most 1024 cells = 32 words, so ONE WARP owns one source cell
// and holds the whole frontier / visited / open bitsets in registers, one word
// per lane: the four shifts are warp shuffles plus funnel shifts, "frontier
// empty?" is one vote, and the lanes scatter the level number to the cells that
// were just reached.  (Round 1 gave every source a thread with its bitsets in
// shared memory: 13 words x ~60 levels walked serially, 165 us for one
// bloxCapture board; a warp per source cuts the dependent chain per level to a
// dozen instructions.)  A CTA of 16 warps numbers the board's open cells once
// (node order = walls.asList(False) = x-major, y-minor, game.py:225-230) and
// stages each source's result row in shared memory so the table is written
// with coalesced 128-bit stores.
//
// Output: dist[l][i][j] = maze distance between nodes i and j, 0xFFFF =
// unreachable (sys.maxsize in the reference) or padding.
#include <cuda_runtime.h>

#include <mutex>

#include "../../include/pacman_b200.h"
#include "pm_format.h"

#define BFS_WARPS 16                      // sources (warps) per CTA
#define BFS_THREADS (32 * BFS_WARPS)

__global__ void __launch_bounds__(BFS_THREADS) pm_bfs_kernel(int W, int H, const uint32_t *__restrict__ layouts, int LS,
                                                             int32_t *__restrict__ n_open_out,
                                                             int16_t *__restrict__ node_of_cell_out,
                                                             uint16_t *__restrict__ dist, int nmax) {
    extern __shared__ __align__(16) uint8_t smraw[];
    const int HW = W * H, NW = pm_nw(HW);
    const int l = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t *walls = layouts + (size_t)l * LS + pm_l_walls(HW);

    // smem carve
    uint16_t *stage = reinterpret_cast<uint16_t *>(smraw);                       // [BFS_WARPS][nmax]
    uint32_t *openm = reinterpret_cast<uint32_t *>(smraw + (size_t)BFS_WARPS * nmax * 2);   // [32]
    int16_t *node_of = reinterpret_cast<int16_t *>(openm + 32);                   // [HW]
    uint16_t *cell_of = reinterpret_cast<uint16_t *>(node_of + ((HW + 1) & ~1)); // [HW]
    int *colbase = reinterpret_cast<int *>(cell_of + ((HW + 1) & ~1));            // [W + 1]

    // ---- node numbering: x-major, y-minor over open cells ----
    if (tid < 32) {
        uint32_t o = 0;
        if (tid < NW) {
            const uint32_t valid = (32 * tid + 32 <= HW) ? 0xffffffffu : ((1u << (HW - 32 * tid)) - 1u);
            o = ~__ldg(walls + tid) & valid;
        }
        openm[tid] = o;
    }
    __syncthreads();
    for (int x = tid; x < W; x += BFS_THREADS) {
        int cnt = 0;
        for (int y = 0; y < H; ++y) { const int c = y * W + x; cnt += (openm[c >> 5] >> (c & 31)) & 1; }
        colbase[x + 1] = cnt;
    }
    __syncthreads();
    if (tid == 0) { colbase[0] = 0; for (int x = 0; x < W; ++x) colbase[x + 1] += colbase[x]; }
    __syncthreads();
    const int n_open = colbase[W];
    for (int x = tid; x < W; x += BFS_THREADS) {
        int k = colbase[x];
        for (int y = 0; y < H; ++y) {
            const int c = y * W + x;
            if ((openm[c >> 5] >> (c & 31)) & 1) { node_of[c] = (int16_t)k; cell_of[k] = (uint16_t)c; ++k; }
            else node_of[c] = -1;
        }
    }
    __syncthreads();
    if (tile == 0) {
        if (tid == 0 && n_open_out) n_open_out[l] = n_open;
        if (node_of_cell_out) for (int c = tid; c < HW; c += BFS_THREADS) node_of_cell_out[(size_t)l * HW + c] = node_of[c];
    }

    // ---- one BFS per warp: lane i holds word i of the open / visited / frontier bitsets ----
    const int src = tile * BFS_WARPS + warp;
    if (src >= nmax) return;                              // (warp-uniform)
    uint16_t *row = stage + (size_t)warp * nmax;
    for (int i = lane; i < nmax / 2; i += 32) reinterpret_cast<uint32_t *>(row)[i] = 0xffffffffu;      // nmax % 8 == 0
    __syncwarp();
    if (src < n_open) {
        const uint32_t open_w = openm[lane];              // zero for lanes >= NW
        const int c0 = cell_of[src];
        uint32_t fr = (lane == (c0 >> 5)) ? 1u << (c0 & 31) : 0u, vis = fr;
        if (lane == 0) row[src] = 0;
        const int ws = W >> 5, bs = W & 31;
        for (int d = 1; d < 0xFFFF; ++d) {
            // +-1: neighbours in the same row; +-W: the rows above / below (W = 32 * ws + bs bits away)
            const uint32_t prev = __shfl_up_sync(0xffffffffu, fr, 1), next = __shfl_down_sync(0xffffffffu, fr, 1);
            uint32_t n = __funnelshift_l(lane > 0 ? prev : 0u, fr, 1) | __funnelshift_r(fr, lane < 31 ? next : 0u, 1);
            const int a = lane - ws, b = lane + ws;
            const uint32_t ua = __shfl_sync(0xffffffffu, fr, a & 31), ua1 = __shfl_sync(0xffffffffu, fr, (a - 1) & 31);
            const uint32_t db = __shfl_sync(0xffffffffu, fr, b & 31), db1 = __shfl_sync(0xffffffffu, fr, (b + 1) & 31);
            n |= __funnelshift_l(a - 1 >= 0 ? ua1 : 0u, a >= 0 ? ua : 0u, bs) | __funnelshift_r(b < 32 ? db : 0u, b + 1 < 32 ? db1 : 0u, bs);
            n &= open_w & ~vis;
            if (!__any_sync(0xffffffffu, n != 0u)) break;
            vis |= n;
            fr = n;
            while (n) {
                const int bit = __ffs(n) - 1;
                n &= n - 1;
                const int k = node_of[32 * lane + bit];
                if (k < nmax) row[k] = (uint16_t)d;
            }
        }
    }
    __syncwarp();

    // ---- coalesced write-out of this warp's row ----
    uint4 *out = reinterpret_cast<uint4 *>(dist + ((size_t)l * nmax + src) * nmax);
    const uint4 *in = reinterpret_cast<const uint4 *>(row);
    for (int i = lane; i < nmax / 8; i += 32) out[i] = in[i];
}

extern "C" int pm_distances(int W, int H, const uint32_t *layouts, int n_layouts, int32_t *n_open,
                            int16_t *node_of_cell, uint16_t *dist, int nmax, pm_stream_t stream) {
    if (!layouts || !dist || n_layouts <= 0 || nmax <= 0 || (nmax & 7)) return PM_EINVAL;
    if (W < 3 || H < 3) return PM_EINVAL;
    if (W > 255 || H > 255 || W * H > PM_MAX_CELLS) return PM_ETOOBIG;
    const int HW = W * H;
    const size_t smem = (size_t)BFS_WARPS * nmax * 2 + 32 * 4 + (size_t)((HW + 1) & ~1) * 4 + (size_t)(W + 1) * 4 + 16;
    if (smem > 200 * 1024) return PM_ETOOBIG;
    static size_t configured[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PM_ENOGPU;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[dev & 63]) {
            cudaError_t e = cudaFuncSetAttribute(pm_bfs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            configured[dev & 63] = smem;
        }
    }
    dim3 grid(n_layouts, (nmax + BFS_WARPS - 1) / BFS_WARPS);
    pm_bfs_kernel<<<grid, BFS_THREADS, smem, (cudaStream_t)stream>>>(W, H, layouts, pm_layout_words_(HW), n_open, node_of_cell,
                                                                    dist, nmax);
    return (int)cudaGetLastError();
}
