// pm_bfs.cu — all-pairs maze distances for L boards at once (sm_100a).
//
// Replaces distanceCalculator.computeDistances (distanceCalculator.py:111-150):
// the reference runs one uniform-cost search per open cell through a Python
// priority queue (9-335 ms per board, rebuilt per bot per reset).  With unit
// edge costs UCS == BFS, so here every thread owns ONE source cell and expands
// its frontier as a flat c-indexed bitset:
//     next = (F<<1 | F>>1 | F<<W | F>>W) & open & ~visited
// (boards have closed borders, so +-1 never wraps a row onto an open cell).
// Bitsets live in shared memory as [word][thread] (conflict-free); each CTA
// handles 64 source nodes of one board and stages its 64 x nmax uint16 result
// rows in shared memory so the table is written with coalesced 128-bit stores.
//
// Output convention: node order = walls.asList(False) = x-major, y-minor
// (game.py:225-230); dist[l][i][j] = maze distance between nodes i and j,
// 0xFFFF = unreachable (sys.maxsize in the reference) or padding.
#include <cuda_runtime.h>

#include "../../include/pacman_b200.h"
#include "pm_format.h"

#define BFS_SRC 64   // sources (threads) per CTA

__global__ void __launch_bounds__(BFS_SRC) pm_bfs_kernel(int W, int H, const uint32_t *__restrict__ layouts, int LS,
                                                         int32_t *__restrict__ n_open_out,
                                                         int16_t *__restrict__ node_of_cell_out,
                                                         uint16_t *__restrict__ dist, int nmax) {
    extern __shared__ __align__(16) uint8_t smraw[];
    const int HW = W * H, NW = pm_nw(HW);
    const int l = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x;
    const uint32_t *walls = layouts + (size_t)l * LS + pm_l_walls(HW);

    // smem carve
    uint16_t *stage = reinterpret_cast<uint16_t *>(smraw);                       // [BFS_SRC][nmax]
    uint32_t *bits = reinterpret_cast<uint32_t *>(smraw + (size_t)BFS_SRC * nmax * 2);   // [3][NW][BFS_SRC]
    uint32_t *openm = bits + 3 * NW * BFS_SRC;                                   // [NW]
    int16_t *node_of = reinterpret_cast<int16_t *>(openm + NW);                   // [HW]
    uint16_t *cell_of = reinterpret_cast<uint16_t *>(node_of + ((HW + 1) & ~1)); // [HW]
    int *colbase = reinterpret_cast<int *>(cell_of + ((HW + 1) & ~1));            // [W + 1]

    // ---- node numbering: x-major, y-minor over open cells ----
    for (int i = tid; i < NW; i += BFS_SRC) {
        uint32_t valid = (32 * i + 32 <= HW) ? 0xffffffffu : ((1u << (HW - 32 * i)) - 1u);
        openm[i] = ~__ldg(walls + i) & valid;
    }
    __syncthreads();
    for (int x = tid; x < W; x += BFS_SRC) {
        int cnt = 0;
        for (int y = 0; y < H; ++y) { const int c = y * W + x; cnt += (openm[c >> 5] >> (c & 31)) & 1; }
        colbase[x + 1] = cnt;
    }
    __syncthreads();
    if (tid == 0) { colbase[0] = 0; for (int x = 0; x < W; ++x) colbase[x + 1] += colbase[x]; }
    __syncthreads();
    const int n_open = colbase[W];
    for (int x = tid; x < W; x += BFS_SRC) {
        int k = colbase[x];
        for (int y = 0; y < H; ++y) {
            const int c = y * W + x;
            if ((openm[c >> 5] >> (c & 31)) & 1) { node_of[c] = (int16_t)k; cell_of[k] = (uint16_t)c; ++k; }
            else node_of[c] = -1;
        }
    }
    for (int i = tid; i < BFS_SRC * nmax / 2; i += BFS_SRC) reinterpret_cast<uint32_t *>(stage)[i] = 0xffffffffu;
    __syncthreads();
    if (tile == 0) {
        if (tid == 0 && n_open_out) n_open_out[l] = n_open;
        if (node_of_cell_out) for (int c = tid; c < HW; c += BFS_SRC) node_of_cell_out[(size_t)l * HW + c] = node_of[c];
    }

    // ---- one BFS per thread ----
    const int src = tile * BFS_SRC + tid;
    if (src < n_open && src < nmax) {
        uint32_t *vis = bits + tid, *fr = vis + NW * BFS_SRC, *nx = fr + NW * BFS_SRC;   // word i at [i * BFS_SRC]
        uint16_t *row = stage + (size_t)tid * nmax;
        const int c0 = cell_of[src];
        for (int i = 0; i < NW; ++i) { vis[i * BFS_SRC] = 0; fr[i * BFS_SRC] = 0; }
        vis[(c0 >> 5) * BFS_SRC] = fr[(c0 >> 5) * BFS_SRC] = 1u << (c0 & 31);
        row[src] = 0;
        const int ws = W >> 5, bs = W & 31;
        for (int d = 1; d < 0xFFFF; ++d) {
            uint32_t any = 0;
            for (int i = 0; i < NW; ++i) {
                const uint32_t cur = fr[i * BFS_SRC];
                const uint32_t lo1 = i > 0 ? fr[(i - 1) * BFS_SRC] : 0u, hi1 = i + 1 < NW ? fr[(i + 1) * BFS_SRC] : 0u;
                uint32_t n = __funnelshift_l(lo1, cur, 1) | __funnelshift_r(cur, hi1, 1);
                const int a = i - ws, b = i + ws;
                const uint32_t ua = (a >= 0) ? fr[a * BFS_SRC] : 0u, ua1 = (a - 1 >= 0) ? fr[(a - 1) * BFS_SRC] : 0u;
                const uint32_t db = (b < NW) ? fr[b * BFS_SRC] : 0u, db1 = (b + 1 < NW) ? fr[(b + 1) * BFS_SRC] : 0u;
                n |= __funnelshift_l(ua1, ua, bs) | __funnelshift_r(db, db1, bs);
                n &= openm[i] & ~vis[i * BFS_SRC];
                nx[i * BFS_SRC] = n;
                any |= n;
            }
            if (!any) break;
            for (int i = 0; i < NW; ++i) {
                uint32_t n = nx[i * BFS_SRC];
                vis[i * BFS_SRC] |= n;
                fr[i * BFS_SRC] = n;
                while (n) {
                    const int b = __ffs(n) - 1;
                    n &= n - 1;
                    const int k = node_of[32 * i + b];
                    if (k < nmax) row[k] = (uint16_t)d;
                }
            }
        }
    }
    __syncthreads();

    // ---- coalesced write-out of this CTA's rows ----
    const int rows = min(BFS_SRC, nmax - tile * BFS_SRC);
    if (rows <= 0) return;
    const int vec_per_row = nmax / 8;                     // nmax % 8 == 0 (checked on the host)
    uint4 *out = reinterpret_cast<uint4 *>(dist + ((size_t)l * nmax + (size_t)tile * BFS_SRC) * nmax);
    const uint4 *in = reinterpret_cast<const uint4 *>(stage);
    for (int i = tid; i < rows * vec_per_row; i += BFS_SRC) out[i] = in[i];
}

extern "C" int pm_distances(int W, int H, const uint32_t *layouts, int n_layouts, int32_t *n_open,
                            int16_t *node_of_cell, uint16_t *dist, int nmax, pm_stream_t stream) {
    if (!layouts || !dist || n_layouts <= 0 || nmax <= 0 || (nmax & 7)) return PM_EINVAL;
    if (W < 3 || H < 3) return PM_EINVAL;
    if (W > 255 || H > 255 || W * H > PM_MAX_CELLS) return PM_ETOOBIG;
    const int HW = W * H, NW = pm_nw(HW);
    const size_t smem = (size_t)BFS_SRC * nmax * 2 + (size_t)(3 * NW * BFS_SRC + NW) * 4 + (size_t)((HW + 1) & ~1) * 4 +
                        (size_t)(W + 1) * 4 + 16;
    if (smem > 200 * 1024) return PM_ETOOBIG;
    static size_t configured[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PM_ENOGPU;
    if (smem > configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(pm_bfs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[dev & 63] = smem;
    }
    dim3 grid(n_layouts, (nmax + BFS_SRC - 1) / BFS_SRC);
    pm_bfs_kernel<<<grid, BFS_SRC, smem, (cudaStream_t)stream>>>(W, H, layouts, pm_layout_words_(HW), n_open, node_of_cell,
                                                                dist, nmax);
    return (int)cudaGetLastError();
}
