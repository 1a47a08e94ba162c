// pm_pack.cpp — host-side packing of boards into the records the kernels read.
// Pure CPU code (no CUDA), part of libpacman_b200.so.
//
//   pm_pack_geom   : half-board plane masks (capture.py:325-350) and the visit
//                    order of the food-dump search (capture.py:635-661)
//   pm_pack_layout : `.lay` text -> bitsets + starts + teams (layout.py:95-130,
//                    capture.py:311-323, game.py:490-508)
#include <algorithm>
#include <cstring>
#include <deque>
#include <vector>

#include "../../include/pacman_b200.h"
#include "pm_format.h"

extern "C" int pm_abi_version(void) { return PM_ABI_VERSION; }

static int check_dims(int W, int H) {
    if (W < 3 || H < 3) return PM_EINVAL;
    if (W > 255 || H > 255 || W * H > PM_MAX_CELLS) return PM_ETOOBIG;
    return 0;
}

extern "C" int pm_geom_words(int W, int H) { return check_dims(W, H) ? check_dims(W, H) : pm_geom_words_(W, H); }
extern "C" int pm_layout_words(int W, int H) { return check_dims(W, H) ? check_dims(W, H) : pm_layout_words_(W * H); }
extern "C" int pm_obs_bits_words(int W, int H) { return check_dims(W, H) ? check_dims(W, H) : pm_obs_bits_words_(W * H); }
extern "C" size_t pm_state_bytes(int W, int H, int64_t n_envs) {
    if (check_dims(W, H) || n_envs <= 0) return 0;
    return pm_state_bytes_(n_envs, W * H);
}

// The reference's dump search pops a FIFO queue, skips cells already seen, and
// appends the 3x3 neighbourhood (dx outer, dy inner, centre included) of EVERY
// newly popped cell, whether or not it is on the board, a wall, or accepted
// (capture.py:643-661).  Expansion therefore never looks at the map, and the
// order in which cells are first popped is one fixed sequence of offsets from
// the death cell.  We run that search once on an unbounded grid, up to the
// Chebyshev radius that covers any board of this size.
static void dump_visit_order(int R, std::vector<uint32_t> &out) {
    const int D = 2 * R + 3;                       // one ring of slack for the last expansion
    std::vector<uint8_t> seen((size_t)D * D, 0);
    std::deque<std::pair<int, int>> q;
    q.emplace_back(0, 0);
    const size_t want = (size_t)(2 * R + 1) * (2 * R + 1);
    while (out.size() < want && !q.empty()) {
        auto [x, y] = q.front();
        q.pop_front();
        if (std::max(std::abs(x), std::abs(y)) > R + 1) continue;   // cannot reach back inside first
        uint8_t &s = seen[(size_t)(x + R + 1) * D + (y + R + 1)];
        if (s) continue;
        s = 1;
        if (std::max(std::abs(x), std::abs(y)) <= R) out.push_back((uint32_t)(x & 0xff) | (uint32_t)(y & 0xff) << 8);
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) q.emplace_back(x + dx, y + dy);
    }
}

extern "C" int pm_pack_geom(int W, int H, uint32_t *out) {
    if (!out) return PM_EINVAL;
    if (int rc = check_dims(W, H)) return rc;
    const int HW = W * H, NCAT = pm_ncat(HW);
    std::memset(out, 0, sizeof(uint32_t) * pm_geom_words_(W, H));
    out[PM_G_W] = W; out[PM_G_H] = H; out[PM_G_HW] = HW; out[PM_G_NW] = pm_nw(HW);
    out[PM_G_NCAT] = NCAT; out[PM_G_ND] = pm_nd(W, H); out[PM_G_LS] = pm_layout_words_(HW);
    out[PM_G_R] = pm_dump_radius(W, H);
    uint32_t *M2 = out + pm_g_mask_off(HW, 0), *M3 = out + pm_g_mask_off(HW, 1);
    uint32_t *M6 = out + pm_g_mask_off(HW, 2), *M7 = out + pm_g_mask_off(HW, 3);
    const int halfway = W / 2;                                 // int(grid.width / 2), capture.py:333
    for (int c = 0; c < HW; ++c) {
        const int x = c % W;
        auto setbit = [&](uint32_t *m, int plane) { int b = plane * HW + c; m[b >> 5] |= 1u << (b & 31); };
        if (2 * x > W) setbit(M2, 2); else setbit(M3, 3);     // halfList: x <= W/2 is red, capture.py:344-350
        if (x >= halfway) setbit(M6, 6); else setbit(M7, 7);
        // mirrored view (canonicalize_obs, pacman_mappo_resnet.py:215-229): canonical cell c holds source column
        // xs = W-1-x; canonical plane 2 <- source plane 3, 3 <- 2, 6 <- 7, 7 <- 6
        const int xs = W - 1 - x;
        uint32_t *Mm = out + pm_g_mask_off(HW, 4);
        if (2 * xs <= W) setbit(Mm + 0 * NCAT, 2); else setbit(Mm + 1 * NCAT, 3);
        if (xs < halfway) setbit(Mm + 2 * NCAT, 6); else setbit(Mm + 3 * NCAT, 7);
    }
    std::vector<uint32_t> order;
    dump_visit_order(pm_dump_radius(W, H), order);
    if ((int)order.size() != pm_nd(W, H)) return PM_EINVAL;
    std::memcpy(out + pm_g_dump_off(HW), order.data(), sizeof(uint32_t) * order.size());
    return 0;
}

extern "C" int pm_pack_layout(const char *text, int W, int H, uint32_t *out, int *has_caps) {
    if (!text || !out) return PM_EINVAL;
    if (int rc = check_dims(W, H)) return rc;
    const int HW = W * H, NW = pm_nw(HW);
    if ((int)strnlen(text, (size_t)HW) < HW) return PM_ELAYOUT;
    std::memset(out, 0, sizeof(uint32_t) * pm_layout_words_(HW));
    uint32_t *walls = out + pm_l_walls(HW), *food = out + pm_l_food(HW), *caps = out + pm_l_caps(HW);
    uint32_t *wallcat = out + pm_l_wallcat(HW), *tail = out + pm_l_tail(HW);
    struct Tagged { int tag, x, y; };
    std::vector<Tagged> ag;
    int ncaps = 0, red0 = 0, blue0 = 0, total = 0;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const char ch = text[(H - 1 - y) * W + x];          // text row 0 is the top row
            const int c = y * W + x;
            if (ch == '%') {
                walls[c >> 5] |= 1u << (c & 31); wallcat[c >> 5] |= 1u << (c & 31);
                const int cm = y * W + (W - 1 - x);
                (out + pm_l_wallcatm(HW))[cm >> 5] |= 1u << (cm & 31);
            }
            else if (ch == '.') {
                food[c >> 5] |= 1u << (c & 31);
                ++total;
                if (x < W / 2) ++red0; else ++blue0;
            } else if (ch == 'o') { caps[c >> 5] |= 1u << (c & 31); ++ncaps; }
            else if (ch == 'P') ag.push_back({0, x, y});
            else if (ch == 'G') ag.push_back({1, x, y});
            else if (ch >= '1' && ch <= '4') ag.push_back({ch - '0', x, y});
        }
    }
    if (ag.size() != 4) return PM_ELAYOUT;          // the gym adapter drives exactly four agents
    for (int x = 0; x < W; ++x)
        if (text[x] != '%' || text[(H - 1) * W + x] != '%') return PM_ELAYOUT;
    for (int r = 0; r < H; ++r)
        if (text[r * W] != '%' || text[r * W + W - 1] != '%') return PM_ELAYOUT;
    std::sort(ag.begin(), ag.end(), [](const Tagged &a, const Tagged &b) {
        if (a.tag != b.tag) return a.tag < b.tag;
        if (a.x != b.x) return a.x < b.x;
        return a.y < b.y;
    });
    uint32_t teams = 0;
    for (int i = 0; i < 4; ++i) {
        tail[PM_L_START + i] = (uint32_t)ag[i].x | (uint32_t)ag[i].y << 8 | (uint32_t)(ag[i].tag == 0) << 16;
        if (2 * ag[i].x < W) teams |= 1u << i;                  // isRed: x < width / 2, capture.py:325-330
    }
    tail[PM_L_TEAMS] = teams;
    tail[PM_L_TOTALFOOD] = (uint32_t)total;
    tail[PM_L_FOODCNT] = (uint32_t)red0 | (uint32_t)blue0 << 16;
    tail[PM_L_HASCAPS] = ncaps > 0;
    if (has_caps) *has_caps = ncaps > 0;
    (void)NW;
    return 0;
}

// L boards of one size at once: texts = L x (H*W) characters back to back, out = L x pm_layout_words words.
// *has_caps (optional) = any board has a capsule; *max_open (optional) = the largest number of open (non-wall) cells of
// any board - the row count pm_distances needs.  Stops at the first malformed board: returns its PM_E* code and its
// index in *bad (optional).
extern "C" int pm_pack_layouts(const char *texts, int W, int H, int n_layouts, uint32_t *out, int *has_caps, int *max_open, int *bad) {
    if (!texts || !out || n_layouts <= 0) return PM_EINVAL;
    if (int rc = check_dims(W, H)) return rc;
    const int HW = W * H, LS = pm_layout_words_(HW), NW = pm_nw(HW);
    int any_caps = 0, mo = 0;
    for (int l = 0; l < n_layouts; ++l) {
        int hc = 0;
        uint32_t *o = out + (size_t)l * LS;
        if (int rc = pm_pack_layout(texts + (size_t)l * HW, W, H, o, &hc)) { if (bad) *bad = l; return rc; }
        any_caps |= hc;
        int walls = 0;
        for (int w = 0; w < NW; ++w) walls += __builtin_popcount(o[pm_l_walls(HW) + w]);
        if (HW - walls > mo) mo = HW - walls;
    }
    if (has_caps) *has_caps = any_caps;
    if (max_open) *max_open = mo;
    return 0;
}
