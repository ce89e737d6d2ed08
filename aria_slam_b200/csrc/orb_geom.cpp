// Host-side geometry: level sizes, per-level quotas, exact-bilinear coefficient tables.
#include "orb_geom.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace aslam {

static inline int cv_round(float v) { return (int)lrintf(v); }   // cvRound: half to even

// cv::resize INTER_LINEAR_EXACT, 8-bit: source coordinate in double, weight in 8.8 fixed point
// rounded half to even (OpenCV imgproc/resize.cpp, interpolationLinear::getCoeffs).
// Normalised so that the weight of p[ofs+1] is c1 in [0,255] and that of p[ofs] is 256-c1:
// a weight of 256 on the right tap becomes (ofs+1, 0); clamped edges are one sample at weight 256.
static void resize_coeff(int src, int dst, int i, int &ofs, int &c1)
{
    const double scale = 1.0 / ((double)dst / (double)src);
    const double f = scale * ((double)i + 0.5) - 0.5;
    const int i0 = (int)std::floor(f);
    if (i0 < 0) { ofs = 0; c1 = 0; }
    else if (i0 >= src - 1) { ofs = src - 1; c1 = 0; }
    else {
        ofs = i0;
        c1 = (int)lrint((f - (double)i0) * 256.0);
        if (c1 == 256) { ofs = i0 + 1; c1 = 0; }
    }
}

// Row table: entry y = ofs | wy1 << 16 with taps (ofs, ofs+1) both inside the image.
static void resize_row_table(int src, int dst, uint32_t *tab)
{
    for (int i = 0; i < dst; ++i) {
        int ofs, c1;
        resize_coeff(src, dst, i, ofs, c1);
        if (ofs >= src - 1) { ofs = src - 2; c1 = 256; }      // keep ofs+1 a real row; same value
        tab[i] = (uint32_t)ofs | ((uint32_t)c1 << 16);
    }
}

// Column table: 8 words per group of 4 output columns (see resize_kernel):
//   [0] aligned source word of the window, [1] funnel shift in bits, [2] PRMT selectors of outputs 0,1,
//   [3] selectors of outputs 2,3, [4..7] weight words (255-c1) | c1 << 8 | 1 << 16.
// Outputs past the level width repeat the last real column (their bytes land in row padding).
// Returns false when a group's window does not fit the 8 bytes the kernel's PRMT selectors can address (taps up
// to 7 bytes past the first one): an effective scale above 2 between two levels.
static bool resize_col_table(int src, int dst, uint32_t *tab)
{
    for (int g = 0; g * 4 < dst; ++g) {
        int ofs[4], c1[4];
        for (int i = 0; i < 4; ++i) resize_coeff(src, dst, std::min(g * 4 + i, dst - 1), ofs[i], c1[i]);
        const int o0 = ofs[0];
        uint32_t *t = tab + 8 * g;
        t[0] = (uint32_t)(o0 >> 2);
        t[1] = (uint32_t)(o0 & 3) * 8u;
        uint32_t sel[4];
        for (int i = 0; i < 4; ++i) {
            const uint32_t k = (uint32_t)(ofs[i] - o0);           // 0..4 at scale 1.2, 0..6 at scale 2
            if (k + 1 > 7) return false;
            sel[i] = k | ((k + 1) << 4) | (k << 8) | (k << 12);
            t[4 + i] = (uint32_t)(255 - c1[i]) | ((uint32_t)c1[i] << 8) | (1u << 16);
        }
        t[2] = sel[0] | (sel[1] << 16);
        t[3] = sel[2] | (sel[3] << 16);
    }
    return true;
}

bool build_orb_geometry(int width, int height, int nfeatures, float scale_factor, int nlevels,
                        int edge, int fast_thr, int blur_mode, OrbGeom &g, std::vector<uint32_t> &tables)
{
    g = OrbGeom();
    g.nlevels = nlevels; g.width = width; g.height = height; g.edge = edge; g.fast_thr = fast_thr;

    // per-level quota: geometric series evaluated in float (computeKeyPoints)
    const float factor = (float)(1.0 / (double)scale_factor);
    float nd = (float)nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        g.lv[l].quota = cv_round(nd);
        sum += g.lv[l].quota;
        nd *= factor;
    }
    g.lv[nlevels - 1].quota = nfeatures - sum > 0 ? nfeatures - sum : 0;

    unsigned int off = 0;
    int tile_off = 0, sel_off = 0, blur_off = 0, desc_off = 0;
    g.nbuilt = 0;
    tables.clear();
    for (int l = 0; l < nlevels; ++l) {
        LevelGeom &L = g.lv[l];
        L.sf = (float)std::pow((double)scale_factor, (double)l);
        L.w = cv_round((float)width / L.sf);
        L.h = cv_round((float)height / L.sf);
        L.pitch = (L.w + 15) / 16 * 16;
        // FAST score blocks start at the even column (edge-1)&~1; their interior columns are outputs
        const int rw = L.w - edge - 1 - ((edge - 1) & ~1), rh = L.h - 2 * edge;
        const bool active = L.w - 2 * edge > 0 && rh > 0 && L.quota > 0 && L.w >= 8 && L.h >= 8;
        L.tiles_x = active ? (rw + FT_OUT_W - 1) / FT_OUT_W : 0;
        L.tiles_y = active ? (rh + FT_OUT_H - 1) / FT_OUT_H : 0;
        L.tile_off = tile_off;
        L.blur_tile_off = blur_off;
        if (active) blur_off += ((L.w + BL_W - 1) / BL_W) * ((L.h + BL_H - 1) / BL_H);
        tile_off += L.tiles_x * L.tiles_y;
        L.sel_off = sel_off;
        L.sel_cap = L.quota + SEL_TIE_ROOM;
        sel_off += L.sel_cap;
        L.desc_block_off = desc_off;
        if (active) desc_off += (L.sel_cap + DESC_WARPS * DESC_KPW - 1) / (DESC_WARPS * DESC_KPW);
        L.inv_sf = 1.f / L.sf;
        L.kp_size = 31.f * L.sf;
        L.row_fused_end = blur_mode == 0 ? 0 : blur_mode == 1 ? L.w : (L.w & ~31);
        L.col_fused_end = blur_mode == 0 ? 0 : blur_mode == 1 ? L.w : (L.w & ~3);
        if (active) g.nbuilt = l + 1;
    }
    // memory only for the levels that must exist (a level needs all levels below it)
    for (int l = 0; l < g.nbuilt; ++l) {
        LevelGeom &L = g.lv[l];
        L.off = off;
        off += (unsigned int)((size_t)L.pitch * L.h + 255) / 256 * 256;
        if (l > 0) {
            const LevelGeom &P = g.lv[l - 1];
            L.xtab_off = (int)tables.size();
            tables.resize(tables.size() + (size_t)(L.w + 3) / 4 * 8, 0);
            if (!resize_col_table(P.w, L.w, tables.data() + L.xtab_off)) return false;
            L.ytab_off = (int)tables.size();
            tables.resize(tables.size() + (size_t)(L.h + 3) / 4 * 4, 0);
            resize_row_table(P.h, L.h, tables.data() + L.ytab_off);
            // strip table of resize_march_kernel: the march walks the source rows s_first, s_first + 1, ... of a strip and
            // bit i of the mask says "source row s_first + 1 + i is the lower tap of the strip's next output row"
            static const int env_small = [] { const char *e = getenv("ASLAM_RS_ROWS_SMALL"); return e ? atoi(e) : 0; }();
            static const int env_big = [] { const char *e = getenv("ASLAM_RS_ROWS_BIG"); return e ? atoi(e) : 0; }();
            L.rs_rows = (long long)L.w * L.h >= 96 * 1024 ? (env_big > 0 ? env_big : 16) : (env_small > 0 ? env_small : 16);
            const int nstrips = (L.h + L.rs_rows - 1) / L.rs_rows;
            L.stab_off = (int)tables.size();
            tables.resize(tables.size() + (size_t)(2 * nstrips + 3) / 4 * 4, 0);
            L.rs_fast = 1;
            for (int st = 0; st < nstrips; ++st) {
                const uint32_t *yt = tables.data() + L.ytab_off;
                const int ya = st * L.rs_rows, yb = std::min(ya + L.rs_rows, L.h);
                const int s_first = (int)(yt[ya] & 0xffffu);
                uint32_t mask = 0;
                for (int y = ya; y < yb; ++y) {
                    const int d = (int)(yt[y] & 0xffffu) - s_first;
                    if (d < 0 || d >= 32 || (mask >> d & 1u)) { L.rs_fast = 0; break; }
                    mask |= 1u << d;
                }
                tables[(size_t)L.stab_off + 2 * st] = mask;
                tables[(size_t)L.stab_off + 2 * st + 1] = (uint32_t)s_first;
            }
        }
    }
    g.pyr_stride = off ? off : 256;
    g.tiles_per_frame = tile_off;
    g.blur_tiles_per_frame = blur_off;
    g.desc_blocks_per_frame = desc_off;
    g.kp_cap = (sel_off + 63) / 64 * 64;
    return true;
}

}  // namespace aslam
