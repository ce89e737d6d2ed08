// Host-side geometry: level sizes, per-level quotas, exact-bilinear coefficient tables.
#include "orb_geom.h"

#include <cmath>

namespace aslam {

static inline int cv_round(float v) { return (int)lrintf(v); }   // cvRound: half to even

// cv::resize INTER_LINEAR_EXACT, 8-bit: source coordinate in double, weight in 8.8 fixed point
// rounded half to even (OpenCV imgproc/resize.cpp, interpolationLinear::getCoeffs).
static void resize_table(int src, int dst, uint32_t *tab)
{
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int i = 0; i < dst; ++i) {
        const double f = scale * ((double)i + 0.5) - 0.5;
        const int i0 = (int)std::floor(f);
        uint32_t ofs, c1;
        if (i0 < 0) { ofs = 0; c1 = 0; }                                  // clamp left: p[0] * 256
        else if (i0 >= src - 1) { ofs = (uint32_t)(src - 2); c1 = 256; }  // clamp right: p[src-1] * 256
        else { ofs = (uint32_t)i0; c1 = (uint32_t)lrint((f - (double)i0) * 256.0); }
        tab[i] = ofs | (c1 << 16);
    }
}

void build_orb_geometry(int width, int height, int nfeatures, float scale_factor, int nlevels,
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
    int tile_off = 0, sel_off = 0;
    g.nbuilt = 0;
    tables.clear();
    for (int l = 0; l < nlevels; ++l) {
        LevelGeom &L = g.lv[l];
        L.sf = (float)std::pow((double)scale_factor, (double)l);
        L.w = cv_round((float)width / L.sf);
        L.h = cv_round((float)height / L.sf);
        L.pitch = (L.w + 15) / 16 * 16;
        const int rw = L.w - 2 * edge, rh = L.h - 2 * edge;
        const bool active = rw > 0 && rh > 0 && L.quota > 0 && L.w >= 8 && L.h >= 8;
        L.tiles_x = active ? (rw + FT_OUT_W - 1) / FT_OUT_W : 0;
        L.tiles_y = active ? (rh + FT_OUT_H - 1) / FT_OUT_H : 0;
        L.tile_off = tile_off;
        tile_off += L.tiles_x * L.tiles_y;
        L.sel_off = sel_off;
        L.sel_cap = L.quota + SEL_TIE_ROOM;
        sel_off += L.sel_cap;
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
            tables.resize(tables.size() + (size_t)(L.w + 3) / 4 * 4, 0);
            resize_table(P.w, L.w, tables.data() + L.xtab_off);
            L.ytab_off = (int)tables.size();
            tables.resize(tables.size() + (size_t)(L.h + 3) / 4 * 4, 0);
            resize_table(P.h, L.h, tables.data() + L.ytab_off);
        }
    }
    g.pyr_stride = off ? off : 256;
    g.tiles_per_frame = tile_off;
    g.kp_cap = (sel_off + 63) / 64 * 64;
}

}  // namespace aslam
