// orb_plan.cu — host-side geometry: scale tables, per-level quotas, cell grid, resize
// coefficient tables and tile lists for one image shape.  Mirrors what the reference computes
// in ORBextractor::ORBextractor (src/ORBextractor.cc:457-511), ComputePyramid (:781-822) and
// the head of ComputeKeyPoints (:527-599); cv::resize's INTER_LINEAR 8-bit coefficient recipe
// is restated from OpenCV 4.x ("K1").
#include "orb_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

static inline int cvRoundF(float v) { return (int)lrintf(v); }
static inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }

int orb_build_tables(orb_ctx* c)
{
    const int nlevels = c->nlevels;
    if (nlevels < 1 || nlevels > ORB_MAX_LEVELS || c->nfeatures < 1) return ORB_ERR_INVALID;
    c->scaleFactor = (double)c->scale_factor_f;                    // member is a double holding the float
    c->mvScaleFactor.assign(nlevels, 1.f);
    c->mvInvScaleFactor.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; i++) c->mvScaleFactor[i] = (float)(c->mvScaleFactor[i - 1] * c->scaleFactor);
    const float invScale = (float)(1.0f / c->scaleFactor);
    for (int i = 1; i < nlevels; i++) c->mvInvScaleFactor[i] = c->mvInvScaleFactor[i - 1] * invScale;

    c->mnFeaturesPerLevel.assign(nlevels, 0);
    const float factor = (float)(1.0 / c->scaleFactor);
    float nDesired = c->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        c->mnFeaturesPerLevel[l] = cvRoundF(nDesired);
        sum += c->mnFeaturesPerLevel[l];
        nDesired *= factor;
    }
    c->mnFeaturesPerLevel[nlevels - 1] = std::max(c->nfeatures - sum, 0);

    // circular patch row extents for the intensity centroid (:495-510)
    int umax[17] = { 0 };
    const int HP = 15;
    int v, v0, vmax = cvFloorD(HP * sqrtf(2.f) / 2 + 1), vmin = cvCeilD(HP * sqrtf(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)HP * HP - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    memcpy(c->umax, umax, sizeof(int) * 16);
    return ORB_OK;
}

// INTER_LINEAR 8U coefficients for one axis: entry = { s0 | s1<<16 , c0 | c1<<16 } (c as int16)
static void axis_table(int ssize, int dsize, bool clamp_frac, std::vector<int2>& out)
{
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cvFloorD(f);
        f -= s;
        int s0, s1;
        if (clamp_frac) {                       // x axis: the fraction is zeroed at the borders
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
            s0 = s; s1 = std::min(s + 1, ssize - 1);
        } else {                                // y axis: row indices are clipped, weights kept
            s0 = std::min(std::max(s, 0), ssize - 1);
            s1 = std::min(std::max(s + 1, 0), ssize - 1);
        }
        short c0 = (short)cvRoundF((1.f - f) * 2048), c1 = (short)cvRoundF(f * 2048);
        int2 e;
        e.x = (s0 & 0xffff) | (s1 << 16);
        e.y = ((int)(unsigned short)c0) | ((int)(unsigned short)c1 << 16);
        out.push_back(e);
    }
}

int orb_build_plan(orb_ctx* c, int w, int h)
{
    if (c->plan_valid && c->plan.w == w && c->plan.h == h) return ORB_OK;
    c->plan_valid = false;
    if (w < 1 || h < 1 || w > 32767 || h > 32767) return ORB_ERR_INVALID;
    if (w > c->max_w || h > c->max_h) return ORB_ERR_CAPACITY;
    Plan& P = c->plan;
    memset(&P, 0, sizeof(P));
    P.nlevels = c->nlevels; P.w = w; P.h = h;
    P.fast_th = c->fast_th; P.th_lo = std::min(c->fast_th, 7); P.harris = c->score_type == ORB_HARRIS_SCORE; P.desc_fma = c->desc_fma;
    c->cells.clear(); c->tiles_fast.clear(); c->tiles_blur.clear(); c->xtab.clear(); c->ytab.clear(); c->fast_coltab.clear(); c->fast_rowtab.clear();

    int off = 0, cand = 0, lvl = 0, kp = 0, border = 0, bm = 0;
    const float imageRatio = (float)w / h;                                   // :527
    for (int l = 0; l < P.nlevels; l++) {
        LevelGeom& L = P.L[l];
        const float scale = c->mvInvScaleFactor[l];
        L.w = cvRoundF((float)w * scale); L.h = cvRoundF((float)h * scale);   // :786
        if (L.w < 1 || L.h < 1) return ORB_ERR_GEOMETRY;
        L.stride = (L.w + 2 * ORB_EDGE + 15) & ~15;
        L.prows = L.h + 2 * ORB_EDGE;
        L.plane_off = off;
        off += (L.stride * L.prows + 255) & ~255;
        L.scale = c->mvScaleFactor[l];
        L.patch_size = (int)(31 * c->mvScaleFactor[l]);                       // :675
        L.xtab_off = (int)c->xtab.size(); L.ytab_off = (int)c->ytab.size();
        if (l > 0) {
            axis_table(P.L[l - 1].w, L.w, true, c->xtab);
            axis_table(P.L[l - 1].h, L.h, false, c->ytab);
            // largest source footprint of one output tile, from a 16-byte aligned origin (k_resize TMA box, <= 256 per side).
            // A thread owns 4 columns x rr rows, a CTA tw columns x (1024 / tw) * rr rows.  Among the tile widths 64..128 the one
            // that leaves the fewest idle threads on this level's size wins (a fixed 128 wastes 18 % of the lanes on a 522-wide
            // level); 64x64 tiles with 4 rows per thread remain the fallback for scale factors whose footprint is too large.
            auto try_tile = [&](int v, int tw, int rr) {
                const int th = (4 * ORB_RESIZE_THREADS / tw) * rr;
                int mw = 16, mr = 1;
                for (int x0 = 0; x0 < L.w; x0 += tw) {
                    const int x1 = std::min(x0 + tw, L.w) - 1;
                    const int lo = (c->xtab[L.xtab_off + x0].x & 0xffff) & ~15, hi = c->xtab[L.xtab_off + x1].x >> 16;
                    mw = std::max(mw, hi - lo + 1);
                }
                for (int y0 = 0; y0 < L.h; y0 += th) {
                    const int y1 = std::min(y0 + th, L.h) - 1;
                    mr = std::max(mr, (c->ytab[L.ytab_off + y1].x >> 16) - (c->ytab[L.ytab_off + y0].x & 0xffff) + 1);
                }
                c->rs_box_w[v][l] = (mw + 15) & ~15; c->rs_box_h[v][l] = mr; c->rs_tile_w[v][l] = tw; c->rs_rows[v][l] = rr;
                return c->rs_box_w[v][l] <= 256 && mr <= 256;
            };
            for (int v = 0; v < 2; v++) {
                const int rows_pref = v == 0 ? c->rs_rows_pref : std::min(c->rs_rows_small, c->rs_rows_pref);
                int best_tw = 128;
                if (c->rs_flex_width) {
                    double best = -1;
                    for (int tw = 64; tw <= 128; tw += 4) {
                        const int th = (4 * ORB_RESIZE_THREADS / tw) * rows_pref;
                        const double ctas = (double)((L.w + tw - 1) / tw) * ((L.h + th - 1) / th);
                        const double eff = (double)L.w * L.h / (ctas * ORB_RESIZE_THREADS * 4 * rows_pref);
                        if (eff >= best) { best = eff; best_tw = tw; }
                    }
                }
                bool fits = try_tile(v, best_tw, rows_pref);
                if (!fits && best_tw != 128) fits = try_tile(v, 128, rows_pref);
                if (!fits) fits = try_tile(v, 64, std::min(4, rows_pref));
                if (!fits) return ORB_ERR_CAPACITY;      // scale factors above ~3.7
            }
            // k_resize_u: per 4-column group the first source byte, the PRMT selectors of the four columns' tap pairs (byte offsets from
            // that first byte, low nibble = tap 0, high nibble = tap 1; two selectors per word) and the four weight pairs
            while (c->xtab.size() & 3) c->xtab.push_back(make_int2(0, 0));        // 32-byte aligned groups
            c->rs_xg_off[l] = (int)c->xtab.size();
            bool packed = true;
            for (int gx = 0; gx < L.w; gx += 4) {
                int2 e[4];
                for (int k = 0; k < 4; k++) e[k] = c->xtab[L.xtab_off + std::min(gx + k, L.w - 1)];
                const int cb = e[0].x & 0xffff;
                unsigned sel[4];
                for (int k = 0; k < 4; k++) {
                    const int d0 = (e[k].x & 0xffff) - cb, d1 = (e[k].x >> 16) - cb;
                    packed = packed && d0 >= 0 && d1 >= 0 && d0 <= 7 && d1 <= 7;
                    sel[k] = (unsigned)(d0 & 7) | ((unsigned)(d1 & 7) << 4);
                }
                c->xtab.push_back(make_int2(cb, (int)(sel[0] | sel[1] << 16)));
                c->xtab.push_back(make_int2((int)(sel[2] | sel[3] << 16), 0));
                c->xtab.push_back(make_int2(e[0].y, e[1].y));
                c->xtab.push_back(make_int2(e[2].y, e[3].y));
            }
            c->rs_packed[l] = packed;
        }
        L.border_base = border;
        // k_border work items (32-bit words): ORB_RING full rows above and below the ROI, and per ROI row the word left of it plus
        // the two words that cover [w, w + ORB_RING) on the right
        L.ring = ORB_RING; L.border_items = 0;       // set below, once the cell grid is known
        // cell grid (:531-547)
        L.nDesired = c->mnFeaturesPerLevel[l];
        L.cols = (int)sqrtf((float)L.nDesired / (5 * imageRatio));
        L.rows = (int)(imageRatio * L.cols);
        if (L.cols < 1 || L.rows < 1) return ORB_ERR_GEOMETRY;               // reference divides by zero
        if (L.cols > ORB_MAX_GRID || L.rows > ORB_MAX_GRID || L.cols * L.rows > ORB_MAX_CELLS_LEVEL) return ORB_ERR_CAPACITY;
        const int minB = ORB_EDGE, maxBX = L.w - ORB_EDGE, maxBY = L.h - ORB_EDGE;
        const int W = maxBX - minB, H = maxBY - minB;
        L.cellW = (int)ceilf((float)W / L.cols);
        L.cellH = (int)ceilf((float)H / L.rows);
        L.ncells = L.cols * L.rows;
        L.nfCell = (int)ceilf((float)L.nDesired / L.ncells);
        L.cell_base = (int)c->cells.size();
        L.xend = minB; L.yend = minB;
        for (int i = 0; i < L.rows; i++) {
            const int iniY = minB + i * L.cellH - 3;
            int hY = L.cellH + 6;
            if (i == L.rows - 1) hY = maxBY + 3 - iniY;
            for (int j = 0; j < L.cols; j++) {
                const int iniX = minB + j * L.cellW - 3;
                int hX = L.cellW + 6;
                if (j == L.cols - 1) hX = maxBX + 3 - iniX;
                CellGeom g;
                g.level = l; g.idx = i * L.cols + j;
                g.inix = iniX; g.iniy = iniY; g.pad = 0;
                // only the LAST row / column is tested for an empty remainder and skipped (:570,:594); an inner cell of a level smaller
                // than its margins (cellW or cellH <= -6) reaches Mat::rowRange / colRange with end < start, which throws like a
                // cell that leaves the ROI does
                const bool skip = (i == L.rows - 1 && hY <= 0) || (j == L.cols - 1 && hX <= 0);
                g.skipped = skip ? 1 : 0;
                if (skip) {                                // never produces keypoints and stays out of the quota loop's first pass
                    g.x0 = g.x1 = g.y0 = g.y1 = 0;
                } else {
                    if (hX < 0 || hY < 0 || iniX < 0 || iniY < 0 || iniX + hX > L.w || iniY + hY > L.h) return ORB_ERR_GEOMETRY;
                    g.x0 = iniX + 3; g.x1 = iniX + hX - 3; g.y0 = iniY + 3; g.y1 = iniY + hY - 3;
                    if (g.x1 < g.x0) g.x1 = g.x0;          // cell image narrower than 7 px: FAST finds nothing
                    if (g.y1 < g.y0) g.y1 = g.y0;
                    if (g.x1 - g.x0 + 6 > 4095 || g.y1 - g.y0 + 6 > 4095) return ORB_ERR_CAPACITY;  // 12-bit cell-local coords
                }
                g.cand_off = cand;
                g.cand_cap = ((g.x1 - g.x0 + 1) / 2) * ((g.y1 - g.y0 + 1) / 2);   // strict 8-neighbour NMS bound
                cand += g.cand_cap;
                if (g.x1 > g.x0 && g.y1 > g.y0) { L.xend = std::max(L.xend, g.x1); L.yend = std::max(L.yend, g.y1); }
                c->cells.push_back(g);
            }
        }
        // Degenerate grids (many features on a small level): cellW = ceil(W / cols), so the inner cells may reach up to cols - 1
        // pixels past size - 16 — the reference detects there (:560-572,:591-596), and a keypoint closer than 16 px to the ROI edge
        // makes IC_Angle and the descriptor pattern read up to 15 px of the reflect-101 frame instead of 3: such a level gets the
        // reference's whole 16 px frame.
        if (L.xend > maxBX || L.yend > maxBY) L.ring = ORB_EDGE;
        L.border_items = 2 * L.ring * (L.stride / 4) + L.h * (L.ring / 2 + 1);   // ring rows above / below; per ROI row ring/4 words left, ring/4 + 1 right
        border += L.border_items;
        L.bm_pitch = ((std::max(L.xend - ORB_EDGE, 0) + ORB_TILE_W - 1) / ORB_TILE_W) * (ORB_TILE_W / 8);   // whole FAST tiles
        L.bm_off = bm;
        bm += ((L.bm_pitch * std::max(((L.yend - ORB_EDGE + ORB_TILE_H - 1) / ORB_TILE_H) * ORB_TILE_H, 0) + 255) & ~255);
        L.lvl_base = lvl;
        // Longest list the quota loop (src/ORBextractor.cc:622-670) can leave before the level's retainBest: ncells * nfCell + ncells.
        // Proof: when an iteration starts with m open cells and deficit D, (retained by closed cells) + D <= (N - m) * (nfCell + 1)
        // by induction (a cell that closes at quota nNew = nfCell + ceil(D / m) contributes its keys plus its own deficit = nNew, and
        // the ceilings add at most one per closed cell); the open cells then receive m * nNew <= m * nfCell + D + m.  The quotas of
        // cells that stay open are REPLACED, not accumulated, by the next iteration.  (Round 1 used ncells^2 / 2 for the ceilings,
        // which refused e.g. 1092 features on one 298x495 level: 209 cells -> 23 K records > the selection kernel's shared memory.)
        // tests/test_kernel_arith_models.py::test_quota_loop_bound checks the bound on adversarial key counts; the kernels still
        // verify it at run time (status ORB_ERR_CAPACITY instead of an overrun).
        L.lvl_cap = L.ncells * L.nfCell + L.ncells + 64;
        if (L.lvl_cap < L.nDesired + 64) L.lvl_cap = L.nDesired + 64;
        lvl += L.lvl_cap;
        L.kp_base = kp;
        kp += L.nDesired;
        {
            // k_fast_nms tables: which detection cell a column / row belongs to (-1: outside [16, w-16) resp. [16, h-16); the last cell
            // takes the remainder, :560-572,:591-596), as byte masks per column (in region, left / right neighbour in the same cell) and
            // as a cell row per row.  A tile reads 72 columns from x0 - 4 and 130 rows from y0 - 1, hence the margins.
            auto cell_of = [](int v, int size, int cell, int n) {
                if (v < ORB_EDGE || cell <= 0) return -1;                                 // cell <= 0: level no larger than its 16-px margins, no tiles
                int cc = (v - ORB_EDGE) / cell;
                if (cc >= n - 1) { cc = n - 1; if (v >= size - ORB_EDGE) cc = -1; }     // degenerate grids: an inner cell may reach past size - 16
                return cc;
            };
            const int nx = (L.w + ORB_TILE_W + 16 + 3) & ~3, ny = L.h + ORB_TILE_H + 8;
            L.ct_off = (int)c->fast_coltab.size(); L.ct_len = nx;
            c->fast_coltab.resize(c->fast_coltab.size() + 3 * (size_t)nx, 0);
            uint8_t* ct = c->fast_coltab.data() + L.ct_off;
            for (int i = 0; i < nx; i++) {
                const int x = i - 4, cc = cell_of(x, L.w, L.cellW, L.cols);
                ct[i] = cc >= 0 ? 0xff : 0;
                ct[nx + i] = (cc >= 0 && cell_of(x - 1, L.w, L.cellW, L.cols) == cc) ? 0xff : 0;
                ct[2 * nx + i] = (cc >= 0 && cell_of(x + 1, L.w, L.cellW, L.cols) == cc) ? 0xff : 0;
            }
            L.rt_off = (int)c->fast_rowtab.size();
            for (int i = 0; i < ny; i++) c->fast_rowtab.push_back((int16_t)cell_of(i - 1, L.h, L.cellH, L.rows));
        }
        for (int y = minB; y < L.yend; y += ORB_TILE_H)
            for (int x = minB; x < L.xend; x += ORB_TILE_W) c->tiles_fast.push_back(Tile{ l, x, y, 0 });
        for (int y = 0; y < L.h; y += ORB_BLUR_TILE_H)
            for (int x = 0; x < L.w; x += ORB_BLUR_TILE_W) c->tiles_blur.push_back(Tile{ l, x, y, 0 });
    }
    P.frame_bytes = off;
    P.ncells = (int)c->cells.size();
    P.cand_total = cand;
    P.lvl_total = lvl;
    P.kp_cap = kp;
    P.border_total = border;
    P.bm_total = bm;
    P.sel_list_cap = 0;
    P.sel_cells_cap = 4;
    for (int l = 0; l < P.nlevels; l++) { P.sel_list_cap = std::max(P.sel_list_cap, P.L[l].lvl_cap); P.sel_cells_cap = std::max(P.sel_cells_cap, (P.L[l].ncells + 3) & ~3); }
    P.ntiles_fast = (int)c->tiles_fast.size();
    P.ntiles_blur = (int)c->tiles_blur.size();
    return ORB_OK;   // device upload happens in orb_api.cu
}
