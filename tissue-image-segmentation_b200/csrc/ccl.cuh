// K3 — connected-component labelling building blocks (union-find, raster-order ids, areas).
//
// Covers skimage.measure.label (unet.py:85, dist.py:107,123, inst_metrics.py:12-13), scipy.ndimage.label
// (hovernet.py:296,358), and the component analysis inside remove_small_objects / binary_fill_holes.
//
// Image access is a functor `Img`:  __device__ bool operator()(int n, long long gi, int& v) const
//   returns whether pixel gi (global index n*P + idx) is foreground; v = the value compared for equality.
//
// par[n*P + idx] = tile-local flat index of the parent (root = lowest index of the component = first pixel
// in raster order, which is what makes skimage's numbering reproducible), -1 for background.
#pragma once
#include <climits>

#include "common.cuh"

namespace tiseg {

// ---- image functors ----------------------------------------------------------------------------
struct ImgEqI32 {               // equal-value components of an int32 image, `bg` is background
    const int32_t* p; int bg;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = p[gi]; return v != bg; }
};
struct ImgEqU8 {                // equal-value components of a uint8 image, bg < 0 => no background
    const uint8_t* p; int bg;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = p[gi]; return v != bg; }
};
struct ImgEqU8Where {           // equal-value components of a uint8 image restricted to the pixels flagged in `on`
    const uint8_t* p; const uint8_t* on;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = p[gi]; return on[gi] != 0; }
};
struct ImgMaskU8 {              // binary: non-zero is foreground
    const uint8_t* p;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = 1; return p[gi] != 0; }
};
struct ImgNotMaskU8 {           // binary complement: zero is foreground (fill-holes background analysis)
    const uint8_t* p;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = 1; return p[gi] == 0; }
};
struct ImgClassU8 {             // binary: pixels of one class id
    const uint8_t* p; int cls;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = 1; return p[gi] == cls; }
};
struct ImgNotClassU8 {          // binary: pixels NOT of one class id
    const uint8_t* p; int cls;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = 1; return p[gi] != cls; }
};
struct ImgEqI32TileBg {         // equal-value components, background value given per tile (arrange_label)
    const int32_t* p; const int* bg;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = p[gi]; return v != bg[n]; }
};
struct ImgAll {                 // every pixel foreground, one value (mask == NULL)
    __device__ __forceinline__ bool operator()(int, long long, int& v) const { v = 1; return true; }
};
struct ImgBelowU8 {             // binary: uint8 value < thr
    const uint8_t* p; int thr;
    __device__ __forceinline__ bool operator()(int, long long gi, int& v) const { v = 1; return p[gi] < thr; }
};
struct ImgOrI32U8 {             // binary: labelled pixel OR foreground pixel (align_foreground's reachable set)
    const int32_t* lab; const uint8_t* fg;
    __device__ __forceinline__ bool operator()(int, long long gi, int& v) const { v = 1; return lab[gi] != 0 || fg[gi] != 0; }
};
struct ImgEqU8Drop {            // equal-value components of a uint8 image, values 0 and `drop` are background
    const uint8_t* p; int drop;
    __device__ __forceinline__ bool operator()(int, long long gi, int& v) const { v = p[gi]; return v != 0 && v != drop; }
};
struct ImgNonZeroI32 {          // binary: non-zero int32
    const int32_t* p;
    __device__ __forceinline__ bool operator()(int n, long long gi, int& v) const { v = 1; return p[gi] != 0; }
};

#ifdef __CUDACC__
// =====================================================================================================
// Three launches per labelling:
//   k_ccl_local   one CTA per 64-row x 32-column tile: values staged in shared memory, union-find entirely in
//                 shared memory (row runs by ballot/clz, merges with the row above by shared atomicMin), flattened
//                 locally, then ONE coalesced global write per pixel: the tile-global index of its local root.
//   k_ccl_border  only the pixels on tile borders (~8 %) merge across tiles with global atomicMin unions.
//   k_ccl_flatten every pixel points at its global root (lowest flat index of the component = first pixel in
//                 raster order); also leaves the bitmap of the roots for the id ranking that usually follows.
// =====================================================================================================
#define CCL_WARPS 8                     // warps per tile CTA
#define CCL_TH (8 * CCL_WARPS)           // tile rows (each warp owns 8 of them); tile width is one warp = 32 columns
#define CCL_BG INT_MIN                  // background marker inside the shared value tile

template <class Img, int CONN, bool LISTED>
__global__ void __launch_bounds__(32 * CCL_WARPS, LISTED ? 1 : 2048 / (32 * CCL_WARPS)) k_ccl_local(Geom g, Img img, int* __restrict__ par) {
    __shared__ int sval[CCL_TH * 32];
    __shared__ int slab[CCL_TH * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ty = blockIdx.x / g.SEG, tx = blockIdx.x - ty * g.SEG;
    const int x = tx * 32 + lane;
    const bool okx = x < g.W;
    // rows are dealt to the warps round-robin (warp w owns rows w, w + CCL_WARPS, ...): a nucleus spreads over all eight warps, so
    // they reach the barriers together instead of one warp doing a blob's unions while seven wait
    const int yt = ty * CCL_TH;
    FOR_TILES(LISTED, g, n) {
    const long long base = (long long)n * g.P;
    // phase A: coalesced loads (8 independent rows in flight per lane), row-run initialisation
    int v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int y = yt + r * CCL_WARPS + warp, vv = 0;
        v[r] = CCL_BG;
        if (okx && y < g.H && img(n, base + (long long)y * g.W + x, vv)) v[r] = vv;
    }
    unsigned fg = 0;                    // bit r: row r of this warp has a foreground pixel (uniform over the warp)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int vl = __shfl_up_sync(0xffffffffu, v[r], 1);
        bool cont = lane > 0 && v[r] != CCL_BG && vl == v[r];
        unsigned m = __ballot_sync(0xffffffffu, cont);
        if (__ballot_sync(0xffffffffu, v[r] != CCL_BG)) fg |= 1u << r;
        int li = (r * CCL_WARPS + warp) * 32 + lane;
        sval[li] = v[r];
        slab[li] = v[r] != CCL_BG ? (r * CCL_WARPS + warp) * 32 + run_start_lane(m, lane) : -1;
    }
    __syncthreads();
    // phase B: merge with the row above (same redundancy rule as documented at k_ccl_border); rows without a
    // foreground pixel are skipped by the whole warp
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int row = r * CCL_WARPS + warp, li = row * 32 + lane;
        if (!((fg >> r) & 1u)) continue;
        if (row == 0 || v[r] == CCL_BG) continue;
        int u = sval[li - 32];
        int ul = lane > 0 ? sval[li - 33] : CCL_BG;
        int ur = lane < 31 ? sval[li - 31] : CCL_BG;
        bool sameL = lane > 0 && sval[li - 1] == v[r];
        if (u == v[r]) {
            if (!(sameL && ul == v[r])) uf_union(slab, li, li - 32);
        } else if (CONN == 2) {
            if (ul == v[r] && !sameL) uf_union(slab, li, li - 33);
            if (ur == v[r]) uf_union(slab, li, li - 31);
        }
    }
    __syncthreads();
    // phase C: local flatten, one global write per pixel
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int y = yt + r * CCL_WARPS + warp;
        if (!okx || y >= g.H) continue;
        int out = -1;
        if (v[r] != CCL_BG) {
            int root = uf_find(slab, (r * CCL_WARPS + warp) * 32 + lane);
            out = (ty * CCL_TH + (root >> 5)) * g.W + tx * 32 + (root & 31);
        }
        par[base + (long long)y * g.W + x] = out;
    }
    if (LISTED) __syncthreads();        // the shared tile is reused by the next listed tile
    }
}

// Cross-tile merges.  Candidates: A) pixels of a tile's top row (y = 64k, k >= 1) look up / up-left / up-right;
// B) pixels of a tile's left column (x = 32k, k >= 1) look left / up-left; C) (8-connectivity) pixels of a tile's
// right column (x = 32k - 1) look up-right.  A diagonal union is skipped when the vertical neighbour has the same
// value (it is then connected through that neighbour's own row).
template <class Img, int CONN>
__device__ __forceinline__ void ccl_border_one(const Geom& g, const Img& img, int* par, int n, int t) {
    const int tilesY = (g.H + CCL_TH - 1) / CCL_TH;
    const int nA = (tilesY - 1) * g.W, nB = (g.SEG - 1) * g.H, nC = CONN == 2 ? nB : 0;
    if (t >= nA + nB + nC) return;
    int y, x, kind;
    if (t < nA) { kind = 0; y = (t / g.W + 1) * CCL_TH; x = t - (t / g.W) * g.W; }
    else if (t < nA + nB) { t -= nA; kind = 1; x = (t / g.H + 1) * 32; y = t - (t / g.H) * g.H; }
    else { t -= nA + nB; kind = 2; x = (t / g.H + 1) * 32 - 1; y = t - (t / g.H) * g.H; }
    const long long base = (long long)n * g.P;
    const int idx = y * g.W + x;
    int v = 0;
    if (!img(n, base + idx, v)) return;
    int* tp = par + base;
    int w = 0;
    bool sU = y > 0 && img(n, base + idx - g.W, w) && w == v;
    if (kind == 0) {
        if (sU) uf_union(tp, idx, idx - g.W);
        else if (CONN == 2) {
            if (x > 0 && img(n, base + idx - g.W - 1, w) && w == v) uf_union(tp, idx, idx - g.W - 1);
            if (x + 1 < g.W && img(n, base + idx - g.W + 1, w) && w == v) uf_union(tp, idx, idx - g.W + 1);
        }
    } else if (kind == 1) {
        if (img(n, base + idx - 1, w) && w == v) uf_union(tp, idx, idx - 1);
        else if (CONN == 2 && !sU && y > 0 && img(n, base + idx - g.W - 1, w) && w == v) uf_union(tp, idx, idx - g.W - 1);
    } else {
        if (!sU && y > 0 && x + 1 < g.W && img(n, base + idx - g.W + 1, w) && w == v) uf_union(tp, idx, idx - g.W + 1);
    }
}
template <class Img, int CONN, bool LISTED>
__global__ void __launch_bounds__(256) k_ccl_border(Geom g, Img img, int* par) {
    FOR_TILES(LISTED, g, n) ccl_border_one<Img, CONN>(g, img, par, n, blockIdx.x * blockDim.x + threadIdx.x);
}

// flatten + bitmap of the roots; defined in ccl.cu
int ccl_flatten(tiseg_ctx* c, const Geom& g, int* par);

// build + flatten.  par: [N*P] int
template <class Img>
int ccl_build(tiseg_ctx* c, const Geom& g, Img img, int conn, int* par) {
    const int tilesY = (g.H + CCL_TH - 1) / CCL_TH;
    dim3 lg((unsigned)(g.SEG * tilesY), grid_tiles(g));
    const int nb = (tilesY - 1) * g.W + (g.SEG - 1) * g.H * (conn == 2 ? 2 : 1);
    if (g.tl) {
        if (conn == 1) {
            TISEG_LAUNCH(c, (k_ccl_local<Img, 1, true>), lg, 32 * CCL_WARPS, 0, g, img, par);
            if (nb > 0) TISEG_LAUNCH(c, (k_ccl_border<Img, 1, true>), dim3((nb + 255) / 256, 1), 256, 0, g, img, par);
        } else {
            TISEG_LAUNCH(c, (k_ccl_local<Img, 2, true>), lg, 32 * CCL_WARPS, 0, g, img, par);
            if (nb > 0) TISEG_LAUNCH(c, (k_ccl_border<Img, 2, true>), dim3((nb + 255) / 256, 1), 256, 0, g, img, par);
        }
    } else if (conn == 1) {
        TISEG_LAUNCH(c, (k_ccl_local<Img, 1, false>), lg, 32 * CCL_WARPS, 0, g, img, par);
        if (nb > 0) TISEG_LAUNCH(c, (k_ccl_border<Img, 1, false>), dim3((nb + 255) / 256, g.N), 256, 0, g, img, par);
    } else {
        TISEG_LAUNCH(c, (k_ccl_local<Img, 2, false>), lg, 32 * CCL_WARPS, 0, g, img, par);
        if (nb > 0) TISEG_LAUNCH(c, (k_ccl_border<Img, 2, false>), dim3((nb + 255) / 256, g.N), 256, 0, g, img, par);
    }
    return ccl_flatten(c, g, par);
}
#endif

// ---- raster-order ranks ------------------------------------------------------------------------
// Selection is a functor  __device__ bool operator()(long long gi, int idx) const  (e.g. "is a root").
// rank[gi] = 1-based raster rank among the selected pixels of its tile (written only where selected);
// counts[n] = number selected (may be null).
// The selection is first reduced to a BITMAP, one 32-bit ballot word per row segment (bits[n, y, seg]); ranking then
// touches P/32 words instead of P pixels: k_rank_rowscan (one CTA per tile: popcount per row, exclusive scan over
// the rows) and k_rank_place_bits (one warp per row: warp scan over the row's words, one store per selected pixel).
struct SelRoot {                // roots of a flattened forest
    const int* par;
    __device__ __forceinline__ bool operator()(long long gi, int idx) const { return par[gi] == idx; }
};
struct SelFlagU8 {
    const uint8_t* f;
    __device__ __forceinline__ bool operator()(long long gi, int) const { return f[gi] != 0; }
};

#ifdef __CUDACC__
// bits[n, y, seg] = ballot of the selected pixels of that 32-pixel row segment
template <class Sel>
__global__ void __launch_bounds__(TISEG_THREADS) k_rank_bits(Geom g, Sel sel, unsigned* __restrict__ bits) {
    Strip s;
    if (!warp_strip(g, s)) return;
    bool f[STRIP_R];
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        int y = s.y0 + r;
        f[r] = s.okx && y < g.H && sel(s.base + (long long)y * g.W + s.x, y * g.W + s.x);
    }
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        unsigned m = __ballot_sync(0xffffffffu, f[r]);
        int y = s.y0 + r;
        if (s.lane == 0 && y < g.H) bits[((long long)s.n * g.H + y) * g.SEG + s.seg] = m;
    }
}

// bitmap -> ranks (defined in ccl.cu)
int rank_from_bits(tiseg_ctx* c, const Geom& g, const unsigned* bits, int* rank, int* counts);

template <class Sel>
int rank_generic(tiseg_ctx* c, const Geom& g, Sel sel, int* rank, int* counts) {
    unsigned* bits = ws<unsigned>(c, (size_t)g.N * g.H * g.SEG);
    if (!bits) return TISEG_ERR_CUDA;
    TISEG_LAUNCH(c, k_rank_bits<Sel>, strip_grid(g), TISEG_THREADS, 0, g, sel, bits);
    return rank_from_bits(c, g, bits, rank, counts);
}
#endif

// roots of a flattened forest: rank[root] = raster rank (1-based); counts[n] = K
int rank_roots(tiseg_ctx* c, const Geom& g, const int* par, int* rank, int* counts);
// out[gi] = par[gi] >= 0 ? rank[par[gi]] : 0
int apply_rank(tiseg_ctx* c, const Geom& g, const int* par, const int* rank, int32_t* out);
// area[root] = component size (this zeroes `area` itself; valid at root positions only)
int ccl_areas(tiseg_ctx* c, const Geom& g, const int* par, int* area);

// full label: img functor -> out ids (1..K raster order), counts
template <class Img>
int ccl_label(tiseg_ctx* c, const Geom& g, Img img, int conn, int32_t* out, int* counts) {
#ifdef __CUDACC__
    size_t total = (size_t)g.N * g.P;
    int* par = ws<int>(c, total);
    int* rank = ws<int>(c, total);
    if (!par || !rank) return TISEG_ERR_CUDA;
    TISEG_TRY(ccl_build(c, g, img, conn, par));
    TISEG_TRY(rank_roots(c, g, par, rank, counts));
    TISEG_TRY(apply_rank(c, g, par, rank, out));
#endif
    return TISEG_OK;
}

}  // namespace tiseg
