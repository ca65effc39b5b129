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
// ---- pass 1: row runs.  Every foreground pixel points at the start of its horizontal run inside its
// 32-pixel segment (ballot + clz: no memory traffic for in-run merging).
template <class Img>
__global__ void __launch_bounds__(TISEG_THREADS) k_ccl_init(Geom g, Img img, int* __restrict__ par) {
    Pix px;
    if (!warp_pixel(g, px)) return;
    int v = 0;
    bool fg = px.ok && img(px.n, px.base + px.idx, v);
    int vl = __shfl_up_sync(0xffffffffu, v, 1);
    bool fgl = __shfl_up_sync(0xffffffffu, (int)fg, 1);
    if (px.lane == 0) fgl = false;                       // cross-segment continuation is merged in pass 2
    bool cont = fg && fgl && vl == v;
    unsigned m = __ballot_sync(0xffffffffu, cont);
    if (px.ok) par[px.base + px.idx] = fg ? px.idx - (px.lane - run_start_lane(m, px.lane)) : -1;
}

// ---- pass 2: merge runs with the row above (and with the previous segment of the same row).
// Redundant unions are skipped: a pixel whose left neighbour continues its run and whose upper-left
// neighbour has the same value is already connected through them.
template <class Img, int CONN>
__global__ void __launch_bounds__(TISEG_THREADS) k_ccl_merge(Geom g, Img img, int* par) {
    Pix px;
    if (!warp_pixel(g, px)) return;
    int v = 0;
    bool fg = px.ok && img(px.n, px.base + px.idx, v);
    // left neighbour (lane 0 reads the last pixel of the previous segment)
    int vl = __shfl_up_sync(0xffffffffu, v, 1);
    bool fgl = __shfl_up_sync(0xffffffffu, (int)fg, 1);
    if (px.lane == 0) {
        fgl = false;
        if (px.x > 0) fgl = img(px.n, px.base + px.idx - 1, vl);
    }
    bool sameL = fg && fgl && vl == v;
    // row above
    int vu = 0; bool fgu = false;
    if (px.y > 0 && px.ok) fgu = img(px.n, px.base + px.idx - g.W, vu);
    int vul = __shfl_up_sync(0xffffffffu, vu, 1);
    bool fgul = __shfl_up_sync(0xffffffffu, (int)fgu, 1);
    int vur = __shfl_down_sync(0xffffffffu, vu, 1);
    bool fgur = __shfl_down_sync(0xffffffffu, (int)fgu, 1);
    if (px.lane == 0) {
        fgul = false;
        if (px.y > 0 && px.x > 0) fgul = img(px.n, px.base + px.idx - g.W - 1, vul);
    }
    if (px.lane == 31) {
        fgur = false;
        if (CONN == 2 && px.y > 0 && px.x + 1 < g.W) fgur = img(px.n, px.base + px.idx - g.W + 1, vur);
    }
    if (px.x + 1 >= g.W) fgur = false;
    if (!fg) return;
    int* tp = par + px.base;
    bool sU = fgu && vu == v, sUL = fgul && vul == v, sUR = fgur && vur == v;
    if (px.lane == 0 && sameL) uf_union(tp, px.idx, px.idx - 1);
    if (sU) {
        if (!(sameL && sUL)) uf_union(tp, px.idx, px.idx - g.W);
    } else if (CONN == 2) {
        if (sUL && !sameL) uf_union(tp, px.idx, px.idx - g.W - 1);
        if (sUR) uf_union(tp, px.idx, px.idx - g.W + 1);
    }
}

// ---- pass 3: flatten (every foreground pixel points directly at its root); defined in ccl.cu
int ccl_flatten(tiseg_ctx* c, const Geom& g, int* par);

// build + flatten.  par: [N*P] int
template <class Img>
int ccl_build(tiseg_ctx* c, const Geom& g, Img img, int conn, int* par) {
    TISEG_LAUNCH(c, k_ccl_init<Img>, warp_grid(g), TISEG_THREADS, 0, g, img, par);
    if (conn == 1) TISEG_LAUNCH(c, (k_ccl_merge<Img, 1>), warp_grid(g), TISEG_THREADS, 0, g, img, par);
    else           TISEG_LAUNCH(c, (k_ccl_merge<Img, 2>), warp_grid(g), TISEG_THREADS, 0, g, img, par);
    return ccl_flatten(c, g, par);
}
#endif

// ---- raster-order ranks ------------------------------------------------------------------------
// Selection is a functor  __device__ bool operator()(long long gi, int idx) const  (e.g. "is a root").
// rank[gi] = 1-based raster rank among the selected pixels of its tile (written only where selected);
// counts[n] = number selected (may be null).  Three launches: per-block counts, per-tile scan of the
// block counts, ballot/popc placement.
struct SelRoot {                // roots of a flattened forest
    const int* par;
    __device__ __forceinline__ bool operator()(long long gi, int idx) const { return par[gi] == idx; }
};
struct SelFlagU8 {
    const uint8_t* f;
    __device__ __forceinline__ bool operator()(long long gi, int) const { return f[gi] != 0; }
};

#ifdef __CUDACC__
template <class Sel>
__global__ void __launch_bounds__(TISEG_THREADS) k_rank_count(Geom g, Sel sel, int* __restrict__ blk) {
    __shared__ int s[TISEG_WARPS_PER_BLOCK];
    Pix px;
    bool act = warp_pixel(g, px);
    bool f = act && px.ok && sel(px.base + px.idx, px.idx);
    unsigned m = __ballot_sync(0xffffffffu, f);
    if (px.lane == 0) s[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < TISEG_WARPS_PER_BLOCK; ++i) t += s[i];
        blk[(long long)blockIdx.y * g.bpt + blockIdx.x] = t;
    }
}

template <class Sel>
__global__ void __launch_bounds__(TISEG_THREADS) k_rank_place(Geom g, Sel sel, const int* __restrict__ blk,
                                                              int* __restrict__ rank) {
    __shared__ int s[TISEG_WARPS_PER_BLOCK];
    Pix px;
    bool act = warp_pixel(g, px);
    bool f = act && px.ok && sel(px.base + px.idx, px.idx);
    unsigned m = __ballot_sync(0xffffffffu, f);
    int w = threadIdx.x >> 5;
    if (px.lane == 0) s[w] = __popc(m);
    __syncthreads();
    if (!f) return;
    int off = blk[(long long)blockIdx.y * g.bpt + blockIdx.x];
    for (int i = 0; i < w; ++i) off += s[i];
    rank[px.base + px.idx] = off + __popc(m & ((1u << px.lane) - 1)) + 1;
}

// in-place exclusive scan of blk[n, 0..bpt) per tile; counts[n] = total (defined in ccl.cu)
int rank_scan(tiseg_ctx* c, int N, int bpt, int* blk, int* counts);

template <class Sel>
int rank_generic(tiseg_ctx* c, const Geom& g, Sel sel, int* rank, int* counts) {
    int* blk = ws<int>(c, (size_t)g.N * g.bpt);
    if (!blk) return TISEG_ERR_CUDA;
    TISEG_LAUNCH(c, k_rank_count<Sel>, warp_grid(g), TISEG_THREADS, 0, g, sel, blk);
    TISEG_TRY(rank_scan(c, g.N, g.bpt, blk, counts));
    TISEG_LAUNCH(c, k_rank_place<Sel>, warp_grid(g), TISEG_THREADS, 0, g, sel, blk, rank);
    return TISEG_OK;
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
