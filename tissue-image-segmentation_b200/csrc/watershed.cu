// K6 — ordered marker-controlled watershed: blob bookkeeping, the uint8 bucket flood, the fp64 heap flood
// and the tiseg_watershed_* entry points.  See watershed.cuh for the algorithm statement.
#include "watershed.cuh"

namespace tiseg {

#define FULL 0xffffffffu

__global__ void k_blob_init(BlobInfo b, int W) {
    int n = blockIdx.y;
    long long o = (long long)n * b.KS;
    int k = b.count[n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= k; i += gridDim.x * blockDim.x) {
        b.root[o + i] = 0; b.ymax[o + i] = -1; b.xmin[o + i] = W; b.xmax[o + i] = -1; b.area[o + i] = 0;
    }
}

__global__ void __launch_bounds__(TISEG_THREADS)
k_blob_roots(long long P, const int* __restrict__ par, const int* __restrict__ rank, BlobInfo b, bool vec) {
    const long long base = (long long)blockIdx.y * P, i = flat4_index();
    if (i >= P) return;
    Pack4<int> p = ld4(par + base, i, P, vec);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i + k < P && p.v[k] == (int)(i + k)) b.root[(long long)blockIdx.y * b.KS + rank[base + i + k]] = (int)(i + k);
}

// bounding box + area per blob, one set of atomics per in-segment run
__global__ void __launch_bounds__(TISEG_THREADS)
k_blob_bbox(Geom g, const int* __restrict__ par, const int* __restrict__ rank, BlobInfo b) {
    Strip s;
    if (!warp_strip(g, s)) return;
    int p[STRIP_R];
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        int y = s.y0 + r;
        p[r] = (s.okx && y < g.H) ? par[s.base + (long long)y * g.W + s.x] : -1;
    }
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        int pl = __shfl_up_sync(FULL, p[r], 1);
        bool cont = s.lane > 0 && p[r] >= 0 && pl == p[r];
        unsigned m = __ballot_sync(FULL, cont);
        if (p[r] >= 0 && !cont) {
            int end = run_end_lane(m, s.lane);
            long long o = (long long)s.n * b.KS + rank[s.base + p[r]];
            atomicMax(&b.ymax[o], s.y0 + r);
            atomicMin(&b.xmin[o], s.x);
            atomicMax(&b.xmax[o], s.x + (end - s.lane));
            atomicAdd(&b.area[o], end - s.lane + 1);
        }
    }
}

// off[1..B] = exclusive prefix sum of area[1..B]
__global__ void k_blob_offsets(BlobInfo b) {
    __shared__ int s[256];
    __shared__ int carry;
    int n = blockIdx.x;
    long long o = (long long)n * b.KS;
    int k = b.count[n];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 1; base <= k; base += 256) {
        int i = base + threadIdx.x;
        int v = i <= k ? b.area[o + i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s[threadIdx.x];
        int c0 = carry;
        if (i <= k) b.off[o + i] = c0 + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c0 + incl;
        __syncthreads();
    }
}

// out = markers * mask (skimage: markers outside the mask are dropped)
__global__ void __launch_bounds__(TISEG_THREADS)
k_ws_seed(long long total, const int32_t* __restrict__ markers, const int* __restrict__ par, int32_t* __restrict__ out, bool vec) {
    const long long i = flat4_index();
    if (i >= total) return;
    Pack4<int> m = ld4(markers, i, total, vec), p = ld4(par, i, total, vec), o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.v[k] = p.v[k] >= 0 ? m.v[k] : 0;
    st4(out, i, total, vec, o);
}

// ---- uint8 levels: 256 FIFO buckets -----------------------------------------------------------------------
// A blob's flood is a chain of dependent accesses (pop -> look at 4 neighbours -> push), so its speed is the
// latency of the memory it runs in.  Each warp therefore STAGES its blob into shared memory first: the bounding box
// plus a one-pixel frame (so the flood needs no bounds checks), 5 bytes per cell — level (u8), FIFO link (u16) and
// the local index of the seed pixel whose label the cell inherits (u16).  The flood then runs entirely in shared
// memory (~30-cycle accesses instead of ~600), and the labels are written back with coalesced stores.  Blobs whose
// framed bounding box exceeds WS_CAP cells take the same algorithm in global memory.
// Persistent grid: one CTA per SM; every CTA walks all tiles (starting at a different one) and drains each
// tile's blob queue with an atomic counter, so the load balances across tiles and blob sizes.
#define WS_CAP 4096
#define WS_NOTIN 0xFFFFu
#define WS_UNLAB 0xFFFEu
#define WS_END 0xFFFFu
extern __shared__ __align__(16) unsigned char ws_smem[];

__global__ void __launch_bounds__(TISEG_THREADS, 1)
k_ws_flood_u8(Geom g, const uint8_t* __restrict__ image, const int* __restrict__ par, BlobInfo b, int* queue,
              int* next, int32_t* out) {
    __shared__ int s_head[TISEG_WARPS_PER_BLOCK][256];
    __shared__ int s_tail[TISEG_WARPS_PER_BLOCK][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned short* lab = reinterpret_cast<unsigned short*>(ws_smem + (size_t)warp * WS_CAP * 5);
    unsigned short* nxs = lab + WS_CAP;
    unsigned char* lvl = reinterpret_cast<unsigned char*>(nxs + WS_CAP);
    int* head = s_head[warp];
    int* tail = s_tail[warp];
    const int W = g.W, H = g.H;
    for (int tn = 0; tn < g.N; ++tn) {
        const int n = (blockIdx.x + tn) % g.N;
        const int B = b.count[n];
        const long long base = (long long)n * g.P, ko = (long long)n * b.KS;
        const uint8_t* I = image + base;
        const int* tp = par + base;
        int32_t* o = out + base;
        int* nx = next + base;
        for (;;) {
            int bid = 0;
            if (lane == 0) bid = atomicAdd(&queue[n], 1) + 1;
            bid = __shfl_sync(FULL, bid, 0);
            if (bid > B) break;
            for (int i = lane; i < 256; i += 32) { head[i] = -1; tail[i] = -1; }
            const int root = b.root[ko + bid];
            const int y0 = root / W, y1 = b.ymax[ko + bid], x0 = b.xmin[ko + bid], x1 = b.xmax[ko + bid];
            const int w = x1 - x0 + 1, h = y1 - y0 + 1, wp = w + 2;
            int cur = 256;
            __syncwarp();
            if (wp * (h + 2) <= WS_CAP) {
                // ---- stage the framed bounding box; seeds enter their buckets in raster order
                for (int ly = 0; ly < h + 2; ++ly) {
                    for (int lxb = 0; lxb < wp; lxb += 32) {
                        const int lx = lxb + lane, j = ly * wp + lx;
                        bool seed = false;
                        unsigned v = 0;
                        if (lx < wp) {
                            unsigned L = WS_NOTIN;
                            if (ly >= 1 && ly <= h && lx >= 1 && lx <= w) {
                                const int gi = (y0 + ly - 1) * W + x0 + lx - 1;
                                if (tp[gi] == root) { v = I[gi]; seed = o[gi] != 0; L = seed ? (unsigned)j : WS_UNLAB; }
                            }
                            lab[j] = (unsigned short)L;
                            lvl[j] = (unsigned char)v;
                        }
                        unsigned m = __ballot_sync(FULL, seed);
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const int slv = __shfl_sync(FULL, (int)v, src);
                            if (lane == 0) {
                                const int pix = ly * wp + lxb + src;
                                nxs[pix] = WS_END;
                                const int t = tail[slv];
                                if (t < 0) head[slv] = pix; else nxs[t] = (unsigned short)pix;
                                tail[slv] = pix;
                                if (slv < cur) cur = slv;
                            }
                        }
                    }
                }
                __syncwarp();
                // ---- the ordered flood, in shared memory
                if (lane == 0) {
                    for (;;) {
                        while (cur < 256 && head[cur] < 0) ++cur;
                        if (cur >= 256) break;
                        const int pix = head[cur];
                        const unsigned nxt = nxs[pix];
                        head[cur] = nxt == WS_END ? -1 : (int)nxt;
                        if (nxt == WS_END) tail[cur] = -1;
                        const unsigned short L = lab[pix];
                        const int nb[4] = {pix - wp, pix - 1, pix + 1, pix + wp};      // up, left, right, down
                        unsigned short ln[4];
                        unsigned char vn[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) { ln[k] = lab[nb[k]]; vn[k] = lvl[nb[k]]; }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (ln[k] == WS_UNLAB) {
                                lab[nb[k]] = L;                                       // labelled at push time
                                nxs[nb[k]] = WS_END;
                                const int t = tail[vn[k]];
                                if (t < 0) head[vn[k]] = nb[k]; else nxs[t] = (unsigned short)nb[k];
                                tail[vn[k]] = nb[k];
                                if (vn[k] < cur) cur = vn[k];
                            }
                        }
                    }
                }
                __syncwarp();
                // ---- write back: every flooded cell takes the marker label of its seed pixel
                for (int ly = 1; ly <= h; ++ly) {
                    for (int lx = 1 + lane; lx <= w; lx += 32) {
                        const int j = ly * wp + lx;
                        const unsigned L = lab[j];
                        if (L < WS_UNLAB && L != (unsigned)j) {
                            const int sy = L / wp, sx = L - sy * wp;
                            o[(y0 + ly - 1) * W + x0 + lx - 1] = o[(y0 + sy - 1) * W + x0 + sx - 1];
                        }
                    }
                }
                __syncwarp();
                continue;
            }
            // ---- fallback: the same flood in global memory (framed bounding box does not fit)
            for (int y = y0; y <= y1; ++y) {
                for (int xb = x0; xb <= x1; xb += 32) {
                    int x = xb + lane;
                    bool seed = false;
                    int lv = 0;
                    if (x <= x1) {
                        int idx = y * W + x;
                        if (tp[idx] == root && o[idx] != 0) { seed = true; lv = I[idx]; }
                    }
                    unsigned m = __ballot_sync(FULL, seed);
                    while (m) {
                        int src = __ffs(m) - 1;
                        m &= m - 1;
                        int slv = __shfl_sync(FULL, lv, src);
                        if (lane == 0) {
                            int pix = y * W + xb + src;
                            nx[pix] = -1;
                            int t = tail[slv];
                            if (t < 0) head[slv] = pix; else nx[t] = pix;
                            tail[slv] = pix;
                            if (slv < cur) cur = slv;
                        }
                    }
                }
            }
            if (lane == 0) {
                for (;;) {
                    while (cur < 256 && head[cur] < 0) ++cur;
                    if (cur >= 256) break;
                    const int pix = head[cur];
                    const int nxt = nx[pix];
                    head[cur] = nxt;
                    if (nxt < 0) tail[cur] = -1;
                    const int lab_g = o[pix];
                    const int y = pix / W, x = pix - y * W;
                    const int nb[4] = {pix - W, pix - 1, pix + 1, pix + W};
                    const bool ok[4] = {y > 0, x > 0, x + 1 < W, y + 1 < H};
                    int pv[4], ov[4], lv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        pv[k] = -1; ov[k] = 1; lv[k] = 0;
                        if (ok[k]) { pv[k] = tp[nb[k]]; ov[k] = o[nb[k]]; lv[k] = I[nb[k]]; }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (pv[k] >= 0 && ov[k] == 0) {
                            o[nb[k]] = lab_g;
                            nx[nb[k]] = -1;
                            int t = tail[lv[k]];
                            if (t < 0) head[lv[k]] = nb[k]; else nx[t] = nb[k];
                            tail[lv[k]] = nb[k];
                            if (lv[k] < cur) cur = lv[k];
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---- fp64 values: binary heap keyed (value, age, index) -------------------------------------------------
struct __align__(16) HeapItem { double v; unsigned age; int idx; };

__device__ __forceinline__ bool heap_less(const HeapItem& a, const HeapItem& b) {
    if (a.v != b.v) return a.v < b.v;
    if (a.age != b.age) return a.age < b.age;
    return a.idx < b.idx;
}
__device__ __forceinline__ void heap_push(HeapItem* h, int& sz, HeapItem it) {
    int i = sz++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        HeapItem hp = h[p];
        if (!heap_less(it, hp)) break;
        h[i] = hp;
        i = p;
    }
    h[i] = it;
}
__device__ __forceinline__ HeapItem heap_pop(HeapItem* h, int& sz) {
    HeapItem top = h[0];
    HeapItem last = h[--sz];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1;
        if (l >= sz) break;
        HeapItem c = h[l];
        if (l + 1 < sz) {
            HeapItem r = h[l + 1];
            if (heap_less(r, c)) { c = r; l = l + 1; }
        }
        if (!heap_less(c, last)) break;
        h[i] = c;
        i = l;
    }
    if (sz > 0) h[i] = last;
    return top;
}

__global__ void __launch_bounds__(TISEG_THREADS)
k_ws_flood_f64(Geom g, const double* __restrict__ image, const int* __restrict__ par, BlobInfo b, int* queue,
               HeapItem* heap, int32_t* out) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const int B = b.count[n];
    const long long base = (long long)n * g.P, ko = (long long)n * b.KS;
    const double* I = image + base;
    const int* tp = par + base;
    int32_t* o = out + base;
    const int W = g.W, H = g.H;
    for (;;) {
        int bid = 0;
        if (lane == 0) bid = atomicAdd(&queue[n], 1) + 1;
        bid = __shfl_sync(FULL, bid, 0);
        if (bid > B) break;
        const int root = b.root[ko + bid];
        const int y0 = root / W, y1 = b.ymax[ko + bid], x0 = b.xmin[ko + bid], x1 = b.xmax[ko + bid];
        HeapItem* h = heap + base + b.off[ko + bid];
        int sz = 0;
        for (int y = y0; y <= y1; ++y) {
            for (int xb = x0; xb <= x1; xb += 32) {
                int x = xb + lane;
                bool seed = false;
                double v = 0.0;
                if (x <= x1) {
                    int idx = y * W + x;
                    if (tp[idx] == root && o[idx] != 0) { seed = true; v = I[idx]; }
                }
                unsigned m = __ballot_sync(FULL, seed);
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    double sv = __shfl_sync(FULL, v, src);
                    if (lane == 0) { HeapItem it; it.v = sv; it.age = 0u; it.idx = y * W + xb + src; heap_push(h, sz, it); }
                }
            }
        }
        if (lane == 0) {
            unsigned age = 0;
            while (sz > 0) {
                HeapItem e = heap_pop(h, sz);
                const int pix = e.idx;
                const int lab = o[pix];
                const int y = pix / W, x = pix - y * W;
                const int nb[4] = {pix - W, pix - 1, pix + 1, pix + W};
                const bool ok[4] = {y > 0, x > 0, x + 1 < W, y + 1 < H};
                int pv[4], ov[4];
                double vv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pv[k] = -1; ov[k] = 1; vv[k] = 0.0;
                    if (ok[k]) { pv[k] = tp[nb[k]]; ov[k] = o[nb[k]]; vv[k] = I[nb[k]]; }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (pv[k] >= 0 && ov[k] == 0) {
                        o[nb[k]] = lab;
                        HeapItem it; it.v = vv[k]; it.age = ++age; it.idx = nb[k];
                        heap_push(h, sz, it);
                    }
                }
            }
        }
        __syncwarp();
    }
}

int ws_seed(tiseg_ctx* c, const Geom& g, const int32_t* markers, const int* par, int32_t* out) {
    long long total = (long long)g.N * g.P;
    TISEG_LAUNCH(c, k_ws_seed, flat4_grid(total), TISEG_THREADS, 0, total, markers, par, out, aligned16(markers, par, out));
    return TISEG_OK;
}

int blobs_describe(tiseg_ctx* c, const Geom& g, const int* par, const int* rank, BlobInfo& b, bool want_offsets) {
    TISEG_LAUNCH(c, k_blob_init, dim3(8, g.N), 256, 0, b, g.W);
    TISEG_LAUNCH(c, k_blob_roots, dim3(flat4_grid(g.P), g.N), TISEG_THREADS, 0, (long long)g.P, par, rank, b, (g.P % 4 == 0) && aligned16(par));
    TISEG_LAUNCH(c, k_blob_bbox, strip_grid(g), TISEG_THREADS, 0, g, par, rank, b);
    if (want_offsets) TISEG_LAUNCH(c, k_blob_offsets, g.N, 256, 0, b);
    return TISEG_OK;
}

static inline int flood_blocks(tiseg_ctx* c, int N) {
    // persistent-style grid: enough warps to fill the chip several times over, split evenly over tiles
    int per_tile = (c->sm_count * 8 * 4 + N - 1) / N;
    if (per_tile < 1) per_tile = 1;
    if (per_tile > 512) per_tile = 512;
    return per_tile;
}

int watershed_u8_dev(tiseg_ctx* c, const Geom& g, const uint8_t* image, const int* par, const int* rank,
                     const BlobInfo& b, int32_t* out) {
    (void)rank;
    int* queue = ws<int>(c, (size_t)g.N);
    int* next = ws<int>(c, (size_t)g.N * g.P);
    if (!queue || !next) return TISEG_ERR_CUDA;
    TISEG_TRY(zero(c, queue, (size_t)g.N * sizeof(int)));
    const size_t smem = (size_t)TISEG_WARPS_PER_BLOCK * WS_CAP * 5;
    static bool attr_set = false;
    if (!attr_set) {
        TISEG_CHECK(cudaFuncSetAttribute(k_ws_flood_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    TISEG_LAUNCH(c, k_ws_flood_u8, c->sm_count, TISEG_THREADS, smem, g, image, par, b, queue, next, out);
    return TISEG_OK;
}

int watershed_f64_dev(tiseg_ctx* c, const Geom& g, const double* image, const int* par, const int* rank,
                      const BlobInfo& b, int32_t* out) {
    (void)rank;
    int* queue = ws<int>(c, (size_t)g.N);
    HeapItem* heap = ws<HeapItem>(c, (size_t)g.N * g.P);
    if (!queue || !heap || !b.off) return TISEG_ERR_CUDA;
    TISEG_TRY(zero(c, queue, (size_t)g.N * sizeof(int)));
    TISEG_LAUNCH(c, k_ws_flood_f64, dim3(flood_blocks(c, g.N), g.N), TISEG_THREADS, 0, g, image, par, b, queue, heap, out);
    return TISEG_OK;
}

template <class T>
static int watershed_entry(tiseg_ctx* c, const T* image, const int32_t* markers, const uint8_t* mask, int N, int H,
                           int W, int32_t* out) {
    if (!c || !image || !markers || !out) { set_error("tiseg_watershed: bad argument"); return TISEG_ERR_ARG; }
    TISEG_TRY(check_geom(N, H, W));
    begin_call(c);
    Geom g = make_geom(N, H, W);
    size_t total = (size_t)N * g.P;
    const T* d_img = in(c, image, total);
    const int32_t* d_mk = in(c, markers, total);
    const uint8_t* d_mask = mask ? in(c, mask, total) : nullptr;
    int32_t* d_out = tiseg::out(c, out, total);
    int* par = ws<int>(c, total);
    int* rank = ws<int>(c, total);
    if (!d_img || !d_mk || !d_out || !par || !rank) return TISEG_ERR_CUDA;
    BlobInfo b;
    constexpr bool F64 = sizeof(T) == 8;
    if (d_mask) TISEG_TRY(blobs_build(c, g, ImgMaskU8{d_mask}, par, rank, b, F64));
    else        TISEG_TRY(blobs_build(c, g, ImgAll{}, par, rank, b, F64));
    TISEG_TRY(ws_seed(c, g, d_mk, par, d_out));
    if constexpr (F64) TISEG_TRY(watershed_f64_dev(c, g, (const double*)d_img, par, rank, b, d_out));
    else               TISEG_TRY(watershed_u8_dev(c, g, (const uint8_t*)d_img, par, rank, b, d_out));
    return end_call(c);
}

}  // namespace tiseg

using namespace tiseg;

extern "C" {

int tiseg_watershed_u8(tiseg_ctx* c, const uint8_t* image, const int32_t* markers, const uint8_t* mask, int N, int H,
                       int W, int32_t* out) {
    return watershed_entry<uint8_t>(c, image, markers, mask, N, H, W, out);
}

int tiseg_watershed_f64(tiseg_ctx* c, const double* image, const int32_t* markers, const uint8_t* mask, int N, int H,
                        int W, int32_t* out) {
    return watershed_entry<double>(c, image, markers, mask, N, H, W, out);
}

}  // extern "C"
