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
// A blob's flood is a chain of dependent steps (pop -> look at 4 neighbours -> push), so its speed is the latency
// of the memory it runs in and the number of floods in flight.  Each warp STAGES its blob into shared memory: the
// bounding box plus a one-pixel frame (no bounds checks in the flood), 5 bytes per cell — level (u8), FIFO link
// (u16) and the local index of the seed pixel whose label the cell inherits (u16) — floods it there, and writes
// the labels back with four loads in flight per lane.
// One persistent CTA per SM holds 22 "small" warps (framed box <= 1024 cells: ~94 % of nuclei blobs) and 4 "large"
// warps (<= 4096 cells); blobs are pre-sorted into a small and a large list per tile; every CTA walks all tiles
// (starting at a different one) and drains the lists through atomic cursors, so work balances across tiles and
// blob sizes.  Boxes beyond 4096 cells are flooded by a large warp directly in global memory.
#define WS_SMALL_WARPS 22
#define WS_LARGE_WARPS 4
#define WS_WARPS (WS_SMALL_WARPS + WS_LARGE_WARPS)
#define WS_CAP_S 1024
#define WS_CAP_L 4096
#define WS_NOTIN 0xFFFFu
#define WS_UNLAB 0xFFFEu
#define WS_END 0xFFFFu
#define WS_CELL_BYTES ((size_t)(WS_SMALL_WARPS * WS_CAP_S + WS_LARGE_WARPS * WS_CAP_L) * 5)
#define WS_SMEM_BYTES (WS_CELL_BYTES + (size_t)WS_WARPS * 1024)
extern __shared__ __align__(16) unsigned char ws_smem[];

struct FloodLists {
    int* small; int* large;   // [N, KS] blob ids
    int* ns; int* nl;         // [N] list lengths
    int* qs; int* ql;         // [N] cursors
};

__global__ void k_blob_classify(BlobInfo b, int W, FloodLists L) {
    int n = blockIdx.y;
    long long ko = (long long)n * b.KS;
    int B = b.count[n];
    for (int bid = 1 + blockIdx.x * blockDim.x + threadIdx.x; bid <= B; bid += gridDim.x * blockDim.x) {
        int h = b.ymax[ko + bid] - b.root[ko + bid] / W + 1, w = b.xmax[ko + bid] - b.xmin[ko + bid] + 1;
        if ((w + 2) * (h + 2) <= WS_CAP_S) L.small[ko + atomicAdd(&L.ns[n], 1)] = bid;
        else L.large[ko + atomicAdd(&L.nl[n], 1)] = bid;
    }
}

// One blob, staged in shared memory.  head/tail: 256 u16 each (0xFFFF = empty bucket).
__device__ __forceinline__ void flood_blob_staged(int lane, int W, const uint8_t* __restrict__ I, const int* __restrict__ tp,
                                                  int32_t* o, int root, int y0, int x0, int w, int h,
                                                  unsigned short* lab, unsigned short* nxs, unsigned char* lvl,
                                                  unsigned short* head, unsigned short* tail) {
    const int wp = w + 2, cells = wp * (h + 2);
    for (int i = lane; i < 256; i += 32) { head[i] = 0xFFFF; tail[i] = 0xFFFF; }
    // stage: cells walked as a flat array (all lanes busy whatever the box width), four cells per lane loaded
    // before any is used (12 loads in flight)
    for (int j0 = 0; j0 < cells; j0 += 128) {
        int jj[4], gi[4];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            jj[u] = j0 + 32 * u + lane;
            const int ly = jj[u] / wp, lx = jj[u] - ly * wp;
            in[u] = jj[u] < cells && ly >= 1 && ly <= h && lx >= 1 && lx <= w;
            gi[u] = in[u] ? (y0 + ly - 1) * W + x0 + lx - 1 : root;
        }
        int tpv[4], ov[4];
        unsigned char iv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { tpv[u] = tp[gi[u]]; iv[u] = I[gi[u]]; ov[u] = o[gi[u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (jj[u] >= cells) continue;
            unsigned L = WS_NOTIN;
            if (in[u] && tpv[u] == root) L = ov[u] != 0 ? (unsigned)jj[u] : WS_UNLAB;
            lab[jj[u]] = (unsigned short)L;
            lvl[jj[u]] = iv[u];
        }
    }
    __syncwarp();
    // seeds (cells that point at themselves) enter their buckets in raster order
    int cur = 256;
    for (int j0 = 0; j0 < cells; j0 += 32) {
        const int j = j0 + lane;
        const bool seed = j < cells && lab[j] == (unsigned short)j;
        const int v = j < cells ? lvl[j] : 0;
        unsigned m = __ballot_sync(FULL, seed);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int slv = __shfl_sync(FULL, v, src);
            if (lane == 0) {
                const int pix = j0 + src;
                nxs[pix] = WS_END;
                const unsigned t = tail[slv];
                if (t == 0xFFFFu) head[slv] = (unsigned short)pix; else nxs[t] = (unsigned short)pix;
                tail[slv] = (unsigned short)pix;
                if (slv < cur) cur = slv;
            }
        }
    }
    __syncwarp();
    // the ordered flood
    if (lane == 0) {
        for (;;) {
            while (cur < 256 && head[cur] == 0xFFFFu) ++cur;
            if (cur >= 256) break;
            const int pix = head[cur];
            const unsigned nxt = nxs[pix];
            head[cur] = (unsigned short)nxt;
            if (nxt == WS_END) tail[cur] = 0xFFFF;
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
                    const unsigned t = tail[vn[k]];
                    if (t == 0xFFFFu) head[vn[k]] = (unsigned short)nb[k]; else nxs[t] = (unsigned short)nb[k];
                    tail[vn[k]] = (unsigned short)nb[k];
                    if (vn[k] < cur) cur = vn[k];
                }
            }
        }
    }
    __syncwarp();
    // write back: every flooded cell takes the marker label of its seed pixel
    for (int j0 = 0; j0 < cells; j0 += 128) {
        int dst[4], src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u + lane;
            dst[u] = -1; src[u] = root;
            if (j < cells) {
                const unsigned L = lab[j];
                if (L < WS_UNLAB && L != (unsigned)j) {
                    const int ly = j / wp, lx = j - ly * wp, sy = L / wp, sx = L - sy * wp;
                    dst[u] = (y0 + ly - 1) * W + x0 + lx - 1;
                    src[u] = (y0 + sy - 1) * W + x0 + sx - 1;
                }
            }
        }
        int val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) val[u] = o[src[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (dst[u] >= 0) o[dst[u]] = val[u];
    }
    __syncwarp();
}

// The same flood in global memory (framed bounding box too large for shared memory); head/tail are int[256].
__device__ __forceinline__ void flood_blob_global(int lane, int W, int H, const uint8_t* __restrict__ I,
                                                  const int* __restrict__ tp, int32_t* o, int* nx, int root, int y0,
                                                  int y1, int x0, int x1, int* head, int* tail) {
    for (int i = lane; i < 256; i += 32) { head[i] = -1; tail[i] = -1; }
    __syncwarp();
    int cur = 256;
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

__global__ void __launch_bounds__(32 * WS_WARPS, 1)
k_ws_flood_u8(Geom g, const uint8_t* __restrict__ image, const int* __restrict__ par, BlobInfo b, FloodLists L,
              int* next, int* gheads, int32_t* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool large = warp >= WS_SMALL_WARPS;
    const size_t off = large ? ((size_t)WS_SMALL_WARPS * WS_CAP_S + (size_t)(warp - WS_SMALL_WARPS) * WS_CAP_L) * 5
                             : (size_t)warp * WS_CAP_S * 5;
    const int cap = large ? WS_CAP_L : WS_CAP_S;
    unsigned short* lab = reinterpret_cast<unsigned short*>(ws_smem + off);
    unsigned short* nxs = lab + cap;
    unsigned char* lvl = reinterpret_cast<unsigned char*>(nxs + cap);
    unsigned short* head = reinterpret_cast<unsigned short*>(ws_smem + WS_CELL_BYTES + (size_t)warp * 1024);
    unsigned short* tail = head + 256;
    const int W = g.W, H = g.H;
    for (int tn = 0; tn < g.N; ++tn) {
        const int n = (blockIdx.x + tn) % g.N;
        const long long base = (long long)n * g.P, ko = (long long)n * b.KS;
        const uint8_t* I = image + base;
        const int* tp = par + base;
        int32_t* o = out + base;
        for (int pass = large ? 0 : 1; pass < 2; ++pass) {
            const int* list = (pass == 0 ? L.large : L.small) + ko;
            int* cursor = (pass == 0 ? L.ql : L.qs) + n;
            const int count = (pass == 0 ? L.nl : L.ns)[n];
            for (;;) {
                int k = 0;
                if (lane == 0) k = atomicAdd(cursor, 1);
                k = __shfl_sync(FULL, k, 0);
                if (k >= count) break;
                const int bid = list[k];
                const int root = b.root[ko + bid];
                const int y0 = root / W, y1 = b.ymax[ko + bid], x0 = b.xmin[ko + bid], x1 = b.xmax[ko + bid];
                const int w = x1 - x0 + 1, h = y1 - y0 + 1;
                if ((w + 2) * (h + 2) <= cap) {
                    flood_blob_staged(lane, W, I, tp, o, root, y0, x0, w, h, lab, nxs, lvl, head, tail);
                } else {
                    // only reachable by large warps: 512 ints of bucket heads per warp in global scratch
                    int* gh = gheads + ((size_t)blockIdx.x * WS_LARGE_WARPS + (warp - WS_SMALL_WARPS)) * 512;
                    flood_blob_global(lane, W, H, I, tp, o, next + base, root, y0, y1, x0, x1, gh, gh + 256);
                }
            }
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
    const int N = g.N;
    const size_t ks = (size_t)N * b.KS;
    FloodLists L;
    L.small = ws<int>(c, ks); L.large = ws<int>(c, ks);
    int* counters = ws<int>(c, 4 * (size_t)N);
    int* next = ws<int>(c, (size_t)N * g.P);
    int* gheads = ws<int>(c, (size_t)c->sm_count * WS_LARGE_WARPS * 512);
    if (!L.small || !L.large || !counters || !next || !gheads) return TISEG_ERR_CUDA;
    L.ns = counters; L.nl = counters + N; L.qs = counters + 2 * N; L.ql = counters + 3 * N;
    TISEG_TRY(zero(c, counters, 4 * (size_t)N * sizeof(int)));
    TISEG_LAUNCH(c, k_blob_classify, dim3(8, N), 256, 0, b, g.W, L);
    static bool attr_set = false;
    if (!attr_set) {
        TISEG_CHECK(cudaFuncSetAttribute(k_ws_flood_u8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
        attr_set = true;
    }
    TISEG_LAUNCH(c, k_ws_flood_u8, c->sm_count, 32 * WS_WARPS, WS_SMEM_BYTES, g, image, par, b, L, next, gheads, out);
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
