// A8 — DIST postprocess: tiseg/models/segmentors/dist.py:275-284 -> dynamic_watershed_alias (:114-129)
// with prepare_prob (:31-40), H_reconstruction_erosion (:43-57; lambda = 0.0 => identity), find_maxima
// (:60-71), arrange_label (:101-111) and generate_wsl (:83-98).
//
//   d  = int32(clip(dist, 0, 255))             (truncation)             k_dist_prep
//   I  = 255 - uint8(d);  b = d > 0.5  <=>  I < 255
//   markers = label(recon_by_erosion(min(255, I+1), I) - I, masked by b)
//           = 8-connected regional-minimum plateaus of I with value < 255, raster ids
//                                                                       plateau CCL + k_plateau_lower + rank
//   ws = watershed(I, markers, mask=b)                                   K6 (bucket flood)
//   arranged = label(ws, background = most frequent value of ws)         k_ws_hist + k_pick_bg + CCL
//   arranged[3x3 window holds >= 2 different non-zero labels] = 0        k_wsl_remove
#include "ccl.cuh"
#include "watershed.cuh"

namespace tiseg {

__global__ void __launch_bounds__(TISEG_THREADS)
k_dist_prep(long long total, const float* __restrict__ dist, uint8_t* __restrict__ I, bool vec) {
    const long long i = flat4_index();
    if (i >= total) return;
    Pack4<float> d = ld4(dist, i, total, vec);
    Pack4<uint8_t> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = d.v[k];
        if (v > 255.f) v = 255.f;      // dist.py:277-278 (comparisons are false for NaN, like numpy)
        if (v < 0.f) v = 0.f;
        int t = (int)v;                // astype('int32'): truncation
        o.v[k] = (uint8_t)(255 - (t & 255));
    }
    st4(I, i, total, vec, o);
}

// Regional-minimum plateaus without labelling every plateau of the image.  A pixel is a CANDIDATE if its value is
// below 255 and no 8-neighbour is strictly lower.  A plateau P (maximal 8-connected set of equal values) is a regional
// minimum iff all its pixels are candidates.  Label the candidates only (equal value, 8-connected): if P is a
// minimum it comes out as one component, none of whose pixels has an equal-valued non-candidate neighbour; if P is
// not, every candidate component C inside it is a proper subset of the connected P, so some pixel of C touches an
// equal-valued pixel outside C — which must be a non-candidate (a candidate would have joined C).  Hence:
// minimum plateaus = candidate components without an "equal-valued non-candidate neighbour" flag.  On a distance map
// the candidates are the few pixels around each nucleus centre, so the labelling pass skips almost every row.
// One thread = four horizontally adjacent pixels (one 32-bit load per row); threads whose four pixels are all
// background (the common case) stop after that single load.
__device__ __forceinline__ unsigned ld_u8x4(const uint8_t* __restrict__ row, int x, int W, unsigned oob, bool vec) {
    if (vec) return *reinterpret_cast<const unsigned*>(row + x);
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) r |= (x + k < W ? (unsigned)row[x + k] : oob) << (8 * k);
    return r;
}
__global__ void __launch_bounds__(TISEG_THREADS)
k_min_candidates(Geom g, const uint8_t* __restrict__ I, uint8_t* __restrict__ cand, bool vec) {
    const int W4 = (g.W + 3) >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)W4 * g.H) return;
    const int y = (int)(t / W4), x = (int)(t - (long long)y * W4) * 4, n = blockIdx.y;
    const uint8_t* It = I + (long long)n * g.P;
    uint8_t* out = cand + (long long)n * g.P + (long long)y * g.W + x;
    const unsigned c = ld_u8x4(It + (long long)y * g.W, x, g.W, 255u, vec);
    unsigned res = 0;
    if (c != 0xffffffffu) {
        // rows y-1, y, y+1, columns x-1 .. x+4; out-of-image taps can never be lower
        unsigned rows[3]; int lft[3], rgt[3];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            const bool ok = yy >= 0 && yy < g.H;
            const uint8_t* rp = It + (long long)yy * g.W;
            rows[dy + 1] = dy == 0 ? c : (ok ? ld_u8x4(rp, x, g.W, 255u, vec) : 0xffffffffu);
            lft[dy + 1] = (ok && x > 0) ? rp[x - 1] : 255;
            rgt[dy + 1] = (ok && x + 4 < g.W) ? rp[x + 4] : 255;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = (c >> (8 * k)) & 255;
            int mn = 255;
#pragma unroll
            for (int r3 = 0; r3 < 3; ++r3) {
                const int a = k == 0 ? lft[r3] : (int)((rows[r3] >> (8 * (k - 1))) & 255);
                const int m = (int)((rows[r3] >> (8 * k)) & 255);
                const int z = k == 3 ? rgt[r3] : (int)((rows[r3] >> (8 * (k + 1))) & 255);
                mn = min(mn, min(a, min(m, z)));
            }
            if (v < 255 && mn >= v) res |= 1u << (8 * k);
        }
    }
    if (vec) *reinterpret_cast<unsigned*>(out) = res;
    else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (x + k < g.W) out[k] = (uint8_t)((res >> (8 * k)) & 255);
    }
}

// low[root] = 1 if a pixel of the candidate component has an equal-valued neighbour that is not a candidate.
// Candidates are rare: one 32-bit load tells a thread that its four pixels hold none; only candidate pixels probe
// their eight neighbours (straight from L1/L2).
__global__ void __launch_bounds__(TISEG_THREADS)
k_cand_invalid(Geom g, const uint8_t* __restrict__ I, const uint8_t* __restrict__ cand, const int* __restrict__ par,
               uint8_t* low, bool vec) {
    const int W4 = (g.W + 3) >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)W4 * g.H) return;
    const int y = (int)(t / W4), x = (int)(t - (long long)y * W4) * 4, n = blockIdx.y;
    const long long base = (long long)n * g.P;
    const uint8_t* It = I + base;
    const uint8_t* Ct = cand + base;
    const unsigned c = ld_u8x4(Ct + (long long)y * g.W, x, g.W, 0u, vec);
    if (!c) return;
    // rows y-1, y, y+1, columns x-1 .. x+4 of the level image and of the candidate mask (out-of-image: level 255,
    // which never equals a candidate's level)
    unsigned iv[3], cv[3];
    int il[3], ir[3], cl[3], cr[3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        const bool ok = yy >= 0 && yy < g.H;
        const long long ro = (long long)yy * g.W;
        iv[dy + 1] = ok ? ld_u8x4(It + ro, x, g.W, 255u, vec) : 0xffffffffu;
        cv[dy + 1] = dy == 0 ? c : (ok ? ld_u8x4(Ct + ro, x, g.W, 1u, vec) : 0x01010101u);
        il[dy + 1] = (ok && x > 0) ? It[ro + x - 1] : 255;
        ir[dy + 1] = (ok && x + 4 < g.W) ? It[ro + x + 4] : 255;
        cl[dy + 1] = (ok && x > 0) ? Ct[ro + x - 1] : 1;
        cr[dy + 1] = (ok && x + 4 < g.W) ? Ct[ro + x + 4] : 1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!((c >> (8 * k)) & 255)) continue;
        const int v = (iv[1] >> (8 * k)) & 255;
        bool bad = false;
#pragma unroll
        for (int r3 = 0; r3 < 3; ++r3) {
            const int ia = k == 0 ? il[r3] : (int)((iv[r3] >> (8 * (k - 1))) & 255);
            const int ca = k == 0 ? cl[r3] : (int)((cv[r3] >> (8 * (k - 1))) & 255);
            const int im = (int)((iv[r3] >> (8 * k)) & 255), cm = (int)((cv[r3] >> (8 * k)) & 255);
            const int iz = k == 3 ? ir[r3] : (int)((iv[r3] >> (8 * (k + 1))) & 255);
            const int cz = k == 3 ? cr[r3] : (int)((cv[r3] >> (8 * (k + 1))) & 255);
            bad |= (ia == v && !ca) || (iz == v && !cz);
            if (r3 != 1) bad |= im == v && !cm;
        }
        if (bad) {
            int root = par[base + (long long)y * g.W + x + k];
            if (!low[base + root]) low[base + root] = 1;
        }
    }
}

// clear, in the bitmap of the roots that ccl_flatten left, the roots whose component is not a minimum plateau
__global__ void __launch_bounds__(TISEG_THREADS)
k_filter_root_bits(Geom g, const uint8_t* __restrict__ low, unsigned* bits) {
    const long long words = (long long)g.H * g.SEG;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= words) return;
    const int n = blockIdx.y;
    unsigned m = bits[(long long)n * words + t];
    if (!m) return;
    const int y = (int)(t / g.SEG), seg = (int)(t - (long long)y * g.SEG);
    unsigned keep = m;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if (low[(long long)n * g.P + (long long)y * g.W + seg * 32 + b]) keep &= ~(1u << b);
    }
    bits[(long long)n * words + t] = keep;
}

struct SelMinimumRoot {         // roots of the candidate components that are regional-minimum plateaus
    const int* par; const uint8_t* low;
    __device__ __forceinline__ bool operator()(long long gi, int idx) const { return par[gi] == idx && !low[gi]; }
};

// markers = raster id of the minimum plateau a pixel belongs to, else 0; written to one or two maps
__global__ void __launch_bounds__(TISEG_THREADS)
k_markers_from_plateaus(long long P, const int* __restrict__ par, const uint8_t* __restrict__ low,
                        const int* __restrict__ rank, int32_t* __restrict__ markers, int32_t* __restrict__ copy, bool vec) {
    const long long base = (long long)blockIdx.y * P, i = flat4_index();
    if (i >= P) return;
    Pack4<int> p = ld4(par + base, i, P, vec), o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o.v[k] = 0;
        if (i + k < P && p.v[k] >= 0) { long long r = base + p.v[k]; if (!low[r]) o.v[k] = rank[r]; }
    }
    st4(markers + base, i, P, vec, o);
    if (copy) st4(copy + base, i, P, vec, o);
}

// histogram of the flood labels (values 0..K) and the first raster pixel of each label, one pair of atomics per
// in-segment run
__global__ void __launch_bounds__(TISEG_THREADS)
k_ws_hist(Geom g, const int32_t* __restrict__ ws, int* hist, int* first, int KS) {
    Strip s;
    if (!warp_strip(g, s)) return;
    int v[STRIP_R];
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        int y = s.y0 + r;
        v[r] = (s.okx && y < g.H) ? ws[s.base + (long long)y * g.W + s.x] : -1;
    }
#pragma unroll
    for (int r = 0; r < STRIP_R; ++r) {
        if (!__ballot_sync(0xffffffffu, v[r] > 0)) continue;       // (uniform) nothing but background here
        int vl = __shfl_up_sync(0xffffffffu, v[r], 1);
        bool cont = s.lane > 0 && v[r] == vl;
        unsigned m = __ballot_sync(0xffffffffu, cont);
        // label 0 (the background, by far the longest runs) is not counted: its total is P minus the others
        if (v[r] > 0 && !cont) {
            const long long o = (long long)s.n * KS + v[r];
            atomicAdd(&hist[o], run_end_lane(m, s.lane) - s.lane + 1);
            atomicMin(&first[o], (s.y0 + r) * g.W + s.x);
        }
    }
}

__global__ void k_init_label_tables(int* hist, int* first, int KS, const int* __restrict__ counts) {
    int n = blockIdx.y;
    int k = counts[n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= k; i += gridDim.x * blockDim.x) {
        hist[(long long)n * KS + i] = 0;
        first[(long long)n * KS + i] = INT_MAX;
    }
}

// arrange_label's background: np.unique(return_counts) + argmax => the most frequent value, smallest value on ties.
// Tiles whose background is NOT 0 (one flood region larger than everything unlabelled) go on the list of the
// general relabelling path.
__global__ void k_pick_bg(const int* __restrict__ hist, int KS, const int* __restrict__ counts, int P, int* bg,
                          int* flagged, int* nflagged) {
    __shared__ unsigned long long s[256];
    __shared__ long long tot[256];
    int n = blockIdx.x;
    int k = counts[n];
    // key = (count << 32) | (0xffffffff - value): max key = largest count, then smallest value
    unsigned long long best = 0;
    long long sum = 0;
    for (int v = 1 + threadIdx.x; v <= k; v += blockDim.x) {
        unsigned cnt = (unsigned)hist[(long long)n * KS + v];
        sum += cnt;
        unsigned long long key = ((unsigned long long)cnt << 32) | (0xffffffffu - (unsigned)v);
        if (key > best) best = key;
    }
    s[threadIdx.x] = best;
    tot[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 128; d; d >>= 1) {
        if (threadIdx.x < d) {
            if (s[threadIdx.x + d] > s[threadIdx.x]) s[threadIdx.x] = s[threadIdx.x + d];
            tot[threadIdx.x] += tot[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long zero_key = ((unsigned long long)(unsigned)(P - tot[0]) << 32) | 0xffffffffu;   // value 0
        unsigned long long w = zero_key > s[0] ? zero_key : s[0];
        int b = (int)(0xffffffffu - (unsigned)(w & 0xffffffffu));
        bg[n] = b;
        if (b != 0) flagged[atomicAdd(nflagged, 1)] = n;
    }
}

// Fast relabelling when the background is 0.  Every flood region is one 8-connected component (its marker plateau is
// 8-connected and the flood grows it by 4-neighbours), so label(ws) only renumbers the regions by their first
// raster pixel: set the bit of each region's first pixel, rank the bitmap, look the ids up.
__global__ void k_first_bits(Geom g, const int* __restrict__ first, int KS, const int* __restrict__ counts, unsigned* bits) {
    int n = blockIdx.y;
    int k = counts[n];
    for (int l = 1 + blockIdx.x * blockDim.x + threadIdx.x; l <= k; l += gridDim.x * blockDim.x) {
        int idx = first[(long long)n * KS + l];
        if (idx == INT_MAX) continue;
        int y = idx / g.W, x = idx - y * g.W;
        atomicOr(&bits[((long long)n * g.H + y) * g.SEG + (x >> 5)], 1u << (x & 31));
    }
}
__global__ void k_arrange_lut(Geom g, const int* __restrict__ first, const int* __restrict__ rank, int KS,
                              const int* __restrict__ counts, int* __restrict__ lut) {
    int n = blockIdx.y;
    int k = counts[n];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= k; l += gridDim.x * blockDim.x) {
        int idx = l ? first[(long long)n * KS + l] : INT_MAX;
        lut[(long long)n * KS + l] = idx == INT_MAX ? 0 : rank[(long long)n * g.P + idx];
    }
}

// generate_wsl + "arranged[wsl > 0] = 0" (dist.py:83-98, 128)
// `lut` (may be null): the id of each label value, applied to the pixels that survive
template <bool LISTED>
__global__ void __launch_bounds__(TISEG_THREADS)
k_wsl_remove(Geom g, const int32_t* __restrict__ lab, const int* __restrict__ lut, int KS, int32_t* __restrict__ out) {
    Strip s;
    if (!warp_strip(g, s)) return;
    FOR_TILES(LISTED, g, n) {
    strip_set_tile(g, s, n);
    int c[STRIP_R + 2], l[STRIP_R + 2], r[STRIP_R + 2];
    strip_load_c<int>(g, s, lab + s.base, 0, c);
    bool any = false;
#pragma unroll
    for (int j = 1; j <= STRIP_R; ++j) any |= c[j] != 0;
    if (!__ballot_sync(0xffffffffu, any)) {                        // (uniform) a strip of background: zeros out
#pragma unroll
        for (int j = 1; j <= STRIP_R; ++j) {
            int y = s.y0 + j - 1;
            if (s.okx && y < g.H) out[s.base + (long long)y * g.W + s.x] = 0;
        }
        continue;
    }
    strip_fill_lr<int>(g, s, lab + s.base, 0, c, l, r);
#pragma unroll
    for (int j = 1; j <= STRIP_R; ++j) {
        int y = s.y0 + j - 1;
        if (!s.okx || y >= g.H) continue;
        int v = c[j];
        bool line = false;
        if (v != 0) {
            const int nb[8] = {l[j - 1], c[j - 1], r[j - 1], l[j], r[j], l[j + 1], c[j + 1], r[j + 1]};
#pragma unroll
            for (int k = 0; k < 8; ++k) line |= (nb[k] != 0 && nb[k] != v);
        }
        if (lut && v != 0 && !line) v = lut[(long long)s.n * KS + v];
        out[s.base + (long long)y * g.W + s.x] = line ? 0 : v;
    }
    }
}

int h_reconstruction_erosion_dev(tiseg_ctx* c, const Geom& g, const uint8_t* img, int h, uint8_t* out);   // recon.cu

int postproc_dist_dev(tiseg_ctx* c, const Geom& g, const float* dist, int lamb, int32_t* inst, int32_t* markers_out,
                      int32_t* ws_out) {
    int N = g.N, KS = g.P + 1;
    size_t total = (size_t)N * g.P;
    uint8_t* I0 = ws<uint8_t>(c, total);
    uint8_t* I = I0;
    uint8_t* low = ws<uint8_t>(c, total);
    int* par = ws<int>(c, total);
    int* rank = ws<int>(c, total);
    int* bpar = ws<int>(c, total);
    int* brank = ws<int>(c, total);
    uint8_t* cand = ws<uint8_t>(c, total);
    int32_t* wsl = ws_out ? ws_out : ws<int32_t>(c, total);
    int32_t* arranged = ws<int32_t>(c, total);
    int* nmark = ws<int>(c, (size_t)N);
    int* bg = ws<int>(c, (size_t)N);
    int* flagged = ws<int>(c, (size_t)N + 1);
    int* hist = ws<int>(c, (size_t)N * KS);
    int* first = ws<int>(c, (size_t)N * KS);
    int* lut = ws<int>(c, (size_t)N * KS);
    unsigned* fbits = ws<unsigned>(c, (size_t)N * g.H * g.SEG);
    if (!I0 || !low || !par || !rank || !bpar || !brank || !cand || !wsl || !arranged || !nmark || !bg || !flagged || !hist ||
        !first || !lut || !fbits) return TISEG_ERR_CUDA;
    int* nflagged = flagged + N;

    TISEG_LAUNCH(c, k_dist_prep, flat4_grid((long long)total), TISEG_THREADS, 0, (long long)total, dist, I0, aligned16(dist) && (((uintptr_t)I0) & 3) == 0);
    // Hrecons (dist.py:120): the identity for the lambda = 0.0 the reference hard-codes (dist.py:281); a real
    // H-minima reconstruction otherwise.  Markers and flood levels come from it, the mask b from the image itself.
    if (lamb > 0) {
        I = ws<uint8_t>(c, total);
        if (!I) return TISEG_ERR_CUDA;
        TISEG_TRY(h_reconstruction_erosion_dev(c, g, I0, lamb, I));
    }
    // markers: regional-minimum plateaus (8-connected, equal value) of I below 255, via the candidate pixels
    const bool vec4 = (g.W % 4 == 0) && ((((uintptr_t)I) | ((uintptr_t)cand)) & 3) == 0;
    const dim3 quad_grid((unsigned)(((long long)((g.W + 3) / 4) * g.H + TISEG_THREADS - 1) / TISEG_THREADS), (unsigned)N);
    TISEG_LAUNCH(c, k_min_candidates, quad_grid, TISEG_THREADS, 0, g, I, cand, vec4);
    TISEG_TRY(ccl_build(c, g, ImgEqU8Where{I, cand}, 2, par));
    unsigned* rbits = (unsigned*)c->rootblk;          // bitmap of the candidate components' roots (left by the flatten)
    c->rootblk_par = nullptr;
    TISEG_TRY(zero(c, low, total));
    TISEG_LAUNCH(c, k_cand_invalid, quad_grid, TISEG_THREADS, 0, g, I, cand, par, low, vec4);
    TISEG_LAUNCH(c, k_filter_root_bits, dim3((unsigned)(((long long)g.H * g.SEG + TISEG_THREADS - 1) / TISEG_THREADS), (unsigned)N),
                 TISEG_THREADS, 0, g, low, rbits);
    TISEG_TRY(rank_from_bits(c, g, rbits, rank, nmark));
    // every marker pixel lies inside the mask b = (I < 255), so the markers are the flood's seed map as they are
    TISEG_LAUNCH(c, k_markers_from_plateaus, dim3(flat4_grid(g.P), N), TISEG_THREADS, 0, (long long)g.P, par, low, rank, wsl,
                 markers_out, (g.P % 4 == 0) && aligned16(par, wsl, markers_out));
    // flood inside b, blob by blob
    BlobInfo b;
    TISEG_TRY(blobs_build(c, g, ImgBelowU8{I0, 255}, bpar, brank, b, false));
    TISEG_TRY(watershed_u8_dev(c, g, I, bpar, brank, b, wsl));
    // arrange_label
    TISEG_TRY(zero(c, nflagged, sizeof(int)));
    TISEG_TRY(zero(c, fbits, (size_t)N * g.H * g.SEG * sizeof(unsigned)));
    TISEG_LAUNCH(c, k_init_label_tables, dim3(8, N), 256, 0, hist, first, KS, nmark);
    TISEG_LAUNCH(c, k_ws_hist, strip_grid(g), TISEG_THREADS, 0, g, wsl, hist, first, KS);
    TISEG_LAUNCH(c, k_pick_bg, N, 256, 0, hist, KS, nmark, g.P, bg, flagged, nflagged);
    //   background 0 (every tile but degenerate ones): ids = rank of each region's first pixel; watershed lines
    //   are found on the flood labels themselves (the renumbering is a bijection)
    TISEG_LAUNCH(c, k_first_bits, dim3(8, N), 256, 0, g, first, KS, nmark, fbits);
    TISEG_TRY(rank_from_bits(c, g, fbits, rank, nullptr));
    TISEG_LAUNCH(c, k_arrange_lut, dim3(8, N), 256, 0, g, first, rank, KS, nmark, lut);
    TISEG_LAUNCH(c, k_wsl_remove<false>, strip_grid(g), TISEG_THREADS, 0, g, wsl, lut, KS, inst);
    //   any other background: the general relabelling, on the listed tiles only (no blocks do anything otherwise)
    Geom gl = listed_geom(g, flagged, nflagged);
    TISEG_TRY(ccl_build(c, gl, ImgEqI32TileBg{wsl, bg}, 2, par));
    TISEG_TRY(rank_roots(c, gl, par, rank, nullptr));
    TISEG_TRY(apply_rank(c, gl, par, rank, arranged));
    TISEG_LAUNCH(c, k_wsl_remove<true>, strip_grid(gl), TISEG_THREADS, 0, gl, arranged, (const int*)nullptr, 0, inst);
    return TISEG_OK;
}

}  // namespace tiseg

using namespace tiseg;

extern "C" int tiseg_postproc_dist_lambda(tiseg_ctx* c, const float* dist, int N, int H, int W, int lamb,
                                          int32_t* inst_out, int32_t* markers_out, int32_t* ws_out);

extern "C" int tiseg_postproc_dist(tiseg_ctx* c, const float* dist, int N, int H, int W, int32_t* inst_out,
                                   int32_t* markers_out, int32_t* ws_out) {
    return tiseg_postproc_dist_lambda(c, dist, N, H, W, 0, inst_out, markers_out, ws_out);
}

extern "C" int tiseg_postproc_dist_lambda(tiseg_ctx* c, const float* dist, int N, int H, int W, int lamb,
                                          int32_t* inst_out, int32_t* markers_out, int32_t* ws_out) {
    if (!c || !dist || !inst_out || lamb < 0 || lamb > 255) { set_error("tiseg_postproc_dist: bad argument (0 <= lambda <= 255)"); return TISEG_ERR_ARG; }
    TISEG_TRY(check_geom(N, H, W));
    begin_call(c);
    Geom g = make_geom(N, H, W);
    size_t total = (size_t)N * g.P;
    const float* d_dist = in(c, dist, total);
    int32_t* d_inst = tiseg::out(c, inst_out, total);
    int32_t* d_mk = markers_out ? tiseg::out(c, markers_out, total) : nullptr;
    int32_t* d_ws = ws_out ? tiseg::out(c, ws_out, total) : nullptr;
    if (!d_dist || !d_inst) return TISEG_ERR_CUDA;
    TISEG_TRY(postproc_dist_dev(c, g, d_dist, lamb, d_inst, d_mk, d_ws));
    return end_call(c);
}
