// K12 / K13 — evaluation kernels: GT x pred pair-overlap histogram -> AJI and PQ, semantic counts.
//
//   tiseg_pair_metrics_bin  replaces pre_eval_bin_aji + pre_eval_bin_pq (tiseg/utils/inst_metrics.py:10-92,
//                           138-229) including the measure.label relabelling they start with.
//   tiseg_sem_counts        replaces pre_eval_all_semantic_metric (tiseg/utils/sem_metrics.py:16-53).
//
// The reference materialises one full-image mask per instance (O(K*P)); here the non-zero entries of the
// [Ng, Np] intersection matrix are accumulated in one pass into a per-tile open-addressing hash table keyed
// by (gt id, pred id) — one atomic per horizontal RUN of equal pairs (warp ballot), not per pixel — and
// everything downstream works on the O(K) non-zero pairs.  Counts are exact integers; the only floating
// point is the fp64 IoU used for argmax / thresholding, evaluated with the reference's expressions.
#include "ccl.cuh"

namespace tiseg {

struct PairTab {
    unsigned long long* key;   // [N, cap]  (g << 32 | p), 0 = empty
    int* cnt;                  // [N, cap]
    int cap;                   // power of two
};

__device__ __forceinline__ unsigned pair_hash(unsigned g, unsigned p) {
    unsigned h = g * 0x9E3779B1u ^ p * 0x85EBCA6Bu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}

__device__ __forceinline__ void pair_add(const PairTab& t, int n, unsigned g, unsigned p, int len, int* overflow) {
    unsigned long long k = ((unsigned long long)g << 32) | p;
    unsigned long long* keys = t.key + (long long)n * t.cap;
    unsigned s = pair_hash(g, p) & (t.cap - 1);
    for (int probe = 0; probe < t.cap; ++probe) {
        unsigned long long cur = keys[s];
        if (cur == 0) {
            cur = atomicCAS(&keys[s], 0ull, k);
            if (cur == 0) cur = k;                         // we claimed the slot
        }
        if (cur == k) { atomicAdd(&t.cnt[(long long)n * t.cap + s], len); return; }
        s = (s + 1) & (t.cap - 1);
    }
    *overflow = 1;
}

__device__ __forceinline__ int pair_lookup(const PairTab& t, int n, unsigned g, unsigned p) {
    unsigned long long k = ((unsigned long long)g << 32) | p;
    const unsigned long long* keys = t.key + (long long)n * t.cap;
    unsigned s = pair_hash(g, p) & (t.cap - 1);
    for (int probe = 0; probe < t.cap; ++probe) {
        unsigned long long cur = keys[s];
        if (cur == k) return t.cnt[(long long)n * t.cap + s];
        if (cur == 0) return 0;
        s = (s + 1) & (t.cap - 1);
    }
    return 0;
}

// per-instance state, dense by id with a per-tile stride of KS = P + 1 entries (only the first K+1 are touched)
struct InstState {
    int* area_g; int* area_p;          // [N, KS]
    unsigned long long* best;          // [N, KS] fp64 bits of the best AJI IoU per gt id
    int* bestp;                        // [N, KS] lowest pred id reaching `best`
    uint8_t* used;                     // [N, KS] pred id chosen by some gt
    double* pqiou;                     // [N, KS] IoU of the (unique) PQ match of a gt id, 0 = none
    const int* ng; const int* np;      // [N]
    int KS;
};

__global__ void k_inst_init(InstState s) {
    int n = blockIdx.y;
    long long o = (long long)n * s.KS;
    int ng = s.ng[n], np = s.np[n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= ng; i += gridDim.x * blockDim.x) {
        s.area_g[o + i] = 0; s.best[o + i] = 0ull; s.bestp[o + i] = 0x7fffffff; s.pqiou[o + i] = 0.0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= np; i += gridDim.x * blockDim.x) {
        s.area_p[o + i] = 0; s.used[o + i] = 0;
    }
}

// one pass over the two relabelled maps: areas by id + pair table
__global__ void __launch_bounds__(TISEG_THREADS)
k_pair_accumulate(Geom g, const int* __restrict__ par_g, const int* __restrict__ rank_g,
                  const int* __restrict__ par_p, const int* __restrict__ rank_p, InstState s, PairTab t, int* overflow) {
    Pix px;
    if (!warp_pixel(g, px)) return;
    int gid = 0, pid = 0;
    if (px.ok) {
        int a = par_g[px.base + px.idx], b = par_p[px.base + px.idx];
        if (a >= 0) gid = rank_g[px.base + a];
        if (b >= 0) pid = rank_p[px.base + b];
    }
    int gl = __shfl_up_sync(0xffffffffu, gid, 1), pl = __shfl_up_sync(0xffffffffu, pid, 1);
    bool first = px.lane == 0;
    bool cg = !first && gid == gl, cp = !first && pid == pl;
    unsigned mg = __ballot_sync(0xffffffffu, cg);
    unsigned mp = __ballot_sync(0xffffffffu, cp);
    unsigned mb = mg & mp;                                  // both continue => the pair continues
    long long o = (long long)px.n * s.KS;
    if (gid && !cg) atomicAdd(&s.area_g[o + gid], run_end_lane(mg, px.lane) - px.lane + 1);
    if (pid && !cp) atomicAdd(&s.area_p[o + pid], run_end_lane(mp, px.lane) - px.lane + 1);
    if (gid && pid && !(cg && cp)) pair_add(t, px.n, gid, pid, run_end_lane(mb, px.lane) - px.lane + 1, overflow);
}

// pass A over the table: best AJI IoU per gt (atomicMax on fp64 bits: positive doubles order like integers),
// and the PQ matches (IoU > 0.5 is unique per gt and per pred)
__global__ void k_pair_best(PairTab t, InstState s, int* tp) {
    int n = blockIdx.y;
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= t.cap) return;
    unsigned long long k = t.key[(long long)n * t.cap + slot];
    if (!k) return;
    unsigned gid = (unsigned)(k >> 32), pid = (unsigned)k;
    long long o = (long long)n * s.KS;
    double inter = (double)t.cnt[(long long)n * t.cap + slot];
    double tot = (double)s.area_g[o + gid] + (double)s.area_p[o + pid];
    double iou_aji = inter / ((tot - inter) + 1.0e-6);      // inst_metrics.py:69
    atomicMax(&s.best[o + gid], (unsigned long long)__double_as_longlong(iou_aji));
    double iou_pq = inter / (tot - inter);                  // inst_metrics.py:194
    if (iou_pq > 0.5) { s.pqiou[o + gid] = iou_pq; atomicAdd(&tp[n], 1); }
}

// pass B: np.argmax tie rule — the lowest pred id among those reaching the best IoU (inst_metrics.py:74)
__global__ void k_pair_argbest(PairTab t, InstState s) {
    int n = blockIdx.y;
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= t.cap) return;
    unsigned long long k = t.key[(long long)n * t.cap + slot];
    if (!k) return;
    unsigned gid = (unsigned)(k >> 32), pid = (unsigned)k;
    long long o = (long long)n * s.KS;
    double inter = (double)t.cnt[(long long)n * t.cap + slot];
    double tot = (double)s.area_g[o + gid] + (double)s.area_p[o + pid];
    double iou_aji = inter / ((tot - inter) + 1.0e-6);
    if ((unsigned long long)__double_as_longlong(iou_aji) == s.best[o + gid]) atomicMin(&s.bestp[o + gid], (int)pid);
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v) {
    __shared__ unsigned long long sh[32];
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0ull;
    if (w == 0) for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;   // valid in thread 0
}

// pass C: per gt — paired inter / union, or its own area when nothing overlaps (inst_metrics.py:76-87)
__global__ void k_aji_gt(PairTab t, InstState s, unsigned long long* IU) {
    int n = blockIdx.y;
    long long o = (long long)n * s.KS;
    int ng = s.ng[n];
    unsigned long long I = 0, U = 0;
    for (int gid = 1 + blockIdx.x * blockDim.x + threadIdx.x; gid <= ng; gid += gridDim.x * blockDim.x) {
        int ag = s.area_g[o + gid];
        if (s.best[o + gid] != 0ull) {
            int pid = s.bestp[o + gid];
            int inter = pair_lookup(t, n, gid, pid);
            I += inter;
            U += (unsigned long long)(ag + s.area_p[o + pid] - inter);
            s.used[o + pid] = 1;
        } else {
            U += ag;
        }
    }
    I = block_sum_u64(I);
    U = block_sum_u64(U);
    if (threadIdx.x == 0) { if (I) atomicAdd(&IU[2 * n], I); if (U) atomicAdd(&IU[2 * n + 1], U); }
}

// pass D: preds never chosen by any gt add their area to the union (inst_metrics.py:88-90)
__global__ void k_aji_pred(InstState s, unsigned long long* IU) {
    int n = blockIdx.y;
    long long o = (long long)n * s.KS;
    int np = s.np[n];
    unsigned long long U = 0;
    for (int pid = 1 + blockIdx.x * blockDim.x + threadIdx.x; pid <= np; pid += gridDim.x * blockDim.x)
        if (!s.used[o + pid]) U += s.area_p[o + pid];
    U = block_sum_u64(U);
    if (threadIdx.x == 0 && U) atomicAdd(&IU[2 * n + 1], U);
}

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum): what
// `paired_iou.sum()` (inst_metrics.py:227) evaluates, reproduced so iou_sum is bit-identical.
__device__ __forceinline__ double np_pairwise_leaf(const double* a, int n) {   // n <= 128
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
// the recursion "n2 = n/2 - (n/2)%8; sum(a, n2) + sum(a+n2, n-n2)" unrolled onto an explicit stack (device
// recursion would need a run-time stack size)
__device__ double np_pairwise_sum(const double* a, int n) {
    struct Frame { int off, n, state; double left; };
    Frame st[32];
    int sp = 0;
    double ret = 0.0;
    st[sp++] = Frame{0, n, 0, 0.0};
    while (sp) {
        Frame& f = st[sp - 1];
        if (f.n <= 128) { ret = np_pairwise_leaf(a + f.off, f.n); --sp; continue; }
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; st[sp++] = Frame{f.off, n2, 0, 0.0}; continue; }
        if (f.state == 1) { f.left = ret; f.state = 2; st[sp++] = Frame{f.off + n2, f.n - n2, 0, 0.0}; continue; }
        ret = f.left + ret;
        --sp;
    }
    return ret;
}

// one warp per tile: compact the PQ match IoUs in gt-id order (== row-major order of np.nonzero), sum them
// the numpy way, and write the two result records
__global__ void k_metrics_final(InstState s, const unsigned long long* IU, const int* tp, double* scratch,
                                double* aji, double* pq) {
    int n = blockIdx.x;
    int lane = threadIdx.x;
    long long o = (long long)n * s.KS;
    int ng = s.ng[n], np = s.np[n];
    double* buf = scratch + o;
    int m = 0;
    for (int base = 1; base <= ng; base += 32) {
        int gid = base + lane;
        double v = gid <= ng ? s.pqiou[o + gid] : 0.0;
        unsigned b = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) buf[m + __popc(b & ((1u << lane) - 1))] = v;
        m += __popc(b);
    }
    __syncwarp();
    if (lane == 0) {
        if (aji) {
            bool empty = ng == 0 || np == 0;                 // inst_metrics.py:72-73: (0., 0.) and nothing else
            aji[2 * n] = empty ? 0.0 : (double)IU[2 * n];
            aji[2 * n + 1] = empty ? 0.0 : (double)IU[2 * n + 1];
        }
        if (pq) {
            int t = tp[n];
            pq[4 * n] = t; pq[4 * n + 1] = np - t; pq[4 * n + 2] = ng - t;
            pq[4 * n + 3] = np_pairwise_sum(buf, m);
        }
    }
}

// ---- K13 semantic counts ---------------------------------------------------------------------------
// counts[n, k, c], k = TP, FP, FN, Pred, GT.  Shared-memory histogram per block, one global atomic per
// non-zero bin per block.  Classes outside [0, C) are not counted (torch.histc range semantics).
#define SEM_MAXC 64
__global__ void __launch_bounds__(TISEG_THREADS)
k_sem_counts(Geom g, const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, int C, int ignore,
             unsigned long long* counts, unsigned long long* valid) {
    __shared__ unsigned h[5 * SEM_MAXC];
    __shared__ unsigned nvalid;
    for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) h[i] = 0;
    if (threadIdx.x == 0) nvalid = 0;
    __syncthreads();
    Pix px;
    bool act = warp_pixel(g, px) && px.ok;
    bool v = false;
    if (act) {
        int p = pred[px.base + px.idx], t = gt[px.base + px.idx];
        if (t != ignore) {
            v = true;
            bool pin = p < C, tin = t < C;
            if (p == t) { if (tin) atomicAdd(&h[0 * C + t], 1u); }
            else { if (pin) atomicAdd(&h[1 * C + p], 1u); if (tin) atomicAdd(&h[2 * C + t], 1u); }
            if (pin) atomicAdd(&h[3 * C + p], 1u);
            if (tin) atomicAdd(&h[4 * C + t], 1u);
        }
    }
    unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&nvalid, __popc(b));
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * C; i += blockDim.x)
        if (h[i]) atomicAdd(&counts[(long long)blockIdx.y * 5 * C + i], (unsigned long long)h[i]);
    if (threadIdx.x == 0 && nvalid) atomicAdd(&valid[blockIdx.y], (unsigned long long)nvalid);
}

static int next_pow2(long long v) { int p = 1; while (p < v) p <<= 1; return p; }

// pair metrics on already-flattened forests + ranks (shared with the multi-class path)
int pair_metrics_core(tiseg_ctx* c, const Geom& g, const int* par_g, const int* rank_g, const int* ng,
                      const int* par_p, const int* rank_p, const int* np, double* d_aji, double* d_pq) {
    int N = g.N, KS = g.P + 1;
    size_t ks = (size_t)N * KS;
    InstState s;
    s.area_g = ws<int>(c, ks); s.area_p = ws<int>(c, ks);
    s.best = ws<unsigned long long>(c, ks); s.bestp = ws<int>(c, ks);
    s.used = ws<uint8_t>(c, ks); s.pqiou = ws<double>(c, ks);
    double* scratch = ws<double>(c, ks);
    s.ng = ng; s.np = np; s.KS = KS;
    unsigned long long* IU = ws<unsigned long long>(c, 2 * (size_t)N);
    int* tp = ws<int>(c, (size_t)N);
    int* overflow = ws<int>(c, 1);
    if (!s.area_g || !s.area_p || !s.best || !s.bestp || !s.used || !s.pqiou || !scratch || !IU || !tp || !overflow)
        return TISEG_ERR_CUDA;
    // the table starts small (instances are compact: O(K) pairs) and is retried at the always-sufficient
    // size 2P if a pathological input overflows it
    int cap = next_pow2(g.P / 16 < 1024 ? 1024 : g.P / 16);
    for (int attempt = 0; attempt < 2; ++attempt) {
        PairTab t;
        t.cap = cap;
        t.key = ws<unsigned long long>(c, (size_t)N * cap);
        t.cnt = ws<int>(c, (size_t)N * cap);
        if (!t.key || !t.cnt) return TISEG_ERR_CUDA;
        TISEG_TRY(zero(c, t.key, (size_t)N * cap * sizeof(unsigned long long)));
        TISEG_TRY(zero(c, t.cnt, (size_t)N * cap * sizeof(int)));
        TISEG_TRY(zero(c, IU, 2 * (size_t)N * sizeof(unsigned long long)));
        TISEG_TRY(zero(c, tp, (size_t)N * sizeof(int)));
        TISEG_TRY(zero(c, overflow, sizeof(int)));
        TISEG_LAUNCH(c, k_inst_init, dim3(16, N), 256, 0, s);
        TISEG_LAUNCH(c, k_pair_accumulate, warp_grid(g), TISEG_THREADS, 0, g, par_g, rank_g, par_p, rank_p, s, t, overflow);
        int hov = 0;
        TISEG_CHECK(cudaMemcpyAsync(&hov, overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TISEG_CHECK(cudaStreamSynchronize(c->stream));
        if (hov) {
            if (attempt == 1) { set_error("pair table overflow"); return TISEG_ERR_LIMIT; }
            cap = next_pow2(2ll * g.P);
            continue;
        }
        dim3 tg((cap + 255) / 256, N);
        TISEG_LAUNCH(c, k_pair_best, tg, 256, 0, t, s, tp);
        TISEG_LAUNCH(c, k_pair_argbest, tg, 256, 0, t, s);
        TISEG_LAUNCH(c, k_aji_gt, dim3(8, N), 256, 0, t, s, IU);
        TISEG_LAUNCH(c, k_aji_pred, dim3(8, N), 256, 0, s, IU);
        TISEG_LAUNCH(c, k_metrics_final, N, 32, 0, s, IU, tp, scratch, d_aji, d_pq);
        break;
    }
    return TISEG_OK;
}

}  // namespace tiseg

using namespace tiseg;

extern "C" {

int tiseg_pair_metrics_bin(tiseg_ctx* c, const int32_t* pred, const int32_t* gt, int N, int H, int W,
                           double* aji, double* pq) {
    if (!c || !pred || !gt) { set_error("tiseg_pair_metrics_bin: bad argument"); return TISEG_ERR_ARG; }
    TISEG_TRY(check_geom(N, H, W));
    begin_call(c);
    Geom g = make_geom(N, H, W);
    size_t total = (size_t)N * g.P;
    const int32_t* d_pred = in(c, pred, total);
    const int32_t* d_gt = in(c, gt, total);
    double* d_aji = aji ? tiseg::out(c, aji, 2 * (size_t)N) : nullptr;
    double* d_pq = pq ? tiseg::out(c, pq, 4 * (size_t)N) : nullptr;
    int* par_g = ws<int>(c, total); int* rank_g = ws<int>(c, total);
    int* par_p = ws<int>(c, total); int* rank_p = ws<int>(c, total);
    int* ng = ws<int>(c, (size_t)N); int* np = ws<int>(c, (size_t)N);
    if (!d_pred || !d_gt || !par_g || !rank_g || !par_p || !rank_p || !ng || !np) return TISEG_ERR_CUDA;
    // measure.label(inst.copy()) on both maps (inst_metrics.py:12-13): equal-value, 8-connected, background 0
    TISEG_TRY(ccl_build(c, g, ImgEqI32{d_gt, 0}, 2, par_g));
    TISEG_TRY(rank_roots(c, g, par_g, rank_g, ng));
    TISEG_TRY(ccl_build(c, g, ImgEqI32{d_pred, 0}, 2, par_p));
    TISEG_TRY(rank_roots(c, g, par_p, rank_p, np));
    TISEG_TRY(pair_metrics_core(c, g, par_g, rank_g, ng, par_p, rank_p, np, d_aji, d_pq));
    return end_call(c);
}

int tiseg_sem_counts(tiseg_ctx* c, const uint8_t* pred, const uint8_t* gt, int N, int H, int W, int C,
                     int ignore_index, int64_t* counts, int64_t* valid) {
    if (!c || !pred || !gt || !counts || !valid || C <= 0 || C > SEM_MAXC) { set_error("tiseg_sem_counts: bad argument (C <= 64)"); return TISEG_ERR_ARG; }
    TISEG_TRY(check_geom(N, H, W));
    begin_call(c);
    Geom g = make_geom(N, H, W);
    size_t total = (size_t)N * g.P;
    const uint8_t* d_pred = in(c, pred, total);
    const uint8_t* d_gt = in(c, gt, total);
    int64_t* d_counts = tiseg::out(c, counts, (size_t)N * 5 * C);
    int64_t* d_valid = tiseg::out(c, valid, (size_t)N);
    if (!d_pred || !d_gt || !d_counts || !d_valid) return TISEG_ERR_CUDA;
    TISEG_TRY(zero(c, d_counts, (size_t)N * 5 * C * sizeof(int64_t)));
    TISEG_TRY(zero(c, d_valid, (size_t)N * sizeof(int64_t)));
    TISEG_LAUNCH(c, k_sem_counts, warp_grid(g), TISEG_THREADS, 0, g, d_pred, d_gt, C, ignore_index,
                 (unsigned long long*)d_counts, (unsigned long long*)d_valid);
    return end_call(c);
}

}  // extern "C"
