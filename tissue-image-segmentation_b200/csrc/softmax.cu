// K1 — softmax over channels + TTA mean + argmax in one pass over the logits.
// Replaces tiseg/models/segmentors/base.py:321-339 (F.softmax per variant, sum/len) and the
// `argmax(dim=1)` + `.cpu().numpy()` that every segmentor does next (unet.py:62-64, dist.py:266).
//
// fp32 arithmetic in the same order as the reference expression: e = exp(x - max), s = e_0 + e_1 + ...,
// p = e / s, acc += p per variant, acc / T.  One thread per pixel, channel planes read coalesced.
#include "common.cuh"

namespace tiseg {

template <int CMAX>
__global__ void __launch_bounds__(TISEG_THREADS)
k_softmax_argmax(long long P, const float* __restrict__ logits, int T, int C, float* __restrict__ prob,
                 uint8_t* __restrict__ cls, bool vec) {
    // one thread = 4 consecutive pixels of one tile (blockIdx.y): float4 per channel plane, uchar4 out
    const int n = blockIdx.y;
    const long long i = flat4_index();
    if (i >= P) return;
    float acc[CMAX][4];
    for (int t = 0; t < T; ++t) {
        const float* src = logits + ((long long)n * T + t) * C * P;
        Pack4<float> x[CMAX];
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) {
                x[c] = ld4(src + c * P, i, P, vec);
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], x[c].v[k]);
            }
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { x[c].v[k] = expf(x[c].v[k] - m[k]); s[k] = s[k] + x[c].v[k]; }
            }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[c][k] = (t == 0) ? x[c].v[k] / s[k] : acc[c][k] + x[c].v[k] / s[k];
            }
    }
    const float tf = (float)T;
    Pack4<uint8_t> best;
    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < 4; ++k) best.v[k] = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) {
            Pack4<float> p;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                p.v[k] = acc[c][k] / tf;
                if (p.v[k] > bv[k]) { bv[k] = p.v[k]; best.v[k] = (uint8_t)c; }
            }
            if (prob) st4(prob + ((long long)n * C + c) * P, i, P, vec, p);
        }
    if (cls) st4(cls + (long long)n * P, i, P, vec, best);
}

// One variant and only the class map wanted: softmax is strictly increasing in the logit, so the class is the first
// maximum of the logits themselves (it can differ from the reference's argmax-of-probabilities only where two
// probabilities round to the same fp32 value, the tie band the parity test already excludes).  Pure streaming.
template <int CMAX>
__global__ void __launch_bounds__(TISEG_THREADS)
k_argmax_logits(long long P, const float* __restrict__ logits, int C, uint8_t* __restrict__ cls, bool vec) {
    const int n = blockIdx.y;
    const long long i = flat4_index();
    if (i >= P) return;
    const float* src = logits + (long long)n * C * P;
    Pack4<float> x[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) if (c < C) x[c] = ld4(src + c * P, i, P, vec);
    Pack4<uint8_t> best;
    float bv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { best.v[k] = 0; bv[k] = x[0].v[k]; }
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
        if (c < C) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x[c].v[k] > bv[k]) { bv[k] = x[c].v[k]; best.v[k] = (uint8_t)c; }
        }
    st4(cls + (long long)n * P, i, P, vec, best);
}

int softmax_argmax_dev(tiseg_ctx* c, const Geom& g, const float* d_in, int T, int C, float* d_prob, uint8_t* d_cls) {
    const long long P = g.P;
    const bool vec = (P % 4 == 0) && aligned16(d_in, d_prob) && (((uintptr_t)d_cls) & 3) == 0;
    dim3 grid(flat4_grid(P), (unsigned)g.N);
    if (T == 1 && !d_prob) {
        if (C <= 4) TISEG_LAUNCH(c, k_argmax_logits<4>, grid, TISEG_THREADS, 0, P, d_in, C, d_cls, vec);
        else if (C <= 8) TISEG_LAUNCH(c, k_argmax_logits<8>, grid, TISEG_THREADS, 0, P, d_in, C, d_cls, vec);
        else TISEG_LAUNCH(c, k_argmax_logits<16>, grid, TISEG_THREADS, 0, P, d_in, C, d_cls, vec);
        return TISEG_OK;
    }
    if (C <= 4) TISEG_LAUNCH(c, k_softmax_argmax<4>, grid, TISEG_THREADS, 0, P, d_in, T, C, d_prob, d_cls, vec);
    else if (C <= 8) TISEG_LAUNCH(c, k_softmax_argmax<8>, grid, TISEG_THREADS, 0, P, d_in, T, C, d_prob, d_cls, vec);
    else TISEG_LAUNCH(c, k_softmax_argmax<16>, grid, TISEG_THREADS, 0, P, d_in, T, C, d_prob, d_cls, vec);
    return TISEG_OK;
}

}  // namespace tiseg

using namespace tiseg;

extern "C" int tiseg_softmax_argmax(tiseg_ctx* c, const float* logits, int N, int T, int C, int H, int W,
                                    float* prob, uint8_t* cls) {
    if (!c || !logits || T <= 0 || C <= 0 || C > 16 || (!prob && !cls)) {
        set_error("tiseg_softmax_argmax: bad argument (1 <= C <= 16)");
        return TISEG_ERR_ARG;
    }
    TISEG_TRY(check_geom(N, H, W));
    begin_call(c);
    Geom g = make_geom(N, H, W);
    size_t total = (size_t)N * g.P;
    const float* d_in = in(c, logits, total * T * C);
    float* d_prob = prob ? tiseg::out(c, prob, total * C) : nullptr;
    uint8_t* d_cls = cls ? tiseg::out(c, cls, total) : nullptr;
    if (!d_in) return TISEG_ERR_CUDA;
    TISEG_TRY(softmax_argmax_dev(c, g, d_in, T, C, d_prob, d_cls));
    return end_call(c);
}
