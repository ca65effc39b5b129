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
k_softmax_argmax(Geom g, const float* __restrict__ logits, int T, int C, float* __restrict__ prob,
                 uint8_t* __restrict__ cls) {
    Pix px;
    if (!warp_pixel(g, px) || !px.ok) return;
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    const long long P = g.P;
    for (int t = 0; t < T; ++t) {
        const float* src = logits + ((long long)px.n * T + t) * C * P + px.idx;
        float x[CMAX];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { x[c] = src[c * P]; m = fmaxf(m, x[c]); }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { x[c] = expf(x[c] - m); s = s + x[c]; }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) acc[c] = (t == 0) ? x[c] / s : acc[c] + x[c] / s;
    }
    float tf = (float)T;
    int best = 0;
    float bv = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) {
            float p = acc[c] / tf;
            if (prob) prob[((long long)px.n * C + c) * P + px.idx] = p;
            if (p > bv) { bv = p; best = c; }
        }
    if (cls) cls[px.base + px.idx] = (uint8_t)best;
}

int softmax_argmax_dev(tiseg_ctx* c, const Geom& g, const float* d_in, int T, int C, float* d_prob, uint8_t* d_cls) {
    if (C <= 4) TISEG_LAUNCH(c, k_softmax_argmax<4>, warp_grid(g), TISEG_THREADS, 0, g, d_in, T, C, d_prob, d_cls);
    else if (C <= 8) TISEG_LAUNCH(c, k_softmax_argmax<8>, warp_grid(g), TISEG_THREADS, 0, g, d_in, T, C, d_prob, d_cls);
    else TISEG_LAUNCH(c, k_softmax_argmax<16>, warp_grid(g), TISEG_THREADS, 0, g, d_in, T, C, d_prob, d_cls);
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
