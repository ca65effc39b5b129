/*
 * oracle/skimage_port.c — TEST INFRASTRUCTURE ONLY (CPU oracle, never the product path).
 *
 * Plain-C restatement of the scikit-image 0.18.3 functions that the reference
 * (clownrat6/Tissue-Image-Segmentation, `tiseg`) calls on its test-time path and that are
 * not installable in this image (requirements.txt:12 pins scikit-image==0.18.3; the source is
 * not under /root/reference).  Also the numba kernel `align_foreground`.
 *
 *   sk_label              skimage.measure.label            (unet.py:85, dist.py:107,123,
 *                                                           inst_metrics.py:12-13,142-143)
 *   sk_watershed          skimage.segmentation.watershed   (dist.py:124, hovernet.py:361)
 *   sk_reconstruction_erosion  skimage.morphology.reconstruction(method='erosion') (dist.py:56)
 *   tiseg_align_foreground     tiseg/models/utils/postprocess.py:123-155
 *
 * PARITY UNPINNED at the scikit-image boundary: the reference ships no tests/golden vectors and
 * the wheel is absent, so these follow the library's published algorithm (two-pass union-find
 * with raster-order ids; (value, age) priority flood with 4-neighbour order -W,-1,+1,+W and
 * label-at-push).  The one behaviour that is DEFINED here rather than reproduced is the
 * tie-break among seed pixels with identical (value, age=0): this oracle orders them by flat
 * index (SURVEY.md §8c "Residual risk").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ union-find */
static int64_t uf_find(int64_t *p, int64_t x) {
    int64_t r = x;
    while (p[r] != r) r = p[r];
    while (p[x] != r) { int64_t n = p[x]; p[x] = r; x = n; }
    return r;
}
static void uf_union(int64_t *p, int64_t a, int64_t b) {
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return;
    if (a < b) p[b] = a; else p[a] = b;   /* root = lowest flat index */
}

/*
 * skimage.measure.label(img, background=bg, connectivity=conn) for 2-D integer images.
 * Equal-valued neighbouring non-background pixels are connected; conn=1 -> 4-neighbourhood,
 * conn=2 -> 8-neighbourhood (the library default connectivity=None == ndim == 2).
 * Output ids 1..K numbered in raster order of each component's first pixel. Returns K.
 */
int64_t sk_label(const int64_t *img, int64_t H, int64_t W, int64_t bg, int conn, int64_t *out) {
    int64_t P = H * W;
    int64_t *par = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    for (int64_t i = 0; i < P; ++i) par[i] = i;
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) {
            int64_t i = y * W + x, v = img[i];
            if (v == bg) continue;
            if (x > 0 && img[i - 1] == v) uf_union(par, i, i - 1);
            if (y > 0) {
                if (img[i - W] == v) uf_union(par, i, i - W);
                if (conn == 2) {
                    if (x > 0 && img[i - W - 1] == v) uf_union(par, i, i - W - 1);
                    if (x + 1 < W && img[i - W + 1] == v) uf_union(par, i, i - W + 1);
                }
            }
        }
    int64_t K = 0;
    for (int64_t i = 0; i < P; ++i) {
        if (img[i] == bg) { out[i] = 0; continue; }
        int64_t r = uf_find(par, i);
        if (r == i) out[i] = ++K;      /* root is the first pixel in raster order */
        else out[i] = out[r];
    }
    free(par);
    return K;
}

/* ------------------------------------------------------------------ watershed */
typedef struct { double value; int64_t age; int64_t index; } ws_item;

static int ws_less(const ws_item *a, const ws_item *b) {
    if (a->value != b->value) return a->value < b->value;
    if (a->age != b->age) return a->age < b->age;
    return a->index < b->index;        /* defined tie-break (see header) */
}
typedef struct { ws_item *d; int64_t n, cap; } ws_heap;
static void heap_push(ws_heap *h, ws_item it) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->d = (ws_item *)realloc(h->d, sizeof(ws_item) * (size_t)h->cap); }
    int64_t c = h->n++;
    h->d[c] = it;
    while (c > 0) {
        int64_t p = (c - 1) / 2;
        if (ws_less(&h->d[c], &h->d[p])) { ws_item t = h->d[c]; h->d[c] = h->d[p]; h->d[p] = t; c = p; }
        else break;
    }
}
static ws_item heap_pop(ws_heap *h) {
    ws_item top = h->d[0];
    h->n--;
    if (h->n > 0) {
        h->d[0] = h->d[h->n];
        int64_t i = 0;
        for (;;) {
            int64_t l = 2 * i + 1, r = l + 1, s = i;
            if (l < h->n && ws_less(&h->d[l], &h->d[s])) s = l;
            if (r < h->n && ws_less(&h->d[r], &h->d[s])) s = r;
            if (s == i) break;
            ws_item t = h->d[i]; h->d[i] = h->d[s]; h->d[s] = t; i = s;
        }
    }
    return top;
}

/*
 * skimage.segmentation.watershed(image, markers, mask=mask) with the defaults the reference
 * uses: connectivity=1, compactness=0, watershed_line=False.
 * image is compared as float64; markers are multiplied by mask; every marker pixel is pushed
 * (value, age=0); pop min (value, age); each unlabelled in-mask 4-neighbour (order up, left,
 * right, down) is labelled when pushed with (image[n], ++age).
 */
void sk_watershed(const double *image, const int32_t *markers, const uint8_t *mask,
                  int64_t H, int64_t W, int32_t *out) {
    int64_t P = H * W;
    ws_heap h = {0, 0, 0};
    for (int64_t i = 0; i < P; ++i) {
        out[i] = mask[i] ? markers[i] : 0;
        if (out[i]) { ws_item it = { image[i], 0, i }; heap_push(&h, it); }
    }
    int64_t age = 1;  /* library starts the counter at 1 and pre-increments: first push has age 2;
                         only the relative order matters */
    while (h.n > 0) {
        ws_item e = heap_pop(&h);
        int64_t y = e.index / W, x = e.index % W;
        int32_t lab = out[e.index];
        int64_t nb[4]; int ok[4];
        nb[0] = e.index - W; ok[0] = y > 0;
        nb[1] = e.index - 1; ok[1] = x > 0;
        nb[2] = e.index + 1; ok[2] = x + 1 < W;
        nb[3] = e.index + W; ok[3] = y + 1 < H;
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
            int64_t n = nb[k];
            if (!mask[n] || out[n]) continue;
            age++;
            out[n] = lab;
            ws_item it = { image[n], age, n };
            heap_push(&h, it);
        }
    }
    free(h.d);
}

/* ------------------------------------------------------------------ reconstruction */
/*
 * skimage.morphology.reconstruction(seed, mask, method='erosion') with the default 3x3
 * footprint: iterate R <- max(erode3x3(R), mask) from R = seed (seed >= mask) to the fixed
 * point.  Computed here with in-place raster / anti-raster sweeps until stable (in-place
 * updates only speed up convergence to the same fixed point); values are doubles (the
 * library returns float64).
 */
void sk_reconstruction_erosion(const double *seed, const double *mask, int64_t H, int64_t W, double *out) {
    int64_t P = H * W;
    memcpy(out, seed, sizeof(double) * (size_t)P);
    int changed = 1;
    /* plain sweeps until stable: simple and obviously correct (oracle, not product) */
    while (changed) {
        changed = 0;
        for (int64_t y = 0; y < H; ++y)
            for (int64_t x = 0; x < W; ++x) {
                int64_t i = y * W + x; double m = out[i];
                for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
                    int64_t yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    double v = out[yy * W + xx]; if (v < m) m = v;
                }
                if (m < mask[i]) m = mask[i];
                if (m != out[i]) { out[i] = m; changed = 1; }
            }
        for (int64_t y = H - 1; y >= 0; --y)
            for (int64_t x = W - 1; x >= 0; --x) {
                int64_t i = y * W + x; double m = out[i];
                for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
                    int64_t yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    double v = out[yy * W + xx]; if (v < m) m = v;
                }
                if (m < mask[i]) m = mask[i];
                if (m != out[i]) { out[i] = m; changed = 1; }
            }
    }
}

/* ------------------------------------------------------------------ align_foreground */
/*
 * tiseg/models/utils/postprocess.py:123-155 (numba).  Ordered multi-source BFS: seeds are all
 * pixels with pred>0 in raster order; at most `time`-1 rounds; a popped pixel claims its
 * still-zero foreground 8-neighbours in the order k=1..8 of (dirx,diry); mutates pred.
 */
void tiseg_align_foreground(int64_t *pred, const uint8_t *fg, int64_t H, int64_t W, int time) {
    static const int dirx[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};
    static const int diry[9] = {0, -1, -1, 0, 1, 1, 1, 0, -1};
    int64_t P = H * W;
    int64_t *Q = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    int64_t *NQ = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    int64_t nq = 0;
    for (int64_t i = 0; i < P; ++i) if (pred[i] > 0) Q[nq++] = i;
    int iter = 1;
    while (nq > 0) {
        int64_t nn = 0;
        if (iter >= time) break;
        iter++;
        for (int64_t ix = 0; ix < nq; ++ix) {
            int64_t x = Q[ix] / W, y = Q[ix] % W;
            for (int k = 1; k < 9; ++k) {
                int64_t nx = x + dirx[k], ny = y + diry[k];
                if (nx >= 0 && nx < H && ny >= 0 && ny < W) {
                    int64_t n = nx * W + ny;
                    if (pred[n] == 0 && fg[n] > 0) { NQ[nn++] = n; pred[n] = pred[Q[ix]]; }
                }
            }
        }
        int64_t *t = Q; Q = NQ; NQ = t; nq = nn;
    }
    free(Q); free(NQ);
}
