// K6 — ordered marker-controlled watershed (skimage.segmentation.watershed, connectivity 1, compactness 0,
// no lines; dist.py:124, hovernet.py:361).
//
// The library's flood is a sequential priority queue ordered by (value, age): every marker pixel enters with
// age 0 (raster order), the minimum is popped, and each still-unlabelled in-mask 4-neighbour (up, left,
// right, down) is labelled AT PUSH TIME with the popped pixel's label and pushed with the next age.  A
// level-synchronous flood is not equivalent (plateau ties are decided by age), so the order is reproduced
// exactly — but per BLOB: 4-connected components of the mask never interact, and the restriction of the
// global order to one blob is the order of that blob flooded alone.  One warp owns one blob at a time
// (dynamic work queue per tile), so parallelism comes from the ~10^3 blobs per 1000^2 tile times the tiles
// of the batch.
//   uint8 images (DIST):   256 FIFO buckets (head/tail in shared memory, links in a global `next` array)
//                          reproduce the (value, age) heap order exactly.
//   fp64 images (HoVer):   per-blob binary heap keyed (value, age) in a global arena slice sized by the
//                          blob's area.
#pragma once
#include "ccl.cuh"

namespace tiseg {

struct BlobInfo {
    int* root;   // [N, KS] flat index of the blob's first pixel (its ymin = root / W)
    int* ymax;   // [N, KS]
    int* xmin;   // [N, KS]
    int* xmax;   // [N, KS]
    int* area;   // [N, KS]
    int* off;    // [N, KS] exclusive prefix of areas (fp64 heap slices)
    const int* count;  // [N] number of blobs
    int KS;
};

// mask functor -> flattened blob forest `par`, blob ids `rank` (at roots), BlobInfo
template <class MaskImg>
int blobs_build(tiseg_ctx* c, const Geom& g, MaskImg mask, int* par, int* rank, BlobInfo& b, bool want_offsets,
                int conn = 1);

int watershed_u8_dev(tiseg_ctx* c, const Geom& g, const uint8_t* image, const int* par, const int* rank,
                     const BlobInfo& b, int32_t* out);
int watershed_f64_dev(tiseg_ctx* c, const Geom& g, const double* image, const int* par, const int* rank,
                      const BlobInfo& b, int32_t* out);

// out = markers where the mask forest has a pixel, else 0 (skimage drops markers outside the mask)
int ws_seed(tiseg_ctx* c, const Geom& g, const int32_t* markers, const int* par, int32_t* out);
// root / bbox / area (/ offsets) of every blob of a flattened + ranked forest
int blobs_describe(tiseg_ctx* c, const Geom& g, const int* par, const int* rank, const unsigned* root_bits, BlobInfo& b,
                   bool want_offsets);

#ifdef __CUDACC__

template <class MaskImg>
int blobs_build(tiseg_ctx* c, const Geom& g, MaskImg mask, int* par, int* rank, BlobInfo& b, bool want_offsets,
                int conn) {
    int N = g.N, KS = g.P + 1;
    size_t ks = (size_t)N * KS;
    int* count = ws<int>(c, (size_t)N);
    b.root = ws<int>(c, ks); b.ymax = ws<int>(c, ks); b.xmin = ws<int>(c, ks); b.xmax = ws<int>(c, ks);
    b.area = ws<int>(c, ks); b.off = want_offsets ? ws<int>(c, ks) : nullptr;
    b.KS = KS; b.count = count;
    if (!count || !b.root || !b.ymax || !b.xmin || !b.xmax || !b.area || (want_offsets && !b.off)) return TISEG_ERR_CUDA;
    TISEG_TRY(ccl_build(c, g, mask, conn, par));
    const unsigned* root_bits = (const unsigned*)c->rootblk;      // left by the flatten; rank_roots consumes the same bitmap
    TISEG_TRY(rank_roots(c, g, par, rank, count));
    return blobs_describe(c, g, par, rank, root_bits, b, want_offsets);
}
#endif

}  // namespace tiseg
