// internal (device-pointer) entry points of morph.cu shared with the other pipelines
#pragma once
#include "common.cuh"

namespace tiseg {
// out = complement-forest background components that do not touch the border, plus the foreground
int fill_from_complement_forest(tiseg_ctx* c, const Geom& g, const int* par, uint8_t* out);
int remove_small_mask(tiseg_ctx* c, const Geom& g, const uint8_t* mask, int min_size, int conn, uint8_t* out);
int remove_small_labels(tiseg_ctx* c, const Geom& g, const int32_t* lab, int min_size, int32_t* out);
int grey_morph(tiseg_ctx* c, const Geom& g, const int32_t* lab, int footprint, int radius, bool dilate, int32_t* out);
int postproc_unet_dev(tiseg_ctx* c, const Geom& g, uint8_t* cls, int max_class, int radius, int edge_id,
                      const uint8_t* kill, uint8_t* sem, int32_t* inst);
}  // namespace tiseg
