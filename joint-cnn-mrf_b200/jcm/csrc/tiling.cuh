// Pixel-tile plans for the tensor-core convolution kernels (host side).
//
// An implicit-GEMM M tile (conv_igemm_kernel: 128 accumulator rows) or a weight-gradient k-block (conv_wgrad_kernel: 64 pixels) is a
// rectangular patch of one image, fetched as ONE TMA box.  A uniform grid of TW x TH patches over the part detector's 60x90 maps
// (84 % of its FLOPs) needs 45 patches of 128 pixels for 5400 pixels: 6.7 % of the MMAs multiply padding (13.8 % on 30x45).  No single
// rectangle of 128 pixels tiles 90 x 60 better, but a MIX does: 16x8 patches over [0,80) x [0,56), an 8-wide and a 2-wide column of
// 8x16 / 2x64 patches for x in [80,90), 32x4 patches for the last four rows - 43 patches, the lower bound ceil(5400 / 128).
//
// plan_tiles() searches that family: a main block of a power-of-two patch shape, the rest as a right strip and a bottom strip that
// are cut into columns / rows of power-of-two width / height, each tiled by patches of exactly `cap` pixels along its long side.
// A patch may hang over the RIGHT or BOTTOM edge of the map (TMA zero-fills / clips out-of-bounds elements) but never over another
// patch: every pixel belongs to exactly one patch (tests/test_cpu_suite.py checks that through jcm_debug_tile_plan).
//   exact_px = true  (weight gradient: the patch pixels are the contraction dimension, so a partial box would leave stale rows in
//                     the shared-memory operand): every box has exactly `cap` pixels;
//   exact_px = false (forward / data gradient: stale rows only produce accumulator rows that are never stored): a patch that would
//                     run into a neighbouring strip is clipped to a smaller box.
#pragma once
#include <stdint.h>

constexpr int kPlanMaxTiles = 96;
constexpr int kPlanMaxShapes = 8;

struct TilePlan {
  int n_tiles, n_shapes;
  uint8_t x0[kPlanMaxTiles], y0[kPlanMaxTiles], shape[kPlanMaxTiles];
  uint8_t sw[kPlanMaxShapes], sh[kPlanMaxShapes];   // box width / height per shape
};

// Returns true and fills `out` when a mixed plan with FEWER patches than the best uniform grid exists within the limits
// (kPlanMaxTiles patches, max_shapes shapes, coordinates < 256).  uniform_tiles = patch count of the caller's uniform grid.
bool plan_tiles(int H, int W, int cap, bool exact_px, int max_shapes, int uniform_tiles, TilePlan* out);
