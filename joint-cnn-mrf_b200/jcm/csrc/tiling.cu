// Host-side search for mixed-shape pixel-tile plans (see tiling.cuh) + a debug export for the CPU tests.
#include "common.cuh"
#include "tiling.cuh"

namespace {

struct Builder {
  TilePlan p;
  int max_shapes;
  bool ok;
  void reset(int ms) { p.n_tiles = 0; p.n_shapes = 0; max_shapes = ms; ok = true; }
  int shape_id(int w, int h) {
    for (int i = 0; i < p.n_shapes; ++i)
      if (p.sw[i] == w && p.sh[i] == h) return i;
    if (p.n_shapes >= max_shapes || p.n_shapes >= kPlanMaxShapes || w > 255 || h > 255 || w <= 0 || h <= 0) { ok = false; return 0; }
    p.sw[p.n_shapes] = (uint8_t)w;
    p.sh[p.n_shapes] = (uint8_t)h;
    return p.n_shapes++;
  }
  void add(int x, int y, int w, int h) {
    if (!ok) return;
    const int s = shape_id(w, h);
    if (!ok) return;
    if (p.n_tiles >= kPlanMaxTiles || x > 255 || y > 255) { ok = false; return; }
    p.x0[p.n_tiles] = (uint8_t)x;
    p.y0[p.n_tiles] = (uint8_t)y;
    p.shape[p.n_tiles] = (uint8_t)s;
    ++p.n_tiles;
  }
};

// Tiles the strip [xs, xs+sw) x [ys, ys+sh) cut into columns of power-of-two width (largest first); a column of width wc is covered
// by boxes (wc, cap / wc) stacked along y.  The last box of a column may hang over the strip's lower end only if that is the map's
// bottom edge (free_end); otherwise it is clipped (exact_px: not allowed -> plan invalid).
void tile_columns(Builder& b, int xs, int ys, int sw, int sh, int cap, bool free_end, bool exact_px) {
  int x = xs;
  for (int wc = cap; wc >= 1 && b.ok; wc >>= 1) {
    if (!(sw & wc) || wc > sw) continue;
    const int bh = cap / wc;
    for (int y = 0; y < sh && b.ok; y += bh) {
      int h = bh;
      if (y + bh > sh && !free_end) {
        if (exact_px) { b.ok = false; return; }
        h = sh - y;
      }
      b.add(x, ys + y, wc, h);
    }
    x += wc;
  }
}
// the same along x, for a strip cut into rows of power-of-two height
void tile_rows(Builder& b, int xs, int ys, int sw, int sh, int cap, bool free_end, bool exact_px) {
  int y = ys;
  for (int hr = cap; hr >= 1 && b.ok; hr >>= 1) {
    if (!(sh & hr) || hr > sh) continue;
    const int bw = cap / hr;
    for (int x = 0; x < sw && b.ok; x += bw) {
      int w = bw;
      if (x + bw > sw && !free_end) {
        if (exact_px) { b.ok = false; return; }
        w = sw - x;
      }
      b.add(xs + x, y, w, hr);
    }
    y += hr;
  }
}

}  // namespace

bool plan_tiles(int H, int W, int cap, bool exact_px, int max_shapes, int uniform_tiles, TilePlan* out) {
  if (H <= 0 || W <= 0 || H > 255 || W > 255) return false;
  bool found = false;
  int best_tiles = uniform_tiles, best_shapes = 1 << 30;
  Builder b;
  for (int tw = cap; tw >= 1; tw >>= 1) {
    const int th = cap / tw;
    const int nx = W / tw, ny = H / th;
    const int rw = W - nx * tw, rh = H - ny * th;
    for (int layout = 0; layout < 2; ++layout) {
      b.reset(max_shapes);
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) b.add(i * tw, j * th, tw, th);
      if (layout == 0) {
        // right strip over the full height (its lower end is the map's bottom edge), bottom strip under the main block only
        if (rw) tile_columns(b, nx * tw, 0, rw, H, cap, true, exact_px);
        if (rh && nx) tile_rows(b, 0, ny * th, nx * tw, rh, cap, rw == 0, exact_px);
      } else {
        // bottom strip over the full width (its right end is the map's right edge), right strip beside the main block only
        if (rh) tile_rows(b, 0, ny * th, W, rh, cap, true, exact_px);
        if (rw && ny) tile_columns(b, nx * tw, 0, rw, ny * th, cap, rh == 0, exact_px);
      }
      if (!b.ok || b.p.n_tiles == 0) continue;
      if (b.p.n_tiles < best_tiles || (found && b.p.n_tiles == best_tiles && b.p.n_shapes < best_shapes)) {
        best_tiles = b.p.n_tiles;
        best_shapes = b.p.n_shapes;
        *out = b.p;
        found = true;
      }
    }
  }
  return found;
}

// Test support: out = [n_tiles, n_shapes, then per tile x0, y0, box_w, box_h]; returns the number of ints written, 0 when no mixed plan
// beats `uniform_tiles`, < 0 when `max_out` is too small.
extern "C" int jcm_debug_tile_plan(int H, int W, int cap, int exact_px, int max_shapes, int uniform_tiles, int* out, int max_out) {
  TilePlan p;
  if (!plan_tiles(H, W, cap, exact_px != 0, max_shapes, uniform_tiles, &p)) return 0;
  const int need = 2 + 4 * p.n_tiles;
  if (!out || max_out < need) return -need;
  out[0] = p.n_tiles;
  out[1] = p.n_shapes;
  for (int t = 0; t < p.n_tiles; ++t) {
    out[2 + 4 * t + 0] = p.x0[t];
    out[2 + 4 * t + 1] = p.y0[t];
    out[2 + 4 * t + 2] = p.sw[p.shape[t]];
    out[2 + 4 * t + 3] = p.sh[p.shape[t]];
  }
  return need;
}
