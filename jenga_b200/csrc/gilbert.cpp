// Generalized-Hilbert ("gilbert") token orders and block adjacency, host side.
//
// Produces bit-identical tables to the reference's gilbert.py (gilbert_mapping :442,
// sliced_gilbert_mapping :332, gilbert_block_neighbor_mapping :597,
// sliced_gilbert_block_neighbor_mapping :679) but walks the curve ONCE in curve order
// (O(N)) instead of solving one root-to-leaf index query per voxel (O(N log N) Python calls,
// 1.5 s + 2.1 s at 32x45x80 in the reference): the recursion below enumerates the same
// sub-boxes, in the same order, that gilbert_xyz2d_r (:68-272) accumulates volumes over, so
// the k-th voxel emitted is the voxel whose index is k.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "jenga_b200.h"

namespace jenga {
int set_error(int code, const char* fmt, ...);
}

namespace {

struct Vec {
  int x, y, z;
};
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator-(Vec a) { return {-a.x, -a.y, -a.z}; }
inline int sgn(int v) { return (v > 0) - (v < 0); }
inline Vec unit(Vec a) { return {sgn(a.x), sgn(a.y), sgn(a.z)}; }
inline int extent(Vec a) { return std::abs(a.x + a.y + a.z); }
// Python floor division by two (gilbert.py:91-93 uses //2 on possibly negative components)
inline Vec half(Vec a) { return {a.x >> 1, a.y >> 1, a.z >> 1}; }

struct Walker {
  int64_t next = 0;
  int W = 0, H = 0;  // linear = z*H*W + y*W + x
  int64_t* to_curve = nullptr;    // [linear] -> curve index
  int64_t* from_curve = nullptr;  // [curve index] -> linear

  void emit(Vec p) {
    const int64_t lin = (static_cast<int64_t>(p.z) * H + p.y) * W + p.x;
    if (to_curve) to_curve[lin] = next;
    if (from_curve) from_curve[next] = lin;
    ++next;
  }
  void line(Vec p, Vec step, int n) {
    for (int i = 0; i < n; ++i) {
      emit(p);
      p = p + step;
    }
  }

  // a: major ("right"), b, c: the two orthogonal extents; p: start corner.
  void walk(Vec p, Vec a, Vec b, Vec c) {
    const int w = extent(a), h = extent(b), d = extent(c);
    const Vec da = unit(a), db = unit(b), dc = unit(c);
    if (h == 1 && d == 1) return line(p, da, w);
    if (w == 1 && d == 1) return line(p, db, h);
    if (w == 1 && h == 1) return line(p, dc, d);

    Vec a2 = half(a), b2 = half(b), c2 = half(c);
    const int w2 = extent(a2), h2 = extent(b2), d2 = extent(c2);
    // prefer even steps
    if ((w2 & 1) && w > 2) a2 = a2 + da;
    if ((h2 & 1) && h > 2) b2 = b2 + db;
    if ((d2 & 1) && d > 2) c2 = c2 + dc;

    if (2 * w > 3 * h && 2 * w > 3 * d) {  // wide: split along a only
      walk(p, a2, b, c);
      walk(p + a2, a - a2, b, c);
    } else if (3 * h > 4 * d) {  // do not split along c
      walk(p, b2, c, a2);
      walk(p + b2, a, b - b2, c);
      walk(p + (a - da) + (b2 - db), -b2, c, -(a - a2));
    } else if (3 * d > 4 * h) {  // do not split along b
      walk(p, c2, a2, b);
      walk(p + c2, a, b, c - c2);
      walk(p + (a - da) + (c2 - dc), -c2, -(a - a2), b);
    } else {  // regular: split along all three
      walk(p, b2, c2, a2);
      walk(p + b2, c, a2, b - b2);
      walk(p + (b2 - db) + (c - dc), a, -b2, -(c - c2));
      walk(p + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2);
      walk(p + (a - da) + (b2 - db), -b2, c2, -(a - a2));
    }
  }

  // Entry: the longest side becomes the major axis (gilbert.py:19-38).
  void run(int width, int height, int depth) {
    const Vec o{0, 0, 0}, X{width, 0, 0}, Y{0, height, 0}, Z{0, 0, depth};
    if (width >= height && width >= depth)
      walk(o, X, Y, Z);
    else if (height >= width && height >= depth)
      walk(o, Y, X, Z);
    else
      walk(o, Z, X, Y);
  }
};

int build_mapping(int t, int h, int w, int sliced, int64_t* l2h, int64_t* h2l) {
  if (t <= 0 || h <= 0 || w <= 0) return jenga::set_error(JENGA_E_INVALID, "gilbert: empty grid");
  if (!sliced) {
    Walker wk;
    wk.W = w;
    wk.H = h;
    wk.to_curve = l2h;
    wk.from_curve = h2l;
    wk.run(w, h, t);
    return wk.next == static_cast<int64_t>(t) * h * w
               ? JENGA_OK
               : jenga::set_error(JENGA_E_INVALID, "gilbert: curve did not cover the grid");
  }
  // Sliced curve (gilbert.py:332-434): one 2-D curve per frame, each frame entered at the
  // corner nearest to where the previous frame ended (mirroring x and/or y).
  const int64_t hw = static_cast<int64_t>(h) * w;
  std::vector<int64_t> base_l2h(hw), base_h2l(hw);
  Walker wk;
  wk.W = w;
  wk.H = h;
  wk.to_curve = base_l2h.data();
  wk.from_curve = base_h2l.data();
  wk.run(w, h, 1);
  if (wk.next != hw) return jenga::set_error(JENGA_E_INVALID, "gilbert: 2-D curve incomplete");
  bool flip_x = false, flip_y = false;
  for (int z = 0; z < t; ++z) {
    const int64_t off = static_cast<int64_t>(z) * hw;
    for (int y = 0; y < h; ++y) {
      const int ay = flip_y ? h - 1 - y : y;
      for (int x = 0; x < w; ++x) {
        const int ax = flip_x ? w - 1 - x : x;
        const int64_t idx = off + base_l2h[static_cast<int64_t>(ay) * w + ax];
        const int64_t lin = off + static_cast<int64_t>(y) * w + x;
        if (l2h) l2h[lin] = idx;
        if (h2l) h2l[idx] = lin;
      }
    }
    // where this frame ends, in unflipped coordinates (:401-405)
    const int64_t last = base_h2l[hw - 1];
    int ey = static_cast<int>(last / w), ex = static_cast<int>(last % w);
    if (flip_x) ex = w - 1 - ex;
    if (flip_y) ey = h - 1 - ey;
    flip_x = 2 * ex >= w;  // end_x >= w/2 -> start from the right edge (:373-385)
    flip_y = 2 * ey >= h;
  }
  return JENGA_OK;
}

}  // namespace

extern "C" int64_t jenga_gilbert_xyz2d(int x, int y, int z, int width, int height, int depth) {
  if (x < 0 || y < 0 || z < 0 || x >= width || y >= height || z >= depth) return -1;
  // The reference answers one voxel per call (gilbert.py:12-38); callers loop over a whole box,
  // so the table of the last box is kept (per thread) and a query is one load.
  thread_local int cw = 0, ch = 0, cd = 0;
  thread_local std::vector<int64_t> table;
  if (cw != width || ch != height || cd != depth) {
    table.assign(static_cast<size_t>(width) * height * depth, 0);
    Walker wk;
    wk.W = width;
    wk.H = height;
    wk.to_curve = table.data();
    wk.run(width, height, depth);
    cw = width, ch = height, cd = depth;
  }
  return table[(static_cast<size_t>(z) * height + y) * width + x];
}

extern "C" int jenga_gilbert_mapping_host(int t, int h, int w, int sliced,
                                          int64_t* linear_to_hilbert, int64_t* hilbert_to_linear) {
  if (!linear_to_hilbert && !hilbert_to_linear)
    return jenga::set_error(JENGA_E_INVALID, "gilbert: no output");
  return build_mapping(t, h, w, sliced, linear_to_hilbert, hilbert_to_linear);
}

namespace {

// 26-neighbourhood + self over the voxel grid for ANY voxel->curve-index table
// (gilbert.py:633-663 / :727-757: both builders colour voxels by idx // block and union).
int neighbours_from_table(int t, int h, int w, int block, const int64_t* l2h, uint8_t* neighbors) {
  const int64_t n = static_cast<int64_t>(t) * h * w;
  const int64_t nb = (n + block - 1) / block;
  std::vector<int32_t> colour(n);
  for (int64_t i = 0; i < n; ++i) {
    if (l2h[i] < 0 || l2h[i] >= n) return jenga::set_error(JENGA_E_INVALID, "gilbert: table entry out of range");
    colour[i] = static_cast<int32_t>(l2h[i] / block);
  }
  for (int64_t i = 0; i < nb * nb; ++i) neighbors[i] = 0;
  for (int z = 0; z < t; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const int32_t me = colour[(static_cast<int64_t>(z) * h + y) * w + x];
        uint8_t* row = neighbors + static_cast<int64_t>(me) * nb;
        for (int dz = -1; dz <= 1; ++dz) {
          const int nz = z + dz;
          if (nz < 0 || nz >= t) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            const int ny = y + dy;
            if (ny < 0 || ny >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
              const int nx = x + dx;
              if (nx < 0 || nx >= w) continue;
              row[colour[(static_cast<int64_t>(nz) * h + ny) * w + nx]] = 1;
            }
          }
        }
      }
  return JENGA_OK;
}

}  // namespace

extern "C" int jenga_gilbert_block_neighbors_host(int t, int h, int w, int block, int sliced,
                                                  uint8_t* neighbors) {
  if (!neighbors || block <= 0) return jenga::set_error(JENGA_E_INVALID, "gilbert: bad argument");
  const int64_t n = static_cast<int64_t>(t) * h * w;
  if (n <= 0) return jenga::set_error(JENGA_E_INVALID, "gilbert: empty grid");
  std::vector<int64_t> l2h(n);
  if (int rc = build_mapping(t, h, w, sliced, l2h.data(), nullptr)) return rc;
  return neighbours_from_table(t, h, w, block, l2h.data(), neighbors);
}

extern "C" int jenga_block_neighbors_from_mapping_host(int t, int h, int w, int block,
                                                       const int64_t* linear_to_hilbert,
                                                       uint8_t* neighbors) {
  if (!neighbors || !linear_to_hilbert || block <= 0 || t <= 0 || h <= 0 || w <= 0)
    return jenga::set_error(JENGA_E_INVALID, "gilbert: bad argument");
  return neighbours_from_table(t, h, w, block, linear_to_hilbert, neighbors);
}

extern "C" int jenga_gilbert_block_neighbors_csr_host(int t, int h, int w, int block, int sliced,
                                                      int32_t* row_ptr, int32_t* col_idx,
                                                      int64_t col_capacity, int64_t* nnz_out) {
  if (!row_ptr || block <= 0) return jenga::set_error(JENGA_E_INVALID, "gilbert: bad argument");
  const int64_t n = static_cast<int64_t>(t) * h * w;
  if (n <= 0) return jenga::set_error(JENGA_E_INVALID, "gilbert: empty grid");
  const int64_t nb = (n + block - 1) / block;
  std::vector<uint8_t> dense(static_cast<size_t>(nb) * nb);
  if (int rc = jenga_gilbert_block_neighbors_host(t, h, w, block, sliced, dense.data())) return rc;
  int64_t nnz = 0;
  for (int64_t i = 0; i < nb; ++i) {
    row_ptr[i] = static_cast<int32_t>(nnz);
    for (int64_t j = 0; j < nb; ++j)
      if (dense[i * nb + j]) {
        if (col_idx && nnz < col_capacity) col_idx[nnz] = static_cast<int32_t>(j);
        ++nnz;
      }
  }
  row_ptr[nb] = static_cast<int32_t>(nnz);
  if (nnz_out) *nnz_out = nnz;
  if (col_idx && nnz > col_capacity)
    return jenga::set_error(JENGA_E_INVALID, "gilbert: col_idx capacity too small (nnz returned)");
  return JENGA_OK;
}
