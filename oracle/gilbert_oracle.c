/*
 * oracle/gilbert_oracle.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * Plain-C restatement of the reference's per-voxel curve-index query
 *   gilbert.py:12-38   gilbert_xyz2d      (axis ordering by the longest side)
 *   gilbert.py:43-65   in_bounds          (half-open box test along the summed direction)
 *   gilbert.py:68-272  gilbert_xyz2d_r    (recursive descent adding sub-box volumes)
 * and of the table builders that call it once per voxel
 *   gilbert.py:442-488 gilbert_mapping, :332-440 sliced_gilbert_mapping,
 *   gilbert.py:597-677 gilbert_block_neighbor_mapping (and the sliced twin :679-766).
 * The product (jenga_b200/csrc/gilbert.cpp) generates the same tables by walking the curve in
 * order; this file answers index queries voxel by voxel, the way the reference does, so the
 * two derivations check each other.  Pinned against the reference's own outputs through
 * tests/golden/gilbert.json (SHA-256 of the int64 tables, tests/golden/make_golden.py).
 *
 * Build: make -C oracle   ->  oracle/_build/libgilbert_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { long x, y, z; } v3;

static v3 vadd(v3 a, v3 b) { v3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static v3 vsub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static v3 vneg(v3 a) { v3 r = {-a.x, -a.y, -a.z}; return r; }
static long sgn1(long v) { return v < 0 ? -1 : (v > 0 ? 1 : 0); }          /* :40-41 */
static v3 vsgn(v3 a) { v3 r = {sgn1(a.x), sgn1(a.y), sgn1(a.z)}; return r; }
static long vsum(v3 a) { return a.x + a.y + a.z; }
static long labs1(long v) { return v < 0 ? -v : v; }
/* Python's // rounds toward -inf (:91-93) */
static long floordiv2(long v) { return (v >= 0) ? v / 2 : -((-v + 1) / 2); }
static v3 vhalf(v3 a) { v3 r = {floordiv2(a.x), floordiv2(a.y), floordiv2(a.z)}; return r; }
static long dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* :43-65 — is p inside the box that starts at s and spans a+b+c (signed, half-open)? */
static int inside(v3 p, v3 s, v3 a, v3 b, v3 c) {
  const v3 d = vadd(vadd(a, b), c);
  const long pc[3] = {p.x, p.y, p.z}, sc[3] = {s.x, s.y, s.z}, dc[3] = {d.x, d.y, d.z};
  for (int i = 0; i < 3; ++i) {
    if (dc[i] < 0) {
      if (pc[i] > sc[i] || pc[i] <= sc[i] + dc[i]) return 0;
    } else {
      if (pc[i] < sc[i] || pc[i] >= sc[i] + dc[i]) return 0;
    }
  }
  return 1;
}

static long volume(v3 a, v3 b, v3 c) { return labs1(vsum(a) * vsum(b) * vsum(c)); }

/* :68-272 */
static long query(long idx, v3 dst, v3 p, v3 a, v3 b, v3 c) {
  for (;;) { /* every reference branch ends in a tail call: iterate instead of recursing */
    const long w = labs1(vsum(a)), h = labs1(vsum(b)), d = labs1(vsum(c));
    const v3 da = vsgn(a), db = vsgn(b), dc = vsgn(c);
    const v3 off = vsub(dst, p);
    if (h == 1 && d == 1) return idx + dot(da, off);   /* :84-85 */
    if (w == 1 && d == 1) return idx + dot(db, off);   /* :87-88 */
    if (w == 1 && h == 1) return idx + dot(dc, off);   /* :90-91 */

    v3 a2 = vhalf(a), b2 = vhalf(b), c2 = vhalf(c);
    const long w2 = labs1(vsum(a2)), h2 = labs1(vsum(b2)), d2 = labs1(vsum(c2));
    if ((w2 % 2) && w > 2) a2 = vadd(a2, da);           /* :101-109 prefer even steps */
    if ((h2 % 2) && h > 2) b2 = vadd(b2, db);
    if ((d2 % 2) && d > 2) c2 = vadd(c2, dc);

    v3 np, na, nb, nc; /* the sub-box we descend into */
    if (2 * w > 3 * h && 2 * w > 3 * d) {               /* :112-132 wide: split a */
      if (inside(dst, p, a2, b, c)) { np = p; na = a2; nb = b; nc = c; }
      else { idx += volume(a2, b, c); np = vadd(p, a2); na = vsub(a, a2); nb = b; nc = c; }
    } else if (3 * h > 4 * d) {                          /* :134-169 keep c whole */
      if (inside(dst, p, b2, c, a2)) { np = p; na = b2; nb = c; nc = a2; }
      else {
        idx += volume(b2, c, a2);
        const v3 q = vadd(p, b2);
        if (inside(dst, q, a, vsub(b, b2), c)) { np = q; na = a; nb = vsub(b, b2); nc = c; }
        else {
          idx += volume(a, vsub(b, b2), c);
          np = vadd(vadd(p, vsub(a, da)), vsub(b2, db));
          na = vneg(b2); nb = c; nc = vneg(vsub(a, a2));
        }
      }
    } else if (3 * d > 4 * h) {                          /* :171-204 keep b whole */
      if (inside(dst, p, c2, a2, b)) { np = p; na = c2; nb = a2; nc = b; }
      else {
        idx += volume(c2, a2, b);
        const v3 q = vadd(p, c2);
        if (inside(dst, q, a, b, vsub(c, c2))) { np = q; na = a; nb = b; nc = vsub(c, c2); }
        else {
          idx += volume(a, b, vsub(c, c2));
          np = vadd(vadd(p, vsub(a, da)), vsub(c2, dc));
          na = vneg(c2); nb = vneg(vsub(a, a2)); nc = b;
        }
      }
    } else {                                             /* :206-272 split all three */
      if (inside(dst, p, b2, c2, a2)) { np = p; na = b2; nb = c2; nc = a2; }
      else {
        idx += volume(b2, c2, a2);
        const v3 q1 = vadd(p, b2);
        if (inside(dst, q1, c, a2, vsub(b, b2))) { np = q1; na = c; nb = a2; nc = vsub(b, b2); }
        else {
          idx += volume(c, a2, vsub(b, b2));
          const v3 q2 = vadd(vadd(p, vsub(b2, db)), vsub(c, dc));
          if (inside(dst, q2, a, vneg(b2), vneg(vsub(c, c2)))) {
            np = q2; na = a; nb = vneg(b2); nc = vneg(vsub(c, c2));
          } else {
            idx += volume(a, vneg(b2), vneg(vsub(c, c2)));
            const v3 q3 = vadd(vadd(vadd(p, vsub(a, da)), b2), vsub(c, dc));
            if (inside(dst, q3, vneg(c), vneg(vsub(a, a2)), vsub(b, b2))) {
              np = q3; na = vneg(c); nb = vneg(vsub(a, a2)); nc = vsub(b, b2);
            } else {
              idx += volume(vneg(c), vneg(vsub(a, a2)), vsub(b, b2));
              np = vadd(vadd(p, vsub(a, da)), vsub(b2, db));
              na = vneg(b2); nb = c2; nc = vneg(vsub(a, a2));
            }
          }
        }
      }
    }
    p = np; a = na; b = nb; c = nc;
  }
}

/* :12-38 */
long oracle_gilbert_xyz2d(long x, long y, long z, long width, long height, long depth) {
  const v3 dst = {x, y, z}, o = {0, 0, 0};
  const v3 X = {width, 0, 0}, Y = {0, height, 0}, Z = {0, 0, depth};
  if (width >= height && width >= depth) return query(0, dst, o, X, Y, Z);
  if (height >= width && height >= depth) return query(0, dst, o, Y, X, Z);
  return query(0, dst, o, Z, X, Y);
}

/* :442-488 (sliced == 0) and :332-434 (sliced != 0) */
int oracle_gilbert_mapping(int t, int h, int w, int sliced, int64_t* l2h, int64_t* h2l) {
  const long hw = (long)h * w;
  if (!sliced) {
    for (long z = 0; z < t; ++z)
      for (long y = 0; y < h; ++y)
        for (long x = 0; x < w; ++x) {
          const long lin = z * hw + y * w + x;
          const long idx = oracle_gilbert_xyz2d(x, y, z, w, h, t);
          l2h[lin] = idx;
          h2l[idx] = lin;
        }
    return 0;
  }
  long* s_l2h = (long*)malloc(sizeof(long) * hw);
  long* s_h2l = (long*)malloc(sizeof(long) * hw);
  if (!s_l2h || !s_h2l) return -1;
  int have_last = 0;
  long end_x = 0, end_y = 0, base = 0;
  for (long z = 0; z < t; ++z) {
    int flip_x = 0, flip_y = 0;                       /* :369-391 nearest corner */
    if (have_last) {
      flip_x = !(2 * end_x < w);   /* end_x >= w/2 (true division) */
      flip_y = !(2 * end_y < h);
    }
    for (long y = 0; y < h; ++y)
      for (long x = 0; x < w; ++x) {
        const long ax = flip_x ? w - 1 - x : x, ay = flip_y ? h - 1 - y : y;
        const long idx = oracle_gilbert_xyz2d(ax, ay, 0, w, h, 1);
        s_l2h[y * w + x] = idx;
        s_h2l[idx] = y * w + x;
      }
    end_y = s_h2l[hw - 1] / w;                         /* :411-415 */
    end_x = s_h2l[hw - 1] % w;
    have_last = 1;
    for (long i = 0; i < hw; ++i) {                    /* :418-431 */
      l2h[z * hw + i] = base + s_l2h[i];
      h2l[base + s_l2h[i]] = z * hw + i;
    }
    base += hw;
  }
  free(s_l2h);
  free(s_h2l);
  return 0;
}

/* :597-677 / :679-766 — neighbours[i*nb + j] = 1 iff a voxel of block i touches (26-nbhd or
 * equals) a voxel of block j */
int oracle_gilbert_block_neighbors(int t, int h, int w, int block, int sliced, uint8_t* nbrs) {
  const long n = (long)t * h * w, nb = (n + block - 1) / block;
  int64_t* l2h = (int64_t*)malloc(sizeof(int64_t) * n);
  int64_t* h2l = (int64_t*)malloc(sizeof(int64_t) * n);
  if (!l2h || !h2l) return -1;
  oracle_gilbert_mapping(t, h, w, sliced, l2h, h2l);
  memset(nbrs, 0, (size_t)(nb * nb));
  for (long x = 0; x < w; ++x)
    for (long y = 0; y < h; ++y)
      for (long z = 0; z < t; ++z) {
        const long cur = l2h[z * (long)h * w + y * w + x] / block;
        nbrs[cur * nb + cur] = 1;
        for (long dx = -1; dx <= 1; ++dx) {
          const long nx = x + dx;
          if (nx < 0 || nx >= w) continue;
          for (long dy = -1; dy <= 1; ++dy) {
            const long ny = y + dy;
            if (ny < 0 || ny >= h) continue;
            for (long dz = -1; dz <= 1; ++dz) {
              const long nz = z + dz;
              if (nz < 0 || nz >= t) continue;
              if (!dx && !dy && !dz) continue;
              nbrs[cur * nb + l2h[nz * (long)h * w + ny * w + nx] / block] = 1;
            }
          }
        }
      }
  free(l2h);
  free(h2l);
  return 0;
}
