/*
 * oracle_c.c — CPU restatement (plain C + OpenMP) of the integer half of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nerf_downstream_b200/ may link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED against MinkowskiEngine itself: the algorithm lives in the third-party
 * dependency MinkowskiEngine (PyPI, un-pinned => 0.5.4; installed by /root/reference/install.sh:50-52
 * and co3d_3d/README.md:13), whose source is NOT under /root/reference and which cannot be
 * installed here.  This file restates ME's published CPU algorithm:
 *   - coordinates are int32 (b,x,y,z); field coordinates are floored per axis (A.2);
 *   - insert_and_map: sequential insert into a hash map, rows in first-occurrence order (A.2);
 *   - stride(): floor(c / ts) * ts per spatial axis, then unique (A.3);
 *   - kernel_map(): for every out row and kernel offset probe the in-map, offsets centred for odd
 *     kernel sizes and [0,k) for even ones, scaled by the input tensor stride, first spatial axis
 *     fastest in the offset index (A.4; index order corroborated in-tree by
 *     co3d_3d/src/models/mink/modules/sparse_conv.py:375-379).
 * It is anchored on the reference's own call sites: sparse_conv.py:80-96 (size / kernel_map
 * arguments and {k: [2,n]} layout), :122-143 (how the pairs are consumed), :397-405 (out-key rules).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t* rows;      /* row index or -1 */
  const int32_t* coords; /* [*,4] the rows point into */
  uint64_t mask;
} table_t;

static uint64_t mix(const int32_t* c) {
  /* murmur-style mixing of the 4 ints */
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < 4; ++i) {
    uint64_t k = (uint32_t)c[i];
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    h ^= k;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
  }
  return h;
}

static int table_init(table_t* t, int64_t n, const int32_t* coords) {
  uint64_t cap = 16;
  while (cap < (uint64_t)(2 * n + 1)) cap <<= 1;
  t->rows = (int32_t*)malloc(cap * sizeof(int32_t));
  if (!t->rows) return -1;
  memset(t->rows, 0xFF, cap * sizeof(int32_t));
  t->mask = cap - 1;
  t->coords = coords;
  return 0;
}

static inline int same4(const int32_t* a, const int32_t* b) {
  return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}

/* returns the row stored for key c, or inserts `row` and returns it */
static int32_t table_find_or_insert(table_t* t, const int32_t* c, int32_t row) {
  uint64_t s = mix(c) & t->mask;
  for (;;) {
    int32_t r = t->rows[s];
    if (r < 0) {
      t->rows[s] = row;
      return row;
    }
    if (same4(t->coords + 4 * (int64_t)r, c)) return r;
    s = (s + 1) & t->mask;
  }
}

static int32_t table_find(const table_t* t, const int32_t* c) {
  uint64_t s = mix(c) & t->mask;
  for (;;) {
    int32_t r = t->rows[s];
    if (r < 0) return -1;
    if (same4(t->coords + 4 * (int64_t)r, c)) return r;
    s = (s + 1) & t->mask;
  }
}

/* floor quantisation of float coordinates: batch floor(b), spatial floor(x/ts)*ts (A.2) */
void orc_quantize_f32(const float* in, int64_t n, const int32_t* ts, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    out[4 * i] = (int32_t)floorf(in[4 * i]);
    for (int a = 0; a < 3; ++a) {
      float x = in[4 * i + 1 + a];
      if (ts[a] == 1) out[4 * i + 1 + a] = (int32_t)floorf(x);
      else out[4 * i + 1 + a] = (int32_t)(floorf(x / (float)ts[a]) * (float)ts[a]);
    }
  }
}

static inline int32_t floor_div(int32_t a, int32_t b) {
  int32_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

/* strided coordinates: floor(c/ts)*ts per spatial axis, batch unchanged (A.3) */
void orc_stride_coords(const int32_t* in, int64_t n, const int32_t* ts, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    out[4 * i] = in[4 * i];
    for (int a = 0; a < 3; ++a) out[4 * i + 1 + a] = floor_div(in[4 * i + 1 + a], ts[a]) * ts[a];
  }
}

/*
 * insert_and_map (A.2): sequential first-occurrence insert.
 *   out_coords [n,4], unique_index [n], inverse [n]; returns the number of unique rows M.
 */
int64_t orc_unique_first(const int32_t* coords, int64_t n, int32_t* out_coords, int32_t* unique_index,
                         int32_t* inverse) {
  table_t t;
  if (table_init(&t, n, out_coords)) return -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    /* tentatively place the candidate at row m so the table can compare against it */
    memcpy(out_coords + 4 * m, coords + 4 * i, 4 * sizeof(int32_t));
    int32_t r = table_find_or_insert(&t, coords + 4 * i, (int32_t)m);
    if (r == (int32_t)m) {
      unique_index[m] = (int32_t)i;
      ++m;
    }
    inverse[i] = r;
  }
  free(t.rows);
  return m;
}

/*
 * kernel_map (A.4): nbr[k*m_out + o] = in row at coord_out[o] + offsets[k], else -1.
 * OpenMP over out rows, like ME's CPU backend iterates the out map in parallel.
 */
int orc_kernel_map(const int32_t* in_coords, int64_t m_in, const int32_t* out_coords, int64_t m_out,
                   const int32_t* offsets, int K, int32_t* nbr) {
  table_t t;
  if (table_init(&t, m_in, in_coords)) return -1;
  for (int64_t i = 0; i < m_in; ++i) table_find_or_insert(&t, in_coords + 4 * i, (int32_t)i);
#pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < m_out; ++o) {
    const int32_t* c = out_coords + 4 * o;
    for (int k = 0; k < K; ++k) {
      int32_t q[4] = {c[0], c[1] + offsets[3 * k], c[2] + offsets[3 * k + 1], c[3] + offsets[3 * k + 2]};
      nbr[(int64_t)k * m_out + o] = table_find(&t, q);
    }
  }
  free(t.rows);
  return 0;
}

/* gather rows: out[j,:] = src[idx[j],:] (zero when idx < 0) */
void orc_gather_rows(const float* src, const int32_t* idx, int64_t n, int C, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    if (idx[j] >= 0) memcpy(out + j * C, src + (int64_t)idx[j] * C, (size_t)C * sizeof(float));
    else memset(out + j * C, 0, (size_t)C * sizeof(float));
  }
}
