// widen.cpp — host side of the packed device-to-host transfer of distance
// fields (mrp_bfs_fields, c_api.cu).  The device sends a field as uint16
// (0xFFFF = MRP_INF) when every finite distance of the batch fits; the host
// threads here expand it to the int32 layout of the ABI
// (ShortestPathHeuristic::getValue, example/shortest_path_heuristic.hpp:58-62:
// `int`, INT_MAX = unreachable) while the next batch is still on the bus.
// This is a transfer format only: the bytes that reach the caller are the same.
#include <cstddef>
#include <cstdint>
#include <immintrin.h>
#include <thread>
#include <vector>

namespace mrp {

static void widenScalar(const uint16_t* s, int32_t* d, size_t n) {
  for (size_t i = 0; i < n; ++i) d[i] = s[i] == 0xFFFFu ? 0x7fffffff : (int32_t)s[i];
}

__attribute__((target("avx2"))) static void widenAvx2(const uint16_t* s, int32_t* d, size_t n) {
  size_t i = 0;
  // head: up to the first 32-byte boundary of the destination
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 31u)) {
    d[i] = s[i] == 0xFFFFu ? 0x7fffffff : (int32_t)s[i];
    ++i;
  }
  const __m256i ffff = _mm256_set1_epi32(0xFFFF), inf = _mm256_set1_epi32(0x7fffffff);
  for (; i + 16 <= n; i += 16) {
    __m256i a = _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)));
    __m256i b = _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 8)));
    a = _mm256_blendv_epi8(a, inf, _mm256_cmpeq_epi32(a, ffff));
    b = _mm256_blendv_epi8(b, inf, _mm256_cmpeq_epi32(b, ffff));
    // the output is written once and not read back here: bypass the caches
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 8), b);
  }
  _mm_sfence();
  for (; i < n; ++i) d[i] = s[i] == 0xFFFFu ? 0x7fffffff : (int32_t)s[i];
}

static void widenSlice(const uint16_t* s, int32_t* d, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2)
    widenAvx2(s, d, n);
  else
    widenScalar(s, d, n);
}

// dst[i] = src[i] == 0xFFFF ? INT_MAX : src[i], on `threads` host threads.
void widenFieldU16(const uint16_t* src, int32_t* dst, size_t n, int threads) {
  const size_t minSlice = (size_t)1 << 16;
  size_t parts = (n + minSlice - 1) / minSlice;
  if (parts > (size_t)threads) parts = (size_t)threads;
  if (parts <= 1) {
    widenSlice(src, dst, n);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(parts - 1);
  const size_t per = ((n + parts - 1) / parts + 15) & ~(size_t)15;
  for (size_t p = 1; p < parts; ++p) {
    const size_t b = p * per, e = b + per < n ? b + per : n;
    if (b >= n) break;
    pool.emplace_back([=] { widenSlice(src + b, dst + b, e - b); });
  }
  widenSlice(src, dst, per < n ? per : n);
  for (auto& t : pool) t.join();
}


// ---- one byte per cell: h = (distance - Manhattan distance to the goal) / 2 ----
// On a 4-connected grid the two have the same parity and the BFS distance is
// never the smaller one, so h is a non-negative integer; 255 = MRP_INF.
static void widenRowsScalar(const uint8_t* s, int32_t* d, int dimx, int y, int gx, int gy) {
  const int base = y > gy ? y - gy : gy - y;
  for (int x = 0; x < dimx; ++x) {
    const int m = base + (x > gx ? x - gx : gx - x);
    d[x] = s[x] == 255 ? 0x7fffffff : 2 * (int)s[x] + m;
  }
}

__attribute__((target("avx2"))) static void widenRowAvx2(const uint8_t* s, int32_t* d, int dimx,
                                                          int y, int gx, int gy) {
  const int base = y > gy ? y - gy : gy - y;
  int x = 0;
  while (x < dimx && (reinterpret_cast<uintptr_t>(d + x) & 31u)) {
    const int m = base + (x > gx ? x - gx : gx - x);
    d[x] = s[x] == 255 ? 0x7fffffff : 2 * (int)s[x] + m;
    ++x;
  }
  const __m256i ramp = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
  const __m256i vff = _mm256_set1_epi32(255), inf = _mm256_set1_epi32(0x7fffffff);
  const __m256i vbase = _mm256_set1_epi32(base), vgx = _mm256_set1_epi32(gx);
  for (; x + 8 <= dimx; x += 8) {
    const __m256i h = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(s + x)));
    const __m256i xs = _mm256_add_epi32(_mm256_set1_epi32(x), ramp);
    const __m256i m = _mm256_add_epi32(vbase, _mm256_abs_epi32(_mm256_sub_epi32(xs, vgx)));
    __m256i v = _mm256_add_epi32(_mm256_slli_epi32(h, 1), m);
    v = _mm256_blendv_epi8(v, inf, _mm256_cmpeq_epi32(h, vff));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + x), v);
  }
  for (; x < dimx; ++x) {
    const int m = base + (x > gx ? x - gx : gx - x);
    d[x] = s[x] == 255 ? 0x7fffffff : 2 * (int)s[x] + m;
  }
}

// rows [r0, r1) of the n_fields*dimy rows of a batch
static void widenRows(const uint8_t* src, int32_t* dst, int dimx, int dimy, const int32_t* goalCell,
                      size_t r0, size_t r1) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  for (size_t r = r0; r < r1; ++r) {
    const size_t k = r / (size_t)dimy;
    const int y = (int)(r - k * (size_t)dimy);
    const int gx = goalCell[k] % dimx, gy = goalCell[k] / dimx;
    if (avx2)
      widenRowAvx2(src + r * dimx, dst + r * dimx, dimx, y, gx, gy);
    else
      widenRowsScalar(src + r * dimx, dst + r * dimx, dimx, y, gx, gy);
  }
  if (avx2) _mm_sfence();
}

// dst[k][y][x] = src == 255 ? INT_MAX : 2*src + |x - gx_k| + |y - gy_k| for the
// n_fields fields of a batch (goalCell[k] = gx + dimx*gy), on `threads` threads.
void widenFieldU8(const uint8_t* src, int32_t* dst, int dimx, int dimy, const int32_t* goalCell,
                  size_t n_fields, int threads) {
  const size_t rows = n_fields * (size_t)dimy;
  size_t parts = (rows * (size_t)dimx + ((size_t)1 << 16) - 1) >> 16;
  if (parts > (size_t)threads) parts = (size_t)threads;
  if (parts <= 1) {
    widenRows(src, dst, dimx, dimy, goalCell, 0, rows);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(parts - 1);
  const size_t per = (rows + parts - 1) / parts;
  for (size_t p = 1; p < parts; ++p) {
    const size_t b = p * per, e = b + per < rows ? b + per : rows;
    if (b >= rows) break;
    pool.emplace_back([=] { widenRows(src, dst, dimx, dimy, goalCell, b, e); });
  }
  widenRows(src, dst, dimx, dimy, goalCell, 0, per < rows ? per : rows);
  for (auto& t : pool) t.join();
}

}  // namespace mrp
