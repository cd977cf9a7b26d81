// yt_lane.cuh — one lane per record for the YouTube line (records whose strings need no escaping).
//
// The line is the templated walk of yt_walk.cuh, instantiated with a writer that streams the lane's own
// line: same byte-stream packing as tg_lane.cuh (16-byte blocks, byte-exact stores where a block is shared
// with the neighbouring line), but the blocks are stored directly (ST.128 from the lane; staging them in
// shared memory for per-lane TMA bulk stores was measured and is slower here: 63 vs 69 M records/s, profiles/README.md)
// and nothing in the writer is warp-collective, so the walk's per-record branches and loops (cached channel or not, 0-5
// thumbnails, n outlinks) may diverge freely between the lanes.  Records with a string that needs escaping
// are left to the warp writer (yt_emit_kernel), which visits only those.
#pragma once
#include "tg_lane.cuh"
#include "yt_walk.cuh"

namespace tgi {

struct LaneDirect {
  uint64_t pos;             // absolute address of the next output byte
  uint32_t c0, c1, c2, c3;  // bytes [head, pos & 15) of the current block, zero elsewhere
  uint32_t head;            // first byte of the current block that belongs to this stream
};

// append bytes [0, n) of src (any address space, any alignment; the 16-byte blocks around it must be readable)
__device__ __noinline__ void ld_copy(LaneDirect* sp, const uint8_t* src, uint32_t n) {
  if (!n) return;
  LaneDirect s = *sp;
  const uint32_t s0 = (uint32_t)(uintptr_t)src & 15u;
  const uint4* A = (const uint4*)(src - s0);
  uint32_t rem = n, first = s0;
  while (rem) {
    uint4 w = *A++;
    if (first) w = shr128_bytes(w, first);
    const uint32_t k = min(rem, 16u - first);
    if (k < 16u) w = mask128(w, k);
    rem -= k;
    first = 0;
    // place the block at the stream's phase (ls_append, tg_lane.cuh)
    const uint32_t ph = (uint32_t)s.pos & 15u, sh = (ph & 3u) * 8u;
    const uint32_t v0 = w.x << sh, v1 = __funnelshift_l(w.x, w.y, sh), v2 = __funnelshift_l(w.y, w.z, sh),
                   v3 = __funnelshift_l(w.z, w.w, sh), v4 = __funnelshift_l(w.w, 0u, sh);
    const bool b0 = (ph & 4u) != 0, b1 = (ph & 8u) != 0;
    const uint32_t z0 = b0 ? 0u : v0, z1 = b0 ? v0 : v1, z2 = b0 ? v1 : v2, z3 = b0 ? v2 : v3, z4 = b0 ? v3 : v4,
                   z5 = b0 ? v4 : 0u;
    s.c0 |= b1 ? 0u : z0;
    s.c1 |= b1 ? 0u : z1;
    s.c2 |= b1 ? z0 : z2;
    s.c3 |= b1 ? z1 : z3;
    if (ph + k >= 16u) {
      const uint64_t blk = s.pos & ~15ull;
      if (s.head == 0) {
        *(uint4*)(uintptr_t)blk = make_uint4(s.c0, s.c1, s.c2, s.c3);
      } else {
        store_bytes(blk, s.c0, s.c1, s.c2, s.c3, s.head, 16u);
        s.head = 0;
      }
      s.c0 = b1 ? z2 : z4;
      s.c1 = b1 ? z3 : z5;
      s.c2 = b1 ? z4 : 0u;
      s.c3 = b1 ? z5 : 0u;
    }
    s.pos += k;
  }
  *sp = s;
}

// append the JSON-escaped form of src[0, n): the string is valid UTF-8 without U+2028 / U+2029 (the size pass checked), so
// only ASCII bytes expand.  16 source bytes at a time: blocks without a special byte (almost all) are appended as they
// are, the others byte by byte.
__device__ __noinline__ void ld_copy_esc(LaneDirect* sp, const uint8_t* src, uint32_t n) {
  __align__(16) uint8_t tmp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // ld_copy reads whole 16-byte blocks
  for (uint32_t i = 0; i < n; i += 16) {
    const uint32_t k = min(16u, n - i);
    bool special = false;
    for (uint32_t j = 0; j < k; j += 4) {
      uint32_t x = ld_u32_unaligned(src + i + j);
      if (k - j < 4) x = (x & ((1u << (8u * (k - j))) - 1u)) | (0x20202020u << (8u * (k - j)));
      const uint32_t ctl = ~(((x & 0x7F7F7F7Fu) + 0x60606060u) | x) & 0x80808080u;
      special = special || (ctl | swar_has_byte(x & 0xFBFBFBFBu, 0x22) | swar_has_byte(x & 0xFDFDFDFDu, 0x3C) | swar_has_byte(x, 0x5C)) != 0;
    }
    if (!special) {
      ld_copy(sp, src + i, k);
      continue;
    }
    for (uint32_t j = 0; j < k; j++) {
      const uint32_t bt = ldb(src + i + j);
      const uint32_t el = bt < 0x80u ? ascii_esc_len(bt) : 1u;
      struct { uint8_t* p; DEVI void st(uint32_t o, uint32_t v) const { p[o] = (uint8_t)v; } } sink{tmp};
      put_escaped(sink, 0, bt, el);
      ld_copy(sp, tmp, el);
    }
  }
}

// leave n bytes to another writer: what the current block holds goes out byte-exact, the stream resumes behind the gap
DEVI void ld_skip(LaneDirect& s, uint32_t n) {
  const uint32_t ph = (uint32_t)s.pos & 15u;
  if (ph > s.head) store_bytes(s.pos & ~15ull, s.c0, s.c1, s.c2, s.c3, s.head, ph);
  s.c0 = s.c1 = s.c2 = s.c3 = 0;
  s.pos += n;
  s.head = (uint32_t)s.pos & 15u;
}

constexpr uint32_t YT_LANE_LONG = 128;  // longer strings are copied by the whole warp after the walk (see YtLaneWriter)
constexpr int YT_LANE_PENDING = 4;

struct YtLaneWriter {
  static constexpr bool kLane = true;
  LaneDirect s;
  // The description is written three times and its length varies from 0 to 5000 bytes: copied by its own lane, every
  // warp would run at the pace of its longest description.  Long strings are therefore skipped by the lane and
  // copied afterwards by the whole warp, lane after lane (flush_pending).
  uint64_t pend_dst[YT_LANE_PENDING];
  const uint8_t* pend_src[YT_LANE_PENDING];
  uint32_t pend_n[YT_LANE_PENDING];
  uint32_t pend_el[YT_LANE_PENDING];  // escaped length (== pend_n: plain copy)
  int npend = 0;
  uint32_t el[2];               // escaped lengths of the description / the title (size pass)
  __align__(16) uint8_t num[64];  // number / time / file-name rendering
  DEVI void begin(uint64_t addr) {
    s.pos = addr;
    s.head = (uint32_t)addr & 15u;
    s.c0 = s.c1 = s.c2 = s.c3 = 0;
  }
  DEVI void end() {  // what the last block holds
    const uint32_t ph = (uint32_t)s.pos & 15u;
    if (ph > s.head) store_bytes(s.pos & ~15ull, s.c0, s.c1, s.c2, s.c3, s.head, ph);
  }
  DEVI void raw(const uint8_t* p, uint32_t n) { ld_copy(&s, p, n); }
  DEVI void esc(const uint8_t* p, uint32_t n) { ld_copy(&s, p, n); }  // nothing to escape on this path
  DEVI void esc_slot(int k, const uint8_t* p, uint32_t n) {  // description / title: may need (ASCII-only) escaping
    const uint32_t e = el[k];
    if (n > YT_LANE_LONG && npend < YT_LANE_PENDING) {
      pend_dst[npend] = s.pos;
      pend_src[npend] = p;
      pend_n[npend] = n;
      pend_el[npend] = e;
      npend++;
      ld_skip(s, e);
    } else if (e == n) {
      ld_copy(&s, p, n);
    } else {
      ld_copy_esc(&s, p, n);
    }
  }
  // warp-collective, after the (possibly divergent) walk: every lane's pending long copies, 512 bytes per step
  DEVI void flush_pending(bool active) {
#pragma unroll 1
    for (int k = 0; k < YT_LANE_PENDING; k++) {
      uint32_t todo = __ballot_sync(FULL, active && k < npend);
      while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint64_t d = __shfl_sync(FULL, pend_dst[k < npend ? k : 0], src_lane);
        const uint8_t* p = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)(uintptr_t)pend_src[k < npend ? k : 0], src_lane);
        const uint32_t n = __shfl_sync(FULL, pend_n[k < npend ? k : 0], src_lane);
        const uint32_t e = __shfl_sync(FULL, pend_el[k < npend ? k : 0], src_lane);
        if (e == n) warp_copy_vec((uint8_t*)(uintptr_t)d, p, n);
        else esc_ascii_to_global((uint8_t*)(uintptr_t)d, p, n);
      }
    }
  }
  DEVI void ch(uint32_t c) {
    num[0] = (uint8_t)c;
    ld_copy(&s, num, 1);
  }
  DEVI void dec(int64_t v) { ld_copy(&s, num, (uint32_t)render_i64(num, v)); }
  DEVI void smem(uint32_t n) { ld_copy(&s, num, n); }
  DEVI uint32_t time_len(int64_t sec, int32_t nsec) { return (uint32_t)render_time(num, sec, nsec, 0); }
  DEVI void time(int64_t sec, int32_t nsec) { ld_copy(&s, num, (uint32_t)render_time(num, sec, nsec, 0)); }
  DEVI void fviews(int64_t v) { ld_copy(&s, num, (uint32_t)yt_render_float_of_int64(num, v)); }
  DEVI void sanitized(const uint8_t* t, uint32_t tn) { ld_copy(&s, num, yt_sanitize(t, tn, num)); }
  DEVI bool duration(const uint8_t* d, uint32_t dn, int64_t& vlen) { return yt_parse_duration(d, dn, vlen); }
};

// thread_esc_len is force-inlined; the sizer's walk has ~15 call sites of it: one shared copy keeps the kernel small
__device__ __noinline__ uint32_t yt_thread_esc_len(const uint8_t* p, uint32_t n) { return thread_esc_len(p, n); }

// length pass of the same walk, one lane per record.  The escaped lengths of the two long strings
// (description, title) come from the warp (el[]); every other string is short and measured by the lane.
struct YtLaneSizer {
  static constexpr bool kLane = true;
  uint64_t total = 0;
  uint32_t el[2];
  bool dirty = false;        // some string needs escaping
  bool small_dirty = false;  // ... and it is not the description / the title: only the warp writer escapes those
  __align__(16) uint8_t num[64];
  DEVI void raw(const uint8_t*, uint32_t n) { total += n; }
  DEVI void esc(const uint8_t* p, uint32_t n) {
    const uint32_t e = yt_thread_esc_len(p, n);
    dirty = dirty || e != n;
    small_dirty = small_dirty || e != n;
    total += e;
  }
  DEVI void esc_slot(int k, const uint8_t*, uint32_t n) {
    dirty = dirty || el[k] != n;
    total += el[k];
  }
  DEVI void ch(uint32_t) { total += 1; }
  DEVI void dec(int64_t v) { total += yt_ndigits(v); }
  DEVI void smem(uint32_t n) { total += n; }
  DEVI uint32_t time_len(int64_t sec, int32_t nsec) { return (uint32_t)render_time(num, sec, nsec, 0); }
  DEVI void time(int64_t sec, int32_t nsec) { total += (uint32_t)render_time(num, sec, nsec, 0); }
  DEVI void fviews(int64_t v) { total += (uint32_t)yt_render_float_of_int64(num, v); }
  DEVI void sanitized(const uint8_t* t, uint32_t tn) { total += yt_sanitize(t, tn, num); }
  DEVI bool duration(const uint8_t* d, uint32_t dn, int64_t& vlen) { return yt_parse_duration(d, dn, vlen); }
};

}  // namespace tgi
