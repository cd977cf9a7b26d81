// tg_tile.cuh — tile emission of the Telegram Post line: lines are ASSEMBLED IN SHARED MEMORY and leave the SM as
// bulk asynchronous stores (cp.async.bulk.global.shared::cta, SASS UBLKCP: the TMA unit's non-tensor path).
//
// The lines of consecutive records are contiguous in the output blob (line_off is an exclusive scan in record order),
// so a warp that walks 32 consecutive records produces ONE contiguous byte stream.  The warp keeps a window of that
// stream in its own shared-memory buffer: byte `a` of the output blob lives at buffer offset (a - gs), gs 16-byte
// aligned, so a byte's alignment is the same in both places and every piece of a line — literal template words,
// rendered numbers, channel / context strings, message strings, escaped strings, maps, comment lists — is written
// with plain byte-granular st.shared at its final offset (no shift network, no byte-exact global stores).  When the
// next line does not fit, the aligned part of the window is retired with one elected-thread bulk store and the few
// tail bytes move to the front of the buffer.  Only the first and the last 16-byte block of a warp's 32-record range
// can be shared with another writer; those are stored byte-exact.
//
// Work split: 16 records at a time are PREPARED lane-parallel (one lane per record renders its numbers and the time
// stamp into a shared row); then the warp emits the records one after the other, all 32 lanes on one line:
//   1. piece table  — lane i owns pieces 2i, 2i+1 of the generated line program (tools/gen_pieces.py), a warp scan of
//                     their lengths gives every piece its offset in the line;
//   2. literals     — the 1360-byte template, word by word with the shift of the owning piece (11 steps);
//   3. fields       — rendered numbers / time / post type, by the owning lanes;
//   4. strings      — channel / context segments and clean message strings: cooperative copies; strings that need
//                     escaping, comment lists, reaction maps, outlinks: the sink-templated walkers of tg_walk.cuh.
// Reference semantics: telegramhelper/tdutils.go:633-717 + encoding/json of model.Post (model/data.go:9-75).
#pragma once
#include "tg_walk.cuh"

namespace tgi {

constexpr int TILE_WARPS = 8;
constexpr uint32_t TILE_BUF = 5888;                       // bytes of line buffer per warp (3 CTAs per SM by shared memory)
constexpr uint32_t TILE_SCRATCH = (sizeof(MapScratch) + 31u) & ~15u;  // map / comment scratch, placed behind the line
constexpr int TILE_GROUP = 16;                            // records prepared at once
constexpr int TILE_ROW_BYTES = 116;                       // 29 words: the lanes' rows start on distinct banks
constexpr uint32_t TILE_CHAN_CACHE = 1024;                // the warp's current channel blob (shared by ~100 consecutive records)
constexpr uint32_t TILE_CFG_CACHE = 256;                  // the context strings (label, created_at, capture_time)
constexpr int TILE_EARLY = 8;                             // 128-byte string chunks loaded before the shared-memory phases
// row layout: msgno @0 (16) | chat id @16 (24) | views @40 (12) | shares @52 (12) | comments @64 (12) | time @76 (28) | lengths @104 (8)
__device__ __constant__ uint8_t kTileFieldOff[8] = {0, 16, 40, 52, 64, 76, 0, 0};

struct TileShared {
  uint32_t tmpl[kTgNWords];
  uint16_t wmeta[kTgNWords];
  uint32_t pieces[kTgNEnt];
  uint32_t ptype[TGI_CT__COUNT * 8];  // MessageContentType() strings, 32 bytes each, zero padded
  uint32_t vshift[TILE_WARPS][kTgNEnt];  // literal pieces: line offset - template offset (VSHIFT_SKIP: absent); others: line offset
  __align__(16) uint8_t cfgc[TILE_CFG_CACHE];
  __align__(16) uint8_t chan[TILE_WARPS][TILE_CHAN_CACHE];
  __align__(4) uint8_t rows[TILE_WARPS][TILE_GROUP][TILE_ROW_BYTES];
  __align__(128) uint8_t buf[TILE_WARPS][TILE_BUF];
};

// does the record need the map scratch behind its line (a non-empty reactions map or a comment list)?
DEVI bool tg_needs_scratch(uint32_t nreact, bool comments_nil, uint32_t c0, uint32_t c1) { return nreact != 0 || (!comments_nil && c1 != c0); }

struct TileStream {
  uint64_t gs;     // global address of buffer byte 0 (16-byte aligned)
  uint64_t pos;    // global address of the next output byte
  uint32_t head;   // bytes at the front of the buffer's first block that belong to another writer
  uint32_t buf_s;  // shared-space address of the buffer
};

DEVI void ts_begin(TileStream& t, uint64_t addr) {
  t.gs = addr & ~15ull;
  t.pos = addr;
  t.head = (uint32_t)addr & 15u;
}

// Retire the window.  end == false: the stream continues (the incomplete last block moves to the front of the
// buffer); end == true: the range ends here (or a line written by somebody else follows): the last block is stored
// byte-exact.  Warp-collective.
DEVI void ts_flush(TileStream& t, bool end) {
  const uint32_t l = lane_id();
  const uint32_t used = (uint32_t)(t.pos - t.gs);
  // every lane's st.shared must be visible to the async proxy before the bulk store reads the buffer
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t lo = 0;
  if (t.head) {  // first block: bytes [head, min(16, used)) are ours
    const uint32_t hi = used < 16u ? used : 16u;
    if (l >= t.head && l < hi) *(uint8_t*)(uintptr_t)(t.gs + l) = (uint8_t)lds8(t.buf_s + l);
    lo = 16;
  }
  const uint32_t full = used & ~15u;
  if (full > lo) {
    if (l == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(t.gs + lo), "r"(t.buf_s + lo), "r"(full - lo) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  const uint32_t tail0 = full > lo ? full : lo;  // first byte not covered above
  if (end) {
    if (used > tail0 && l < used - tail0) *(uint8_t*)(uintptr_t)(t.gs + tail0 + l) = (uint8_t)lds8(t.buf_s + tail0 + l);
    if (l == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer is refilled next
    __syncwarp();
    return;
  }
  // continue: carry [tail0, used) to the front.  The bulk store may still be reading the front of the buffer.
  uint32_t carry = 0;
  if (used > tail0 && l < used - tail0) carry = lds8(t.buf_s + tail0 + l);
  if (l == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  if (used > tail0 && l < used - tail0) sts8(t.buf_s + l, carry);
  t.gs += tail0;
  t.head = 0;
  __syncwarp();
}

// lane-parallel preparation of one record: the rendered fields of its line; its strings are requested from HBM now
// (L2 prefetch), long before the warp copies them
DEVI void tile_prep(uint8_t* row, const tgi_tg_rec* rec, const uint8_t* strs, bool comments_nil, uint32_t ncomments, int32_t tz) {
  {
    const uint8_t* p = strs + rec->str_off;
    const uint32_t tot = rec->text_len + rec->alt_len + rec->media_len + rec->handle_len;
    for (uint32_t o = 0; o < tot + 127u && o < 1024u; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
#pragma unroll 1
  for (int f = 0; f < 5; f++) {
    const int64_t v = f == 0 ? rec->id / 1048576  // tdutils.go:1008
                             : f == 1 ? rec->chat_id
                                      : f == 2 ? (int64_t)rec->view_count
                                               : f == 3 ? (int64_t)rec->share_count : (comments_nil ? 0 : (int64_t)ncomments);
    row[104 + f] = (uint8_t)render_i64(row + kTileFieldOff[f], v);
  }
  row[104 + F_TIME] = (uint8_t)render_time(row + 76, rec->date, 0, tz);  // tdutils.go:417
}

struct TileIn {
  const uint32_t* xlen;       // [n][8]
  const uint32_t* link_start;
  const uint32_t* link_count;
  const tgi_link* arena;
  int* err;
};

// n bytes shared -> shared, any alignment (line pieces that come from the warp's caches)
DEVI void copy_ss(uint32_t dst, uint32_t src, uint32_t n) {
  for (uint32_t i = lane_id(); i < n; i += 32) sts8(dst + i, lds8(src + i));
}
// n bytes global -> shared with eight loads in flight per lane (the plain byte loop pays one memory latency per 32 bytes)
DEVI void copy_gs_deep(uint32_t dst, const uint8_t* src, uint32_t n) {
  const uint32_t l = lane_id();
  for (uint32_t base = 0; base < n; base += 1024) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t o = base + 128u * k + 4u * l;
      v[k] = o < n ? ld_u32_unaligned(src + o) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t o = base + 128u * k + 4u * l;
      if (o < n) {
        sts8(dst + o, v[k]);
        if (o + 1 < n) sts8(dst + o + 1, v[k] >> 8);
        if (o + 2 < n) sts8(dst + o + 2, v[k] >> 16);
        if (o + 3 < n) sts8(dst + o + 3, v[k] >> 24);
      }
    }
  }
}

// map[string]int that the size pass flagged XLF_SIMPLE_MAP: <= LANE_MAP_MAX entries, keys of 1..8 bytes that need no
// escaping (repeated keys allowed: the last one wins).  One lane per entry: two loads (entry, key), rank and offset by shuffles, direct byte stores.
DEVI void tile_emit_simple_map(uint32_t dst, const tgi_reaction* reacts, uint32_t r0, uint32_t nr, const uint8_t* aux) {
  const uint32_t l = lane_id();
  uint32_t k0 = 0, k1 = 0, kl = 0, len = 0;
  int32_t cnt = 0;
  if (l < nr) {
    const tgi_reaction rc = reacts[r0 + l];
    const uint8_t* kp = aux + rc.emoji_off;
    kl = rc.emoji_len;
    cnt = rc.count;
    k0 = ld_u32_unaligned(kp);
    k1 = kl > 4 ? ld_u32_unaligned(kp + 4) : 0u;
    if (kl < 4) k0 &= (1u << (8u * kl)) - 1u;
    if (kl > 4 && kl < 8) k1 &= (1u << (8u * (kl - 4u))) - 1u;
    len = 4u + kl + ndigits_i64(cnt);  // "key":n and a comma or the closing brace
  }
  const uint64_t ck = ((uint64_t)__byte_perm(k0, 0, 0x0123) << 32) | __byte_perm(k1, 0, 0x0123);  // bytewise order
  // a later entry with the same key overwrites this one (Go map assignment, tdutils.go:598); keys hold no NUL byte,
  // so equal compare keys are equal keys
  bool live = l < nr;
  for (uint32_t j = 1; j < nr; j++) {
    const uint64_t cj = __shfl_sync(FULL, ck, (int)j);
    if (l < j && cj == ck) live = false;
  }
  if (!live) len = 0;
  const uint32_t nlive = __popc(__ballot_sync(FULL, live));
  uint32_t off = 1, rank = 0;
  for (uint32_t j = 0; j < nr; j++) {
    const uint64_t cj = __shfl_sync(FULL, ck, (int)j);
    const uint32_t lj = __shfl_sync(FULL, len, (int)j);
    if (lj && cj < ck) {
      off += lj;
      rank++;
    }
  }
  if (l == 0) sts8(dst, '{');
  if (live) {
    uint32_t d = dst + off;
    sts8(d, '"');
    for (uint32_t i = 0; i < kl; i++) sts8(d + 1 + i, (i < 4 ? k0 >> (8u * i) : k1 >> (8u * (i - 4u))));
    d += 1 + kl;
    sts8(d, '"');
    sts8(d + 1, ':');
    d += 2;
    uint32_t nd = len - 4u - kl;
    uint32_t v = cnt < 0 ? (uint32_t)0 - (uint32_t)cnt : (uint32_t)cnt;
    if (cnt < 0) {
      sts8(d, '-');
      d++;
      nd--;
    }
    for (uint32_t i = nd; i-- > 0;) {
      const uint32_t q = v / 10u;
      sts8(d + i, '0' + (v - q * 10u));
      v = q;
    }
    sts8(d + nd, rank + 1 == nlive ? '}' : ',');
  }
}

// per-warp state that survives from record to record
struct TileWarp {
  uint32_t chan_idx;   // channel whose blob sits in the warp's cache (0xffffffff: none)
  bool chan_cached;    // ... and it fitted
  bool cfg_cached;     // the context strings fitted the CTA's cache
};

// One record, the whole warp.  The line [t.pos, t.pos + total) is assembled in the buffer; t.pos advances.
DEVI void tile_emit_record(TileShared& sh, int wid, TileStream& t, TileWarp& tw, const TgBatchDev& b, const CfgDev& cfg, uint64_t r,
                           uint32_t row_s, uint32_t total, const TileIn& in, uint64_t& bytes_in) {
  const int l = lane_id();
  const tgi_tg_rec* rec = &b.recs[r];
  TgWalkArgs a;
  a.b = &b;
  a.cfg = &cfg;
  a.r = r;
  a.v.rec = rec;
  a.v.text = b.strs + rec->str_off;
  a.v.text_len = rec->text_len;
  a.v.alt = a.v.text + a.v.text_len;
  a.v.alt_len = rec->alt_len;
  a.v.media = a.v.alt + a.v.alt_len;
  a.v.media_len = rec->media_len;
  a.v.handle = a.v.media + a.v.media_len;
  a.v.handle_len = rec->handle_len;
  a.v.ct = rec->content_type;
  a.v.flags = rec->flags;
  a.v.e0 = a.v.e1 = 0;
  const uint32_t chan_idx = rec->chan_idx;
  const ChanDerived cd = b.chan_derived[chan_idx];
  const TgDerived d = tg_derive(a, cd);
  const uint32_t condmask = tg_condmask(a, d);
  const uint32_t ct = a.v.ct;
  const uint4 xa = *(const uint4*)(in.xlen + r * 8), xb = *(const uint4*)(in.xlen + r * 8 + 4);
  const uint32_t r0 = b.react_off[r], nr = b.react_off[r + 1] - r0;
  const bool scratch = tg_needs_scratch(nr, d.comments_nil, d.c0, d.c1);

  // ---- 0. the record's clean strings: up to TILE_EARLY 128-byte chunks are requested NOW and stored after the
  //         shared-memory phases (one memory latency per record instead of one per 32 bytes) ----
  const bool desc_clean = xa.x != 0 && xa.x == d.desc_len, media_clean = xa.y != 0 && xa.y == a.v.media_len,
             handle_clean = xa.z != 0 && xa.z == a.v.handle_len;
  const uint32_t desc_early = desc_clean ? min(d.desc_len, 128u * (TILE_EARLY - 2)) : 0u;
  const uint32_t media_early = media_clean ? min((uint32_t)a.v.media_len, 128u) : 0u;
  const uint32_t handle_early = handle_clean ? min((uint32_t)a.v.handle_len, 128u) : 0u;
  uint32_t ev[TILE_EARLY];
#pragma unroll
  for (int k = 0; k < TILE_EARLY - 2; k++) {
    const uint32_t o = 128u * k + 4u * (uint32_t)l;
    ev[k] = o < desc_early ? ld_u32_unaligned(d.desc + o) : 0u;
  }
  ev[TILE_EARLY - 2] = 4u * (uint32_t)l < media_early ? ld_u32_unaligned(a.v.media + 4 * l) : 0u;
  ev[TILE_EARLY - 1] = 4u * (uint32_t)l < handle_early ? ld_u32_unaligned(a.v.handle + 4 * l) : 0u;

  uint32_t used = (uint32_t)(t.pos - t.gs);
  if (used + total + (scratch ? TILE_SCRATCH : 0u) > TILE_BUF) {
    ts_flush(t, false);
    used = (uint32_t)(t.pos - t.gs);
  }
  const uint32_t line_s = t.buf_s + used;
  MapScratch* ms = (MapScratch*)(sh.buf[wid] + ((used + total + 15u) & ~15u));

  // the channel's pre-rendered strings: kept in shared memory while consecutive records share the channel
  const uint32_t chan_bytes = pad16(cd.user_len) + pad16(cd.name_len) + pad16(cd.title_len) + pad16(cd.cdata_len);
  const uint32_t chan_s = smem_addr(sh.chan[wid]);
  if (tw.chan_idx != chan_idx) {
    tw.chan_idx = chan_idx;
    tw.chan_cached = chan_bytes <= TILE_CHAN_CACHE;
    if (tw.chan_cached) {
      const uint4* src = (const uint4*)(b.chan_blob + cd.off);  // segments are 16-byte aligned and zero padded
      __syncwarp();
      for (uint32_t i = l; i < chan_bytes / 16u; i += 32) {
        const uint4 v = __ldg(src + i);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(chan_s + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      __syncwarp();
    }
  }

  // ---- 1. piece table ----
  const uint32_t vs_s = smem_addr(sh.vshift[wid]);
  uint32_t en[kTgEPL], len[kTgEPL], off[kTgEPL];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    const uint32_t e = sh.pieces[kTgEPL * l + k];
    const uint32_t kind = e & 15u, arg = (e >> 4) & 15u;
    uint32_t ln = 0;
    if ((condmask >> ((e >> 8) & 15u)) & 1u) {
      if (kind == K_LIT) ln = e >> 23;
      else if (kind == K_FIELD) ln = lds8(row_s + 104u + arg);
      else if (kind == K_POSTTYPE) ln = kPostTypeLen[ct];
      else if (kind == K_CHAN) ln = arg == 0 ? cd.user_len : arg == 1 ? cd.name_len : arg == 2 ? cd.title_len : cd.cdata_len;
      else if (kind == K_CFG) ln = arg == 0 ? cfg.label_len : arg == 1 ? cfg.created_tg_len : arg == 2 ? cfg.created_yt_len : cfg.capture_len;
      else if (kind == K_ESC) ln = arg == 0 ? xa.x : arg == 1 ? xa.y : arg == 2 ? xa.z : xa.w;
      else if (kind == K_COMMENTS) ln = xb.x;
      else if (kind == K_REACTIONS) ln = xb.y;
      else if (kind == K_OUTLINKS) ln = xb.z;
    }
    en[k] = e;
    len[k] = ln;
    sum += ln;
  }
  const uint32_t incl = warp_incl_scan(sum);
  uint32_t run = incl - sum;
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    off[k] = run;
    run += len[k];
  }
  if (__shfl_sync(FULL, incl, 31) != total) {  // sizing and emission disagree: never expected; the host reports it
    if (l == 0) atomicOr(in.err, 16);
    t.pos += total;
    return;
  }
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    const uint32_t kind = en[k] & 15u;
    uint32_t v = off[k];
    if (kind == K_LIT) v = len[k] ? off[k] - ((en[k] >> 12) & 0x7FFu) : VSHIFT_SKIP;
    sts32(vs_s + 4u * (uint32_t)(kTgEPL * l + k), v);
  }
  __syncwarp();

  // ---- 2. literals: the template, word by word ----
  {
    const uint32_t wm_s = smem_addr(sh.wmeta), tm_s = smem_addr(sh.tmpl);
#pragma unroll 2
    for (uint32_t j = l; j < (uint32_t)kTgNWords; j += 32) {
      const uint32_t m = lds16(wm_s + 2u * j);
      const uint32_t sh_ = lds32(vs_s + 4u * (m >> 3));
      if (sh_ != VSHIFT_SKIP) {
        const uint32_t v = lds32(tm_s + 4u * j);
        const uint32_t dst = line_s + 4u * j + sh_;  // sh_ may be "negative" mod 2^32
        const uint32_t nv = m & 7u;
        sts8(dst, v);
        if (nv > 1) sts8(dst + 1, v >> 8);
        if (nv > 2) sts8(dst + 2, v >> 16);
        if (nv > 3) sts8(dst + 3, v >> 24);
      }
    }
  }
  // ---- 3. rendered fields, by the owning lanes (sources are word aligned in shared memory) ----
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    const uint32_t kind = en[k] & 15u, arg = (en[k] >> 4) & 15u;
    if ((kind == K_FIELD || kind == K_POSTTYPE) && len[k]) {
      const uint32_t src = kind == K_FIELD ? row_s + kTileFieldOff[arg] : smem_addr(sh.ptype) + 32u * ct;
      const uint32_t dst = line_s + off[k], n = len[k];
      for (uint32_t w = 0; w < n; w += 4) {
        const uint32_t v = lds32(src + w);
        sts8(dst + w, v);
        if (w + 1 < n) sts8(dst + w + 1, v >> 8);
        if (w + 2 < n) sts8(dst + w + 2, v >> 16);
        if (w + 3 < n) sts8(dst + w + 3, v >> 24);
      }
    }
  }
  // ---- 4. strings, maps, lists: the cooperative pieces in line order ----
  uint32_t copied = 0;
#pragma unroll 1
  for (int bi = 0; bi < kTgNBig; bi++) {
    const uint32_t idx = kTgBig[bi];
    const uint32_t e = sh.pieces[idx];
    const uint32_t kind = e & 15u, arg = (e >> 4) & 15u;
    if (!((condmask >> ((e >> 8) & 15u)) & 1u)) continue;
    const uint32_t dst_s = line_s + lds32(vs_s + 4u * idx);
    const DstS dst{dst_s};
    if (kind == K_CHAN) {
      const uint32_t o = arg == 0 ? 0u : arg == 1 ? pad16(cd.user_len) : arg == 2 ? pad16(cd.user_len) + pad16(cd.name_len)
                                                             : pad16(cd.user_len) + pad16(cd.name_len) + pad16(cd.title_len);
      const uint32_t n = arg == 0 ? cd.user_len : arg == 1 ? cd.name_len : arg == 2 ? cd.title_len : cd.cdata_len;
      if (tw.chan_cached) copy_ss(dst_s, chan_s + o, n);
      else copy_gs_deep(dst_s, b.chan_blob + cd.off + o, n);
      copied += n;
    } else if (kind == K_CFG) {
      const uint32_t n = arg == 0 ? cfg.label_len : arg == 1 ? cfg.created_tg_len : arg == 2 ? cfg.created_yt_len : cfg.capture_len;
      if (tw.cfg_cached) copy_ss(dst_s, smem_addr(sh.cfgc) + cfg.off[arg], n);
      else copy_gs_deep(dst_s, cfg.blob + cfg.off[arg], n);
      copied += n;
    } else if (kind == K_ESC) {
      const uint8_t* sp = arg == XL_DESC ? d.desc : arg == XL_MEDIA ? a.v.media : arg == XL_HANDLE ? a.v.handle : a.v.alt;
      const uint32_t sn = arg == XL_DESC ? d.desc_len : arg == XL_MEDIA ? a.v.media_len : arg == XL_HANDLE ? a.v.handle_len : a.v.alt_len;
      const uint32_t xl = arg == 0 ? xa.x : arg == 1 ? xa.y : arg == 2 ? xa.z : xa.w;
      if (xl == 0) continue;
      copied += sn;
      if (xl == sn) {  // clean: the chunks requested at the top, then whatever is left
        const uint32_t early = arg == XL_DESC ? desc_early : arg == XL_MEDIA ? media_early : arg == XL_HANDLE ? handle_early : 0u;
        if (arg == XL_DESC) {
#pragma unroll
          for (int k = 0; k < TILE_EARLY - 2; k++) {
            const uint32_t o = 128u * k + 4u * (uint32_t)l;
            if (o < early) {
              sts8(dst_s + o, ev[k]);
              if (o + 1 < early) sts8(dst_s + o + 1, ev[k] >> 8);
              if (o + 2 < early) sts8(dst_s + o + 2, ev[k] >> 16);
              if (o + 3 < early) sts8(dst_s + o + 3, ev[k] >> 24);
            }
          }
        } else if (arg == XL_MEDIA || arg == XL_HANDLE) {
          const uint32_t v = arg == XL_MEDIA ? ev[TILE_EARLY - 2] : ev[TILE_EARLY - 1];
          const uint32_t o = 4u * (uint32_t)l;
          if (o < early) {
            sts8(dst_s + o, v);
            if (o + 1 < early) sts8(dst_s + o + 1, v >> 8);
            if (o + 2 < early) sts8(dst_s + o + 2, v >> 16);
            if (o + 3 < early) sts8(dst_s + o + 3, v >> 24);
          }
        }
        if (sn > early) copy_gs_deep(dst_s + early, sp + early, sn - early);
      } else if (arg == XL_DESC && !(xb.w & XLF_DESC_EXACT)) {
        esc_ascii_to(dst, sp, sn);
      } else {
        esc_to(dst, sp, sn);
      }
    } else if (kind == K_COMMENTS) {
      if (d.comments_nil) {
        if (l < 4) sts8(dst_s + l, (0x6c6c756eu >> (8u * l)) & 0xFFu);  // null
      } else if (d.c1 == d.c0) {
        put2(dst, '[', ']');
      } else {
        emit_tg_comments_to(dst, ms, b, d.c0, d.c1);
      }
    } else if (kind == K_REACTIONS) {
      if (nr == 0) put2(dst, '{', '}');
      else if (xb.w & XLF_SIMPLE_MAP) tile_emit_simple_map(dst_s, b.reacts, r0, nr, b.aux);
      else emit_reaction_map_to(dst, ms, b.reacts, r0, r0 + nr, b.aux);
    } else {  // K_OUTLINKS
      const uint32_t nl = in.link_count[r];
      if (nl) emit_tg_outlinks_to(dst, in.arena + in.link_start[r], nl);
      copied += xb.z;
    }
  }
  if (l == 0) bytes_in += copied + 64u + 32u + 8u;  // + header, piece lengths, line offset
  t.pos += total;
  __syncwarp();
}

}  // namespace tgi
