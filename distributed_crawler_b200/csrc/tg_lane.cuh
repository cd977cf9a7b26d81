// tg_lane.cuh — lane-per-record emission of the fixed part of the Telegram Post line.
//
// One warp takes 32 consecutive records and every lane streams ITS OWN line: the scalar work
// (number / time rendering, piece lengths, offsets) that kept 31 lanes idle in the warp-per-record
// walker now runs 32 records wide, and the bytes leave the SM as 16-byte aligned vector stores.
// The line is the same generated piece program as everywhere else (tools/gen_pieces.py); control
// flow is uniform across the warp (same piece sequence for every record), only lengths, sources
// and the output phase differ per lane.
//
// A lane's output is a byte stream at an arbitrary address: LaneStream keeps the bytes of the
// current 16-byte block in four registers.  Full blocks are staged in the lane's own shared-memory
// row and drained to HBM by per-lane TMA bulk stores (cp.async.bulk global <- shared::cta): 32 lanes
// writing 32 different lines with ST.128 cost the LSU 32 wavefronts per instruction and reached
// 1.2 TB/s in a micro-benchmark, the same rows drained by bulk stores 4.8-6.1 TB/s
// (scratch/tma_bench.cu, profiles/README.md).  Blocks shared with somebody else (the neighbouring
// line, or a variable piece written by the esc / maps kernels) are stored byte-exact (store_bytes),
// so the kernels never overwrite each other's bytes.
//
// Reference semantics: telegramhelper/tdutils.go:633-717 (field map) + encoding/json of model.Post
// (model/data.go:9-75); see tg_walk.cuh.
#pragma once
#include "tg_walk.cuh"

namespace tgi {

constexpr uint32_t LANE_STAGE_HALF = 128;                       // bytes per staging half (one bulk store)
constexpr uint32_t LANE_STAGE_ROW = 2 * LANE_STAGE_HALF + 16;   // two halves; +16 spreads the rows over the banks

struct LaneStream {
  uint64_t pos;             // absolute address of the next output byte
  uint32_t c0, c1, c2, c3;  // bytes [head, pos & 15) of the current block, zero elsewhere
  uint32_t head;            // first byte of the current block that belongs to this stream
  uint64_t seg;             // global address of the first block staged in the active half
  uint32_t row_s, half_s;   // shared-space address of the lane's staging row / of its active half
  uint32_t fill;            // bytes staged in the active half
};

// hand the staged blocks to the TMA and switch halves; the half we switch to was read by the group
// before the one committed here, so at most that one may still be pending
DEVI void ls_drain(LaneStream& s) {
  if (s.fill) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(s.seg), "r"(s.half_s), "r"(s.fill) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    s.half_s = s.half_s == s.row_s ? s.row_s + LANE_STAGE_HALF : s.row_s;
    s.fill = 0;
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  }
}
DEVI void ls_stage(LaneStream& s, uint64_t blk) {
  if (s.fill == 0) s.seg = blk;
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(s.half_s + s.fill), "r"(s.c0), "r"(s.c1), "r"(s.c2), "r"(s.c3) : "memory");
  s.fill += 16;
  if (s.fill == LANE_STAGE_HALF) ls_drain(s);
}

// bytes [lo, hi) of the 16-byte block at blk (16-byte aligned)
__device__ __noinline__ void store_bytes(uint64_t blk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t lo,
                                         uint32_t hi) {
  const uint32_t w[4] = {c0, c1, c2, c3};
#pragma unroll
  for (uint32_t i = 0; i < 4; i++) {
    const uint32_t a = 4u * i;
    if (lo <= a && a + 4u <= hi) {
      *(uint32_t*)(uintptr_t)(blk + a) = w[i];
    } else {
#pragma unroll
      for (uint32_t k = 0; k < 4; k++)
        if (lo <= a + k && a + k < hi) *(uint8_t*)(uintptr_t)(blk + a + k) = (uint8_t)(w[i] >> (8u * k));
    }
  }
}

// once per kernel: the active half must survive from one record to the next, because the bulk store
// of the previous record may still be reading the other one
DEVI void ls_init(LaneStream& s, uint32_t row_s) {
  s.row_s = s.half_s = row_s;
  s.fill = 0;
  s.seg = 0;
}
DEVI void ls_begin(LaneStream& s, uint64_t addr) {
  s.pos = addr;
  s.head = (uint32_t)addr & 15u;
  s.c0 = s.c1 = s.c2 = s.c3 = 0;
}

// append n (1..16) bytes held little-endian in w (zero beyond n)
DEVI void ls_append(LaneStream& s, uint4 w, uint32_t n) {
  const uint32_t ph = (uint32_t)s.pos & 15u, sh = (ph & 3u) * 8u;
  // shift left by ph bytes into an 8-word window: first the byte part ...
  const uint32_t v0 = w.x << sh, v1 = __funnelshift_l(w.x, w.y, sh), v2 = __funnelshift_l(w.y, w.z, sh),
                 v3 = __funnelshift_l(w.z, w.w, sh), v4 = __funnelshift_l(w.w, 0u, sh);
  // ... then the word part (two select levels)
  const bool b0 = (ph & 4u) != 0, b1 = (ph & 8u) != 0;
  const uint32_t z0 = b0 ? 0u : v0, z1 = b0 ? v0 : v1, z2 = b0 ? v1 : v2, z3 = b0 ? v2 : v3, z4 = b0 ? v3 : v4,
                 z5 = b0 ? v4 : 0u;
  s.c0 |= b1 ? 0u : z0;
  s.c1 |= b1 ? 0u : z1;
  s.c2 |= b1 ? z0 : z2;
  s.c3 |= b1 ? z1 : z3;
  if (ph + n >= 16u) {
    const uint64_t blk = s.pos & ~15ull;
    if (s.head == 0) {
      ls_stage(s, blk);
    } else {
      store_bytes(blk, s.c0, s.c1, s.c2, s.c3, s.head, 16u);
      s.head = 0;
    }
    s.c0 = b1 ? z2 : z4;
    s.c1 = b1 ? z3 : z5;
    s.c2 = b1 ? z4 : 0u;
    s.c3 = b1 ? z5 : 0u;
  }
  s.pos += n;
}

// write out what the current block holds (before a gap / at the end of the line)
DEVI void ls_flush(LaneStream& s) {
  ls_drain(s);
  const uint32_t ph = (uint32_t)s.pos & 15u;
  if (ph > s.head) store_bytes(s.pos & ~15ull, s.c0, s.c1, s.c2, s.c3, s.head, ph);
  s.c0 = s.c1 = s.c2 = s.c3 = 0;
  s.head = ph;
}

// leave g bytes to another writer
DEVI void ls_skip(LaneStream& s, uint32_t g) {
  if (g) {
    ls_flush(s);
    s.pos += g;
    s.head = (uint32_t)s.pos & 15u;
  }
}

// ---- the kernel's shared state ------------------------------------------------------------------------
constexpr int LANE_ROW_WORDS = 36;  // per lane: 8 blocks of rendered fields + 8 length bytes; 36 = 4 * odd
                                    // keeps the lanes' LDS.128 on distinct banks
constexpr int LANE_WARPS = 8;
struct LaneShared {
  uint4 tmpl[kTgLaneTemplateLen / 16];
  uint4 ptype[TGI_CT__COUNT * 2];  // MessageContentType() strings, 32 bytes each, zero padded
  uint32_t pieces[64];
  uint32_t rows[LANE_WARPS][32][LANE_ROW_WORDS];
  uint4 stage[LANE_WARPS][32][LANE_STAGE_ROW / 16];
};

DEVI void lane_shared_fill(LaneShared& sh) {
  for (int i = threadIdx.x; i < kTgLaneTemplateLen / 16; i += blockDim.x) sh.tmpl[i] = ((const uint4*)kTgLaneTemplate)[i];
  for (int i = threadIdx.x; i < kTgLaneNPieces; i += blockDim.x) sh.pieces[i] = kTgLanePieces[i];
  uint32_t* pt = (uint32_t*)sh.ptype;
  for (int i = threadIdx.x; i < TGI_CT__COUNT * 8; i += blockDim.x) {
    const int ct = i >> 3, w = i & 7;
    pt[i] = w < 7 ? ((const uint32_t*)kPostType[ct])[w] : 0u;
  }
}

// One lane, one record.  `active` lanes emit; the others only keep the warp's control flow company.
DEVI void emit_tg_lane(LaneShared& sh, uint32_t* row, LaneStream& s, const TgBatchDev& b, const CfgDev& cfg, uint64_t r, bool active,
                       uint8_t* out, const uint64_t* line_off, const uint32_t* xlen_g, uint32_t* xpos_g, int* err) {
  TgWalkArgs a;
  a.b = &b;
  a.cfg = &cfg;
  a.r = r;
  a.links = nullptr;
  a.n_links = 0;
  {  // the record header, as load_rec_view (kernels.cuh)
    const tgi_tg_rec* rec = &b.recs[r];
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
  }
  const tgi_tg_rec* rec = a.v.rec;
  const ChanDerived cd = b.chan_derived[rec->chan_idx];
  const TgDerived d = tg_derive(a, cd);
  const uint32_t condmask = active ? tg_condmask(a, d) : 0u;
  const uint32_t ct = a.v.ct;

  // rendered fields: block 0 msgno, 1-2 chat id, 3 views, 4 shares, 5 comments, 6-7 time; lengths at byte 128+
  uint8_t* rb = (uint8_t*)row;
#pragma unroll
  for (int k = 0; k < 8; k++) ((uint4*)row)[k] = make_uint4(0, 0, 0, 0);
#pragma unroll 1
  for (int f = 0; f < 5; f++) {
    const int64_t v = f == 0 ? rec->id / 1048576  // tdutils.go:1008
                             : f == 1 ? rec->chat_id
                                      : f == 2 ? (int64_t)rec->view_count : f == 3 ? (int64_t)rec->share_count : d.ncomments;
    const uint32_t fo = f ? 16u + (f >= 2 ? 16u * f : 0u) : 0u;
    rb[128 + f] = (uint8_t)render_i64(rb + fo, v);
  }
  rb[128 + F_TIME] = (uint8_t)render_time(rb + 96, rec->date, 0, cfg.tz);  // tdutils.go:417

  const uint64_t line_start = (uint64_t)(uintptr_t)out + line_off[r];
  const uint32_t total = (uint32_t)(line_off[r + 1] - line_off[r]);
  ls_begin(s, active ? line_start : 0ull);

#pragma unroll 1
  for (int p = 0; p < kTgLaneNPieces; p++) {
    const uint32_t en = sh.pieces[p];
    const uint32_t kind = en & 15u, arg = (en >> 4) & 15u;
    const bool on = ((condmask >> ((en >> 8) & 15u)) & 1u) != 0;
    if (kind == K_LIT || kind == K_FIELD || kind == K_POSTTYPE) {  // shared-memory sources
      const uint4* src;
      uint32_t n;
      if (kind == K_LIT) {
        src = sh.tmpl + ((en >> 12) & 0x7FFu);
        n = en >> 23;
      } else if (kind == K_FIELD) {
        src = (const uint4*)(rb + (arg ? 16u + (arg >= 2 ? 16u * arg : 0u) : 0u));
        n = rb[128 + arg];
      } else {
        src = sh.ptype + 2u * ct;
        n = kPostTypeLen[ct];
      }
      if (!on) n = 0;
      const uint32_t nmax = __reduce_max_sync(FULL, n);
      for (uint32_t i = 0; i < nmax; i += 16, src++)
        if (i < n) ls_append(s, *src, min(16u, n - i));
    } else if (kind == K_CHAN || kind == K_CFG) {  // global sources, 16-byte aligned and zero padded
      const uint4* src;
      uint32_t n;
      if (kind == K_CHAN) {
        const uint32_t o = arg == 0 ? 0u : arg == 1 ? pad16(cd.user_len) : arg == 2 ? pad16(cd.user_len) + pad16(cd.name_len)
                                                               : pad16(cd.user_len) + pad16(cd.name_len) + pad16(cd.title_len);
        src = (const uint4*)(b.chan_blob + cd.off + o);
        n = arg == 0 ? cd.user_len : arg == 1 ? cd.name_len : arg == 2 ? cd.title_len : cd.cdata_len;
      } else {
        src = (const uint4*)(cfg.blob + (arg == 0 ? cfg.off[0] : arg == 1 ? cfg.off[1] : arg == 2 ? cfg.off[2] : cfg.off[3]));
        n = arg == 0 ? cfg.label_len : arg == 1 ? cfg.created_tg_len : arg == 2 ? cfg.created_yt_len : cfg.capture_len;
      }
      if (!on) n = 0;
      const uint32_t nmax = __reduce_max_sync(FULL, n);
      for (uint32_t i = 0; i < nmax; i += 16, src++)
        if (i < n) ls_append(s, __ldg(src), min(16u, n - i));
    } else {  // pieces written by the esc / maps kernels: remember where they go
      const uint32_t xi = kind == K_ESC ? arg : kind == K_COMMENTS ? (uint32_t)XL_COMMENTS
                                                 : kind == K_REACTIONS ? (uint32_t)XL_REACTIONS : (uint32_t)XL_OUTLINKS;
      uint32_t g = 0;
      if (active) {
        xpos_g[xi] = (uint32_t)(s.pos - line_start);
        if (on) g = xlen_g[xi];
      }
      ls_skip(s, g);
    }
  }
  ls_flush(s);
  if (active && (uint32_t)(s.pos - line_start) != total) atomicOr(err, 16);  // sizing and emission disagree: never expected
}

}  // namespace tgi
