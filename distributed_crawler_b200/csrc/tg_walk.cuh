// tg_walk.cuh — the Telegram Post line: closed-form length (parse kernel) and table-driven emission
// (emit kernel), both derived from the same generated piece table (tools/gen_pieces.py).
//
// Replaces telegramhelper/tdutils.go:380-732 ParseMessage (field map :633-717) followed by
// json.Marshal(post)+'\n' (state/storageproviders.go:276-282, state/daprstate.go:1118-1120) for
// model.Post (model/data.go:9-75).  Key order = struct declaration order; see SURVEY Appendix A.6
// for the encoding/json rules restated in dev_common.cuh.
#pragma once
#include <cstddef>

#include "dev_common.cuh"
#include "tg_links.cuh"

namespace tgi {

// per-channel strings pre-rendered once per batch by the channel job
struct ChanDerived {
  uint64_t off;        // into chan_blob: esc_user | esc_name | "esc_title" | cdata
  uint32_t user_len;   // JSON-escaped ActiveUsernames[0] (0 = no public link)
  uint32_t name_len;   // JSON-escaped channelName
  uint32_t title_len;  // "esc(chat.Title)" with quotes
  uint32_t cdata_len;  // ,"channel_name":...,"published_at":"0001-01-01T00:00:00Z"}
};

struct TgBatchDev {
  uint64_t n;
  const tgi_tg_rec* recs;
  const uint8_t* strs;
  const uint32_t* ent_off;
  const tgi_entity* ents;
  const uint32_t* react_off;
  const tgi_reaction* reacts;
  const uint32_t* comment_off;
  const tgi_comment* comments;
  const uint8_t* aux;
  uint32_t n_chans;
  const tgi_tg_chan* chans;
  const uint8_t* chan_strs;
  const ChanDerived* chan_derived;
  const uint8_t* chan_blob;
};

struct CfgDev {             // per-context constants in a small device blob
  const uint8_t* blob;      // label_esc | created_tg | created_yt | capture
  uint32_t label_len;       // JSON-escaped crawl_label (no quotes)
  uint32_t created_tg_len;  // quoted RFC3339 of created_at.UTC().Truncate(s)
  uint32_t created_yt_len;  // quoted RFC3339Nano of created_at in the local zone
  uint32_t capture_len;     // quoted RFC3339Nano of capture_time
  uint32_t flags;           // TGI_CFG_*; bit 31: injected clock not representable (Marshal error)
  int32_t tz;
  int64_t min_post_date;
};
#define CFGDEV_CLOCK_INVALID 0x80000000u

__device__ const char kPostType[TGI_CT__COUNT][28] = {
    "unknown",          "messageText",          "messageVideo",           "messagePhoto",
    "messageAnimation", "messageAnimatedEmoji", "messagePoll",            "messageGiveaway",
    "messagePaidMedia", "messageSticker",       "messageGiveawayWinners", "messageGiveawayCompleted",
    "messageVideoNote", "messageDocument",      "messageAudio",           "messageVoiceNote",
    ""};
__device__ const uint8_t kPostTypeLen[TGI_CT__COUNT] = {7, 11, 12, 12, 16, 20, 11, 15, 16, 14, 22, 24, 16, 15, 12, 16, 0};

// per-warp shared scratch of the emit kernel
struct WarpScratch {
  uint8_t field[8][40];    // rendered numeric / time fields
  uint32_t flen[8];        // their lengths
  uint64_t src_ptr[8];     // global sources: chan segments 0..3, cfg segments 4..7
  uint32_t src_len[8];
  uint32_t xlen[8];        // emitted lengths of the variable pieces (XL_*), computed by the parse kernel
  uint32_t vshift[64];     // per piece: output offset - template offset (literal pieces), or VSHIFT_SKIP
  uint8_t rslot[32][64];   // per-lane rendered map entries  "key":count
};

enum { K_LIT, K_FIELD, K_CHAN, K_CFG, K_ESC, K_POSTTYPE, K_COMMENTS, K_REACTIONS, K_OUTLINKS };
enum { C_NONE, C_USER, C_ALBUM, C_CT_OTHER, C_NOT_CT_OTHER, C_HAS_MEDIA };
enum { F_MSGNO, F_CHAT, F_VIEW, F_SHARE, F_NCOMM, F_TIME, F_POSTTYPE };
enum { XL_DESC, XL_MEDIA, XL_HANDLE, XL_ALT, XL_COMMENTS, XL_REACTIONS, XL_OUTLINKS, XL_COUNT };
constexpr uint32_t K_NOP = 15;
#include "tg_pieces.inc"

// block-shared copies of the template and the lane-parallel entry table (filled once per CTA)
#define VSHIFT_SKIP 0x80000000u
struct CtaShared {
  uint32_t ents[kTgNEnt];
  uint32_t tmpl[kTgNWords];
  uint16_t wmeta[kTgNWords];
};

// ---- map[string]int (reactions) ------------------------------------------------------------------
// encoding/json sorts map keys bytewise; later duplicates of a key overwrite earlier ones (Go map
// assignment, tdutils.go:598).  Up to 32 entries per map (checked in the parse kernel).
DEVI int key_cmp(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
  uint32_t m = la < lb ? la : lb;
  for (uint32_t i = 0; i < m; i++) {
    uint32_t x = ldb(a + i), y = ldb(b + i);
    if (x != y) return x < y ? -1 : 1;
  }
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

struct MapLane {  // one map entry per lane
  const uint8_t* kp;
  uint32_t kl;
  int32_t cnt;
  bool live;       // last occurrence of its key
  uint32_t rank;   // position among the live keys in bytewise order
  uint32_t nlive;
};
DEVI MapLane warp_map_prepare(const tgi_reaction* reacts, uint32_t r0, uint32_t r1, const uint8_t* aux, bool want_rank) {
  int l = lane_id();
  uint32_t n = r1 - r0;
  if (n > 32) n = 32;
  MapLane m;
  m.kp = nullptr;
  m.kl = 0;
  m.cnt = 0;
  if ((uint32_t)l < n) {
    tgi_reaction rc = reacts[r0 + l];
    m.kp = aux + rc.emoji_off;
    m.kl = rc.emoji_len;
    m.cnt = rc.count;
  }
  m.live = (uint32_t)l < n;
  for (uint32_t j = 1; j < n; j++) {
    const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)m.kp, j);
    uint32_t lj = __shfl_sync(FULL, m.kl, j);
    if ((uint32_t)l < j && m.live && key_cmp(m.kp, m.kl, pj, lj) == 0) m.live = false;
  }
  uint32_t livemask = __ballot_sync(FULL, m.live);
  m.nlive = __popc(livemask);
  m.rank = 0;
  if (want_rank) {
    for (uint32_t j = 0; j < n; j++) {
      if (!((livemask >> j) & 1u)) continue;
      const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)m.kp, j);
      uint32_t lj = __shfl_sync(FULL, m.kl, j);
      if (m.live && (uint32_t)l != j && key_cmp(pj, lj, m.kp, m.kl) < 0) m.rank++;
    }
  }
  return m;
}

__device__ __noinline__ uint32_t size_reaction_map(const tgi_reaction* reacts, uint32_t r0, uint32_t r1, const uint8_t* aux) {
  if (r1 == r0) return 2;
  MapLane m = warp_map_prepare(reacts, r0, r1, aux, false);
  uint32_t mine = m.live ? 3u + thread_esc_len(m.kp, m.kl) + ndigits_i64(m.cnt) : 0u;  // "key":n
  return 2u + warp_sum(mine) + (m.nlive - 1);
}

template <bool STREAM>
__device__ __noinline__ Em emit_reaction_map(Em e, WarpScratch* ws, const tgi_reaction* reacts, uint32_t r0, uint32_t r1,
                                             const uint8_t* aux) {
  if (r1 == r0) {
    em_ch2(e, '{', '}');
    return e;
  }
  int l = lane_id();
  MapLane m = warp_map_prepare(reacts, r0, r1, aux, true);
  // every live lane renders its own  "key":count  into its slot; long keys fall back to the warp path
  uint32_t slot = smem_addr(ws->rslot[l]);
  uint32_t sl = 0;
  __syncwarp();
  if (m.live) {
    sts8(slot, '"');
    uint32_t k = thread_esc(m.kp, m.kl, slot + 1, 64 - 1 - 2 - 11);
    if (k != ~0u) {
      sts8(slot + 1 + k, '"');
      sts8(slot + 2 + k, ':');
      sl = 3 + k + (uint32_t)render_i64(ws->rslot[l] + 3 + k, m.cnt);
    } else {
      sl = ~0u;
    }
  }
  __syncwarp();
  em_ch(e, '{');
  for (uint32_t r = 0; r < m.nlive; r++) {
    uint32_t who = __ballot_sync(FULL, m.live && m.rank == r);
    int src = __ffs(who) - 1;
    uint32_t len = __shfl_sync(FULL, sl, src);
    if (r) em_ch(e, ',');
    if (len != ~0u) {
      em_copy_s(e, smem_addr(ws->rslot[src]), len);
    } else {  // key too long for a slot: escape it cooperatively
      const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)m.kp, src);
      uint32_t lj = __shfl_sync(FULL, m.kl, src);
      int32_t cj = __shfl_sync(FULL, m.cnt, src);
      em_ch(e, '"');
      if (STREAM) em_esc_stream(e, pj, lj); else em_esc_fit(e, pj, lj);
      em_ch2(e, '"', ':');
      uint32_t dl = 0;
      __syncwarp();
      if (l == 0) dl = (uint32_t)render_i64(ws->field[7], cj);
      __syncwarp();
      dl = __shfl_sync(FULL, dl, 0);
      em_copy_s(e, smem_addr(ws->field[7]), dl);
      __syncwarp();
    }
  }
  em_ch(e, '}');
  return e;
}

// ---- []model.Comment --------------------------------------------------------------------------------
__device__ const char kCm0[] = "{\"text\":\"";
__device__ const char kCm1[] = "\",\"reactions\":";
__device__ const char kCm2[] = ",\"view_count\":";
__device__ const char kCm3[] = ",\"reply_count\":";
__device__ const char kCm4[] = ",\"handle\":\"";
__device__ const char kCm5[] = "\"}";
__device__ const char kNullLit[] = "null";
constexpr uint32_t kCmFixed = sizeof(kCm0) + sizeof(kCm1) + sizeof(kCm2) + sizeof(kCm3) + sizeof(kCm4) + sizeof(kCm5) - 6;

__device__ __noinline__ uint32_t size_tg_comments(const TgBatchDev& b, uint32_t c0, uint32_t c1) {
  uint32_t tot = 2 + (c1 > c0 ? c1 - c0 - 1 : 0);  // [ ] and commas
  for (uint32_t k = c0; k < c1; k++) {
    tgi_comment cm = b.comments[k];
    tot += kCmFixed + warp_esc_len(b.aux + cm.text_off, cm.text_len) + warp_esc_len(b.aux + cm.handle_off, cm.handle_len) +
           ndigits_i64(cm.view_count) + ndigits_i64(cm.reply_count) +
           ((cm.flags & 1) ? size_reaction_map(b.reacts, cm.react_start, cm.react_start + cm.react_count, b.aux) : 4u);
  }
  return tot;
}

template <bool STREAM>
__device__ __noinline__ Em emit_tg_comments(Em e, WarpScratch* ws, const TgBatchDev& b, uint32_t c0, uint32_t c1) {
  int l = lane_id();
  em_ch(e, '[');
  for (uint32_t k = c0; k < c1; k++) {
    tgi_comment cm = b.comments[k];
    uint32_t dl = 0;
    __syncwarp();
    if (l < 2) dl = (uint32_t)render_i64(ws->field[6 + l], l == 0 ? cm.view_count : cm.reply_count);
    __syncwarp();
    uint32_t d0 = __shfl_sync(FULL, dl, 0), d1 = __shfl_sync(FULL, dl, 1);
    if (k > c0) em_ch(e, ',');
    em_copy_g(e, (const uint8_t*)kCm0, sizeof(kCm0) - 1);
    if (STREAM) em_esc_stream(e, b.aux + cm.text_off, cm.text_len); else em_esc_fit(e, b.aux + cm.text_off, cm.text_len);
    em_copy_g(e, (const uint8_t*)kCm1, sizeof(kCm1) - 1);
    if (cm.flags & 1) e = emit_reaction_map<STREAM>(e, ws, b.reacts, cm.react_start, cm.react_start + cm.react_count, b.aux);
    else em_copy_g(e, (const uint8_t*)kNullLit, 4);
    em_copy_g(e, (const uint8_t*)kCm2, sizeof(kCm2) - 1);
    em_copy_s(e, smem_addr(ws->field[6]), d0);
    em_copy_g(e, (const uint8_t*)kCm3, sizeof(kCm3) - 1);
    em_copy_s(e, smem_addr(ws->field[7]), d1);
    em_copy_g(e, (const uint8_t*)kCm4, sizeof(kCm4) - 1);
    if (STREAM) em_esc_stream(e, b.aux + cm.handle_off, cm.handle_len); else em_esc_fit(e, b.aux + cm.handle_off, cm.handle_len);
    em_copy_g(e, (const uint8_t*)kCm5, sizeof(kCm5) - 1);
    __syncwarp();
  }
  em_ch(e, ']');
  return e;
}

__device__ __noinline__ Em emit_tg_outlinks(Em e, const tgi_link* links, uint32_t n) {
  for (uint32_t k = 0; k < n; k++) {
    if (k) em_ch2(e, ',', '"'); else em_ch(e, '"');
    em_copy_g(e, links[k].name, links[k].len);  // [a-z0-9_] only: no escaping needed
    em_ch(e, '"');
  }
  return e;
}
DEVI uint32_t size_tg_outlinks(const tgi_link* links, uint32_t n) {
  if (!n) return 0;
  uint32_t s = 0;
  for (uint32_t k = lane_id(); k < n; k += 32) s += links[k].len + 2u;
  return warp_sum(s) + (n - 1);
}

// ---- record ---------------------------------------------------------------------------------------
struct TgWalkArgs {
  const TgBatchDev* b;
  const CfgDev* cfg;
  uint64_t r;
  TgRecView v;
  const tgi_link* links;  // this record's outlinks (arena)
  uint32_t n_links;
};

struct TgDerived {  // what both passes need to know about a record
  const uint8_t* desc;
  uint32_t desc_len;
  bool has_media, has_user, album, comments_nil;
  uint32_t c0, c1;
  int64_t ncomments;
};
DEVI TgDerived tg_derive(const TgWalkArgs& a, const ChanDerived& cd) {
  TgDerived d;
  const TgBatchDev& b = *a.b;
  d.c0 = b.comment_off[a.r];
  d.c1 = b.comment_off[a.r + 1];
  d.comments_nil = (a.v.flags & TGI_RF_COMMENTS_NIL) != 0;
  d.ncomments = d.comments_nil ? 0 : (int64_t)(d.c1 - d.c0);
  // description / media by content type (tdutils.go:443-587)
  d.desc = nullptr;
  d.desc_len = 0;
  uint32_t ct = a.v.ct;
  if (ct == TGI_CT_TEXT || ct == TGI_CT_VIDEO || ct == TGI_CT_PHOTO || ct == TGI_CT_ANIMATION) {
    if (a.v.flags & TGI_RF_HAS_TEXT) { d.desc = a.v.text; d.desc_len = a.v.text_len; }
  } else if (ct == TGI_CT_ANIMATED_EMOJI || ct == TGI_CT_POLL || ct == TGI_CT_GIVEAWAY ||
             ct == TGI_CT_PAID_MEDIA || ct == TGI_CT_DOCUMENT) {
    d.desc = a.v.alt; d.desc_len = a.v.alt_len;
  }
  d.has_media = ct == TGI_CT_VIDEO || ct == TGI_CT_VIDEO_NOTE || ct == TGI_CT_DOCUMENT;
  d.has_user = cd.user_len != 0;
  d.album = d.has_user && a.v.rec->media_album_id != 0;
  return d;
}

// line length in bytes; 0 if a time field is not representable (Marshal error -> TGI_ST_NOLINE).
// The formula's coefficients come from the generated piece table, so it cannot drift from emit.
DEVI uint32_t size_tg_record(const TgWalkArgs& a, uint32_t* xl) {
  const TgBatchDev& b = *a.b;
  const CfgDev& cfg = *a.cfg;
  const tgi_tg_rec* rec = a.v.rec;
  const ChanDerived cd = b.chan_derived[rec->chan_idx];
  TgDerived d = tg_derive(a, cd);
  if (cfg.flags & CFGDEV_CLOCK_INVALID) return 0;
  // int32 dates are always inside year [0,9999]: RFC3339 with quotes, 'Z' or a +hh:mm offset
  uint32_t L[8] = {ndigits_i64(rec->id / 1048576), ndigits_i64(rec->chat_id), ndigits_i64(rec->view_count),
                   ndigits_i64(rec->share_count), ndigits_i64(d.ncomments), cfg.tz == 0 ? 22u : 27u, 0, 0};
  uint32_t chan[4] = {cd.user_len, cd.name_len, cd.title_len, cd.cdata_len};
  uint32_t cf[4] = {cfg.label_len, cfg.created_tg_len, cfg.created_yt_len, cfg.capture_len};
  uint32_t tot = tg_size_fixed(L, chan, cf, d.has_user, d.album);
  xl[XL_DESC] = warp_esc_len(d.desc, d.desc_len);
  xl[XL_ALT] = a.v.ct == TGI_CT_OTHER ? warp_esc_len(a.v.alt, a.v.alt_len) : 0u;
  xl[XL_MEDIA] = d.has_media ? warp_esc_len(a.v.media, a.v.media_len) : 0u;
  xl[XL_HANDLE] = warp_esc_len(a.v.handle, a.v.handle_len);
  xl[XL_COMMENTS] = d.comments_nil ? 4u : size_tg_comments(b, d.c0, d.c1);
  xl[XL_REACTIONS] = size_reaction_map(b.reacts, b.react_off[a.r], b.react_off[a.r + 1], b.aux);
  xl[XL_OUTLINKS] = size_tg_outlinks(a.links, a.n_links);
  tot += a.v.ct == TGI_CT_OTHER ? 0u : (uint32_t)kPostTypeLen[a.v.ct];
  for (int j = 0; j < XL_COUNT; j++) tot += xl[j];
  return tot;
}

// prologue of both emit paths: lanes render the numeric / time fields and fill the source tables
DEVI void emit_tg_prologue(WarpScratch* ws, const TgWalkArgs& a, const ChanDerived& cd, const TgDerived& d,
                           const uint32_t* xlen_g) {
  const TgBatchDev& b = *a.b;
  const CfgDev& cfg = *a.cfg;
  const tgi_tg_rec* rec = a.v.rec;
  int l = lane_id();
  __syncwarp();
  if (l < 5) {
    int64_t v = l == 0 ? rec->id / 1048576                                     // tdutils.go:1008
                       : l == 1 ? rec->chat_id
                                : l == 2 ? (int64_t)rec->view_count
                                         : l == 3 ? (int64_t)rec->share_count : d.ncomments;
    ws->flen[l] = (uint32_t)render_i64(ws->field[l], v);
  } else if (l == 5) {
    ws->flen[5] = (uint32_t)render_time(ws->field[5], rec->date, 0, cfg.tz);  // :417
  } else if (l == 6) {  // MessageContentType() string (28-byte rows, 4-byte aligned)
    const uint32_t* src = (const uint32_t*)kPostType[a.v.ct];
    uint32_t* dst = (uint32_t*)ws->field[F_POSTTYPE];
#pragma unroll
    for (int w = 0; w < 7; w++) dst[w] = src[w];
    ws->flen[F_POSTTYPE] = kPostTypeLen[a.v.ct];
  } else if (l >= 8 && l < 12) {
    int k = l - 8;
    uint32_t o = k == 0 ? 0u : k == 1 ? cd.user_len : k == 2 ? cd.user_len + cd.name_len : cd.user_len + cd.name_len + cd.title_len;
    ws->src_ptr[k] = (uint64_t)(uintptr_t)(b.chan_blob + cd.off + o);
    ws->src_len[k] = k == 0 ? cd.user_len : k == 1 ? cd.name_len : k == 2 ? cd.title_len : cd.cdata_len;
  } else if (l >= 12 && l < 16) {
    int k = l - 12;
    uint32_t o = k == 0 ? 0u : k == 1 ? cfg.label_len : k == 2 ? cfg.label_len + cfg.created_tg_len
                                                               : cfg.label_len + cfg.created_tg_len + cfg.created_yt_len;
    ws->src_ptr[4 + k] = (uint64_t)(uintptr_t)(cfg.blob + o);
    ws->src_len[4 + k] = k == 0 ? cfg.label_len : k == 1 ? cfg.created_tg_len : k == 2 ? cfg.created_yt_len : cfg.capture_len;
  } else if (l >= 16 && l < 16 + XL_COUNT && xlen_g) {
    ws->xlen[l - 16] = xlen_g[l - 16];
  }
  __syncwarp();
}

DEVI uint32_t tg_condmask(const TgWalkArgs& a, const TgDerived& d) {
  return 1u | (d.has_user ? 1u << C_USER : 0) | (d.album ? 1u << C_ALBUM : 0) |
         (a.v.ct == TGI_CT_OTHER ? 1u << C_CT_OTHER : 1u << C_NOT_CT_OTHER) | (d.has_media ? 1u << C_HAS_MEDIA : 0);
}

// sequential path (any line length): walk the piece table, streaming through the staging buffer
__device__ __noinline__ Em emit_tg_record_seq(Em e, WarpScratch* ws, const TgWalkArgs& a) {
  const TgBatchDev& b = *a.b;
  const ChanDerived cd = b.chan_derived[a.v.rec->chan_idx];
  TgDerived d = tg_derive(a, cd);
  const uint32_t ws_s = smem_addr(ws);
  emit_tg_prologue(ws, a, cd, d, nullptr);
  const uint32_t ct = a.v.ct;
  const uint32_t condmask = tg_condmask(a, d);
  for (int pi = 0; pi < kTgNPieces; pi++) {
    const uint32_t pc = kTgPieces[pi];
    const uint32_t kind = pc & 15u, arg = (pc >> 4) & 15u;
    if (!((condmask >> ((pc >> 8) & 15u)) & 1u)) continue;
    if (kind == K_LIT) {
      em_copy_g(e, (const uint8_t*)kTgTemplate + ((pc >> 12) & 0x7FFu), pc >> 23);
    } else if (kind == K_FIELD) {
      em_copy_s(e, ws_s + (uint32_t)offsetof(WarpScratch, field) + 40u * arg,
                lds32(ws_s + (uint32_t)offsetof(WarpScratch, flen) + 4u * arg));
    } else if (kind == K_CHAN || kind == K_CFG) {
      uint32_t idx = (kind == K_CFG ? 4u : 0u) + arg;
      em_copy_g_long(e, (const uint8_t*)(uintptr_t)ws->src_ptr[idx], ws->src_len[idx]);
    } else if (kind == K_ESC) {
      const uint8_t* p = arg == 0 ? d.desc : arg == 1 ? a.v.media : arg == 2 ? a.v.handle : a.v.alt;
      uint32_t n = arg == 0 ? d.desc_len : arg == 1 ? a.v.media_len : arg == 2 ? a.v.handle_len : a.v.alt_len;
      em_esc_stream(e, p, n);
    } else if (kind == K_POSTTYPE) {
      em_copy_g(e, (const uint8_t*)kPostType[ct], kPostTypeLen[ct]);
    } else if (kind == K_COMMENTS) {
      if (d.comments_nil) em_copy_g(e, (const uint8_t*)kNullLit, 4);
      else e = emit_tg_comments<true>(e, ws, b, d.c0, d.c1);
    } else if (kind == K_REACTIONS) {
      e = emit_reaction_map<true>(e, ws, b.reacts, b.react_off[a.r], b.react_off[a.r + 1], b.aux);
    } else {
      e = emit_tg_outlinks(e, a.links, a.n_links);
    }
  }
  return e;
}

// ---- lane-parallel path, in phases ------------------------------------------------------------------
// The whole line (total bytes, fill + total <= EMIT_FLUSH_AT) is assembled out of order in the staging
// buffer.  Lane i owns pieces kTgEPL*i ..; an exclusive scan over the lanes' length sums gives each
// piece its output offset.  Literal pieces only publish their shift (output offset - template
// offset); the template is then copied word by word (340 words = 11 steps) with the shift of the
// owning piece.  Rendered fields are copied by their owning lane, the variable pieces cooperatively.
//
// The work is cut into phases that the emit kernel separates with __syncthreads(): all 24 warps of
// the (single) CTA of an SM execute the same few KB of code at any time.  The B200 instruction caches
// are small (L0 ~6 KB per sub-partition, L1.5 32 KB per SM); without this, 24 warps at 24 different
// places of a 70 KB kernel stall mostly on instruction fetch (ncu: stall_no_instruction).
struct FastRec {
  TgDerived d;
  uint32_t off[kTgEPL], len[kTgEPL];
  uint32_t base;   // shared address of the line's first byte
  bool ok;
};

DEVI void fast_phase_prologue(FastRec& f, WarpScratch* ws, const TgWalkArgs& a, const uint32_t* xlen_g) {
  const ChanDerived cd = a.b->chan_derived[a.v.rec->chan_idx];
  f.d = tg_derive(a, cd);
  emit_tg_prologue(ws, a, cd, f.d, xlen_g);
}

// entries + scan + shifts + lane-owned field copies + template + CHAN/CFG copies
DEVI void fast_phase_fixed(FastRec& f, const Em& e, WarpScratch* ws, const CtaShared* cs, const TgWalkArgs& a,
                           uint32_t total, int* err) {
  const uint32_t ws_s = smem_addr(ws), ents_s = smem_addr(cs->ents), tmpl_s = smem_addr(cs->tmpl);
  const uint32_t vs_s = ws_s + (uint32_t)offsetof(WarpScratch, vshift);
  const uint32_t condmask = tg_condmask(a, f.d);
  const int l = lane_id();
  uint32_t ent[kTgEPL];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    uint32_t en = lds32(ents_s + 4u * (uint32_t)(kTgEPL * l + k));
    uint32_t kind = en & 15u, arg = (en >> 4) & 15u;
    uint32_t ln = 0;
    if ((condmask >> ((en >> 8) & 15u)) & 1u) {
      if (kind == K_LIT) ln = en >> 23;
      else if (kind == K_FIELD) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, flen) + 4u * arg);
      else if (kind == K_POSTTYPE) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, flen) + 4u * F_POSTTYPE);
      else if (kind == K_CHAN) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, src_len) + 4u * arg);
      else if (kind == K_CFG) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, src_len) + 4u * (4u + arg));
      else if (kind == K_ESC) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, xlen) + 4u * arg);
      else if (kind == K_COMMENTS) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, xlen) + 4u * XL_COMMENTS);
      else if (kind == K_REACTIONS) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, xlen) + 4u * XL_REACTIONS);
      else if (kind == K_OUTLINKS) ln = lds32(ws_s + (uint32_t)offsetof(WarpScratch, xlen) + 4u * XL_OUTLINKS);
    }
    ent[k] = en;
    f.len[k] = ln;
    sum += ln;
  }
  uint32_t incl = warp_incl_scan(sum);
  uint32_t run = incl - sum;
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    f.off[k] = run;
    run += f.len[k];
  }
  f.base = e.sbuf + e.fill;
  f.ok = __shfl_sync(FULL, incl, 31) == total;
  if (!f.ok) {  // sizing and emission disagree: never expected; the host reports it
    if (l == 0) atomicOr(err, 16);
    return;
  }
  const uint32_t base = f.base;
#pragma unroll
  for (int k = 0; k < kTgEPL; k++) {
    uint32_t kind = ent[k] & 15u;
    if (kind == K_LIT) {
      sts32(vs_s + 4u * (uint32_t)(kTgEPL * l + k), f.len[k] ? f.off[k] - ((ent[k] >> 12) & 0x7FFu) : VSHIFT_SKIP);
    } else if (f.len[k] && (kind == K_FIELD || kind == K_POSTTYPE)) {
      uint32_t src = ws_s + (uint32_t)offsetof(WarpScratch, field) + 40u * (kind == K_FIELD ? (ent[k] >> 4) & 15u : F_POSTTYPE);
      uint32_t dst = base + f.off[k], n = f.len[k];
      for (uint32_t w = 0; w < n; w += 4) {
        uint32_t v = lds32(src + w);
        sts8(dst + w, v);
        if (w + 1 < n) sts8(dst + w + 1, v >> 8);
        if (w + 2 < n) sts8(dst + w + 2, v >> 16);
        if (w + 3 < n) sts8(dst + w + 3, v >> 24);
      }
    }
  }
  __syncwarp();
  {  // the template, word by word
    const uint32_t wm_s = smem_addr(cs->wmeta);
#pragma unroll 2
    for (uint32_t j = l; j < (uint32_t)kTgNWords; j += 32) {
      uint32_t m = lds16(wm_s + 2u * j);
      uint32_t sh = lds32(vs_s + 4u * (m >> 3));
      if (sh != VSHIFT_SKIP) {
        uint32_t v = lds32(tmpl_s + 4u * j);
        uint32_t dst = base + 4u * j + sh, nv = m & 7u;
        sts8(dst, v);
        if (nv > 1) sts8(dst + 1, v >> 8);
        if (nv > 2) sts8(dst + 2, v >> 16);
        if (nv > 3) sts8(dst + 3, v >> 24);
      }
    }
  }
  for (int bi = 0; bi < kTgNCopy; bi++) {  // per-channel / per-context strings
    const uint32_t idx = kTgCopy[bi];
    const uint32_t owner = idx / kTgEPL, kk = idx % kTgEPL;
    uint32_t o_sel = f.off[0], l_sel = f.len[0];
#pragma unroll
    for (int k = 1; k < kTgEPL; k++)
      if (kk == (uint32_t)k) { o_sel = f.off[k]; l_sel = f.len[k]; }
    const uint32_t ln = __shfl_sync(FULL, l_sel, owner);
    if (ln == 0) continue;
    const uint32_t o = __shfl_sync(FULL, o_sel, owner);
    const uint32_t en = kTgPieces[idx];
    uint32_t si = ((en & 15u) == K_CFG ? 4u : 0u) + ((en >> 4) & 15u);
    copy_g_to(base + o, (const uint8_t*)(uintptr_t)ws->src_ptr[si], ln);
  }
}

DEVI void fast_piece_pos(const FastRec& f, uint32_t idx, uint32_t& o, uint32_t& ln) {
  const uint32_t owner = idx / kTgEPL, kk = idx % kTgEPL;
  uint32_t o_sel = f.off[0], l_sel = f.len[0];
#pragma unroll
  for (int k = 1; k < kTgEPL; k++)
    if (kk == (uint32_t)k) { o_sel = f.off[k]; l_sel = f.len[k]; }
  ln = __shfl_sync(FULL, l_sel, owner);
  o = __shfl_sync(FULL, o_sel, owner);
}

DEVI void fast_phase_esc(const FastRec& f, const TgWalkArgs& a) {
  if (!f.ok) return;
  for (int bi = 0; bi < kTgNEsc; bi++) {
    const uint32_t idx = kTgEsc[bi];
    uint32_t o, ln;
    fast_piece_pos(f, idx, o, ln);
    if (ln == 0) continue;
    const uint32_t arg = (kTgPieces[idx] >> 4) & 15u;
    const uint8_t* p = arg == 0 ? f.d.desc : arg == 1 ? a.v.media : arg == 2 ? a.v.handle : a.v.alt;
    uint32_t n = arg == 0 ? f.d.desc_len : arg == 1 ? a.v.media_len : arg == 2 ? a.v.handle_len : a.v.alt_len;
    uint32_t carry = 0;
    esc_range(f.base + o, p, n, 0, n, carry);
  }
}

DEVI void fast_phase_maps(const FastRec& f, const Em& e, WarpScratch* ws, const TgWalkArgs& a) {
  if (!f.ok) return;
  const TgBatchDev& b = *a.b;
  for (int bi = 0; bi < kTgNMap; bi++) {
    const uint32_t idx = kTgMap[bi];
    uint32_t o, ln;
    fast_piece_pos(f, idx, o, ln);
    if (ln == 0) continue;
    const uint32_t kind = kTgPieces[idx] & 15u;
    Em t = e;
    t.fill = e.fill + o;  // never reaches EMIT_FLUSH_AT: the caller checked fill + total
    if (kind == K_COMMENTS) {  // records with comments never reach this kernel: "null" or "[]"
      if (f.d.comments_nil) em_copy_g(t, (const uint8_t*)kNullLit, 4);
      else em_ch2(t, '[', ']');
    } else if (kind == K_REACTIONS) {
      t = emit_reaction_map<false>(t, ws, b.reacts, b.react_off[a.r], b.react_off[a.r + 1], b.aux);
    } else {
      t = emit_tg_outlinks(t, a.links, a.n_links);
    }
  }
}

// ---- channel job: the per-channel constant strings, rendered once per batch ----------------------
// blob layout per channel: esc(username) | esc(channelName) | "esc(title)" | channel_data tail
__device__ const char kCd0[] = ",\"channel_name\":\"";
__device__ const char kCd1[] = "\",\"channel_description\":\"\",\"channel_profile_image\":\"\",\"channel_engagement_data\":{\"follower_count\":";
__device__ const char kCd2[] = ",\"following_count\":0,\"like_count\":0,\"post_count\":";
__device__ const char kCd3[] = ",\"views_count\":";
__device__ const char kCd4[] = ",\"comment_count\":0,\"share_count\":0},\"channel_url_external\":\"https://t.me/c/";
__device__ const char kCd5[] = "\",\"channel_url\":\"https://t.me/c/";
__device__ const char kCd6[] = "\",\"country_code\":\"\",\"published_at\":\"0001-01-01T00:00:00Z\"}";
constexpr uint32_t kCdFixed = sizeof(kCd0) + sizeof(kCd1) + sizeof(kCd2) + sizeof(kCd3) + sizeof(kCd4) + sizeof(kCd5) + sizeof(kCd6) - 7;

DEVI ChanDerived size_tg_chan(const TgBatchDev& b, uint32_t c) {
  const tgi_tg_chan ch = b.chans[c];
  const uint8_t* cs = b.chan_strs + ch.str_off;
  uint32_t et = warp_esc_len(cs, ch.title_len), en = warp_esc_len(cs + ch.title_len, ch.name_len),
           eu = warp_esc_len(cs + ch.title_len + ch.name_len, ch.user_len);
  ChanDerived d;
  d.off = 0;
  d.user_len = eu;
  d.name_len = en;
  d.title_len = et + 2;
  d.cdata_len = kCdFixed + et + 2 * en + ndigits_i64(ch.member_count) + ndigits_i64(ch.post_count) + ndigits_i64(ch.view_count);
  return d;
}

DEVI Em emit_tg_chan(Em e, WarpScratch* ws, const TgBatchDev& b, uint32_t c) {
  const tgi_tg_chan ch = b.chans[c];
  const uint8_t* cs = b.chan_strs + ch.str_off;
  const uint8_t *title = cs, *name = cs + ch.title_len, *user = name + ch.name_len;
  int l = lane_id();
  __syncwarp();
  if (l < 3) ws->flen[l] = (uint32_t)render_i64(ws->field[l], l == 0 ? ch.member_count : l == 1 ? ch.post_count : ch.view_count);
  __syncwarp();
  uint32_t L0 = ws->flen[0], L1 = ws->flen[1], L2 = ws->flen[2];
  em_esc_stream(e, user, ch.user_len);
  em_esc_stream(e, name, ch.name_len);
  em_ch(e, '"');
  em_esc_stream(e, title, ch.title_len);
  em_ch(e, '"');
  em_copy_g(e, (const uint8_t*)kCd0, sizeof(kCd0) - 1);
  em_esc_stream(e, title, ch.title_len);
  em_copy_g(e, (const uint8_t*)kCd1, sizeof(kCd1) - 1);
  em_copy_s(e, smem_addr(ws->field[0]), L0);
  em_copy_g(e, (const uint8_t*)kCd2, sizeof(kCd2) - 1);
  em_copy_s(e, smem_addr(ws->field[1]), L1);
  em_copy_g(e, (const uint8_t*)kCd3, sizeof(kCd3) - 1);
  em_copy_s(e, smem_addr(ws->field[2]), L2);
  em_copy_g(e, (const uint8_t*)kCd4, sizeof(kCd4) - 1);
  em_esc_stream(e, name, ch.name_len);
  em_copy_g(e, (const uint8_t*)kCd5, sizeof(kCd5) - 1);
  em_esc_stream(e, name, ch.name_len);
  em_copy_g(e, (const uint8_t*)kCd6, sizeof(kCd6) - 1);
  __syncwarp();
  return e;
}

}  // namespace tgi
