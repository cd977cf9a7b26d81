// tg_walk.cuh — the Telegram Post as a piece walk: one templated function that visits every byte
// range of the JSONL line in order.  Instantiated with tgi::Sizer in the parse kernel (line length)
// and with tgi::Emitter in the emit kernel (bytes).
//
// Replaces telegramhelper/tdutils.go:380-732 ParseMessage (field map :633-717) followed by
// json.Marshal(post)+'\n' (state/storageproviders.go:276-282, state/daprstate.go:1118-1120) for
// model.Post (model/data.go:9-75).  Key order = struct declaration order; see SURVEY Appendix A.6
// for the encoding/json rules restated in dev_common.cuh.
#pragma once
#include "dev_common.cuh"
#include "tg_links.cuh"

namespace tgi {

#define LIT(w, str)                                   \
  do {                                                \
    static __device__ const char _lit[] = str;        \
    (w).lit(_lit, (const uint8_t*)_lit);              \
  } while (0)

// per-channel strings pre-rendered once per batch by the channel job (see ChanWalk below)
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

struct CfgDev {           // per-context constants in a small device blob
  const uint8_t* blob;    // label_esc | created_tg | created_yt | capture
  uint32_t label_len;     // JSON-escaped crawl_label (no quotes)
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

// per-warp shared scratch used by the walkers
struct WarpScratch {
  uint8_t field[8][40];   // rendered numeric / time fields
  uint32_t flen[8];       // their lengths
  uint8_t rslot[32][12];  // per-lane rendered reaction counts
};

// the line as data: see tools/gen_pieces.py
struct Piece {
  uint8_t kind, arg, cond, pad;
  uint16_t off, len;
};
enum { K_LIT, K_FIELD, K_CHAN, K_CFG, K_ESC, K_POSTTYPE, K_COMMENTS, K_REACTIONS, K_OUTLINKS };
enum { C_NONE, C_USER, C_ALBUM, C_CT_OTHER, C_NOT_CT_OTHER, C_HAS_MEDIA };
enum { F_MSGNO, F_CHAT, F_VIEW, F_SHARE, F_NCOMM, F_TIME };
#include "tg_pieces.inc"

// ---- map[string]int (reactions) ------------------------------------------------------------------
// encoding/json sorts map keys bytewise; later duplicates of a key overwrite earlier ones (Go map
// assignment, tdutils.go:598).  Up to 32 entries per map (checked by the caller).
DEVI int key_cmp(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb) {
  uint32_t m = la < lb ? la : lb;
  for (uint32_t i = 0; i < m; i++) {
    uint32_t x = ldb(a + i), y = ldb(b + i);
    if (x != y) return x < y ? -1 : 1;
  }
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

template <class W>
__device__ __noinline__ void walk_reaction_map(W& w, WarpScratch* ws, const tgi_reaction* reacts, uint32_t r0,
                                               uint32_t r1, const uint8_t* aux) {
  int l = lane_id();
  uint32_t n = r1 - r0;
  if (n == 0) {
    w.ch('{');
    w.ch('}');
    return;
  }
  if (n > 32) n = 32;
  const uint8_t* kp = nullptr;
  uint32_t kl = 0;
  int32_t cnt = 0;
  if ((uint32_t)l < n) {
    tgi_reaction rc = reacts[r0 + l];
    kp = aux + rc.emoji_off;
    kl = rc.emoji_len;
    cnt = rc.count;
  }
  // last occurrence of each key wins
  bool live = (uint32_t)l < n;
  for (uint32_t j = 1; j < n; j++) {
    const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)kp, j);
    uint32_t lj = __shfl_sync(FULL, kl, j);
    if ((uint32_t)l < j && live && key_cmp(kp, kl, pj, lj) == 0) live = false;
  }
  uint32_t livemask = __ballot_sync(FULL, live);
  uint32_t rank = 0;
  for (uint32_t j = 0; j < n; j++) {
    if (!((livemask >> j) & 1u)) continue;
    const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)kp, j);
    uint32_t lj = __shfl_sync(FULL, kl, j);
    if (live && (uint32_t)l != j && key_cmp(pj, lj, kp, kl) < 0) rank++;
  }
  uint32_t dl = 0;
  __syncwarp();
  if (live) dl = (uint32_t)render_i64(ws->rslot[l], cnt);
  __syncwarp();
  w.ch('{');
  uint32_t m = __popc(livemask);
  for (uint32_t r = 0; r < m; r++) {
    uint32_t who = __ballot_sync(FULL, live && rank == r);
    int src = __ffs(who) - 1;
    const uint8_t* pj = (const uint8_t*)__shfl_sync(FULL, (unsigned long long)kp, src);
    uint32_t lj = __shfl_sync(FULL, kl, src);
    uint32_t dj = __shfl_sync(FULL, dl, src);
    if (r) w.ch(',');
    w.ch('"');
    w.esc(pj, lj);
    w.ch('"');
    w.ch(':');
    w.raw_smem(ws->rslot[src], dj);
  }
  w.ch('}');
}

// ---- record walk ----------------------------------------------------------------------------------
struct TgWalkArgs {
  const TgBatchDev* b;
  const CfgDev* cfg;
  uint64_t r;
  TgRecView v;
  const tgi_link* links;  // this record's outlinks (arena)
  uint32_t n_links;
};

template <class W>
__device__ __noinline__ void walk_tg_comments(W& w, WarpScratch* ws, const TgBatchDev& b, uint32_t c0, uint32_t c1) {
  static __device__ const char k0[] = "{\"text\":\"";
  static __device__ const char k1[] = "\",\"reactions\":";
  static __device__ const char k2[] = ",\"view_count\":";
  static __device__ const char k3[] = ",\"reply_count\":";
  static __device__ const char k4[] = ",\"handle\":\"";
  static __device__ const char k5[] = "\"}";
  static __device__ const char kNull[] = "null";
  int l = lane_id();
  w.ch('[');
  for (uint32_t k = c0; k < c1; k++) {
    tgi_comment cm = b.comments[k];
    uint32_t dl = 0;
    __syncwarp();
    if (l < 2) dl = (uint32_t)render_i64(ws->field[6 + l], l == 0 ? cm.view_count : cm.reply_count);
    __syncwarp();
    uint32_t d0 = __shfl_sync(FULL, dl, 0), d1 = __shfl_sync(FULL, dl, 1);
    if (k > c0) w.ch(',');
    w.raw((const uint8_t*)k0, sizeof(k0) - 1);
    w.esc(b.aux + cm.text_off, cm.text_len);
    w.raw((const uint8_t*)k1, sizeof(k1) - 1);
    if (cm.flags & 1) walk_reaction_map(w, ws, b.reacts, cm.react_start, cm.react_start + cm.react_count, b.aux);
    else w.raw((const uint8_t*)kNull, 4);
    w.raw((const uint8_t*)k2, sizeof(k2) - 1);
    w.raw_smem(ws->field[6], d0);
    w.raw((const uint8_t*)k3, sizeof(k3) - 1);
    w.raw_smem(ws->field[7], d1);
    w.raw((const uint8_t*)k4, sizeof(k4) - 1);
    w.esc(b.aux + cm.handle_off, cm.handle_len);
    w.raw((const uint8_t*)k5, sizeof(k5) - 1);
  }
  w.ch(']');
}

template <class W>
__device__ __noinline__ void walk_tg_outlinks(W& w, const tgi_link* links, uint32_t n) {
  for (uint32_t k = 0; k < n; k++) {
    if (k) w.ch(',');
    w.ch('"');
    w.raw(links[k].name, links[k].len);  // [a-z0-9_] only: no escaping needed
    w.ch('"');
  }
}

// returns false if a time field is not representable (Marshal error -> TGI_ST_NOLINE)
template <class W>
DEVI bool walk_tg_record(W& w, WarpScratch* ws, const TgWalkArgs& a) {
  const TgBatchDev& b = *a.b;
  const CfgDev& cfg = *a.cfg;
  const tgi_tg_rec* rec = a.v.rec;
  int l = lane_id();
  const ChanDerived cd = b.chan_derived[rec->chan_idx];
  const uint8_t* cb = b.chan_blob + cd.off;
  uint32_t c0 = b.comment_off[a.r], c1 = b.comment_off[a.r + 1];
  bool comments_nil = (a.v.flags & TGI_RF_COMMENTS_NIL) != 0;
  int64_t ncomments = comments_nil ? 0 : (int64_t)(c1 - c0);

  // prologue: lanes 0..5 render the numeric / time fields of this record into shared scratch
  __syncwarp();
  if (l < 5) {
    int64_t v = l == 0 ? rec->id / 1048576                                     // tdutils.go:1008
                       : l == 1 ? rec->chat_id
                                : l == 2 ? (int64_t)rec->view_count
                                         : l == 3 ? (int64_t)rec->share_count : ncomments;
    ws->flen[l] = (uint32_t)render_i64(ws->field[l], v);
  } else if (l == 5) {
    ws->flen[5] = (uint32_t)render_time(ws->field[5], rec->date, 0, cfg.tz);  // :417
  }
  __syncwarp();
  if (ws->flen[F_TIME] == 0 || (cfg.flags & CFGDEV_CLOCK_INVALID)) return false;

  // description / media by content type (tdutils.go:443-587)
  const uint8_t* desc = nullptr;
  uint32_t desc_len = 0;
  uint32_t ct = a.v.ct;
  if (ct == TGI_CT_TEXT || ct == TGI_CT_VIDEO || ct == TGI_CT_PHOTO || ct == TGI_CT_ANIMATION) {
    if (a.v.flags & TGI_RF_HAS_TEXT) { desc = a.v.text; desc_len = a.v.text_len; }
  } else if (ct == TGI_CT_ANIMATED_EMOJI || ct == TGI_CT_POLL || ct == TGI_CT_GIVEAWAY ||
             ct == TGI_CT_PAID_MEDIA || ct == TGI_CT_DOCUMENT) {
    desc = a.v.alt; desc_len = a.v.alt_len;
  }
  const bool has_media = ct == TGI_CT_VIDEO || ct == TGI_CT_VIDEO_NOTE || ct == TGI_CT_DOCUMENT;
  const bool has_user = cd.user_len != 0, album = has_user && rec->media_album_id != 0;
  static __device__ const char kNullLit[] = "null";

  for (int pi = 0; pi < kTgNPieces; pi++) {
    const Piece pc = kTgPieces[pi];
    switch (pc.cond) {
      case C_USER: if (!has_user) continue; break;
      case C_ALBUM: if (!album) continue; break;
      case C_CT_OTHER: if (ct != TGI_CT_OTHER) continue; break;
      case C_NOT_CT_OTHER: if (ct == TGI_CT_OTHER) continue; break;
      case C_HAS_MEDIA: if (!has_media) continue; break;
      default: break;
    }
    const uint8_t* src = nullptr;
    uint32_t len = 0;
    int mode = 0;  // 0 copy from global, 1 copy from shared, 2 escape, 3 composite
    switch (pc.kind) {
      case K_LIT: src = (const uint8_t*)kTgTemplate + pc.off; len = pc.len; break;
      case K_FIELD: src = ws->field[pc.arg]; len = ws->flen[pc.arg]; mode = 1; break;
      case K_CHAN: {
        uint32_t o = pc.arg == 0 ? 0u : pc.arg == 1 ? cd.user_len : pc.arg == 2 ? cd.user_len + cd.name_len
                                                                                : cd.user_len + cd.name_len + cd.title_len;
        len = pc.arg == 0 ? cd.user_len : pc.arg == 1 ? cd.name_len : pc.arg == 2 ? cd.title_len : cd.cdata_len;
        src = cb + o;
        break;
      }
      case K_CFG: {
        uint32_t o = pc.arg == 0 ? 0u : pc.arg == 1 ? cfg.label_len : pc.arg == 2 ? cfg.label_len + cfg.created_tg_len
                                                                                  : cfg.label_len + cfg.created_tg_len + cfg.created_yt_len;
        len = pc.arg == 0 ? cfg.label_len : pc.arg == 1 ? cfg.created_tg_len : pc.arg == 2 ? cfg.created_yt_len : cfg.capture_len;
        src = cfg.blob + o;
        break;
      }
      case K_ESC:
        mode = 2;
        if (pc.arg == 0) { src = desc; len = desc_len; }
        else if (pc.arg == 1) { src = a.v.media; len = a.v.media_len; }
        else if (pc.arg == 2) { src = a.v.handle; len = a.v.handle_len; }
        else { src = a.v.alt; len = a.v.alt_len; }
        break;
      case K_POSTTYPE: src = (const uint8_t*)kPostType[ct]; len = kPostTypeLen[ct]; break;
      default: mode = 3; break;
    }
    if (mode == 0) w.raw(src, len);
    else if (mode == 1) w.raw_smem(src, len);
    else if (mode == 2) w.esc(src, len);
    else if (pc.kind == K_COMMENTS) {
      if (comments_nil) w.raw((const uint8_t*)kNullLit, 4);
      else walk_tg_comments(w, ws, b, c0, c1);
    } else if (pc.kind == K_REACTIONS) {
      walk_reaction_map(w, ws, b.reacts, b.react_off[a.r], b.react_off[a.r + 1], b.aux);
    } else {
      walk_tg_outlinks(w, a.links, a.n_links);
    }
  }
  return true;
}

// ---- channel walk: the per-channel constant strings, rendered once per batch -------------------
// segment 0: esc(username)  1: esc(channelName)  2: "esc(title)"  3: channel_data tail
template <class W>
DEVI void walk_tg_chan(W& w, WarpScratch* ws, const TgBatchDev& b, uint32_t c, int seg) {
  const tgi_tg_chan ch = b.chans[c];
  const uint8_t* cs = b.chan_strs + ch.str_off;
  const uint8_t *title = cs, *name = cs + ch.title_len, *user = name + ch.name_len;
  int l = lane_id();
  if (seg == 0) {
    w.esc(user, ch.user_len);
  } else if (seg == 1) {
    w.esc(name, ch.name_len);
  } else if (seg == 2) {
    w.ch('"');
    w.esc(title, ch.title_len);
    w.ch('"');
  } else {
    uint32_t flen = 0;
    __syncwarp();
    if (l == 0) flen = (uint32_t)render_i64(ws->field[0], ch.member_count);
    else if (l == 1) flen = (uint32_t)render_i64(ws->field[1], ch.post_count);
    else if (l == 2) flen = (uint32_t)render_i64(ws->field[2], ch.view_count);
    __syncwarp();
    uint32_t L0 = __shfl_sync(FULL, flen, 0), L1 = __shfl_sync(FULL, flen, 1), L2 = __shfl_sync(FULL, flen, 2);
    LIT(w, ",\"channel_name\":\"");
    w.esc(title, ch.title_len);
    LIT(w, "\",\"channel_description\":\"\",\"channel_profile_image\":\"\",\"channel_engagement_data\":{"
           "\"follower_count\":");
    w.raw_smem(ws->field[0], L0);
    LIT(w, ",\"following_count\":0,\"like_count\":0,\"post_count\":");
    w.raw_smem(ws->field[1], L1);
    LIT(w, ",\"views_count\":");
    w.raw_smem(ws->field[2], L2);
    LIT(w, ",\"comment_count\":0,\"share_count\":0},\"channel_url_external\":\"https://t.me/c/");
    w.esc(name, ch.name_len);
    LIT(w, "\",\"channel_url\":\"https://t.me/c/");
    w.esc(name, ch.name_len);
    LIT(w, "\",\"country_code\":\"\",\"published_at\":\"0001-01-01T00:00:00Z\"}");
  }
}

}  // namespace tgi
