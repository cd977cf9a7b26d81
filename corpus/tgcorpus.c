/*
 * tgcorpus.c — deterministic synthetic corpus generator for the BASELINE.json configs
 * (SURVEY.md §8d).  Host-side measurement infrastructure: it fabricates the already-fetched
 * messages the hot path consumes, directly in the packed batch layout of include/tgingest.h.
 *
 * Determinism: every record draws from its own splitmix64 stream keyed by (seed, global record
 * index), so any shard [first, first+n) can be generated independently on any rank with any thread
 * count and is bit-identical to the same range of a larger run.
 *
 * Profiles: 1 = config 1 (text-only), 2 = config 2 (mixed text/photo/video/... metadata),
 *           3 = configs 3/5 (text-bearing types only, link extraction + frontier).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tgingest.h"

typedef struct {
  uint8_t* p;
  size_t len, cap;
} buf_t;
static void b_reserve(buf_t* b, size_t extra) {
  if (b->len + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : (1 << 16);
  while (nc < b->len + extra) nc *= 2;
  b->p = (uint8_t*)realloc(b->p, nc);
  if (!b->p) abort();
  b->cap = nc;
}
static void b_put(buf_t* b, const void* s, size_t n) {
  b_reserve(b, n);
  memcpy(b->p + b->len, s, n);
  b->len += n;
}
static void b_putc(buf_t* b, uint8_t c) {
  b_reserve(b, 1);
  b->p[b->len++] = c;
}

/* ---- rng -------------------------------------------------------------------------------------- */
typedef struct { uint64_t s; } rng_t;
static uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static uint64_t rnd(rng_t* r) { return mix64(r->s += 0x9e3779b97f4a7c15ull); }
static double rnd01(rng_t* r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }
static uint32_t rnd_n(rng_t* r, uint32_t n) { return (uint32_t)((rnd(r) >> 32) * (uint64_t)n >> 32); }
static double rnd_normal(rng_t* r) {
  double u = rnd01(r), v = rnd01(r);
  if (u < 1e-300) u = 1e-300;
  return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
}
static uint32_t rnd_geom(rng_t* r, double p) { /* >= 1 */
  double u = rnd01(r);
  if (u < 1e-300) u = 1e-300;
  uint32_t k = 1 + (uint32_t)(log(u) / log(1.0 - p));
  return k > 1000000 ? 1000000 : k;
}

/* ---- utf-8 ------------------------------------------------------------------------------------ */
static int put_cp(buf_t* b, uint32_t c) { /* returns UTF-16 units */
  if (c < 0x80) {
    b_putc(b, (uint8_t)c);
  } else if (c < 0x800) {
    b_putc(b, (uint8_t)(0xC0 | (c >> 6)));
    b_putc(b, (uint8_t)(0x80 | (c & 0x3F)));
  } else if (c < 0x10000) {
    b_putc(b, (uint8_t)(0xE0 | (c >> 12)));
    b_putc(b, (uint8_t)(0x80 | ((c >> 6) & 0x3F)));
    b_putc(b, (uint8_t)(0x80 | (c & 0x3F)));
  } else {
    b_putc(b, (uint8_t)(0xF0 | (c >> 18)));
    b_putc(b, (uint8_t)(0x80 | ((c >> 12) & 0x3F)));
    b_putc(b, (uint8_t)(0x80 | ((c >> 6) & 0x3F)));
    b_putc(b, (uint8_t)(0x80 | (c & 0x3F)));
    return 2;
  }
  return 1;
}

enum { SC_LATIN, SC_CYR, SC_ARAB, SC_CJK, SC_EMOJI };
static uint32_t script_cp(rng_t* r, int sc) {
  switch (sc) {
    case SC_CYR: return 0x0430 + rnd_n(r, 32);
    case SC_ARAB: return 0x0627 + rnd_n(r, 36);
    case SC_CJK: return 0x4E00 + rnd_n(r, 0x4000);
    case SC_EMOJI: return rnd_n(r, 3) ? (uint32_t)('a' + rnd_n(r, 26)) : 0x1F600 + rnd_n(r, 80);
    default: {
      uint32_t k = rnd_n(r, 60);
      return k < 52 ? (k < 26 ? 'a' + k : 'A' + (k - 26)) : (k < 58 ? '0' + (k - 52) : (k == 58 ? '-' : '\''));
    }
  }
}

/* ---- names ------------------------------------------------------------------------------------ */
#define NAME_UNIVERSE 20000000u
static const char B32[] = "abcdefghijklmnopqrstuvwxyz012345";
static const char WORDC[] = "abcdefghijklmnopqrstuvwxyz0123456789_";
/* injective map rank -> username: letter + 5 base-32 digits of the rank + hash-derived tail */
static int make_name(uint32_t rank, char* out) {
  uint64_t h = mix64(0xC0FFEEull + rank);
  int n = 0;
  out[n++] = (char)('a' + h % 26);
  h /= 26;
  for (int i = 0; i < 5; i++) out[n++] = B32[(rank >> (5 * i)) & 31];
  int tail = (int)(h % 11);
  h /= 11;
  for (int i = 0; i < tail; i++) {
    out[n++] = WORDC[h % 37];
    h = mix64(h);
  }
  if (out[n - 1] == '_') out[n - 1] = 'x';
  if (n >= 3 && out[n - 3] == 'b' && out[n - 2] == 'o' && out[n - 1] == 't') out[n - 1] = 's';
  return n;
}
/* continuous approximation of Zipf(s) over [1,N] by inverse CDF */
static uint32_t zipf(rng_t* r, uint32_t N, double s) {
  double u = rnd01(r);
  double a = 1.0 - s;
  double x = pow((pow((double)N, a) - 1.0) * u + 1.0, 1.0 / a);
  uint32_t k = (uint32_t)x;
  if (k < 1) k = 1;
  if (k > N) k = N;
  return k - 1;
}
static const char* RESERVED[] = {"joinchat", "addlist", "addstickers", "addtheme", "setlanguage",
                                 "share",    "proxy",   "socks",       "login",    "confirm"};

static int make_link_name(rng_t* r, char* out) {
  double u = rnd01(r);
  int n;
  if (u < 0.06) {
    const char* w = RESERVED[rnd_n(r, 10)];
    n = (int)strlen(w);
    memcpy(out, w, (size_t)n);
    if (rnd_n(r, 5) == 0) out[n++] = 'x'; /* "sharex": NOT reserved (exact match only) */
  } else if (u < 0.10) {
    n = 1 + (int)rnd_n(r, 4);
    for (int i = 0; i < n; i++) out[i] = (char)('a' + rnd_n(r, 26));
  } else if (u < 0.13) {
    n = 33 + (int)rnd_n(r, 12);
    for (int i = 0; i < n; i++) out[i] = WORDC[rnd_n(r, 36)];
    out[0] = (char)('a' + rnd_n(r, 26));
  } else {
    n = make_name(zipf(r, NAME_UNIVERSE, 1.05), out);
    if (u < 0.18) {
      if (n > 28) n = 28;
      if (rnd_n(r, 2)) out[n++] = '_';
      out[n++] = 'b'; out[n++] = 'o'; out[n++] = 't';
    }
  }
  if (rnd01(r) < 0.30)
    for (int i = 0; i < n; i++)
      if (out[i] >= 'a' && out[i] <= 'z' && rnd_n(r, 3) == 0) out[i] = (char)(out[i] - 32);
  return n;
}

/* ---- per-thread output ------------------------------------------------------------------------ */
typedef struct {
  uint64_t seed, first, r0, r1;
  int profile;
  uint32_t n_chans;
  buf_t recs, strs, ents, reacts, comments, aux;
  buf_t ent_cnt, react_cnt, comment_cnt; /* uint32 per record */
  buf_t creacts;                         /* comment reactions (appended after message reactions) */
} gen_t;

static const char* EMOJI[] = {"\xF0\x9F\x91\x8D", "\xE2\x9D\xA4\xEF\xB8\x8F", "\xF0\x9F\x94\xA5", "\xF0\x9F\x98\x82",
                              "\xF0\x9F\x98\xA2", "\xF0\x9F\x8E\x89", "\xF0\x9F\x91\x8F", "\xF0\x9F\xA4\x94",
                              "\xF0\x9F\x98\xA1", "\xF0\x9F\x99\x8F", "\xE2\x9D\xA4", "\xF0\x9F\x91\x8E",
                              "\xF0\x9F\xA4\xA1", "\xF0\x9F\x92\xAF", "\xF0\x9F\x98\x81", "\xF0\x9F\x95\x8A",
                              "\xF0\x9F\x90\xB3", "\xF0\x9F\x8C\x9A", "\xF0\x9F\x8D\xBE", "\xE2\x9A\xA1",
                              "\xF0\x9F\x8F\x86", "\xF0\x9F\x92\x94", "\xF0\x9F\xA4\xA8", "\xF0\x9F\x98\x90",
                              "\xF0\x9F\x8D\x93", "\xF0\x9F\x92\x8B", "\xF0\x9F\x96\x95", "\xF0\x9F\x98\x88",
                              "\xF0\x9F\x98\xB4", "\xF0\x9F\x98\xAD", "\xF0\x9F\xA4\x93", "\xF0\x9F\x91\xBB",
                              "\xF0\x9F\x91\x80", "\xF0\x9F\x8E\x83", "\xF0\x9F\x99\x88", "\xF0\x9F\x98\x87",
                              "\xF0\x9F\x98\xA8", "\xF0\x9F\xA4\x9D", "\xE2\x9C\x8D", "\xF0\x9F\xA4\x97"};

static void gen_reactions(rng_t* r, buf_t* dst, buf_t* aux, uint32_t* count) {
  uint32_t k = 1 + rnd_n(r, 6);
  for (uint32_t i = 0; i < k; i++) {
    const char* e = EMOJI[rnd_n(r, 40)];
    tgi_reaction rc = {(uint32_t)aux->len, (uint16_t)strlen(e), 0, (int32_t)rnd_geom(r, 0.02)};
    b_put(aux, e, strlen(e));
    b_put(dst, &rc, sizeof rc);
  }
  *count += k;
}

/* words of one script until `target` UTF-16 units; returns units written.  Specials inject the
 * JSON-escape alphabet, `bad` injects invalid UTF-8. */
static int gen_words(rng_t* r, buf_t* t, int sc, int target, int specials, int bad) {
  int u = 0;
  while (u < target) {
    int wl = 2 + (int)rnd_n(r, 9);
    for (int i = 0; i < wl && u < target; i++) u += put_cp(t, script_cp(r, sc));
    if (u >= target) break;
    if (specials && rnd_n(r, 6) == 0) {
      static const uint32_t sp[] = {'<', '>', '&', '"', '\\', 1, 8, 12, '\n', '\r', '\t', 0x1f, 0x2028, 0x2029, 0x7f};
      u += put_cp(t, sp[rnd_n(r, 15)]);
    }
    if (bad && rnd_n(r, 8) == 0) {
      static const uint8_t bads[][3] = {{0xFF, 0, 0}, {0xC0, 0x80, 0}, {0x80, 0, 0}, {0xE2, 0x82, 0}, {0xF0, 0x9F, 0x98}, {0xED, 0xA0, 0x80}};
      const uint8_t* bb = bads[rnd_n(r, 6)];
      for (int i = 0; i < 3 && bb[i]; i++) { b_putc(t, bb[i]); u++; }
    }
    uint32_t sep = rnd_n(r, 40);
    u += put_cp(t, sep == 0 ? '\n' : (sep < 4 ? ',' : (sep < 6 ? '.' : ' ')));
    if (sep > 0 && sep < 6 && u < target) u += put_cp(t, ' ');
  }
  return u;
}

static void trim_sep(buf_t* t, size_t floor) {
  while (t->len > floor) {
    uint8_t c = t->p[t->len - 1];
    if (c == ' ' || c == ',' || c == '.' || c == '\n') t->len--; else break;
  }
}

static int pick_script(rng_t* r) {
  double u = rnd01(r);
  return u < 0.50 ? SC_LATIN : u < 0.80 ? SC_CYR : u < 0.88 ? SC_ARAB : u < 0.95 ? SC_CJK : SC_EMOJI;
}

/* FormattedText with links; appends text to t and entities to g->ents / URLs to g->aux */
static void gen_formatted_text(rng_t* r, gen_t* g, buf_t* t, double len_scale, uint32_t* nent) {
  size_t t0 = t->len;
  double ln = exp(log(180.0) + 1.01 * rnd_normal(r)) * len_scale;
  int target = ln > 4096 ? 4096 : (int)ln;
  int sc = pick_script(r);
  int specials = rnd01(r) < 0.02, bad = rnd01(r) < 0.001;
  double pu = rnd01(r);
  int nlinks = pu < 0.7047 ? 0 : pu < 0.9513 ? 1 : pu < 0.9945 ? 2 : pu < 0.9995 ? 3 : 4 + (int)rnd_n(r, 4);
  int u = 0;
  for (int k = 0; k <= nlinks; k++) {
    int seg = (target - u) / (nlinks - k + 1);
    if (seg > 0) u += gen_words(r, t, sc, seg, specials, bad);
    if (k == nlinks) break;
    if (u > 4096 - 64) break;
    char name[64];
    int nn = make_link_name(r, name);
    double kind = rnd01(r);
    tgi_entity e;
    memset(&e, 0, sizeof e);
    if (kind < 0.45) { /* @mention entity */
      e.offset = u; e.length = 1 + nn; e.type = TGI_ENT_MENTION;
      if (nn > 32) { nn = 32; e.length = 33; } /* Telegram mentions are <= 32 */
      b_putc(t, '@'); b_put(t, name, (size_t)nn); u += 1 + nn;
      b_put(&g->ents, &e, sizeof e); (*nent)++;
    } else if (kind < 0.65) { /* text_url entity over a normal word */
      e.offset = u; e.type = TGI_ENT_TEXT_URL;
      int w = gen_words(r, t, sc, 3 + (int)rnd_n(r, 8), 0, 0);
      e.length = w; u += w;
      e.url_off = (uint32_t)g->aux.len;
      char url[128];
      int ul = snprintf(url, sizeof url, "%st.me/%.*s%s", rnd_n(r, 8) ? "https://" : "http://", nn, name,
                        rnd_n(r, 3) == 0 ? "/123" : "");
      if (rnd_n(r, 10) == 0) ul = snprintf(url, sizeof url, "https://example.com/%.*s", nn, name);
      e.url_len = (uint16_t)ul;
      b_put(&g->aux, url, (size_t)ul);
      b_put(&g->ents, &e, sizeof e); (*nent)++;
    } else { /* bare url (entity) or plaintext */
      int scheme = (int)rnd_n(r, 3); /* 0 none, 1 https, 2 http */
      const char* pre = scheme == 0 ? "" : scheme == 1 ? "https://" : "http://";
      e.offset = u; e.type = TGI_ENT_URL;
      b_put(t, pre, strlen(pre)); b_put(t, "t.me/", 5); b_put(t, name, (size_t)nn);
      int l = (int)strlen(pre) + 5 + nn;
      if (rnd_n(r, 4) == 0) { b_put(t, "/42", 3); l += 3; }
      e.length = l; u += l;
      if (kind < 0.80) { b_put(&g->ents, &e, sizeof e); (*nent)++; }
    }
    u += put_cp(t, ' ');
  }
  /* rare adversarial shapes: regex chains and broken entity offsets */
  double adv = rnd01(r);
  if (adv < 0.0005) {
    static const char* chain = " t.me/abcdt.me/efght.me/ijklmt.me/nopqr xt.me/chain_end";
    b_put(t, chain, strlen(chain));
  } else if (adv < 0.0008) { /* offsets that never resolve / overrun / hit a surrogate pair */
    tgi_entity e;
    memset(&e, 0, sizeof e);
    e.type = rnd_n(r, 2) ? TGI_ENT_MENTION : TGI_ENT_URL;
    uint32_t m = rnd_n(r, 3);
    e.offset = m == 0 ? u + 5 : (m == 1 ? (u > 4 ? u - 3 : 0) : 1);
    e.length = m == 0 ? 9 : (m == 1 ? 40 : 7);
    if (m == 2) { /* astral first rune: offset 1 is inside the pair -> Go panics */
      buf_t tmp = {0};
      b_put(&tmp, t->p + t0, t->len - t0);
      t->len = t0;
      put_cp(t, 0x1F600);
      b_put(t, tmp.p, tmp.len);
      free(tmp.p);
      /* shift previously emitted entities of this text by 2 units */
      tgi_entity* ev = (tgi_entity*)(g->ents.p + g->ents.len) - *nent;
      for (uint32_t i = 0; i < *nent; i++) ev[i].offset += 2;
    }
    b_put(&g->ents, &e, sizeof e); (*nent)++;
  }
}

static void gen_record(gen_t* g, uint64_t k) {
  rng_t r = {mix64(g->seed ^ mix64(k + 0x1234567ull))};
  tgi_tg_rec rec;
  memset(&rec, 0, sizeof rec);
  rec.id = (int64_t)(k + 1) << 20;
  rec.chan_idx = (uint32_t)((k / 100) % g->n_chans);
  rec.chat_id = -1000000000000ll - (int64_t)rec.chan_idx;
  rec.date = 1672531200 + (int32_t)rnd_n(&r, 94608000u);
  rec.media_album_id = rnd01(&r) < 0.05 ? (int64_t)(rnd(&r) >> 8) + 1 : 0;
  double v = exp(7.0 + 2.0 * rnd_normal(&r));
  rec.view_count = v > 2147483647.0 ? 2147483647 : (int32_t)v;
  rec.share_count = rec.view_count / 80;
  rec.str_off = g->strs.len;
  /* content type */
  int ct = TGI_CT_TEXT;
  if (g->profile == 2) {
    double u = rnd01(&r);
    if (u < 0.60) ct = TGI_CT_TEXT;
    else if (u < 0.82) ct = TGI_CT_PHOTO;
    else if (u < 0.94) ct = TGI_CT_VIDEO;
    else if (u < 0.97) ct = TGI_CT_DOCUMENT;
    else {
      static const int other[] = {TGI_CT_ANIMATION, TGI_CT_STICKER, TGI_CT_POLL, TGI_CT_VOICE_NOTE, TGI_CT_AUDIO,
                                  TGI_CT_ANIMATED_EMOJI, TGI_CT_GIVEAWAY, TGI_CT_VIDEO_NOTE, TGI_CT_PAID_MEDIA,
                                  TGI_CT_OTHER, TGI_CT_NONE, TGI_CT_GIVEAWAY_WINNERS};
      ct = other[rnd_n(&r, 12)];
    }
  } else if (g->profile == 3) {
    double u = rnd01(&r);
    ct = u < 0.7 ? TGI_CT_TEXT : u < 0.85 ? TGI_CT_PHOTO : u < 0.95 ? TGI_CT_VIDEO : TGI_CT_DOCUMENT;
  }
  rec.content_type = (uint8_t)ct;
  uint32_t nent = 0, nreact = 0, ncomment = 0;
  int carrier = ct == TGI_CT_TEXT || ct == TGI_CT_PHOTO || ct == TGI_CT_VIDEO || ct == TGI_CT_DOCUMENT ||
                ct == TGI_CT_ANIMATION || ct == TGI_CT_AUDIO || ct == TGI_CT_VOICE_NOTE;
  size_t s0 = g->strs.len;
  if (carrier && !(ct != TGI_CT_TEXT && rnd01(&r) < 0.02)) { /* 2% of media have a nil caption */
    rec.flags |= TGI_RF_HAS_TEXT;
    if (!(ct != TGI_CT_TEXT && rnd01(&r) < 0.3)) /* 30% of media captions are empty */
      gen_formatted_text(&r, g, &g->strs, ct == TGI_CT_TEXT ? 1.0 : 0.5, &nent);
  }
  rec.text_len = (uint32_t)(g->strs.len - s0);
  /* alt */
  s0 = g->strs.len;
  switch (ct) {
    case TGI_CT_ANIMATED_EMOJI: { const char* e = EMOJI[rnd_n(&r, 40)]; b_put(&g->strs, e, strlen(e)); break; }
    case TGI_CT_POLL: gen_words(&r, &g->strs, pick_script(&r), 10 + (int)rnd_n(&r, 60), 0, 0); b_putc(&g->strs, '?'); break;
    case TGI_CT_GIVEAWAY: { const char* p = rnd_n(&r, 2) ? "giveawayPrizePremium" : "giveawayPrizeStars"; b_put(&g->strs, p, strlen(p)); break; }
    case TGI_CT_PAID_MEDIA: gen_words(&r, &g->strs, pick_script(&r), 5 + (int)rnd_n(&r, 100), rnd_n(&r, 20) == 0, 0); break;
    case TGI_CT_DOCUMENT: {
      gen_words(&r, &g->strs, rnd_n(&r, 3) ? SC_LATIN : SC_CYR, 4 + (int)rnd_n(&r, 20), 0, 0);
      trim_sep(&g->strs, s0);
      static const char* ext[] = {".pdf", ".docx", ".zip", ".mp3", ".apk"};
      const char* e = ext[rnd_n(&r, 5)];
      b_put(&g->strs, e, strlen(e));
      break;
    }
    case TGI_CT_OTHER: { static const char* ty[] = {"messageLocation", "messageContact", "messageDice", "messageStory"}; const char* p = ty[rnd_n(&r, 4)]; b_put(&g->strs, p, strlen(p)); break; }
    default: break;
  }
  rec.alt_len = (uint32_t)(g->strs.len - s0);
  /* media remote id */
  s0 = g->strs.len;
  if ((ct == TGI_CT_VIDEO && rnd01(&r) < 0.95) || ct == TGI_CT_VIDEO_NOTE || ct == TGI_CT_DOCUMENT) {
    static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
    int n = 60 + (int)rnd_n(&r, 20);
    for (int i = 0; i < n; i++) b_putc(&g->strs, (uint8_t)B64[rnd_n(&r, 64)]);
  }
  rec.media_len = (uint16_t)(g->strs.len - s0);
  /* handle: mostly the channel title (sender = chat); we only know its index here, so synthesize
   * the same title bytes the channel table holds */
  s0 = g->strs.len;
  {
    double u = rnd01(&r);
    if (u < 0.05) b_put(&g->strs, "unknown", 7);
    else if (u < 0.10) { gen_words(&r, &g->strs, pick_script(&r), 4 + (int)rnd_n(&r, 12), 0, 0); trim_sep(&g->strs, s0); }
    else {
      rng_t cr = {mix64(g->seed ^ (0xC4A7ull + rec.chan_idx))};
      gen_words(&cr, &g->strs, pick_script(&cr), 6 + (int)rnd_n(&cr, 24), rnd_n(&cr, 50) == 0, 0);
      trim_sep(&g->strs, s0);
    }
  }
  rec.handle_len = (uint16_t)(g->strs.len - s0);
  /* reactions / comments */
  if (rnd01(&r) < 0.40) gen_reactions(&r, &g->reacts, &g->aux, &nreact);
  if (g->profile != 1) {
    double u = rnd01(&r);
    if (u < 0.01) rec.flags |= TGI_RF_COMMENTS_NIL;
    else if (u < 0.04) {
      ncomment = 1 + rnd_n(&r, 3);
      for (uint32_t i = 0; i < ncomment; i++) {
        tgi_comment c;
        memset(&c, 0, sizeof c);
        buf_t tmp = {0};
        gen_words(&r, &tmp, pick_script(&r), 5 + (int)rnd_n(&r, 80), rnd_n(&r, 30) == 0, 0);
        c.text_off = (uint32_t)g->aux.len; c.text_len = (uint32_t)tmp.len;
        b_put(&g->aux, tmp.p, tmp.len);
        tmp.len = 0;
        gen_words(&r, &tmp, SC_LATIN, 4 + (int)rnd_n(&r, 10), 0, 0);
        trim_sep(&tmp, 0);
        c.handle_off = (uint32_t)g->aux.len; c.handle_len = (uint16_t)tmp.len;
        b_put(&g->aux, tmp.p, tmp.len);
        free(tmp.p);
        c.view_count = (int32_t)rnd_n(&r, 5000); c.reply_count = rnd_n(&r, 4) ? 0 : (int32_t)rnd_n(&r, 20);
        if (rnd_n(&r, 2)) {
          c.flags = 1;
          c.react_start = (uint32_t)(g->creacts.len / sizeof(tgi_reaction)); /* fixed up at merge */
          uint32_t cnt = 0;
          gen_reactions(&r, &g->creacts, &g->aux, &cnt);
          c.react_count = cnt;
        }
        b_put(&g->comments, &c, sizeof c);
      }
    }
  }
  if (rnd01(&r) < 0.00002) rec.flags |= TGI_RF_PANIC;
  b_put(&g->recs, &rec, sizeof rec);
  b_put(&g->ent_cnt, &nent, 4);
  b_put(&g->react_cnt, &nreact, 4);
  b_put(&g->comment_cnt, &ncomment, 4);
}

static void* gen_worker(void* arg) {
  gen_t* g = (gen_t*)arg;
  for (uint64_t k = g->r0; k < g->r1; k++) gen_record(g, g->first + k);
  return NULL;
}

/* ---- public API ------------------------------------------------------------------------------- */
typedef struct tgc_corpus {
  tgi_tg_batch b;
  uint64_t total_bytes;
} tgc_corpus;

static void* xalloc(size_t n) {
  void* p = NULL;
  if (posix_memalign(&p, 4096, (n + 4095 + 64) & ~(size_t)4095)) abort();
  return p;
}

void tgc_free(tgc_corpus* c) {
  if (!c) return;
  free((void*)c->b.recs); free((void*)c->b.strs); free((void*)c->b.ent_off); free((void*)c->b.ents);
  free((void*)c->b.react_off); free((void*)c->b.reacts); free((void*)c->b.comment_off);
  free((void*)c->b.comments); free((void*)c->b.aux); free((void*)c->b.chans); free((void*)c->b.chan_strs);
  free(c);
}

tgc_corpus* tgc_telegram(uint64_t seed, uint64_t first, uint64_t n, int profile, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > n) nthreads = n ? (int)n : 1;
  uint64_t total = first + n;
  uint32_t n_chans = (uint32_t)((total + 99) / 100);
  if (n_chans > 100000) n_chans = 100000;
  if (n_chans < 1) n_chans = 1;
  gen_t* g = (gen_t*)calloc((size_t)nthreads, sizeof(gen_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    g[t].seed = seed; g[t].first = first; g[t].profile = profile; g[t].n_chans = n_chans;
    g[t].r0 = n * (uint64_t)t / (uint64_t)nthreads;
    g[t].r1 = n * (uint64_t)(t + 1) / (uint64_t)nthreads;
    if (nthreads > 1) pthread_create(&th[t], NULL, gen_worker, &g[t]); else gen_worker(&g[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  size_t S = 0, E = 0, R = 0, CR = 0, Cm = 0, A = 0;
  for (int t = 0; t < nthreads; t++) {
    S += g[t].strs.len; E += g[t].ents.len; R += g[t].reacts.len; CR += g[t].creacts.len;
    Cm += g[t].comments.len; A += g[t].aux.len;
  }
  tgc_corpus* c = (tgc_corpus*)calloc(1, sizeof *c);
  tgi_tg_rec* recs = (tgi_tg_rec*)xalloc(n * sizeof(tgi_tg_rec));
  uint8_t* strs = (uint8_t*)xalloc(S);
  uint32_t* ent_off = (uint32_t*)xalloc((n + 1) * 4);
  uint32_t* react_off = (uint32_t*)xalloc((n + 1) * 4);
  uint32_t* comment_off = (uint32_t*)xalloc((n + 1) * 4);
  uint8_t* ents = (uint8_t*)xalloc(E);
  uint8_t* reacts = (uint8_t*)xalloc(R + CR);
  uint8_t* comments = (uint8_t*)xalloc(Cm);
  uint8_t* aux = (uint8_t*)xalloc(A);
  size_t so = 0, eo = 0, ro = 0, cro = R, co = 0, ao = 0;
  uint64_t ri = 0;
  uint32_t ecount = 0, rcount = 0, ccount = 0;
  for (int t = 0; t < nthreads; t++) {
    uint64_t m = g[t].r1 - g[t].r0;
    memcpy(strs + so, g[t].strs.p, g[t].strs.len);
    memcpy(ents + eo, g[t].ents.p, g[t].ents.len);
    memcpy(reacts + ro, g[t].reacts.p, g[t].reacts.len);
    memcpy(reacts + cro, g[t].creacts.p, g[t].creacts.len);
    memcpy(comments + co, g[t].comments.p, g[t].comments.len);
    memcpy(aux + ao, g[t].aux.p, g[t].aux.len);
    memcpy(recs + ri, g[t].recs.p, m * sizeof(tgi_tg_rec));
    /* fix-ups: absolute offsets */
    tgi_entity* ev = (tgi_entity*)(ents + eo);
    for (size_t i = 0; i < g[t].ents.len / sizeof(tgi_entity); i++)
      if (ev[i].type == TGI_ENT_TEXT_URL) ev[i].url_off += (uint32_t)ao;
    tgi_reaction* rv = (tgi_reaction*)(reacts + ro);
    for (size_t i = 0; i < g[t].reacts.len / sizeof(tgi_reaction); i++) rv[i].emoji_off += (uint32_t)ao;
    rv = (tgi_reaction*)(reacts + cro);
    for (size_t i = 0; i < g[t].creacts.len / sizeof(tgi_reaction); i++) rv[i].emoji_off += (uint32_t)ao;
    tgi_comment* cv = (tgi_comment*)(comments + co);
    for (size_t i = 0; i < g[t].comments.len / sizeof(tgi_comment); i++) {
      cv[i].text_off += (uint32_t)ao; cv[i].handle_off += (uint32_t)ao;
      if (cv[i].flags & 1) cv[i].react_start += (uint32_t)(cro / sizeof(tgi_reaction));
    }
    const uint32_t *ec = (const uint32_t*)g[t].ent_cnt.p, *rc = (const uint32_t*)g[t].react_cnt.p,
                   *cc = (const uint32_t*)g[t].comment_cnt.p;
    for (uint64_t i = 0; i < m; i++) {
      recs[ri + i].str_off += so;
      ent_off[ri + i] = ecount; react_off[ri + i] = rcount; comment_off[ri + i] = ccount;
      ecount += ec[i]; rcount += rc[i]; ccount += cc[i];
    }
    so += g[t].strs.len; eo += g[t].ents.len; ro += g[t].reacts.len; cro += g[t].creacts.len;
    co += g[t].comments.len; ao += g[t].aux.len; ri += m;
    free(g[t].recs.p); free(g[t].strs.p); free(g[t].ents.p); free(g[t].reacts.p); free(g[t].comments.p);
    free(g[t].aux.p); free(g[t].ent_cnt.p); free(g[t].react_cnt.p); free(g[t].comment_cnt.p); free(g[t].creacts.p);
  }
  ent_off[n] = ecount; react_off[n] = rcount; comment_off[n] = ccount;
  /* channel table (global, generated from channel index only) */
  tgi_tg_chan* chans = (tgi_tg_chan*)xalloc(n_chans * sizeof(tgi_tg_chan));
  buf_t cs = {0};
  for (uint32_t i = 0; i < n_chans; i++) {
    tgi_tg_chan ch;
    memset(&ch, 0, sizeof ch);
    ch.str_off = (uint32_t)cs.len;
    rng_t cr = {mix64(seed ^ (0xC4A7ull + i))}; /* same stream as the handle above */
    size_t a = cs.len;
    gen_words(&cr, &cs, pick_script(&cr), 6 + (int)rnd_n(&cr, 24), rnd_n(&cr, 50) == 0, 0);
    trim_sep(&cs, a);
    ch.title_len = (uint16_t)(cs.len - a);
    char name[64];
    int nn = make_name(i * 7u + 3u, name);
    b_put(&cs, name, (size_t)nn);
    ch.name_len = (uint16_t)nn;
    rng_t c2 = {mix64(seed ^ (0xBEEFull + i))};
    if (rnd01(&c2) < 0.9) { b_put(&cs, name, (size_t)nn); ch.user_len = (uint16_t)nn; }
    ch.member_count = (int64_t)exp(6.0 + 2.5 * rnd_normal(&c2));
    ch.post_count = 100;
    ch.view_count = (int64_t)exp(11.0 + 2.0 * rnd_normal(&c2));
    chans[i] = ch;
  }
  uint8_t* chan_strs = (uint8_t*)xalloc(cs.len);
  memcpy(chan_strs, cs.p, cs.len);
  c->b.n = n; c->b.recs = recs; c->b.strs = strs; c->b.strs_len = S;
  c->b.ent_off = ent_off; c->b.ents = (tgi_entity*)ents;
  c->b.react_off = react_off; c->b.reacts = (tgi_reaction*)reacts; c->b.n_reacts = (R + CR) / sizeof(tgi_reaction);
  c->b.comment_off = comment_off; c->b.comments = (tgi_comment*)comments; c->b.n_comments = Cm / sizeof(tgi_comment);
  c->b.aux = aux; c->b.aux_len = A; c->b.n_chans = n_chans; c->b.chans = chans;
  c->b.chan_strs = chan_strs; c->b.chan_strs_len = cs.len;
  c->total_bytes = n * sizeof(tgi_tg_rec) + S + 3 * (n + 1) * 4 + E + R + CR + Cm + A +
                   n_chans * sizeof(tgi_tg_chan) + cs.len;
  free(cs.p); free(g); free(th);
  return c;
}

const tgi_tg_batch* tgc_batch(const tgc_corpus* c) { return &c->b; }
uint64_t tgc_total_bytes(const tgc_corpus* c) { return c->total_bytes; }

/* ---- YouTube (BASELINE config 4; shape: SURVEY.md §8d) ---------------------------------------------
 * id 11 base64url chars; title lognormal median 45 B; description lognormal median 400 B clipped at 5000 with
 * Poisson(1.5) URLs of which 5 % are youtube.com/channel/UC..., 5 % youtube.com/@handle; views lognormal(8, 3),
 * likes = views / 30, comments = views / 300; durations PT#H#M#S (1 % P0D, 0.5 % empty, 0.5 % P1DT2H);
 * 3-5 thumbnails; 1000 cached channels.  2 % of the descriptions carry characters that need escaping, 10 % are
 * Cyrillic (valid multi-byte UTF-8 that passes through unescaped).  Same determinism rule as the Telegram
 * generator: record k draws from its own stream keyed by (seed, k).                                        */
typedef struct tgc_yt_corpus {
  tgi_yt_batch b;
  uint64_t total_bytes;
} tgc_yt_corpus;

typedef struct {
  uint64_t seed, first, r0, r1;
  uint32_t n_chans;
  buf_t recs, strs;
} ytgen_t;

static const char B64U[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
static const char* YT_PLAIN[] = {"the", "video", "about", "channel", "new", "watch", "and", "more", "from", "this",
                                 "week", "episode", "review", "how", "to", "guide", "music", "official", "live", "part",
                                 "best", "of", "2024", "full", "with", "our", "your", "for", "you", "in"};

static uint32_t rnd_poisson(rng_t* r, double lam) {
  const double L = exp(-lam);
  uint32_t k = 0;
  double p = 1.0;
  for (;;) {
    p *= rnd01(r);
    if (p <= L) return k;
    k++;
  }
}
static void yt_words(rng_t* r, buf_t* t, size_t nbytes, int cyr) {
  const size_t t0 = t->len;
  while (t->len - t0 < nbytes) {
    if (cyr) {
      int wl = 2 + (int)rnd_n(r, 9);
      for (int i = 0; i < wl; i++) put_cp(t, 0x0430 + rnd_n(r, 32));
    } else {
      const char* w = YT_PLAIN[rnd_n(r, 30)];
      b_put(t, w, strlen(w));
    }
    b_putc(t, ' ');
  }
  if (t->len > t0) t->len--; /* no trailing blank */
}
static void yt_gen_record(ytgen_t* g, uint64_t k) {
  rng_t r = {mix64(g->seed ^ mix64(k + 0x7654321ull))};
  tgi_yt_rec rec;
  memset(&rec, 0, sizeof rec);
  rec.str_off = g->strs.len;
  char id[12];
  for (int i = 0; i < 11; i++) id[i] = B64U[rnd_n(&r, 64)];
  b_put(&g->strs, id, 11);
  rec.id_len = 11;
  size_t s0 = g->strs.len;
  const int cyr = rnd01(&r) < 0.10;
  double tl = exp(3.8 + 0.5 * rnd_normal(&r));
  yt_words(&r, &g->strs, (size_t)(tl > 300 ? 300 : tl), cyr);
  rec.title_len = (uint16_t)(g->strs.len - s0);
  s0 = g->strs.len;
  double dl = exp(5.99 + 0.9 * rnd_normal(&r));
  const size_t target = (size_t)(dl > 5000 ? 5000 : dl);
  const uint32_t nurls = rnd_poisson(&r, 1.5);
  for (uint32_t u = 0; u <= nurls; u++) {
    const size_t done = g->strs.len - s0;
    const size_t seg = done < target ? (target - done) / (nurls - u + 1) : 0;
    if (seg) yt_words(&r, &g->strs, seg, cyr);
    if (u == nurls) break;
    if (g->strs.len > s0) b_putc(&g->strs, ' ');
    const double w = rnd01(&r);
    char link[96];
    int ll;
    if (w < 0.05) {
      ll = snprintf(link, sizeof link, "https://www.youtube.com/channel/UC");
      for (int i = 0; i < 22; i++) link[ll++] = B64U[zipf(&r, 2000000u, 1.05) * 2654435761u >> (i & 7) & 63];
    } else if (w < 0.10) {
      ll = snprintf(link, sizeof link, "https://youtube.com/@");
      char nm[64];
      int nn = make_name(zipf(&r, 2000000u, 1.05), nm);
      memcpy(link + ll, nm, (size_t)nn);
      ll += nn;
    } else {
      ll = snprintf(link, sizeof link, "https://example.com/");
      int tail = 4 + (int)rnd_n(&r, 16);
      for (int i = 0; i < tail; i++) link[ll++] = B64U[rnd_n(&r, 64)];
    }
    b_put(&g->strs, link, (size_t)ll);
    b_putc(&g->strs, ' ');
  }
  while (g->strs.len > s0 && g->strs.p[g->strs.len - 1] == ' ') g->strs.len--;
  if (rnd01(&r) < 0.02) {
    static const char esc[] = "\n\"quoted\" <tag> & more\n";
    b_put(&g->strs, esc, sizeof esc - 1);
  }
  rec.desc_len = (uint32_t)(g->strs.len - s0);
  s0 = g->strs.len;
  {
    const double u = rnd01(&r);
    char d[32];
    int n = u < 0.01 ? snprintf(d, sizeof d, "P0D") : u < 0.015 ? 0 : u < 0.02 ? snprintf(d, sizeof d, "P1DT2H")
            : snprintf(d, sizeof d, "PT%uH%uM%uS", rnd_n(&r, 3), rnd_n(&r, 60), rnd_n(&r, 60));
    b_put(&g->strs, d, (size_t)n);
  }
  rec.duration_len = (uint16_t)(g->strs.len - s0);
  b_put(&g->strs, "en", 2);
  rec.lang_len = 2;
  static const char* KEY[5] = {"default", "medium", "high", "standard", "maxres"};
  const uint32_t nth = 3 + rnd_n(&r, 3);
  for (uint32_t t = 0; t < 5; t++) {
    if (t < nth) {
      char url[96];
      int n = snprintf(url, sizeof url, "https://i.ytimg.com/vi/%.11s/%s.jpg", id, KEY[t]);
      b_put(&g->strs, url, (size_t)n);
      rec.thumb_len[t] = (uint16_t)n;
    } else {
      rec.thumb_len[t] = TGI_YT_THUMB_ABSENT;
    }
  }
  rec.published_sec = 1300000000ll + (int64_t)rnd_n(&r, 460000000u);
  const double v = exp(8.0 + 3.0 * rnd_normal(&r));
  rec.view_count = v > 9e18 ? (int64_t)9e18 : (int64_t)v;
  rec.like_count = rec.view_count / 30;
  rec.comment_count = rec.view_count / 300;
  rec.chan_idx = rnd_n(&r, g->n_chans);
  b_put(&g->recs, &rec, sizeof rec);
}
static void* yt_gen_worker(void* arg) {
  ytgen_t* g = (ytgen_t*)arg;
  for (uint64_t k = g->r0; k < g->r1; k++) yt_gen_record(g, g->first + k);
  return NULL;
}

void tgc_yt_free(tgc_yt_corpus* c) {
  if (!c) return;
  free((void*)c->b.recs); free((void*)c->b.strs); free((void*)c->b.chans); free((void*)c->b.chan_strs);
  free(c);
}

tgc_yt_corpus* tgc_youtube(uint64_t seed, uint64_t first, uint64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > n) nthreads = n ? (int)n : 1;
  const uint32_t n_chans = 1000;
  ytgen_t* g = (ytgen_t*)calloc((size_t)nthreads, sizeof(ytgen_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    g[t].seed = seed; g[t].first = first; g[t].n_chans = n_chans;
    g[t].r0 = n * (uint64_t)t / (uint64_t)nthreads;
    g[t].r1 = n * (uint64_t)(t + 1) / (uint64_t)nthreads;
    if (nthreads > 1) pthread_create(&th[t], NULL, yt_gen_worker, &g[t]); else yt_gen_worker(&g[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  size_t S = 0;
  for (int t = 0; t < nthreads; t++) S += g[t].strs.len;
  tgc_yt_corpus* c = (tgc_yt_corpus*)calloc(1, sizeof *c);
  tgi_yt_rec* recs = (tgi_yt_rec*)xalloc(n * sizeof(tgi_yt_rec));
  uint8_t* strs = (uint8_t*)xalloc(S);
  size_t so = 0;
  uint64_t ri = 0;
  for (int t = 0; t < nthreads; t++) {
    const uint64_t m = g[t].r1 - g[t].r0;
    memcpy(strs + so, g[t].strs.p, g[t].strs.len);
    memcpy(recs + ri, g[t].recs.p, m * sizeof(tgi_yt_rec));
    for (uint64_t i = 0; i < m; i++) recs[ri + i].str_off += so;
    so += g[t].strs.len; ri += m;
    free(g[t].recs.p); free(g[t].strs.p);
  }
  tgi_yt_chan* chans = (tgi_yt_chan*)xalloc(n_chans * sizeof(tgi_yt_chan));
  buf_t cs = {0};
  for (uint32_t i = 0; i < n_chans; i++) {
    tgi_yt_chan ch;
    memset(&ch, 0, sizeof ch);
    rng_t cr = {mix64(seed ^ (0xCAFEull + i))};
    ch.str_off = (uint32_t)cs.len;
    char tmp[96];
    int n2 = 2;
    tmp[0] = 'U'; tmp[1] = 'C';
    for (int k = 0; k < 22; k++) tmp[n2++] = B64U[rnd_n(&cr, 64)];
    b_put(&cs, tmp, (size_t)n2); ch.id_len = (uint16_t)n2;
    n2 = snprintf(tmp, sizeof tmp, "Channel %u", i);
    b_put(&cs, tmp, (size_t)n2); ch.title_len = (uint16_t)n2;
    size_t a = cs.len;
    yt_words(&cr, &cs, 100, 0);
    ch.desc_len = (uint32_t)(cs.len - a);
    n2 = snprintf(tmp, sizeof tmp, "https://yt3.ggpht.com/");
    for (int k = 0; k < 30; k++) tmp[n2++] = B64U[rnd_n(&cr, 64)];
    b_put(&cs, tmp, (size_t)n2); ch.thumb_len = (uint16_t)n2;
    b_put(&cs, "US", 2); ch.country_len = 2;
    ch.subscriber_count = (int64_t)rnd_n(&cr, 10000000u);
    ch.view_count = (int64_t)(rnd(&cr) % 10000000000ull);
    ch.video_count = (int64_t)rnd_n(&cr, 10000u);
    ch.published_sec = 1100000000ll + (int64_t)rnd_n(&cr, 600000000u);
    ch.cached = 1;
    chans[i] = ch;
  }
  uint8_t* chan_strs = (uint8_t*)xalloc(cs.len);
  memcpy(chan_strs, cs.p, cs.len);
  c->b.n = n; c->b.recs = recs; c->b.strs = strs; c->b.strs_len = S;
  c->b.n_chans = n_chans; c->b.chans = chans; c->b.chan_strs = chan_strs; c->b.chan_strs_len = cs.len;
  c->total_bytes = n * sizeof(tgi_yt_rec) + S + n_chans * sizeof(tgi_yt_chan) + cs.len;
  free(cs.p); free(g); free(th);
  return c;
}
const tgi_yt_batch* tgc_yt_batch(const tgc_yt_corpus* c) { return &c->b; }
uint64_t tgc_yt_total_bytes(const tgc_yt_corpus* c) { return c->total_bytes; }
