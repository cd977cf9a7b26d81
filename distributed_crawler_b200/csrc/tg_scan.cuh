// tg_scan.cuh — ONE pass over a message text: JSON-escaped length, UTF-8 sanity, "t.me/" candidates.
//
// Replaces, for the common case, the three walks the text used to get (link scan in the parse kernel, escape length in
// the size kernel over 128-byte strips with full UTF-8 bookkeeping): 512-byte strips, 16 bytes per lane, SWAR only.
// The scanner does not have to DECIDE hard cases, it only has to notice them: whenever a strip holds anything the
// cheap tests cannot vouch for (structurally broken UTF-8, a lead byte with a restricted second-byte range, C0 / C1 /
// F5.., an E2 .. A8/A9 combination that could be U+2028 / U+2029), `exact` is raised and the caller re-measures the
// string with the exact path (warp_esc_len, dev_common.cuh), which restates utf8.DecodeRuneInString byte by byte.
//
// Reference semantics: encoding/json string escaping, channelLinkRegex candidates
// (telegramhelper/tdutils.go:23).
#pragma once
#include "tg_links.cuh"

namespace tgi {

struct TextScan {
  uint32_t esc;      // escaped length (valid when !exact)
  uint32_t tme;      // number of "t.me/" occurrences: upper bound on plaintext channel links
  bool exact;        // the cheap tests cannot vouch for this string
  bool non_ascii;    // some byte >= 0x80
};

// zero-byte "any" test: non-zero iff some byte of v is 0x00 (false positives only above a true zero byte)
DEVI uint32_t swar_zero_any(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

// Warp-collective.  s[0..n); the blob padding covers the 16-byte over-read.
template <bool WANT_ESC, bool WANT_TME>
DEVI TextScan warp_text_scan(const uint8_t* s, uint32_t n) {
  const uint32_t M = 0x80808080u;
  const int l = lane_id();
  uint32_t esc = 0, tme = 0, bad = 0, hi_any = 0, seen = 0;
  uint32_t carryP = 0;  // P flags (lead / lead3 / lead4 in bits 7 / 6 / 5 of every byte) of the previous strip's last word
  // one extra strip when the text ends within 3 bytes of a strip end: a sequence cut off by the end of the string
  // shows up as a missing continuation byte one position PAST the end
  for (uint32_t base = 0; base < n + (WANT_ESC ? 3u : 0u); base += 512) {
    const uint32_t p0 = base + 16u * (uint32_t)l;
    const uint32_t nv = p0 >= n ? 0u : min(16u, n - p0);
    uint32_t w[4] = {0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u};
    if (nv) {
      const uint4 v = load16_lane(s + p0);
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
      if (nv < 16) {  // bytes past the end read as spaces: neither special nor part of a sequence
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
          const uint32_t nj = nv > 4u * j ? min(4u, nv - 4u * j) : 0u;
          if (nj < 4) w[j] = (nj ? (w[j] & ((1u << (8u * nj)) - 1u)) : 0u) | (0x20202020u << (8u * nj));
        }
      }
    }
    if (WANT_TME && nv) {  // '/' preceded by "t.me" (tg_links.cuh strip16_tme, on the words already loaded)
      const uint32_t s0 = swar_eq(w[0], '/'), s1 = swar_eq(w[1], '/'), s2 = swar_eq(w[2], '/'), s3 = swar_eq(w[3], '/');
      if (s0 | s1 | s2 | s3) {
        uint32_t m = swar_movemask(s0) | (swar_movemask(s1) << 4) | (swar_movemask(s2) << 8) | (swar_movemask(s3) << 12);
        if (nv < 16) m &= (1u << nv) - 1u;
        while (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1;
          if (p0 + k >= 4 && ld_u32_unaligned(s + p0 + k - 4) == 0x656D2E74u) tme++;  // "t.me"
        }
      }
    }
    if (!WANT_ESC) continue;
    // ---- escape length: bytes that become 2 or 6 bytes ----
    uint32_t spec = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t x = w[j];
      const uint32_t ctl = ~(((x & 0x7F7F7F7Fu) + 0x60606060u) | x) & M;                    // < 0x20
      spec |= ctl | swar_has_byte(x & 0xFBFBFBFBu, 0x22) | swar_has_byte(x & 0xFDFDFDFDu, 0x3C) | swar_has_byte(x, 0x5C);  // " & < > backslash
    }
    uint32_t e = nv;
    if (spec && nv) {  // rare: count exactly, byte by byte
#pragma unroll 1
      for (uint32_t k = 0; k < nv; k++) {
        const uint32_t bt = (w[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
        if (bt < 0x80u) e += ascii_esc_len(bt) - 1u;
      }
    }
    esc += e;
    // ---- UTF-8: only strips that hold a byte >= 0x80 (or follow one that ended inside a sequence) ----
    const uint32_t hi = (w[0] | w[1] | w[2] | w[3]) & M;
    const bool any_hi = __ballot_sync(FULL, hi != 0) != 0;
    if (!any_hi && carryP == 0) continue;
    hi_any |= any_hi ? 1u : 0u;
    uint32_t P[4], c[4];
    uint32_t susp = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t x = w[j];
      c[j] = x & ((~x) << 1) & M;          // 10xxxxxx
      const uint32_t ld = x & (x << 1) & M;  // 11xxxxxx
      const uint32_t l3 = ld & (x << 2);     // 111xxxxx
      const uint32_t l4 = l3 & (x << 3);     // 1111xxxx
      P[j] = ld | (l3 >> 1) | (l4 >> 2);
      // what the cheap tests cannot vouch for: F0.. (second-byte ranges, > F4), E0 / ED (second-byte ranges), E2 with an
      // A8 / A9 somewhere (U+2028/9 candidates), C0 / C1 (overlong)
      susp |= l4;
      if (l3) {
        const uint32_t keep = ~(l3 >> 7) & 0x01010101u;  // non-lead3 bytes are forced non-zero
        const uint32_t lo = x & 0x0F0F0F0Fu;
        susp |= swar_zero_any(lo | keep) | swar_zero_any((lo ^ 0x0D0D0D0Du) | keep);                       // E0, ED
        if (swar_zero_any((lo ^ 0x02020202u) | keep)) susp |= 0x100u;                                       // E2 seen
      }
      if (swar_has_byte(x & 0xFEFEFEFEu, 0xA8)) susp |= 0x200u;                                              // A8 / A9 seen
      const uint32_t l2 = ld & ~l3;
      if (l2) susp |= swar_zero_any((x & 0x3E3E3E3Eu) | (~(l2 >> 7) & 0x01010101u));                        // C0, C1
    }
    // expected continuation bytes from the flags of the three preceding bytes (across words, lanes and strips)
    uint32_t Pprev = __shfl_up_sync(FULL, P[3], 1);
    if (l == 0) Pprev = carryP;
    uint32_t bad_l = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t pv = j ? P[j - 1] : Pprev;
      const uint32_t e1 = __funnelshift_l(pv, P[j], 8) & M;
      const uint32_t e2 = (__funnelshift_l(pv, P[j], 16) << 1) & M;
      const uint32_t e3 = (__funnelshift_l(pv, P[j], 24) << 2) & M;
      bad_l |= c[j] ^ (e1 | e2 | e3);
    }
    // lanes past the end still check the three positions behind it (their words are spaces: an expected continuation
    // there means the string ended inside a sequence); lanes further out see nothing
    if (p0 >= n + 3u) bad_l = 0;
    seen |= __reduce_or_sync(FULL, (susp & 0x300u));  // sticky over the string: E2 80 | A8 may straddle two strips
    bad |= bad_l | (susp & M) | (seen == 0x300u ? 1u : 0u);
    carryP = __shfl_sync(FULL, P[3], 31);
    if (!(carryP & 0x80402000u)) carryP = 0;  // only a sequence that is still open at the strip end matters
  }
  TextScan r;
  r.esc = WANT_ESC ? warp_sum(esc) : 0u;
  r.tme = WANT_TME ? warp_sum(tme) : 0u;
  r.exact = WANT_ESC && __any_sync(FULL, bad != 0);
  r.non_ascii = hi_any != 0;
  return r;
}

}  // namespace tgi
