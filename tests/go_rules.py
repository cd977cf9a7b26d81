"""Independent pure-Python restatement of the Go standard-library rules the path depends on
(SURVEY.md Appendix A.6), written from the rules — NOT from oracle/tgoracle.c — so that the oracle is
cross-checked by a second implementation.  Pure-Python loops: small inputs only."""
from __future__ import annotations

import datetime as _dt
import re


def go_decode_rune(b: bytes, i: int):
    """unicode/utf8.DecodeRune: returns (rune, width); (0xFFFD, 1) for invalid."""
    n = len(b) - i
    b0 = b[i]
    if b0 < 0x80:
        return b0, 1
    # python's strict decoder rejects exactly what Go rejects (overlongs, surrogates, > U+10FFFF)
    for w in (2, 3, 4):
        if n >= w:
            try:
                s = b[i:i + w].decode("utf-8")
                if len(s) == 1:
                    return ord(s), w
            except UnicodeDecodeError:
                continue
    return 0xFFFD, 1


def go_json_string(b: bytes) -> bytes:
    out = bytearray(b'"')
    i = 0
    while i < len(b):
        c = b[i]
        if c < 0x80:
            if c >= 0x20 and c not in b'"\\<>&':
                out.append(c)
            elif c in b'"\\':
                out += b"\\" + bytes([c])
            elif c == 8: out += b"\\b"
            elif c == 12: out += b"\\f"
            elif c == 10: out += b"\\n"
            elif c == 13: out += b"\\r"
            elif c == 9: out += b"\\t"
            else: out += b"\\u00%02x" % c
            i += 1
            continue
        r, w = go_decode_rune(b, i)
        if r == 0xFFFD and w == 1:
            out += b"\\ufffd"
        elif r in (0x2028, 0x2029):
            out += b"\\u%04x" % r
        else:
            out += b[i:i + w]
        i += w
    out += b'"'
    return bytes(out)


def go_time_json(sec: int, nsec: int = 0, tz: int = 0):
    t = _dt.datetime(1970, 1, 1) + _dt.timedelta(seconds=sec + tz)
    s = t.strftime("%Y-%m-%dT%H:%M:%S")
    if nsec:
        s += "." + ("%09d" % nsec).rstrip("0")
    if tz == 0:
        s += "Z"
    else:
        a = abs(tz)
        s += ("-" if tz < 0 else "+") + "%02d:%02d" % (a // 3600, a % 3600 // 60)
    return ('"' + s + '"').encode()


def utf16_offset_to_bytes(s: bytes, off: int, length: int):
    """telegramhelper/tdutils.go:55-78 restated literally."""
    i, u16, rune_start = 0, 0, -1
    stop = ((off + length + 2 ** 31) % 2 ** 32) - 2 ** 31
    while i < len(s):
        if u16 == off:
            rune_start = i
        if u16 == stop:
            return rune_start, i
        r, w = go_decode_rune(s, i)
        u16 += 2 if r >= 0x10000 else 1
        i += w
    if rune_start == -1:
        return 0, 0
    return rune_start, len(s)


_CHANNEL_RE = re.compile(rb"(https?://)?t\.me/([a-zA-Z][a-zA-Z0-9_]{4,31})")
_USER_RE = re.compile(rb"(?:@)?([a-zA-Z][a-zA-Z0-9_]{4,31})")
_RESERVED = {b"joinchat", b"addlist", b"addstickers", b"addtheme", b"setlanguage", b"share", b"c", b"s",
             b"iv", b"proxy", b"socks", b"login", b"confirm", b"bg"}


def extract_links(text: bytes | None, entities, aux_urls=None):
    """tdutils.go:897-949 with python's `re` standing in for Go's regexp (both leftmost, greedy; the
    patterns have no alternation whose priority could differ).  Returns [(name, src)] in first-
    insertion order, or None where Go would panic."""
    out, seen = [], set()

    def add(name, src):
        name = name.lower()
        if name not in seen:
            seen.add(name)
            out.append((name, src))

    def chan(m, src):
        if m and m.group(2).lower() not in _RESERVED:
            add(m.group(2), src)

    if text is None:
        return out
    for off, ln, typ, url in entities:
        if typ == "text_url":
            chan(_CHANNEL_RE.search(url.encode() if isinstance(url, str) else url), "text_url")
        elif typ in ("mention", "url"):
            st, en = utf16_offset_to_bytes(text, off, ln)
            if st < en and en <= len(text):
                if st < 0:
                    return None
                sl = text[st:en]
                if typ == "mention":
                    m = _USER_RE.search(sl)
                    if m:
                        add(m.group(1), "mention")
                else:
                    chan(_CHANNEL_RE.search(sl), "url")
    for m in _CHANNEL_RE.finditer(text):
        chan(m, "plaintext")
    return out


def filter_username(u: bytes) -> str:
    """username_filter.go:26-68."""
    if len(u) < 5: return "too_short"
    if len(u) > 32: return "too_long"
    if not (65 <= u[0] <= 90 or 97 <= u[0] <= 122): return "invalid_start_char"
    if u[-1:] == b"_": return "ends_with_underscore"
    if not re.fullmatch(rb"[A-Za-z0-9_]+", u): return "invalid_char"
    if u.lower().endswith(b"bot"): return "bot_suffix"
    return ""


_ISO = re.compile(rb"^P(?:(\d+)D)?(?:T(?:(\d+)H)?(?:(\d+)M)?(?:(\d+)S)?)?$")


def parse_iso8601_duration(s: bytes):
    m = _ISO.match(s)
    if not m or s.endswith(b"\n"):
        return None
    tot = 0
    for g, mul in zip(m.groups(), (86400, 3600, 60, 1)):
        if g:
            v = min(int(g), 2 ** 63 - 1)
            tot += v * mul
    tot &= 2 ** 64 - 1
    return tot - 2 ** 64 if tot >= 2 ** 63 else tot
