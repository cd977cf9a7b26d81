#!/usr/bin/env python3
"""Generates distributed_crawler_b200/csrc/tg_pieces.inc: the JSONL line of a Telegram / YouTube Post
as a flat table of pieces (literal template ranges + field references) in model.Post declaration
order (model/data.go:9-75).  The device walker (tg_walk.cuh) interprets the table; keeping the
line as DATA instead of inlined code keeps the emit kernel's instruction footprint inside the
instruction cache.  Re-run after editing: python tools/gen_pieces.py
"""
import os

# piece kinds (must match tg_walk.cuh)
K_LIT, K_FIELD, K_CHAN, K_CFG, K_ESC, K_POSTTYPE, K_COMMENTS, K_REACTIONS, K_OUTLINKS = range(9)
C_NONE, C_USER, C_ALBUM, C_CT_OTHER, C_NOT_CT_OTHER, C_HAS_MEDIA = range(6)
F_MSGNO, F_CHAT, F_VIEW, F_SHARE, F_NCOMM, F_TIME = range(6)
CH_USER, CH_NAME, CH_TITLEQ, CH_CDATA = range(4)
CFG_LABEL, CFG_CREATED_TG, CFG_CREATED_YT, CFG_CAPTURE = range(4)
E_DESC, E_MEDIA, E_HANDLE, E_ALT = range(4)


def L(s, c=C_NONE): return (K_LIT, 0, c, s)
def F(i, c=C_NONE): return (K_FIELD, i, c, "")
def CH(i, c=C_NONE): return (K_CHAN, i, c, "")
def CFG(i): return (K_CFG, i, C_NONE, "")
def E(i, c=C_NONE): return (K_ESC, i, c, "")


LINK = [L("https://t.me/", C_USER), CH(CH_USER, C_USER), L("/", C_USER), F(F_MSGNO, C_USER),
        L("?single", C_ALBUM)]

NULLS = (',"search_terms":null,"search_term_ids":null,"project_ids":null,"exercise_ids":null,'
         '"label_data":null,"labels_metadata":null,"project_labeled_post_ids":null,'
         '"labeler_ids":null,"all_labels":null,"label_ids":null,"is_ad":false,'
         '"transcript_text":"","image_text":"","video_length":null,"is_verified":null,'
         '"channel_data":{"channel_id":"')
MID = (',"platform_name":"Telegram","shared_id":null,"quoted_id":null,"replied_id":null,'
       '"ai_label":null,"root_post_id":null,"engagement_steps_count":0,"ocr_data":null,'
       '"performance_scores":{"likes":null,"shares":null,"comments":null,"views":0},'
       '"has_embed_media":null,"description":"')

TG = ([L('{"post_link":"')] + LINK + [L('","channel_id":"'), F(F_CHAT), L('","post_uid":"'), F(F_MSGNO),
      L('-'), CH(CH_NAME), L('","url":"')] + LINK + [
      L('","published_at":'), F(F_TIME), L(',"created_at":'), CFG(CFG_CREATED_TG),
      L(',"language_code":"","engagement":'), F(F_VIEW), L(',"view_count":'), F(F_VIEW),
      L(',"like_count":0,"share_count":'), F(F_SHARE), L(',"comment_count":'), F(F_NCOMM),
      L(',"crawl_label":"'), CFG(CFG_LABEL), L('","list_ids":null,"channel_name":'), CH(CH_TITLEQ),
      L(NULLS), F(F_CHAT), L('"'), CH(CH_CDATA), L(MID), E(E_DESC),
      L('","repost_channel_data":null,"post_type":["'), E(E_ALT, C_CT_OTHER),
      (K_POSTTYPE, 0, C_NOT_CT_OTHER, ""),
      L('"],"inner_link":{},"post_title":null,"media_data":{"document_name":""},'
        '"is_reply":null,"ad_fields":null,"likes_count":0,"shares_count":'), F(F_SHARE),
      L(',"comments_count":'), F(F_NCOMM), L(',"views_count":'), F(F_VIEW),
      L(',"searchable_text":"","all_text":"","contrast_agent_project_ids":null,'
        '"agent_ids":null,"segment_ids":null,"thumb_url":"","media_url":"'), E(E_MEDIA, C_HAS_MEDIA),
      L('","comments":'), (K_COMMENTS, 0, C_NONE, ""), L(',"reactions":'), (K_REACTIONS, 0, C_NONE, ""),
      L(',"outlinks":['), (K_OUTLINKS, 0, C_NONE, ""), L('],"capture_time":'), CFG(CFG_CAPTURE),
      L(',"handle":"'), E(E_HANDLE), L('"}\n')])


def c_string(b: bytes) -> str:
    out = []
    for ch in b:
        if ch == ord('"'): out.append('\\"')
        elif ch == ord('\\'): out.append('\\\\')
        elif ch == ord('\n'): out.append('\\n')
        elif 32 <= ch < 127: out.append(chr(ch))
        else: out.append('\\%03o' % ch)
    return '"' + ''.join(out) + '"'


K_NOP = 15
LANE_KINDS = (K_LIT, K_FIELD, K_POSTTYPE)


def emit_table(name, prog):
    """Generated from ONE program (the Post line in declaration order):
    (1) the piece table  kind | arg << 4 | cond << 8 | off << 12 | len << 23; literal pieces start
        4-byte aligned in the template blob and carry their ordinal (`arg`), so the emit kernel can
        copy the whole template word by word with a per-piece shift (output offset - template offset);
        entry e belongs to lane e // EPL: an exclusive warp scan over the per-lane length sums gives
        every piece its output offset;
    (2) per template word: owning piece index << 3 | valid bytes;
    (3) the list of pieces that are written cooperatively (variable length);
    (4) the closed-form length of the fixed part of the line."""
    blob = bytearray()
    rows, wmeta = [], []
    const_len = user_lit = album_lit = 0
    nfield = [[0] * 8, [0] * 8]      # [cond NONE, cond USER][field]
    nchan = [[0] * 4, [0] * 4]
    ncfg = [0] * 4
    nesc = {}
    nlit = 0
    for kind, arg, cond, text in prog:
        if kind == K_LIT:
            tb = text.encode()
            if cond == C_NONE: const_len += len(tb)
            elif cond == C_USER: user_lit += len(tb)
            elif cond == C_ALBUM: album_lit += len(tb)
            else: raise ValueError("literal under unsupported condition")
            while len(blob) % 4: blob.append(0)
            base = len(blob)
            blob += tb
            assert len(tb) < 512 and base + len(tb) < 2048
            for o in range(0, len(tb), 4):
                wmeta.append((len(rows) << 3) | min(4, len(tb) - o))  # piece index << 3 | valid bytes
            rows.append(kind | (cond << 8) | (base << 12) | (len(tb) << 23))
            nlit += 1
            continue
        if kind == K_FIELD:
            assert cond in (C_NONE, C_USER)
            nfield[cond == C_USER][arg] += 1
        elif kind == K_CHAN:
            assert cond in (C_NONE, C_USER)
            nchan[cond == C_USER][arg] += 1
        elif kind == K_CFG:
            assert cond == C_NONE
            ncfg[arg] += 1
        elif kind == K_ESC:
            nesc[arg] = nesc.get(arg, 0) + 1
            assert nesc[arg] == 1 and cond == {E_DESC: C_NONE, E_MEDIA: C_HAS_MEDIA, E_HANDLE: C_NONE, E_ALT: C_CT_OTHER}[arg]
        rows.append(kind | (arg << 4) | (cond << 8))
    while len(blob) % 4: blob.append(0)
    assert len(wmeta) == len(blob) // 4
    npieces = len(rows)
    epl = (npieces + 31) // 32
    big = [i for i, e in enumerate(rows) if (e & 15) not in LANE_KINDS]
    ents = rows + [K_NOP] * (32 * epl - npieces)
    lines = [f"// generated by tools/gen_pieces.py — do not edit",
             f"constexpr int k{name}NPieces = {npieces};",
             f"constexpr int k{name}EPL = {epl};            // entries per lane in the lane-parallel path",
             f"constexpr int k{name}NEnt = {32 * epl};",
             f"constexpr int k{name}NBig = {len(big)};          // pieces written cooperatively, in line order",
             f"constexpr int k{name}NLit = {nlit};",
             f"constexpr int k{name}TemplateLen = {len(blob)};   // multiple of 4",
             f"constexpr int k{name}NWords = {len(blob) // 4};",
             f"// closed-form length of the fixed part of the line: L = field lengths, chan / cf = segment lengths",
             f"DEVI uint32_t {name.lower()}_size_fixed(const uint32_t* L, const uint32_t* chan, const uint32_t* cf, bool has_user, bool album) {{",
             f"  uint32_t t = {const_len}u" + "".join(f" + {c}u * L[{j}]" for j, c in enumerate(nfield[0]) if c)
             + "".join(f" + {c}u * chan[{j}]" for j, c in enumerate(nchan[0]) if c)
             + "".join(f" + {c}u * cf[{j}]" for j, c in enumerate(ncfg) if c) + ";",
             f"  if (has_user) t += {user_lit}u" + "".join(f" + {c}u * L[{j}]" for j, c in enumerate(nfield[1]) if c)
             + "".join(f" + {c}u * chan[{j}]" for j, c in enumerate(nchan[1]) if c) + ";",
             f"  if (album) t += {album_lit}u;",
             f"  return t;",
             f"}}",
             f"__device__ __align__(16) const char k{name}Template[] ="]
    bb = bytes(blob)
    for i in range(0, len(bb), 64):
        lines.append("    " + c_string(bb[i:i + 64]))
    lines[-1] += ";"
    lines.append(f"__device__ const uint32_t k{name}Pieces[{len(ents)}] = {{")
    for i in range(0, len(ents), 6):
        lines.append("    " + ", ".join("0x%08xu" % r for r in ents[i:i + 6]) + ",")
    lines.append("};")
    lines.append(f"__device__ const uint16_t k{name}WordMeta[{len(wmeta)}] = {{")
    for i in range(0, len(wmeta), 24):
        lines.append("    " + ", ".join(map(str, wmeta[i:i + 24])) + ",")
    lines.append("};")
    lines.append(f"__device__ const uint8_t k{name}Big[{len(big)}] = {{" + ", ".join(map(str, big)) + "};")
    groups = {"Copy": (K_CHAN, K_CFG), "Esc": (K_ESC,), "Map": (K_COMMENTS, K_REACTIONS, K_OUTLINKS)}
    for gname, kinds in groups.items():  # the phased emit kernel visits the cooperative pieces by kind
        sel = [i for i in big if (rows[i] & 15) in kinds]
        lines.append(f"constexpr int k{name}N{gname} = {len(sel)};")
        lines.append(f"__device__ const uint8_t k{name}{gname}[{len(sel)}] = {{" + ", ".join(map(str, sel)) + "};")
    return "\n".join(lines) + "\n"


def emit_lane_table(name, prog):
    """The same program for the lane-per-record emitter (tg_lane.cuh): every lane streams its own
    line, so literals are fetched as whole 16-byte blocks: they start 16-byte aligned in this blob
    and are zero padded.  Entry = kind | arg << 4 | cond << 8 | (off / 16) << 12 | len << 23."""
    blob = bytearray()
    rows = []
    for kind, arg, cond, text in prog:
        if kind == K_LIT:
            tb = text.encode()
            while len(blob) % 16: blob.append(0)
            base = len(blob)
            blob += tb
            assert len(tb) < 512 and base // 16 < 2048
            rows.append(kind | (cond << 8) | ((base // 16) << 12) | (len(tb) << 23))
        else:
            rows.append(kind | (arg << 4) | (cond << 8))
    while len(blob) % 16: blob.append(0)
    lines = [f"constexpr int k{name}LaneNPieces = {len(rows)};",
             f"constexpr int k{name}LaneTemplateLen = {len(blob)};   // multiple of 16",
             f"__device__ __align__(16) const char k{name}LaneTemplate[] ="]
    bb = bytes(blob)
    for i in range(0, len(bb), 64):
        lines.append("    " + c_string(bb[i:i + 64]))
    lines[-1] += ";"
    lines.append(f"__device__ const uint32_t k{name}LanePieces[{len(rows)}] = {{")
    for i in range(0, len(rows), 6):
        lines.append("    " + ", ".join("0x%08xu" % r for r in rows[i:i + 6]) + ",")
    lines.append("};")
    return "\n".join(lines) + "\n"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "distributed_crawler_b200", "csrc", "tg_pieces.inc")
    with open(out, "w") as f:
        f.write(emit_table("Tg", TG))
        f.write(emit_lane_table("Tg", TG))
    print("wrote", os.path.normpath(out), "pieces:", len(TG))


if __name__ == "__main__":
    main()
