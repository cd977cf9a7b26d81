#!/usr/bin/env python3
"""Writes tests/golden/reference_vectors.json: the known-answer vectors the reference's OWN tests
hold for this path, transcribed literal by literal with file:line provenance.

The reference is Go and cannot be executed here (no Go toolchain), so these are transcriptions of
the test sources, not outputs of a run:
  /root/reference/telegramhelper/channel_links_test.go      (extractChannelLinksFromMessage)
  /root/reference/telegramhelper/username_filter_test.go    (FilterUsername)
  /root/reference/crawl/runner_tandem_test.go               (tandem frontier fan-out)
  /root/reference/crawl/validator_test.go                   (validateSingleEdge: the two cache look-ups)
  /root/reference/chunk/main_test.go                        (processBatches: the combiner's batching rule)
Nothing under /root/reference is read at test time.
"""
import json
import os

LINKS = [  # (test name, go line, content_type, text, entities[(off,len,type,url)], expected set)
    ("PlainTextTmeLink", 64, "messageText", "Check out https://t.me/channelname for news", [], ["channelname"]),
    ("PlainTextTmeLinkNoScheme", 70, "messageText", "Visit t.me/somechan today", [], ["somechan"]),
    ("PlainTextMultipleLinks", 76, "messageText", "t.me/chanone and t.me/chantwo", [], ["chanone", "chantwo"]),
    ("TextEntityTypeTextUrl", 84, "messageText", "click here", [(0, 10, "text_url", "https://t.me/linkedchan")], ["linkedchan"]),
    ("TextEntityTypeTextUrl_NonTme", 98, "messageText", "link", [(0, 4, "text_url", "https://example.com/page")], []),
    ("Mention_ASCII", 116, "messageText", "Hello @testchan!", [(6, 9, "mention", "")], ["testchan"]),
    ("Mention_CyrillicPrefix_UTF16Regression", 134, "messageText", "Привет @testchan", [(7, 9, "mention", "")], ["testchan"]),
    ("Mention_EmojiPrefix_UTF16Regression", 150, "messageText", "😀 @testchan", [(3, 9, "mention", "")], ["testchan"]),
    ("Mention_ArabicPrefix_UTF16Regression", 162, "messageText", "مرحبا @testchan", [(6, 9, "mention", "")], ["testchan"]),
    ("TextEntityTypeUrl_TmeLink", 178, "messageText", "See https://t.me/urlchan for details", [(4, 20, "url", "")], ["urlchan"]),
    ("PhotoCaption", 193, "messagePhoto", "t.me/photochan", [], ["photochan"]),
    ("VideoCaption", 199, "messageVideo", "t.me/videochan", [], ["videochan"]),
    ("DocumentCaption", 205, "messageDocument", "t.me/docchan", [], ["docchan"]),
    ("AnimationCaption", 211, "messageAnimation", "t.me/animchan", [], ["animchan"]),
    ("AudioCaption", 217, "messageAudio", "t.me/audiochan", [], ["audiochan"]),
    ("VoiceNoteCaption", 223, "messageVoiceNote", "t.me/voicechan", [], ["voicechan"]),
    ("ReservedPath_Joinchat", 231, "messageText", "https://t.me/joinchat/abc123", [], []),
    ("ReservedPath_Share", 239, "messageText", "https://t.me/share/url?url=x", [], []),
    ("ReservedPath_Proxy", 247, "messageText", "https://t.me/proxy?server=x", [], []),
    ("Deduplication", 257, "messageText", "Check t.me/samechan", [(6, 13, "text_url", "https://t.me/samechan")], ["samechan"]),
    ("CaseNormalization", 279, "messageText", "t.me/MixedCase", [], ["mixedcase"]),
    ("UnknownContentType", 287, "messageSticker", None, [], []),
    ("TooShortName", 298, "messageText", "t.me/abc", [], []),
    ("Mention_AtEndOfString", 309, "messageText", "@endchan", [(0, 8, "mention", "")], ["endchan"]),
]

FILTER = [  # username_filter_test.go:13-51 (name, username, valid, reason)
    ("valid simple", "testchannel", True, ""), ("valid with underscore", "test_channel", True, ""),
    ("valid with numbers", "channel123", True, ""), ("valid min length", "abcde", True, ""),
    ("valid 32 chars", "abcdefghijklmnopqrstuvwxyz123456", True, ""),
    ("too short 4 chars", "abcd", False, "too_short"), ("too short 1 char", "a", False, "too_short"),
    ("too short empty", "", False, "too_short"),
    ("too long 33 chars", "abcdefghijklmnopqrstuvwxyz1234567", False, "too_long"),
    ("starts with number", "1channel", False, "invalid_start_char"),
    ("starts with underscore", "_channel", False, "invalid_start_char"),
    ("starts with non-ASCII letter", "échannel", False, "invalid_start_char"),
    ("ends with underscore", "channel_", False, "ends_with_underscore"),
    ("contains space", "test channel", False, "invalid_char"), ("contains dash", "test-channel", False, "invalid_char"),
    ("contains dot", "test.channel", False, "invalid_char"), ("contains unicode", "téstchannel", False, "invalid_char"),
    ("ends with _bot", "some_bot", False, "bot_suffix"), ("ends with Bot", "SomeBot", False, "bot_suffix"),
    ("ends with BOT", "SomeBOT", False, "bot_suffix"), ("ends with _Bot", "Test_Bot", False, "bot_suffix"),
    ("looks like path", "usr/local", False, "invalid_char"), ("contains tilde", "home~user", False, "invalid_char"),
    ("contains dot path", "file.name", False, "invalid_char"),
]

TANDEM = {  # crawl/runner_tandem_test.go:15-93 (WithEdges) and :161-214 (invalid channel)
    "with_edges": {
        "go_line": 15, "owner_url": "source_channel",
        "text": "Check out @valid_channel and @another_chan",
        "entities": [(10, 14, "mention", ""), (29, 13, "mention", "")],
        "expected_edges": ["valid_channel", "another_chan"],  # 2 InsertPendingEdge, 1 batch
    },
    "no_edges": {  # TestTandemMode_NoEdges: nothing to insert, no batch created
        "go_line": 97, "owner_url": "source_channel",
        "text": "Just a regular message with no channels", "entities": [], "expected_edges": [],
    },
    "invalid_channel_skipped": {  # TestTandemMode_InvalidChannelSkipped: IsInvalidChannel("invalid_chan") == true
        "go_line": 161, "owner_url": "source_channel",
        "text": "Check @invalid_chan", "entities": [(6, 13, "mention", "")],
        "invalid": ["invalid_chan"], "expected_edges": [],  # InsertPendingEdge / CreatePendingBatch never called
    },
}

VALIDATOR = [  # crawl/validator_test.go: what validateSingleEdge answers from the two caches, before any request
    # (test name, go line, destination_channel, in invalid cache, in discovered cache, status, reason)
    ("Valid", 28, "testchan", False, False, "pending", ""),              # goes on to the HTTP validation
    ("AlreadyInvalid", 145, "badchan", True, False, "invalid", "cached_invalid"),  # IsChannelDiscovered is not consulted
    ("AlreadyDiscovered", 167, "known_chan", False, True, "duplicate", ""),
]

CHUNK = [  # chunk/main_test.go TestProcessBatches_*: (test name, go line, trigger, hard cap, file sizes, expectation)
    ("FlushOnTrigger", 313, 100, 1000, [40, 40, 40], {"batches": [[0, 1, 2]]}),
    ("FlushOnTrigger_ExactSplit", 335, 60, 1000, [50, 50, 50], {"min_batches": 2}),
    ("HardCapDropsOversizedFile", 356, 1000, 100, [200, 50], {"files_in_batches": 1, "dropped": [0]}),
    ("HardCapForceFlush", 392, 10000, 100, [60, 60], {"min_batches": 2, "first_batches": [[0], [1]]}),
    ("FinalPartialBatchFlushed", 415, 10000, 100000, [10, 10], {"batches": [[0, 1]]}),
    ("EmptyChannelProducesNoBatches", 434, 100, 1000, [], {"batches": []}),
    ("TotalUploadSizeTracked", 448, 1, 1000000, [50, 75], {"total_size": 125, "files_in_batches": 2}),
]


def main():
    out = {
        "provenance": "transcribed from the reference's Go test sources; see this script's docstring",
        "channel_links": [dict(name=n, go_file="telegramhelper/channel_links_test.go", go_line=ln, content_type=ct,
                               text=tx, entities=[list(e) for e in en], expected=ex)
                          for n, ln, ct, tx, en, ex in LINKS],
        "filter_username": [dict(name=n, go_file="telegramhelper/username_filter_test.go", username=u, valid=v, reason=r)
                            for n, u, v, r in FILTER],
        "tandem": TANDEM,
        "validator_cache": [dict(name=n, go_file="crawl/validator_test.go", go_line=ln, destination=d, invalid=i, discovered=k,
                                 status=st, reason=r) for n, ln, d, i, k, st, r in VALIDATOR],
        "chunk_batches": [dict(name=n, go_file="chunk/main_test.go", go_line=ln, trigger=t, hard_cap=h, sizes=sz, expect=ex)
                          for n, ln, t, h, sz, ex in CHUNK],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print("wrote", path, len(LINKS), "link vectors,", len(FILTER), "filter vectors,", len(TANDEM), "tandem,", len(VALIDATOR),
          "validator-cache,", len(CHUNK), "chunk-batching vectors")


if __name__ == "__main__":
    main()
