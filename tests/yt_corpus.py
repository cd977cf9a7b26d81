"""Small deterministic YouTube corpus for the config-4 parity tests (SURVEY.md §8d shape, Python
generated: sizes the oracle finishes in seconds)."""
import random

from distributed_crawler_b200.pack import YouTubeChannel, YouTubeVideo, pack_youtube

_B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"
_WORDS = ["video", "новости", "смотрите", "канал", "subscribe", "плейлист", "中文", "字幕", "😀", "live", "&", "<3", "\"quoted\"",
          "line\nbreak", "tab\there", "обзор", "2024", "часть", "#shorts", "مرحبا", " ", "a\\b"]


def make_youtube(n: int, seed: int = 7, n_chans: int = 37):
    rng = random.Random(seed)
    chans = []
    for c in range(n_chans):
        cid = ("@" + "".join(rng.choice("abcdefghij_.-") for _ in range(rng.randrange(3, 20)))) if c % 7 == 3 else \
              "UC" + "".join(rng.choice(_B64) for _ in range(22))
        chans.append(YouTubeChannel(
            id=cid, title=" ".join(rng.choice(_WORDS) for _ in range(rng.randrange(1, 5))),
            description=" ".join(rng.choice(_WORDS) for _ in range(rng.randrange(0, 30))),
            thumb_default="https://yt3.ggpht.com/" + "".join(rng.choice(_B64) for _ in range(30)) if c % 5 else "",
            country=rng.choice(["", "US", "RU", "DE"]), subscriber_count=rng.randrange(0, 10 ** 8),
            view_count=rng.randrange(0, 10 ** 11), video_count=rng.randrange(0, 10 ** 5),
            published_sec=rng.randrange(1_100_000_000, 1_700_000_000), published_nsec=rng.choice([0, 0, 123_000_000]),
            cached=(c % 6 != 5)))
    vids = []
    for i in range(n):
        words = []
        for _ in range(int(rng.lognormvariate(3.5, 1.0)) % 600):
            u = rng.random()
            if u < 0.05:
                tail = rng.choice(["", ".", ",", ")!", "?\"", "'", "/path?q=1&x=<y>", ":"])
                words.append(rng.choice(["http://", "https://"]) + rng.choice(["example.com/", "t.me/chan", "bit.ly/", "youtu.be/"]) +
                             "".join(rng.choice(_B64) for _ in range(rng.randrange(0, 9))) + tail)
            elif u < 0.065:
                words.append("https://www.youtube.com/channel/UC" + "".join(rng.choice(_B64) for _ in range(rng.choice([22, 22, 40]))))
            elif u < 0.08:
                words.append("youtube.com/@" + "".join(rng.choice("abcxyz019_.-") for _ in range(rng.choice([5, 12, 45]))) + rng.choice(["", "/videos", "!"]))
            elif u < 0.085:
                words.append(rng.choice(["http://", "https:// x", "https://", "http", "xhttps://a.b", "https://dup.example/1", "https://dup.example/1"]))
            else:
                words.append(rng.choice(_WORDS))
        desc = " ".join(words)
        if i % 97 == 0:
            desc = desc.encode()[: rng.randrange(0, 40)] + b"\xff\xe2\x80" + desc.encode()[:50]
        views = int(rng.lognormvariate(8, 3)) if i % 41 else rng.choice([0, 2 ** 53, 2 ** 53 + 1, 2 ** 62 + 12345, 9_223_372_036_854_775_807, 12345678901234567890 // 2])
        dur = rng.choice(["PT%dM%dS" % (rng.randrange(60), rng.randrange(60)), "PT%dH%dM%dS" % (rng.randrange(5), rng.randrange(60), rng.randrange(60)),
                          "P0D", "", "P1DT2H", "PT", "P", "PT15M", "PT1H2M3", "PT99999999999999999999S", "bogus", "PT5S "])
        keys = [k for k in ("default", "medium", "high", "standard", "maxres") if rng.random() < 0.8]
        thumbs = {k: ("https://i.ytimg.com/vi/%s/%s.jpg" % (i, k) if rng.random() < 0.95 else "") for k in keys}
        vids.append(YouTubeVideo(
            id="".join(rng.choice(_B64) for _ in range(11)), title=" ".join(rng.choice(_WORDS) for _ in range(rng.randrange(0, 12))),
            description=desc, published_sec=rng.randrange(1_100_000_000, 1_760_000_000) if i % 53 else rng.choice([-62135596800, 253402300800]),
            published_nsec=rng.choice([0, 0, 0, 500_000_000]), view_count=views, like_count=views // 30, comment_count=views // 300,
            duration=dur, thumbnails=thumbs, language=rng.choice(["", "en", "ru", "zh-Hans"]), channel=rng.randrange(n_chans)))
    return pack_youtube(vids, chans), vids, chans


_PLAIN = ["the", "video", "about", "channel", "new", "watch", "and", "more", "from", "this", "week", "episode", "review", "how", "to",
          "guide", "music", "official", "live", "part", "best", "of", "2024", "full", "with", "our", "your", "for", "you", "in"]


def make_youtube_config4(n: int, seed: int = 0x5EED0004, n_chans: int = 1000):
    """BASELINE config 4 shape: title lognormal median 45 B, description lognormal median 400 B
    clipped at 5000 with URLs Poisson(1.5) of which 10 % are youtube.com/channel/UC… or youtube.com/@handle,
    views lognormal(8, 3), likes = views/30, comments = views/300, PT#H#M#S durations (1 % P0D, 0.5 % empty,
    0.5 % P1DT…), 3-5 thumbnails.  2 % of the descriptions carry characters that need escaping."""
    rng = random.Random(seed)
    chans = [YouTubeChannel(id="UC" + "".join(rng.choice(_B64) for _ in range(22)), title="Channel %d" % c,
                            description=" ".join(rng.choice(_PLAIN) for _ in range(20)),
                            thumb_default="https://yt3.ggpht.com/" + "".join(rng.choice(_B64) for _ in range(30)), country="US",
                            subscriber_count=rng.randrange(0, 10 ** 7), view_count=rng.randrange(0, 10 ** 10),
                            video_count=rng.randrange(0, 10 ** 4), published_sec=rng.randrange(1_100_000_000, 1_700_000_000),
                            cached=True) for c in range(n_chans)]

    def text(nbytes, urls):
        words, size = [], 0
        while size < nbytes:
            w = rng.choice(_PLAIN)
            words.append(w)
            size += len(w) + 1
        for _ in range(urls):
            u = rng.random()
            if u < 0.05:
                link = "https://www.youtube.com/channel/UC" + "".join(rng.choice(_B64) for _ in range(22))
            elif u < 0.10:
                link = "https://youtube.com/@" + "".join(rng.choice("abcdefghij_") for _ in range(rng.randrange(5, 15)))
            else:
                link = "https://example.com/" + "".join(rng.choice(_B64) for _ in range(rng.randrange(4, 20)))
            words.insert(rng.randrange(len(words) + 1), link)
        return " ".join(words)

    def poisson(lam):
        k, p, L = 0, 1.0, 2.718281828 ** -lam
        while True:
            p *= rng.random()
            if p <= L:
                return k
            k += 1

    vids = []
    for i in range(n):
        desc = text(min(int(rng.lognormvariate(5.99, 0.9)), 5000), poisson(1.5))
        if rng.random() < 0.02:
            desc = desc.replace(" ", "\n", 3) + ' "quoted" <tag> & more'
        views = int(rng.lognormvariate(8, 3))
        u = rng.random()
        dur = "P0D" if u < 0.01 else "" if u < 0.015 else "P1DT2H" if u < 0.02 else "PT%dH%dM%dS" % (rng.randrange(3), rng.randrange(60), rng.randrange(60))
        keys = ["default", "medium", "high", "standard", "maxres"][: rng.randrange(3, 6)]
        vids.append(YouTubeVideo(
            id="".join(rng.choice(_B64) for _ in range(11)), title=text(int(rng.lognormvariate(3.8, 0.5)), 0), description=desc,
            published_sec=rng.randrange(1_300_000_000, 1_760_000_000), view_count=views, like_count=views // 30, comment_count=views // 300,
            duration=dur, thumbnails={k: "https://i.ytimg.com/vi/%s/%s.jpg" % (i, k) for k in keys}, language="en",
            channel=rng.randrange(n_chans)))
    return pack_youtube(vids, chans), vids, chans
