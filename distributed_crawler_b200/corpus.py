"""Synthetic corpus — ctypes wrapper of corpus/libtgcorpus.so.

Measurement infrastructure: fabricates already-fetched messages in the packed batch layout.
Arrays are zero-copy numpy views on the generator's C memory (freed with the Corpus object).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import abi
from .pack import TgBatch, YtBatch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "corpus")
_LIB = None

PROFILE_TEXT_ONLY, PROFILE_MIXED, PROFILE_LINKS = 1, 2, 3


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "libtgcorpus.so")
    src = os.path.join(_DIR, "tgcorpus.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.tgc_telegram.restype = C.c_void_p
        L.tgc_telegram.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.tgc_batch.restype = C.POINTER(abi.TgBatchC)
        L.tgc_batch.argtypes = [C.c_void_p]
        L.tgc_free.argtypes = [C.c_void_p]
        L.tgc_total_bytes.restype = C.c_uint64
        L.tgc_total_bytes.argtypes = [C.c_void_p]
        L.tgc_youtube.restype = C.c_void_p
        L.tgc_youtube.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.tgc_yt_batch.restype = C.POINTER(abi.YtBatchC)
        L.tgc_yt_batch.argtypes = [C.c_void_p]
        L.tgc_yt_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _view(p, nbytes, dt):
    if not nbytes:
        return np.zeros(0, dt)
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (nbytes,)).view(dt)


class Corpus:
    """Telegram corpus shard [first, first+n). `.batch` is a TgBatch over generator-owned memory."""

    def __init__(self, n: int, *, seed: int = 0x5EED0002, first: int = 0,
                 profile: int = PROFILE_MIXED, nthreads: int | None = None):
        nthreads = nthreads or min(os.cpu_count() or 1, 64)
        self._h = _lib().tgc_telegram(seed, first, n, profile, nthreads)
        b = _lib().tgc_batch(self._h).contents
        n_ent = int(_view(b.ent_off, (n + 1) * 4, np.uint32)[n]) if n else 0
        self.batch = TgBatch(
            recs=_view(b.recs, n * 64, abi.TG_REC), strs=_view(b.strs, b.strs_len, np.uint8),
            ent_off=_view(b.ent_off, (n + 1) * 4, np.uint32), ents=_view(b.ents, n_ent * 16, abi.ENTITY),
            react_off=_view(b.react_off, (n + 1) * 4, np.uint32),
            reacts=_view(b.reacts, b.n_reacts * 12, abi.REACTION),
            comment_off=_view(b.comment_off, (n + 1) * 4, np.uint32),
            comments=_view(b.comments, b.n_comments * 32, abi.COMMENT),
            aux=_view(b.aux, b.aux_len, np.uint8), chans=_view(b.chans, b.n_chans * 40, abi.TG_CHAN),
            chan_strs=_view(b.chan_strs, b.chan_strs_len, np.uint8))
        self.total_bytes = int(_lib().tgc_total_bytes(self._h))
        self.batch._owner = self  # the views borrow the generator's memory: keep it alive with them

    def close(self):
        if getattr(self, "_h", None):
            self.batch = None
            _lib().tgc_free(self._h)
            self._h = None

    __del__ = close


class YtCorpus:
    """YouTube corpus shard [first, first+n) in the BASELINE config-4 shape. `.batch` is a YtBatch over
    generator-owned memory."""

    def __init__(self, n: int, *, seed: int = 0x5EED0004, first: int = 0, nthreads: int | None = None):
        nthreads = nthreads or min(os.cpu_count() or 1, 64)
        self._h = _lib().tgc_youtube(seed, first, n, nthreads)
        b = _lib().tgc_yt_batch(self._h).contents
        self.batch = YtBatch(recs=_view(b.recs, n * 80, abi.YT_REC), strs=_view(b.strs, b.strs_len, np.uint8),
                             chans=_view(b.chans, b.n_chans * 64, abi.YT_CHAN),
                             chan_strs=_view(b.chan_strs, b.chan_strs_len, np.uint8))
        self.batch._owner = self

    def close(self):
        if getattr(self, "_h", None):
            self.batch = None
            _lib().tgc_yt_free(self._h)
            self._h = None

    __del__ = close
