"""CPU tests: the C ABI header, its Python mirror and the built library agree."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from distributed_crawler_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tgingest.h")


def test_struct_sizes_match_header():
    probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "tgingest.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(tgi_tg_rec), sizeof(tgi_entity),
         sizeof(tgi_reaction), sizeof(tgi_comment), sizeof(tgi_tg_chan), sizeof(tgi_yt_rec), sizeof(tgi_yt_chan),
         sizeof(tgi_link), sizeof(tgi_tg_batch), sizeof(tgi_yt_batch), sizeof(tgi_config), sizeof(tgi_result),
         sizeof(tgi_stats));
  printf("%zu %zu %zu\n", offsetof(tgi_config, crawl_label), offsetof(tgi_result, kernel_ms), offsetof(tgi_tg_rec, content_type));
  printf("%zu %zu %zu %zu\n", sizeof(tgi_gm_rec), sizeof(tgi_gm_reaction), sizeof(tgi_gm_batch), offsetof(tgi_result, frontier_ms));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    got = list(map(int, out))
    want = [abi.TG_REC.itemsize, abi.ENTITY.itemsize, abi.REACTION.itemsize, abi.COMMENT.itemsize,
            abi.TG_CHAN.itemsize, abi.YT_REC.itemsize, abi.YT_CHAN.itemsize, abi.LINK.itemsize,
            C.sizeof(abi.TgBatchC), C.sizeof(abi.YtBatchC), C.sizeof(abi.ConfigC), C.sizeof(abi.ResultC),
            C.sizeof(abi.StatsC), abi.ConfigC.crawl_label.offset, abi.ResultC.kernel_ms.offset,
            abi.TG_REC.fields["content_type"][1], abi.GM_REC.itemsize, abi.GM_REACTION.itemsize, C.sizeof(abi.GmBatchC),
            abi.ResultC.frontier_ms.offset]
    assert got == want


def test_library_exports_every_declared_symbol(engine_lib):
    decl = set(re.findall(r"^(?:int|void|const char\*)\s+(tgi_\w+)\(", open(HEADER).read(), re.M))
    assert decl == set(engine.EXPORTED_SYMBOLS), decl ^ set(engine.EXPORTED_SYMBOLS)
    for s in decl:
        assert hasattr(engine_lib, s), s


def test_no_cpu_fallback_without_gpu(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine()
    assert ei.value.code == abi.E_NODEVICE


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing in the shipped package may import, link or include it"""
    pkg = os.path.join(ROOT, "distributed_crawler_b200")
    bad = re.compile(r"(^|\n)\s*(import|from)\s+oracle|libtgoracle|pyoracle|#include\s+\"[^\"]*oracle")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", "Makefile")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_generated_piece_table_is_current(tmp_path):
    """csrc/tg_pieces.inc is generated from the Post line program in tools/gen_pieces.py."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_pieces", os.path.join(root, "tools", "gen_pieces.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    want = g.emit_table("Tg", g.TG) + g.emit_lane_table("Tg", g.TG)
    have = open(os.path.join(root, "distributed_crawler_b200", "csrc", "tg_pieces.inc")).read()
    assert have == want, "run python tools/gen_pieces.py"
