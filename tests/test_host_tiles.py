"""host/tiles.h mirrors the reference's tile geometry (tiles.h, tiles.cpp, UpdateTileInfo v4.cpp:1505-1555);
compile a tiny program against it and compare with the formulas of SURVEY.md 8a a11."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include <cstdio>
#include "tiles.h"
int main() {
    printf("%d %d\n", RoundIntegerToNextMultiple(1279, 8), RoundIntegerToNextMultiple(1280, 8));
    TileSet t = MakeTiles(1920, 1080, 192, 72);
    printf("%d %d\n", t.NumTilesX, t.NumTilesY);
    t = MakeTiles(1000, 500, 192, 72);
    printf("%d %d\n", t.NumTilesX, t.NumTilesY);
    RenderTileInfo a = MakeTileInfo(1920, 1080, 192, 72, 3, 7);
    printf("%d %d %d %d %d %d %lld\n", a.TileMinX, a.TileMaxX, a.TileMinY, a.TileMaxY, a.TileWidth, a.TileHeight,
           (long long)TileBufferOffset(1920, 3, a));
    RenderTileInfo b = MakeTileInfo(1000, 500, 192, 72, 5, 6);  // clamped edge tile
    printf("%d %d %d %d %d %d\n", b.TileMinX, b.TileMaxX, b.TileMinY, b.TileMaxY, b.TileWidth, b.TileHeight);
}
'''


def test_tiles_header(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(PROG)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "cpuperformanceraytracer_b200", "host"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "1280 1280"
    assert out[1] == "10 15"
    assert out[2] == "6 7"
    # tile (3,7): float offset = ty*TH*W*3 + tx*TW*TH*3
    assert out[3] == "576 767 504 575 192 72 %d" % (7 * 72 * 1920 * 3 + 3 * 192 * 72 * 3)
    assert out[4] == "960 999 432 499 40 68"
