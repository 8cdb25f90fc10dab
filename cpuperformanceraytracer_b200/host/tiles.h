// tiles.h -- tile geometry of the reference (tiles.h:5-25, tiles.cpp:4-37, and the live table
// UpdateTileInfo in demofox_path_tracing_optimization_v4.cpp:1505-1555), plus where a tile lives in
// the accumulation buffer and in the kernel's work-item space.  Header-only; used by the C ABI
// (b200pt_set_tile_range) and by hosts that schedule per tile like the reference's work queue.
#pragma once
#include <cstdint>

typedef int32_t i32;

struct RenderTileInfo {  // v4.cpp:1141-1147 (tiles.h:5-10 without the size fields)
    i32 TileX, TileY;
    i32 TileWidth, TileHeight;
    i32 TileMinX, TileMaxX;
    i32 TileMinY, TileMaxY;
};

struct TileSet {  // tiles.h:12-17
    i32 MaxDepth;
    i32 NumTilesX, NumTilesY;
    i32 TileWidth, TileHeight;
};

inline i32 RoundIntegerToNextMultiple(i32 i, i32 Multiple) { return ((i + Multiple - 1) / Multiple) * Multiple; }  // tiles.h:21-24

inline TileSet MakeTiles(i32 BufferWidth, i32 BufferHeight, i32 TileWidth, i32 TileHeight)  // tiles.cpp:4-37
{
    TileSet t;
    t.MaxDepth = 0;
    t.NumTilesX = (BufferWidth + TileWidth - 1) / TileWidth;
    t.NumTilesY = (BufferHeight + TileHeight - 1) / TileHeight;
    t.TileWidth = TileWidth;
    t.TileHeight = TileHeight;
    return t;
}

// entry FlatTileIndex = TileX + NumTilesX * TileY of the table UpdateTileInfo fills (v4.cpp:1519-1552)
inline RenderTileInfo MakeTileInfo(i32 BufferWidth, i32 BufferHeight, i32 TileWidth, i32 TileHeight, i32 TileX, i32 TileY)
{
    RenderTileInfo ti;
    ti.TileX = TileX;
    ti.TileY = TileY;
    ti.TileMinX = TileX * TileWidth;
    const i32 maxx = ti.TileMinX + (TileWidth - 1);
    ti.TileMaxX = maxx < BufferWidth ? maxx : (BufferWidth - 1);
    ti.TileMinY = TileY * TileHeight;
    const i32 maxy = ti.TileMinY + (TileHeight - 1);
    ti.TileMaxY = maxy < BufferHeight ? maxy : (BufferHeight - 1);
    ti.TileHeight = ti.TileMaxY - ti.TileMinY + 1;
    ti.TileWidth = ti.TileMaxX - ti.TileMinX + 1;
    return ti;
}

// float offset of a tile in the accumulation buffer: TileXOffset + TileYOffset of RenderTile
// (v4.cpp:1189-1194); tiles follow one another in FlatTileIndex order when the tiling is exact
inline int64_t TileBufferOffset(i32 BufferWidth, i32 NumChannels, const RenderTileInfo& ti)
{
    const int64_t TileSize = (int64_t)ti.TileHeight * ti.TileWidth * NumChannels;
    const int64_t TileYOffset = (int64_t)ti.TileY * ti.TileHeight * BufferWidth * NumChannels;
    return TileSize * ti.TileX + TileYOffset;
}
