/*
 * asset_tool.cpp -- TEST INFRASTRUCTURE.  Runs the reference's own asset loader/writer
 * (asset_loading.cpp:9-54 = stb_image v2.26 / stb_image_write v1.15 as vendored in the reference
 * tree) so the product's Radiance .hdr reader and BMP writer can be pinned against it.
 *   ref_asset_tool equirect in.hdr out.f32          -> LoadTexture (vertical flip)
 *   ref_asset_tool cubemap px nx py ny pz nz out.f32 -> LoadCubemapTexture (W x 6H atlas)
 *   ref_asset_tool writebmp in.rgba W H out.bmp      -> WriteImage(..., 4, ...)
 * Prints "W H C" for the loaders.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "asset_loading.h"

int oracle_num_threads = 1;

int main(int argc, char** argv)
{
    if (argc >= 4 && !strcmp(argv[1], "equirect")) {
        texture t = LoadTexture(argv[2]);
        if (!t.Data) { fprintf(stderr, "load failed\n"); return 1; }
        FILE* f = fopen(argv[3], "wb");
        fwrite(t.Data, 4, (size_t)t.Width * t.Height * t.Components, f);
        fclose(f);
        printf("%d %d %d\n", t.Width, t.Height, t.Components);
        return 0;
    }
    if (argc >= 9 && !strcmp(argv[1], "cubemap")) {
        char* names[6] = {argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]};
        texture t = LoadCubemapTexture(names);
        if (!t.Data) { fprintf(stderr, "load failed\n"); return 1; }
        FILE* f = fopen(argv[8], "wb");
        fwrite(t.Data, 4, (size_t)t.Width * t.Height * t.Components, f);
        fclose(f);
        printf("%d %d %d\n", t.Width, t.Height, t.Components);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "writebmp")) {
        int W = atoi(argv[3]), H = atoi(argv[4]);
        std::vector<unsigned char> px((size_t)W * H * 4);
        FILE* f = fopen(argv[2], "rb");
        if (!f || fread(px.data(), 1, px.size(), f) != px.size()) { fprintf(stderr, "read failed\n"); return 1; }
        fclose(f);
        WriteImage(argv[5], W, H, 4, px.data());
        return 0;
    }
    fprintf(stderr, "usage: see asset_tool.cpp\n");
    return 2;
}
