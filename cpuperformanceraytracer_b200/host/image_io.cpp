// image_io.cpp -- see image_io.h.  Own implementation of the two file formats; pinned against the
// reference's loader/writer by tests/test_host_io.py (bit-exact texels, byte-exact BMP).
#include "image_io.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b200pt {
namespace {

bool read_file(const std::string& path, std::vector<uint8_t>* out, std::string* err)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    size_t got = out->empty() ? 0 : std::fread(out->data(), 1, out->size(), f);
    std::fclose(f);
    if (got != out->size()) {
        if (err) *err = "short read on " + path;
        return false;
    }
    return true;
}

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool overrun = false;
    int get()
    {
        if (p < end) return *p++;
        overrun = true;  // the decoder fails on truncated pixel data instead of zero-filling it
        return 0;
    }
    bool eof() const { return p >= end; }
    std::string line()
    {
        std::string s;
        while (p < end) {
            char c = (char)*p++;
            if (c == '\n') break;
            if (s.size() < 1022) s.push_back(c);
        }
        return s;
    }
};

// RGBE -> float: mantissa * 2^(e - 136), zero when e == 0
inline void rgbe_to_float(const uint8_t* px, float* out)
{
    if (px[3] != 0) {
        const float scale = (float)std::ldexp(1.0f, (int)px[3] - (128 + 8));
        out[0] = px[0] * scale;
        out[1] = px[1] * scale;
        out[2] = px[2] * scale;
    } else {
        out[0] = out[1] = out[2] = 0.f;
    }
}

}  // namespace

bool DecodeRadianceHDR(const uint8_t* data, size_t size, HostImage* out, std::string* err)
{
    auto fail = [&](const char* m) {
        if (err) *err = m;
        return false;
    };
    Cursor c{data, data + size, false};
    const std::string magic = c.line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE") return fail("not a Radiance HDR file");
    bool rle_rgbe = false;
    for (;;) {
        if (c.eof()) return fail("truncated header");
        const std::string ln = c.line();
        if (ln.empty()) break;
        if (ln == "FORMAT=32-bit_rle_rgbe") rle_rgbe = true;
    }
    if (!rle_rgbe) return fail("unsupported HDR format (need FORMAT=32-bit_rle_rgbe)");
    const std::string dims = c.line();
    int h = 0, w = 0;
    if (std::sscanf(dims.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0 || w > (1 << 24) || h > (1 << 24))
        return fail("unsupported HDR data layout (need -Y h +X w)");

    bool flat = (w < 8 || w >= 32768);
    if (!flat) {
        // new-style RLE scanlines start with 2, 2, hi, lo (hi < 128); anything else means flat data
        if (c.end - c.p >= 4 && !(c.p[0] == 2 && c.p[1] == 2 && !(c.p[2] & 0x80))) flat = true;
    }
    // Size the image against the bytes that are actually there BEFORE allocating: a flat file holds 4 bytes per pixel, an
    // RLE one at least 4 marker bytes + 2 bytes per channel per scanline.  The renderer cannot use more than 2^24 floats
    // either (the reference indexes texels through binary32 arithmetic, texture.cpp:56-65; b200pt_set_env enforces it).
    const size_t remaining = (size_t)(c.end - c.p);
    const unsigned long long pixels = (unsigned long long)w * (unsigned long long)h;
    if (pixels * 3ull >= (1ull << 24)) return fail("HDR image too large (>= 2^24 floats)");
    if (flat ? pixels * 4ull > remaining : (unsigned long long)h * 12ull > remaining) return fail("truncated HDR pixel data");
    std::vector<float> top_down((size_t)w * h * 3);
    std::vector<uint8_t> scan((size_t)w * 4);
    if (flat) {
        for (size_t i = 0; i < (size_t)w * h; i++) {
            uint8_t px[4] = {(uint8_t)c.get(), (uint8_t)c.get(), (uint8_t)c.get(), (uint8_t)c.get()};
            rgbe_to_float(px, &top_down[i * 3]);
        }
    } else {
        for (int j = 0; j < h; j++) {
            const int c1 = c.get(), c2 = c.get(), hi = c.get(), lo = c.get();
            if (c1 != 2 || c2 != 2 || (hi & 0x80)) return fail("mixed flat / RLE scanlines are not supported");
            if (((hi << 8) | lo) != w) return fail("invalid decoded scanline length");
            for (int k = 0; k < 4; k++) {
                int i = 0;
                while (i < w) {
                    int count = c.get();
                    if (count > 128) {  // run
                        const uint8_t value = (uint8_t)c.get();
                        count -= 128;
                        if (count > w - i) return fail("bad RLE data in HDR");
                        for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = value;
                    } else {  // literal
                        if (count > w - i) return fail("bad RLE data in HDR");
                        if (count == 0) return fail("bad RLE data in HDR");
                        for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = (uint8_t)c.get();
                    }
                }
            }
            if (c.overrun) return fail("truncated HDR pixel data");
            for (int i = 0; i < w; i++) rgbe_to_float(&scan[(size_t)i * 4], &top_down[((size_t)j * w + i) * 3]);
        }
    }
    if (c.overrun) return fail("truncated HDR pixel data");
    // stbi_set_flip_vertically_on_load(true), asset_loading.cpp:12
    out->width = w;
    out->height = h;
    out->rgb.resize(top_down.size());
    const size_t row = (size_t)w * 3;
    for (int j = 0; j < h; j++) std::memcpy(&out->rgb[(size_t)j * row], &top_down[(size_t)(h - 1 - j) * row], row * sizeof(float));
    return true;
}

bool LoadRadianceHDR(const std::string& path, HostImage* out, std::string* err)
{
    std::vector<uint8_t> bytes;
    if (!read_file(path, &bytes, err)) return false;
    return DecodeRadianceHDR(bytes.data(), bytes.size(), out, err);
}

bool LoadCubemapAtlas(const std::string paths[6], HostImage* out, std::string* err)
{
    HostImage face;
    out->rgb.clear();
    for (int i = 0; i < 6; i++) {
        if (!LoadRadianceHDR(paths[i], &face, err)) return false;
        if (i == 0) {
            out->width = face.width;
            out->height = face.height * 6;
            out->rgb.reserve(face.rgb.size() * 6);
        } else if (face.width != out->width || face.height * 6 != out->height) {
            if (err) *err = "cubemap faces differ in size";
            return false;
        }
        out->rgb.insert(out->rgb.end(), face.rgb.begin(), face.rgb.end());  // faces stacked along rows
    }
    return true;
}

std::vector<uint8_t> EncodeBMP32(int width, int height, const uint32_t* rgba)
{
    // What stb_image_write v1.15 (the version vendored by the reference) produces for a 4-component
    // buffer: a 24-bit BI_RGB bitmap, rows bottom-up and padded to 4 bytes, pixels B,G,R after
    // compositing RGBA against a pink (255,0,255) background with integer arithmetic.
    const uint32_t pad = (uint32_t)(-(width * 3)) & 3u;
    const uint32_t header = 14 + 40;
    const uint32_t total = header + ((uint32_t)width * 3u + pad) * (uint32_t)height;
    std::vector<uint8_t> b;
    b.reserve(total);
    auto u16 = [&](uint32_t v) { b.push_back(v & 0xFF); b.push_back((v >> 8) & 0xFF); };
    auto u32 = [&](uint32_t v) { u16(v & 0xFFFF); u16(v >> 16); };
    b.push_back('B'); b.push_back('M');
    u32(total); u16(0); u16(0); u32(header);                 // BITMAPFILEHEADER
    u32(40); u32((uint32_t)width); u32((uint32_t)height);    // BITMAPINFOHEADER
    u16(1); u16(24); u32(0); u32(0); u32(0); u32(0); u32(0); u32(0);
    const int bg[3] = {255, 0, 255};
    for (int y = height - 1; y >= 0; y--) {
        const uint32_t* row = rgba + (size_t)y * width;
        for (int x = 0; x < width; x++) {
            const uint32_t p = row[x];  // memory order R, G, B, A
            const int d[3] = {(int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)((p >> 16) & 0xFF)};
            const int a = (int)(p >> 24);
            uint8_t px[3];
            for (int k = 0; k < 3; k++) px[k] = (uint8_t)(bg[k] + ((d[k] - bg[k]) * a) / 255);
            b.push_back(px[2]); b.push_back(px[1]); b.push_back(px[0]);
        }
        for (uint32_t k = 0; k < pad; k++) b.push_back(0);
    }
    return b;
}

bool WriteBMP32(const std::string& path, int width, int height, const uint32_t* rgba, std::string* err)
{
    const std::vector<uint8_t> b = EncodeBMP32(width, height, rgba);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    const size_t n = std::fwrite(b.data(), 1, b.size(), f);
    std::fclose(f);
    if (n != b.size()) {
        if (err) *err = "short write on " + path;
        return false;
    }
    return true;
}

}  // namespace b200pt

namespace {
int export_image(const b200pt::HostImage& img, float** data, int* width, int* height)
{
    *data = static_cast<float*>(std::malloc(img.rgb.size() * sizeof(float)));
    if (!*data) return 2;
    std::memcpy(*data, img.rgb.data(), img.rgb.size() * sizeof(float));
    *width = img.width;
    *height = img.height;
    return 0;
}
}  // namespace

extern "C" {
int b200pt_io_load_hdr(const char* path, float** data, int* width, int* height)
{
    b200pt::HostImage img;
    std::string err;
    try {  // no exception may cross the C boundary (std::bad_alloc on a hostile header, ...)
        if (!path || !data || !width || !height || !b200pt::LoadRadianceHDR(path, &img, &err)) return 1;
        return export_image(img, data, width, height);
    } catch (...) {
        return 1;
    }
}
int b200pt_io_load_cubemap(const char* const paths[6], float** data, int* width, int* height)
{
    b200pt::HostImage img;
    std::string err, p[6];
    if (!paths || !data || !width || !height) return 1;
    try {
        for (int i = 0; i < 6; i++) p[i] = paths[i] ? paths[i] : "";
        if (!b200pt::LoadCubemapAtlas(p, &img, &err)) return 1;
        return export_image(img, data, width, height);
    } catch (...) {
        return 1;
    }
}
int b200pt_io_write_bmp32(const char* path, int width, int height, const uint32_t* rgba)
{
    std::string err;
    return (path && rgba && width > 0 && height > 0 && b200pt::WriteBMP32(path, width, height, rgba, &err)) ? 0 : 1;
}
void b200pt_io_free(void* p) { std::free(p); }
}
