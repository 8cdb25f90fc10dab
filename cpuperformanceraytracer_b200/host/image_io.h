// image_io.h -- on-disk formats either side of the hot path:
// Radiance .hdr (RGBE) reader and BMP writer with the semantics of the reference's asset loader
// (asset_loading.cpp:9-54, which uses stb_image / stb_image_write): vertical flip on load so row 0
// is the bottom row, RGB f32, cubemap faces stacked px,nx,py,ny,pz,nz into one W x 6H atlas;
// BMP = what stb_image_write v1.15 writes for 4 components: 24-bit BI_RGB, rows bottom-up, BGR,
// alpha composited against pink.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200pt {

struct HostImage {
    std::vector<float> rgb;  // row-major, 3 floats per texel, row 0 = bottom
    int width = 0, height = 0;
};

// LoadTexture (asset_loading.cpp:9-16).  Returns false and sets *err on failure.
bool LoadRadianceHDR(const std::string& path, HostImage* out, std::string* err);
bool DecodeRadianceHDR(const uint8_t* data, size_t size, HostImage* out, std::string* err);
// LoadCubemapTexture (asset_loading.cpp:18-44): faces in the order px, nx, py, ny, pz, nz
bool LoadCubemapAtlas(const std::string paths[6], HostImage* out, std::string* err);
// WriteImage(filename, w, h, 4, data) (asset_loading.cpp:48-54): `rgba` is the row-major u32 buffer
// that CopyOutputToFile fills (A<<24 | B<<16 | G<<8 | R, row 0 = top)
bool WriteBMP32(const std::string& path, int width, int height, const uint32_t* rgba, std::string* err);
std::vector<uint8_t> EncodeBMP32(int width, int height, const uint32_t* rgba);

}  // namespace b200pt

// C entry points (used by the Python tests and by non-C++ hosts); *data is malloc'ed, free with
// b200pt_io_free.  Return 0 on success.
extern "C" {
int b200pt_io_load_hdr(const char* path, float** data, int* width, int* height);
int b200pt_io_load_cubemap(const char* const paths[6], float** data, int* width, int* height);
int b200pt_io_write_bmp32(const char* path, int width, int height, const uint32_t* rgba);
void b200pt_io_free(void* p);
}
