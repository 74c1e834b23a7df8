// image.cpp -- see image.h.  Conversion rule of the reference (src/image.cpp:22-39):
// clamp(pixel, 0, 1) * 255.f, then a truncating cast to unsigned char, RGB8 PNG named <base>.png.
// The reference encodes with stb_image_write; this file writes the PNG container itself over zlib.
#include "image.h"
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>
#include <zlib.h>

image::image(int x, int y) : xSize(x), ySize(y), pixels(new ptbglm::vec3[(size_t)x * y]) {}
image::~image() { delete[] pixels; }

void image::setPixel(int x, int y, const ptbglm::vec3 &pixel) {
    assert(x >= 0 && y >= 0 && x < xSize && y < ySize);
    pixels[(size_t)y * xSize + x] = pixel;
}

static void putChunk(FILE *f, const char type[4], const unsigned char *data, unsigned len) {
    unsigned char hdr[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                            (unsigned char)type[0], (unsigned char)type[1], (unsigned char)type[2], (unsigned char)type[3]};
    fwrite(hdr, 1, 8, f);
    if (len) fwrite(data, 1, len, f);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, data, len);
    unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    fwrite(c, 1, 4, f);
}

bool ptbWritePngRgb8(const std::string &filename, int w, int h, const unsigned char *rgb) {
    std::vector<unsigned char> raw((size_t)h * ((size_t)w * 3 + 1));
    for (int y = 0; y < h; ++y) {
        unsigned char *row = raw.data() + (size_t)y * ((size_t)w * 3 + 1);
        row[0] = 0;                                            // filter type: none
        memcpy(row + 1, rgb + (size_t)y * w * 3, (size_t)w * 3);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) return false;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    unsigned char ihdr[13] = {(unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
                              (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h,
                              8, 2, 0, 0, 0};                  // 8 bit, colour type 2 (RGB)
    putChunk(f, "IHDR", ihdr, 13);
    putChunk(f, "IDAT", comp.data(), (unsigned)clen);
    putChunk(f, "IEND", nullptr, 0);
    return fclose(f) == 0;
}

bool image::savePNG(const std::string &baseFilename, bool quiet) {
    std::vector<unsigned char> bytes((size_t)3 * xSize * ySize);
    for (int y = 0; y < ySize; y++) {
        for (int x = 0; x < xSize; x++) {
            const size_t i = (size_t)y * xSize + x;
            const float c[3] = {pixels[i].x, pixels[i].y, pixels[i].z};
            for (int k = 0; k < 3; ++k) {
                // glm::clamp(v, 0, 1) = min(max(v, 0), 1) with glm's min/max (func_common.inl:409-435)
                float v = c[k];
                v = (v > 0.0f) ? v : 0.0f;
                v = (v < 1.0f) ? v : 1.0f;
                bytes[3 * i + k] = (unsigned char)(v * 255.f);
            }
        }
    }
    const std::string filename = baseFilename + ".png";
    const bool ok = ptbWritePngRgb8(filename, xSize, ySize, bytes.data());
    if (!quiet) std::cout << (ok ? "Saved " : "FAILED to save ") << filename << "." << std::endl;
    return ok;
}

// Radiance RGBE, flat (non-RLE) scanlines (the reference calls stbi_write_hdr, src/image.cpp:41-45)
bool image::saveHDR(const std::string &baseFilename, bool quiet) {
    const std::string filename = baseFilename + ".hdr";
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", ySize, xSize);
    std::vector<unsigned char> row((size_t)xSize * 4);
    for (int y = 0; y < ySize; ++y) {
        for (int x = 0; x < xSize; ++x) {
            const ptbglm::vec3 &p = pixels[(size_t)y * xSize + x];
            float m = std::fmax(p.x, std::fmax(p.y, p.z));
            unsigned char *o = &row[(size_t)x * 4];
            if (!(m > 1e-32f)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            int e;
            float s = std::frexp(m, &e) * 256.0f / m;
            o[0] = (unsigned char)(p.x * s); o[1] = (unsigned char)(p.y * s); o[2] = (unsigned char)(p.z * s);
            o[3] = (unsigned char)(e + 128);
        }
        fwrite(row.data(), 1, row.size(), f);
    }
    const bool ok = fclose(f) == 0;
    if (!quiet) std::cout << "Saved " + filename + "." << std::endl;
    return ok;
}
