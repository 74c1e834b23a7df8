// image.h -- float RGB image with PNG / Radiance HDR output; same interface as the reference's
// src/image.h (image(w,h), setPixel, savePNG, saveHDR).
#pragma once
#include <string>
#include "glm_lite.h"

class image {
private:
    int xSize;
    int ySize;
    ptbglm::vec3 *pixels;
public:
    image(int x, int y);
    ~image();
    image(const image &) = delete;
    image &operator=(const image &) = delete;
    void setPixel(int x, int y, const ptbglm::vec3 &pixel);
    bool savePNG(const std::string &baseFilename, bool quiet = false);
    bool saveHDR(const std::string &baseFilename, bool quiet = false);
};

// RGB8 rows -> PNG file (zlib deflate); used by image::savePNG and by the device-resolved path
bool ptbWritePngRgb8(const std::string &filename, int w, int h, const unsigned char *rgb);
