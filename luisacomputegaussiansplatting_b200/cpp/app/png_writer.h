// png_writer.h -- minimal 8-bit RGB PNG encoder on zlib (the reference uses stb_image_write from
// LuisaCompute's ext tree, app/main.cpp:24,339, which is not available offline).
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace lcgs
{

inline bool write_png_rgb8(const std::string& path, int w, int h, const uint8_t* rgb)
{
    std::vector<uint8_t> raw((size_t)h * (1 + (size_t)w * 3));
    for (int y = 0; y < h; y++) {
        uint8_t* row = raw.data() + (size_t)y * (1 + (size_t)w * 3);
        row[0]       = 0;  // filter: none
        std::copy(rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3, row + 1);
    }
    uLongf               zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto be32 = [](uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
    auto chunk = [&](const char* tag, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        be32(hdr, len);
        std::copy(tag, tag + 4, hdr + 4);
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(data, 1, len, f);
        uLong crc = crc32(0L, reinterpret_cast<const Bytef*>(tag), 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t c[4];
        be32(c, (uint32_t)crc);
        std::fwrite(c, 1, 4, f);
    };
    const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    std::fwrite(sig, 1, 8, f);
    uint8_t ihdr[13];
    be32(ihdr, (uint32_t)w);
    be32(ihdr + 4, (uint32_t)h);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
    return std::fclose(f) == 0;
}

}  // namespace lcgs
