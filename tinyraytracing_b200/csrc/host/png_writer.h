// Minimal PNG writer (8-bit RGB / RGBA, stored deflate blocks, no compression) — fills the role of the
// third-party svpng.inc the reference includes (main.cpp:1,40); written from the PNG / zlib specifications.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>

namespace trt
{
namespace pngdetail
{
inline uint32_t crc32(const uint8_t *p, size_t n, uint32_t c = 0xFFFFFFFFu)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init)
    {
        for (uint32_t i = 0; i < 256; ++i)
        {
            uint32_t r = i;
            for (int k = 0; k < 8; ++k)
                r = (r & 1) ? 0xEDB88320u ^ (r >> 1) : r >> 1;
            table[i] = r;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i)
        c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c;
}
inline void be32(std::vector<uint8_t> &o, uint32_t v)
{
    for (int s = 24; s >= 0; s -= 8)
        o.push_back((uint8_t)(v >> s));
}
inline void chunk(std::vector<uint8_t> &file, const char *type, const std::vector<uint8_t> &body)
{
    be32(file, (uint32_t)body.size());
    size_t at = file.size();
    file.insert(file.end(), type, type + 4);
    file.insert(file.end(), body.begin(), body.end());
    be32(file, ~crc32(&file[at], file.size() - at));
}
} // namespace pngdetail

// img: h rows of w pixels, 3 (alpha == 0) or 4 bytes each, rows top to bottom.
inline bool writePNG(FILE *fp, unsigned w, unsigned h, const unsigned char *img, int alpha)
{
    using namespace pngdetail;
    const size_t bpp = alpha ? 4 : 3, stride = (size_t)w * bpp + 1;
    std::vector<uint8_t> raw(stride * h);
    for (unsigned y = 0; y < h; ++y)
    {
        raw[y * stride] = 0; // filter: none
        for (size_t k = 0; k < stride - 1; ++k)
            raw[y * stride + 1 + k] = img[(size_t)y * (stride - 1) + k];
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0; // adler32
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535)
    {
        const size_t len = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + len >= raw.size() ? 1 : 0);
        z.push_back((uint8_t)(len & 0xFF)), z.push_back((uint8_t)(len >> 8));
        z.push_back((uint8_t)(~len & 0xFF)), z.push_back((uint8_t)((~len >> 8) & 0xFF));
        for (size_t i = 0; i < len; ++i)
        {
            a = (a + raw[off + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        if (raw.empty())
            break;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> file = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'}, ihdr;
    be32(ihdr, w), be32(ihdr, h);
    ihdr.insert(ihdr.end(), {8, (uint8_t)(alpha ? 6 : 2), 0, 0, 0});
    chunk(file, "IHDR", ihdr);
    chunk(file, "IDAT", z);
    chunk(file, "IEND", {});
    return std::fwrite(file.data(), 1, file.size(), fp) == file.size();
}
} // namespace trt
