// png.cpp -- minimal PNG encoder standing in for massiv-io's writeImage (Lib.hs:75): 8-bit RGB, zlib "stored"
// blocks (no compression library needed), CRC-32 and Adler-32 computed here.
#include <cstdio>

#include "squigly.hpp"

namespace squigly {
namespace {
uint32_t crcTable[256];
bool crcReady = false;
void initCrc() {
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        crcTable[n] = c;
    }
    crcReady = true;
}
uint32_t crc32(const uint8_t *p, size_t n, uint32_t c = 0xffffffffu) {
    if (!crcReady) initCrc();
    for (size_t i = 0; i < n; ++i) c = crcTable[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
void be32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data) {
    be32(out, (uint32_t)data.size());
    std::vector<uint8_t> body(type, type + 4);
    body.insert(body.end(), data.begin(), data.end());
    out.insert(out.end(), body.begin(), body.end());
    be32(out, crc32(body.data(), body.size()) ^ 0xffffffffu);
}
}  // namespace

void writeImage(const std::string &path, const uint8_t *rgb8, int rows, int cols) {
    const size_t dot = path.rfind('.');
    const std::string ext = dot == std::string::npos ? "" : path.substr(dot);
    if (ext != ".png" && ext != ".PNG") throw std::runtime_error("writeImage: unsupported image format for " + path + " (only .png)");
    // raw scanlines: filter byte 0 + RGB row
    std::vector<uint8_t> raw;
    raw.reserve((size_t)rows * ((size_t)cols * 3 + 1));
    for (int y = 0; y < rows; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb8 + (size_t)y * cols * 3, rgb8 + (size_t)(y + 1) * cols * 3);
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + n == raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += n;
    } while (pos < raw.size());
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    be32(ihdr, (uint32_t)cols); be32(ihdr, (uint32_t)rows);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error(path + ": openBinaryFile: does not exist (No such file or directory)");
    const size_t w = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (w != out.size()) throw std::runtime_error(path + ": short write");
}

}  // namespace squigly
