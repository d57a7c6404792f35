// ImageIO.cpp -- image export of the headless renderer.
//
// The reference's RendererCore::saveImage (src/RendererCore.cpp:608-646) reads the GL renderbuffer and hands it to
// stb_image_write: ".hdr" from the float image (GL_RGB, GL_FLOAT), ".png" / ".jpg" (quality 100) from the 8-bit read-back
// (GL_UNSIGNED_BYTE: clamp to [0, 1], scale by 255, round), flipped vertically so that row 0 of the file is the top row.
// This file writes the same formats from the RGBA float read-backs of the C ABI (rows bottom-up) with small encoders of
// its own -- Radiance RGBE (flat scanlines), PNG (zlib "stored" blocks), baseline JPEG (4:4:4, quantisation tables of ones
// = what quality 100 means in stb, the standard Huffman tables) -- plus PFM and binary PPM.  No third-party code.
#include "ImageIO.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace yune {

namespace {

typedef std::vector<unsigned char> Bytes;

unsigned char to_u8(float v) { v = v != v ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v)); return (unsigned char)(v * 255.0f + 0.5f); }

// top-down 8-bit RGB from the bottom-up float RGBA image
Bytes rgb8_top_down(const float* rgba, int w, int h)
{
    Bytes out((size_t)w * h * 3);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float* p = rgba + 4 * ((size_t)(h - 1 - y) * w + x);
            unsigned char* q = &out[3 * ((size_t)y * w + x)];
            q[0] = to_u8(p[0]); q[1] = to_u8(p[1]); q[2] = to_u8(p[2]);
        }
    return out;
}

bool write_file(const std::string& path, const Bytes& b, std::string& err)
{
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "Error opening image file for writing."; return false; }
    const bool ok = b.empty() || std::fwrite(b.data(), 1, b.size(), f) == b.size();
    if (std::fclose(f) != 0 || !ok) { err = "Error writing image file."; return false; }
    return true;
}

void put_str(Bytes& b, const std::string& s) { b.insert(b.end(), s.begin(), s.end()); }
void put_be32(Bytes& b, uint32_t v) { b.push_back(v >> 24); b.push_back(v >> 16); b.push_back(v >> 8); b.push_back(v); }
void put_be16(Bytes& b, unsigned v) { b.push_back((v >> 8) & 255); b.push_back(v & 255); }

// ---- Radiance RGBE, flat scanlines, top-down ----
Bytes encode_hdr(const float* rgba, int w, int h)
{
    Bytes b;
    put_str(b, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " + std::to_string(h) + " +X " + std::to_string(w) + "\n");
    for (int y = h - 1; y >= 0; y--)
        for (int x = 0; x < w; x++) {
            const float* p = rgba + 4 * ((size_t)y * w + x);
            const float m = std::fmax(p[0], std::fmax(p[1], p[2]));
            unsigned char px[4] = {0, 0, 0, 0};
            if (m > 1e-32f && std::isfinite(m)) {
                int e; const float s = std::frexp(m, &e) * 256.0f / m;
                for (int k = 0; k < 3; k++) { const float v = p[k] * s; px[k] = (unsigned char)(v > 0.0f ? v : 0.0f); }
                px[3] = (unsigned char)(e + 128);
            }
            b.insert(b.end(), px, px + 4);
        }
    return b;
}

// ---- PFM: bottom-up float RGB, little endian (our row order as is) ----
Bytes encode_pfm(const float* rgba, int w, int h)
{
    Bytes b;
    put_str(b, "PF\n" + std::to_string(w) + " " + std::to_string(h) + "\n-1.0\n");
    const size_t n = (size_t)w * h, at = b.size();
    b.resize(at + n * 12);
    for (size_t i = 0; i < n; i++) std::memcpy(&b[at + 12 * i], rgba + 4 * i, 12);
    return b;
}

Bytes encode_ppm(const Bytes& rgb, int w, int h)
{
    Bytes b;
    put_str(b, "P6\n" + std::to_string(w) + " " + std::to_string(h) + "\n255\n");
    b.insert(b.end(), rgb.begin(), rgb.end());
    return b;
}

// ---- PNG: 8-bit RGB, filter 0, deflate "stored" blocks ----
uint32_t crc32(const unsigned char* p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 255] ^ (crc >> 8);
    return ~crc;
}
void png_chunk(Bytes& out, const char* type, const Bytes& data)
{
    put_be32(out, (uint32_t)data.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32(&out[at], out.size() - at));
}
Bytes encode_png(const Bytes& rgb, int w, int h)
{
    Bytes raw; raw.reserve((size_t)h * (3 * (size_t)w + 1));
    for (int y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgb.begin() + 3 * (size_t)y * w, rgb.begin() + 3 * (size_t)(y + 1) * w); }
    Bytes z; z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, bsum = 0;
    for (size_t i = 0; i < raw.size(); ) {
        const size_t len = std::min<size_t>(65535, raw.size() - i);
        z.push_back(i + len == raw.size() ? 1 : 0);
        z.push_back(len & 255); z.push_back(len >> 8); z.push_back(~len & 255); z.push_back((~len >> 8) & 255);
        z.insert(z.end(), raw.begin() + i, raw.begin() + i + len);
        for (size_t k = i; k < i + len; k++) { a = (a + raw[k]) % 65521; bsum = (bsum + a) % 65521; }
        i += len;
    }
    if (raw.empty()) { const unsigned char e[5] = {1, 0, 0, 0xff, 0xff}; z.insert(z.end(), e, e + 5); }
    put_be32(z, (bsum << 16) | a);
    Bytes out; const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.insert(out.end(), sig, sig + 8);
    Bytes ihdr; put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr); png_chunk(out, "IDAT", z); png_chunk(out, "IEND", Bytes());
    return out;
}

// ---- baseline JPEG, YCbCr 4:4:4, quantisation tables of ones (stb's quality 100), standard Huffman tables (T.81 annex K) ----
const unsigned char kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const unsigned char kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const unsigned char kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const unsigned char kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const unsigned char kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const unsigned char kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const unsigned char kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Huff { uint16_t code[256]; unsigned char len[256]; };
Huff make_huff(const unsigned char* bits, const unsigned char* vals)
{
    Huff h; std::memset(&h, 0, sizeof(h));
    unsigned code = 0; int k = 0;
    for (int l = 1; l <= 16; l++) { for (int i = 0; i < bits[l - 1]; i++, k++) { h.code[vals[k]] = (uint16_t)code++; h.len[vals[k]] = (unsigned char)l; } code <<= 1; }
    return h;
}
struct BitWriter {
    Bytes& out; uint32_t acc = 0; int n = 0;
    explicit BitWriter(Bytes& o) : out(o) {}
    void put(unsigned v, int bits)
    {
        acc = (acc << bits) | (v & ((1u << bits) - 1u)); n += bits;
        while (n >= 8) { const unsigned char c = (unsigned char)(acc >> (n - 8)); out.push_back(c); if (c == 0xff) out.push_back(0); n -= 8; }
    }
    void flush() { if (n > 0) put(0x7f, 8 - n); }
};
void dht(Bytes& b, int cls_id, const unsigned char* bits, const unsigned char* vals, int n_vals)
{
    b.push_back(0xff); b.push_back(0xc4); put_be16(b, 2 + 1 + 16 + n_vals); b.push_back((unsigned char)cls_id);
    b.insert(b.end(), bits, bits + 16); b.insert(b.end(), vals, vals + n_vals);
}
int bit_size(int v) { v = v < 0 ? -v : v; int n = 0; while (v) { n++; v >>= 1; } return n; }

Bytes encode_jpg(const Bytes& rgb, int w, int h)
{
    int zig[64];                                             // zig[k] = raster index of the k-th coefficient in zigzag order
    { int k = 0; for (int s = 0; s < 15; s++) for (int i = 0; i <= s; i++) { const int r = (s & 1) ? i : s - i, c = s - r; if (r < 8 && c < 8) zig[k++] = r * 8 + c; } }
    float cs[8][8];                                          // cs[u][x] = C(u) / 2 * cos((2x + 1) u pi / 16)
    for (int u = 0; u < 8; u++) for (int x = 0; x < 8; x++) cs[u][x] = (float)((u == 0 ? std::sqrt(0.5) : 1.0) * 0.5 * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    const Huff hdc[2] = {make_huff(kDcLumBits, kDcVals), make_huff(kDcChrBits, kDcVals)};
    const Huff hac[2] = {make_huff(kAcLumBits, kAcLumVals), make_huff(kAcChrBits, kAcChrVals)};

    Bytes b;
    const unsigned char head[] = {0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    b.insert(b.end(), head, head + sizeof(head));
    for (int t = 0; t < 2; t++) { b.push_back(0xff); b.push_back(0xdb); put_be16(b, 67); b.push_back((unsigned char)t); b.insert(b.end(), 64, 1); }
    b.push_back(0xff); b.push_back(0xc0); put_be16(b, 17); b.push_back(8); put_be16(b, (unsigned)h); put_be16(b, (unsigned)w); b.push_back(3);
    for (int c = 0; c < 3; c++) { b.push_back((unsigned char)(c + 1)); b.push_back(0x11); b.push_back(c == 0 ? 0 : 1); }
    dht(b, 0x00, kDcLumBits, kDcVals, 12); dht(b, 0x10, kAcLumBits, kAcLumVals, 162);
    dht(b, 0x01, kDcChrBits, kDcVals, 12); dht(b, 0x11, kAcChrBits, kAcChrVals, 162);
    const unsigned char sos[] = {0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    b.insert(b.end(), sos, sos + sizeof(sos));

    BitWriter bw(b);
    int pred[3] = {0, 0, 0};
    for (int by = 0; by < h; by += 8)
        for (int bx = 0; bx < w; bx += 8) {
            float comp[3][64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    const int yy = std::min(by + y, h - 1), xx = std::min(bx + x, w - 1);      // edge blocks repeat the border pixel
                    const unsigned char* p = &rgb[3 * ((size_t)yy * w + xx)];
                    const float r = p[0], g = p[1], bl = p[2];
                    comp[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * bl - 128.0f;
                    comp[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * bl;
                    comp[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * bl;
                }
            for (int c = 0; c < 3; c++) {
                float tmp[64]; int q[64];
                for (int y = 0; y < 8; y++) for (int u = 0; u < 8; u++) { float s = 0; for (int x = 0; x < 8; x++) s += comp[c][y * 8 + x] * cs[u][x]; tmp[y * 8 + u] = s; }
                for (int v = 0; v < 8; v++) for (int u = 0; u < 8; u++) { float s = 0; for (int y = 0; y < 8; y++) s += tmp[y * 8 + u] * cs[v][y]; q[v * 8 + u] = (int)std::lrintf(s); }
                const Huff& dc = hdc[c ? 1 : 0]; const Huff& ac = hac[c ? 1 : 0];
                const int diff = q[0] - pred[c]; pred[c] = q[0];
                int nb = bit_size(diff);
                bw.put(dc.code[nb], dc.len[nb]);
                if (nb) bw.put((unsigned)(diff < 0 ? diff - 1 : diff), nb);
                int run = 0;
                for (int k = 1; k < 64; k++) {
                    const int v = q[zig[k]];
                    if (v == 0) { run++; continue; }
                    while (run > 15) { bw.put(ac.code[0xf0], ac.len[0xf0]); run -= 16; }
                    nb = bit_size(v);
                    bw.put(ac.code[(run << 4) | nb], ac.len[(run << 4) | nb]);
                    bw.put((unsigned)(v < 0 ? v - 1 : v), nb);
                    run = 0;
                }
                if (run) bw.put(ac.code[0], ac.len[0]);
            }
        }
    bw.flush();
    b.push_back(0xff); b.push_back(0xd9);
    return b;
}

} // namespace

std::string imageExtension(const std::string& path)
{
    const size_t dot = path.find_last_of('.'), slash = path.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    std::string e = path.substr(dot);
    for (char& c : e) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    return e;
}

bool imageIsLdr(const std::string& ext) { return ext == ".png" || ext == ".jpg" || ext == ".jpeg" || ext == ".ppm"; }
bool imageIsHdr(const std::string& ext) { return ext == ".hdr" || ext == ".pfm"; }

bool writeImage(const std::string& path, const std::string& ext, const float* rgba, int w, int h, std::string& err)
{
    if (!rgba || w <= 0 || h <= 0 || w > 65535 || h > 65535) { err = "writeImage: bad image"; return false; }
    if (ext == ".hdr") return write_file(path, encode_hdr(rgba, w, h), err);
    if (ext == ".pfm") return write_file(path, encode_pfm(rgba, w, h), err);
    if (imageIsLdr(ext)) {
        const Bytes rgb = rgb8_top_down(rgba, w, h);
        return write_file(path, ext == ".png" ? encode_png(rgb, w, h) : (ext == ".ppm" ? encode_ppm(rgb, w, h) : encode_jpg(rgb, w, h)), err);
    }
    err = "unsupported image extension (use .hdr, .png, .jpg, .pfm or .ppm)";
    return false;
}

} // namespace yune
