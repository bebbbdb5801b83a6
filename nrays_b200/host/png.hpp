// png.hpp — minimal PNG codec on zlib for the C++ host (the reference uses servo's rust-png / stb_image,
// Cargo.toml:18-22; both are outside the render path).  Decoder: 8-bit gray / gray+alpha / RGB / RGBA /
// palette, non-interlaced.  Encoder: RGB8.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace nrays {
namespace png {

struct Decoded {
  uint32_t w = 0, h = 0;
  int depth = 0;  // channels per pixel after decoding (1..4), 8 bits each
  std::vector<uint8_t> data;
};

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool load(const std::string &path, Decoded &out) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  std::fclose(f);
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) return false;
  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int bit_depth = 0, color = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  while (pos + 12 <= buf.size()) {
    uint32_t len = be32(&buf[pos]);
    std::string type((const char *)&buf[pos + 4], 4);
    const uint8_t *d = &buf[pos + 8];
    if (pos + 12 + len > buf.size()) return false;
    if (type == "IHDR") {
      w = be32(d), h = be32(d + 4), bit_depth = d[8], color = d[9], interlace = d[12];
    } else if (type == "IDAT") {
      idat.insert(idat.end(), d, d + len);
    } else if (type == "PLTE") {
      plte.assign(d, d + len);
    } else if (type == "tRNS") {
      trns.assign(d, d + len);
    } else if (type == "IEND") {
      break;
    }
    pos += 12 + len;
  }
  if (!w || !h || bit_depth != 8 || interlace != 0) return false;
  int ch = color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : color == 6 ? 4 : 0;
  if (!ch) return false;
  size_t stride = (size_t)w * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf rawlen = raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
  std::vector<uint8_t> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t *src = &raw[(stride + 1) * y];
    uint8_t *dst = &img[stride * y];
    const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
    int ft = src[0];
    for (size_t x = 0; x < stride; ++x) {
      int a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
      int v = src[1 + x];
      switch (ft) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: {
          int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: break;
      }
      dst[x] = (uint8_t)v;
    }
  }
  out.w = w, out.h = h;
  if (color == 3) {  // palette -> RGB(A)
    bool alpha = !trns.empty();
    out.depth = alpha ? 4 : 3;
    out.data.resize((size_t)w * h * out.depth);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      size_t k = img[i];
      for (int c = 0; c < 3; ++c) out.data[i * out.depth + c] = 3 * k + c < plte.size() ? plte[3 * k + c] : 0;
      if (alpha) out.data[i * 4 + 3] = k < trns.size() ? trns[k] : 255;
    }
  } else {
    out.depth = ch;
    out.data.swap(img);
  }
  return true;
}

inline void chunk(std::vector<uint8_t> &o, const char *type, const std::vector<uint8_t> &d) {
  uint32_t len = (uint32_t)d.size();
  uint8_t l[4] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len};
  o.insert(o.end(), l, l + 4);
  size_t start = o.size();
  o.insert(o.end(), type, type + 4);
  o.insert(o.end(), d.begin(), d.end());
  uint32_t c = (uint32_t)crc32(0L, &o[start], (uInt)(o.size() - start));
  uint8_t cc[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
  o.insert(o.end(), cc, cc + 4);
}

// store_png of an RGB8 image (src/image.rs:79-89)
inline bool save_rgb8(const std::string &path, const uint8_t *rgb, uint32_t w, uint32_t h) {
  std::vector<uint8_t> raw((size_t)(w * 3 + 1) * h);
  for (uint32_t y = 0; y < h; ++y) {
    raw[(size_t)(w * 3 + 1) * y] = 0;
    std::memcpy(&raw[(size_t)(w * 3 + 1) * y + 1], rgb + (size_t)w * 3 * y, (size_t)w * 3);
  }
  uLongf clen = compressBound(raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) return false;
  comp.resize(clen);
  std::vector<uint8_t> o = {137, 80, 78, 71, 13, 10, 26, 10};
  std::vector<uint8_t> ihdr = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                               (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
  chunk(o, "IHDR", ihdr);
  chunk(o, "IDAT", comp);
  chunk(o, "IEND", {});
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = std::fwrite(o.data(), 1, o.size(), f) == o.size();
  std::fclose(f);
  return ok;
}

}  // namespace png
}  // namespace nrays
