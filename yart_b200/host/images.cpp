// images.cpp — image decoders of the host layer: what the reference obtains from stb_image v2.30 (vendored under
// stb-image/, a third-party public-domain library), restated so that every decoded byte is the one stb_image yields.
//
//   decodeImageRGBA8  ↔ stbi_load_from_memory(data, len, &w, &h, nullptr, 4)     src/core/texture.hpp:62-90 (loadTexture)
//        PNG  (8 / 16-bit and 1 / 2 / 4-bit grey + palette, tRNS, Adam7 interlace)      stb_image.h "png" section
//        JPEG (baseline + progressive Huffman, 8-bit, 1 or 3 components, restart intervals, h/v factors 1..4,
//              JFIF / Adobe APP14 RGB)                                                   stb_image.h "jpeg" section
//   decodeRadianceHdr ↔ stbi_loadf(filename, &w, &h, nullptr, 4) for ".hdr"             src/core/texture.cpp:21-35 (loadTextureHDR)
//
// The pieces whose ARITHMETIC decides the bytes are restated operation for operation: the 8x8 inverse DCT (integer
// DCT_ISLOW derivative with 12-bit constants: two extra bits kept after the column pass, +65536 + (128 << 17) before
// the final >> 17), chroma upsampling (h2, v2, the 3:1 / 9:3:3:1 "hv2" filter, nearest for other factors), the
// YCbCr → RGB fixed-point rows (20-bit, the Cb contribution to green masked to its high 16 bits), 16 → 8-bit
// reduction by the high byte, low-bit-depth grey scaling (0xff, 0x55, 0x11), and RGBE → float by an exact power of two.
// Entropy decoding is the JPEG / deflate standards: any conforming decoder yields the same coefficients.  stb's SSE2
// kernels are bit-identical to the scalar forms restated here (stb_image.h: "produces bit-identical results").
// Pinned byte for byte against the reference's own loadTexture / loadTextureHDR by tests/test_image_decoders.py
// through `oracle_ref texload` / `oracle_ref hdrload`.
//
// Not handled (an error, where stb_image decodes): 4-component (CMYK / YCCK) and arithmetic-coded JPEGs, the other
// container formats stb knows (BMP, GIF, PSD, PIC, PNM, TGA).
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace yartb {

namespace {

// ------------------------------------------------------------------------------------------------------
// PNG
// ------------------------------------------------------------------------------------------------------
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

// Undoes the scanline filters of one (sub-)image of `w` x `h` pixels whose rows are `rowBytes` long and whose filter
// unit is `bpp` bytes (PNG spec §9).  `raw` holds h x (1 + rowBytes) bytes.
bool pngUnfilter(const uint8_t* raw, size_t rowBytes, size_t h, size_t bpp, std::vector<uint8_t>& img, std::string& err) {
  img.assign(rowBytes * h, 0);
  for (size_t y = 0; y < h; y++) {
    const uint8_t ft = raw[(rowBytes + 1) * y];
    const uint8_t* in = &raw[(rowBytes + 1) * y + 1];
    uint8_t* out = &img[rowBytes * y];
    const uint8_t* up = y ? &img[rowBytes * (y - 1)] : nullptr;
    for (size_t i = 0; i < rowBytes; i++) {
      const int a = i >= bpp ? out[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: err = "bad PNG filter type"; return false;
      }
      out[i] = uint8_t(in[i] + pred);
    }
  }
  return true;
}

bool decodePng(const uint8_t* data, size_t len, int& w, int& h, std::vector<uint8_t>& rgba, std::string& err) {
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  bool gotHdr = false;
  while (pos + 12 <= len) {
    const uint32_t n = be32(data + pos);
    const uint8_t* tag = data + pos + 4;
    const uint8_t* body = data + pos + 8;
    if (n > len || pos + 12 + n > len) break;
    if (!memcmp(tag, "IHDR", 4) && n >= 13) {
      w = int(be32(body)), h = int(be32(body + 4));
      depth = body[8], ctype = body[9], interlace = body[12];
      gotHdr = true;
    } else if (!memcmp(tag, "PLTE", 4)) {
      plte.assign(body, body + n);
    } else if (!memcmp(tag, "tRNS", 4)) {
      trns.assign(body, body + n);
    } else if (!memcmp(tag, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + n);
    } else if (!memcmp(tag, "IEND", 4)) {
      break;
    }
    pos += 12 + size_t(n);
  }
  if (!gotHdr || w <= 0 || h <= 0) return err = "PNG without IHDR", false;
  if (w > (1 << 24) || h > (1 << 24)) return err = "PNG dimensions too large", false;  // STBI_MAX_DIMENSIONS
  if (interlace > 1) return err = "bad PNG interlace method", false;
  const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!ch) return err = "bad PNG colour type", false;
  const bool depthOk = ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                       : ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                                    : (depth == 8 || depth == 16);
  if (!depthOk) return err = "bad PNG bit depth for its colour type", false;
  if (uint64_t(w) * uint64_t(h) > (1ull << 28)) return err = "PNG too large", false;

  const size_t bitsPerPixel = size_t(ch) * depth, bpp = std::max<size_t>(1, bitsPerPixel / 8);
  auto rowBytesOf = [&](size_t pw) { return (pw * bitsPerPixel + 7) / 8; };
  // sub-images: the whole picture, or the seven Adam7 passes
  struct Pass {
    int x0, y0, dx, dy;
  };
  static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
  std::vector<Pass> passes;
  if (interlace) passes.assign(adam7, adam7 + 7);
  else passes.push_back({0, 0, 1, 1});
  size_t rawSize = 0;
  for (const Pass& p : passes) {
    const size_t pw = (size_t(w) - p.x0 + p.dx - 1) / p.dx, ph = (size_t(h) - p.y0 + p.dy - 1) / p.dy;
    if (int(pw) > 0 && int(ph) > 0 && p.x0 < w && p.y0 < h) rawSize += (rowBytesOf(pw) + 1) * ph;
  }
  std::vector<uint8_t> raw(rawSize);
  uLongf outLen = uLongf(raw.size());
  const int zrc = uncompress(raw.data(), &outLen, idat.data(), uLong(idat.size()));
  if ((zrc != Z_OK && zrc != Z_BUF_ERROR) || outLen < raw.size()) return err = "PNG inflate failed", false;  // stb ignores trailing data

  // samples of every pixel, 8 bits each (16-bit: the high byte; sub-byte grey: scaled; palette: the index), plus a
  // colour-key alpha decided on the ORIGINAL sample values
  const size_t nPix = size_t(w) * h;
  std::vector<uint8_t> samples(nPix * ch), keyAlpha;
  const bool hasKey = !trns.empty() && (ctype == 0 || ctype == 2) && trns.size() >= size_t(ch) * 2;
  uint16_t key[3] = {0, 0, 0};
  if (hasKey) {
    keyAlpha.assign(nPix, 255);
    for (int c = 0; c < ch; c++) key[c] = uint16_t((trns[2 * c] << 8) | trns[2 * c + 1]);
  }
  static const uint8_t depthScale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
  size_t off = 0;
  std::vector<uint8_t> img;
  for (const Pass& p : passes) {
    if (p.x0 >= w || p.y0 >= h) continue;
    const size_t pw = (size_t(w) - p.x0 + p.dx - 1) / p.dx, ph = (size_t(h) - p.y0 + p.dy - 1) / p.dy;
    if (!pw || !ph) continue;
    const size_t rb = rowBytesOf(pw);
    if (!pngUnfilter(&raw[off], rb, ph, bpp, img, err)) return false;
    off += (rb + 1) * ph;
    for (size_t y = 0; y < ph; y++)
      for (size_t x = 0; x < pw; x++) {
        const size_t dst = (size_t(p.y0) + y * p.dy) * w + (size_t(p.x0) + x * p.dx);
        for (int c = 0; c < ch; c++) {
          uint32_t v;  // the sample as stored
          if (depth == 16) v = uint32_t(img[rb * y + (x * ch + c) * 2] << 8) | img[rb * y + (x * ch + c) * 2 + 1];
          else if (depth == 8) v = img[rb * y + x * ch + c];
          else {
            const size_t bit = x * depth;  // ch == 1 for sub-byte depths
            v = (img[rb * y + bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1u);
          }
          samples[dst * ch + c] = depth == 16 ? uint8_t(v >> 8) : depth == 8 ? uint8_t(v) : ctype == 3 ? uint8_t(v) : uint8_t(v * depthScale[depth]);
        }
        if (hasKey) {
          bool all = true;
          for (int c = 0; c < ch; c++) {
            uint32_t v;
            if (depth == 16) v = uint32_t(img[rb * y + (x * ch + c) * 2] << 8) | img[rb * y + (x * ch + c) * 2 + 1];
            else if (depth == 8) v = img[rb * y + x * ch + c];
            else {
              const size_t bit = x * depth;
              v = (img[rb * y + bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1u);
            }
            // stb compares 16-bit images on the full sample, 8-bit and narrower ones on the (scaled) byte
            all = all && (depth == 16 ? v == key[c] : uint8_t(depth == 8 ? v : v * depthScale[depth]) == uint8_t(key[c] * (depth < 8 ? depthScale[depth] : 1)));
          }
          keyAlpha[dst] = all ? 0 : 255;
        }
      }
  }

  rgba.resize(nPix * 4);
  for (size_t i = 0; i < nPix; i++) {
    const uint8_t* px = &samples[i * ch];
    uint8_t r, g, b, a = 255;
    if (ctype == 3) {
      const size_t k = px[0];
      if (k * 3 + 2 >= plte.size()) return err = "PNG palette index out of range", false;
      r = plte[k * 3], g = plte[k * 3 + 1], b = plte[k * 3 + 2];
      if (k < trns.size()) a = trns[k];
    } else if (ctype == 0 || ctype == 4) {
      r = g = b = px[0];
      if (ctype == 4) a = px[1];
    } else {
      r = px[0], g = px[1], b = px[2];
      if (ctype == 6) a = px[3];
    }
    if (hasKey) a = keyAlpha[i];
    rgba[i * 4] = r, rgba[i * 4 + 1] = g, rgba[i * 4 + 2] = b, rgba[i * 4 + 3] = a;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------------
// JPEG
// ------------------------------------------------------------------------------------------------------
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {
  // canonical code: for each length 1..16 the first code, the index of its first symbol, and the count
  int firstCode[17] = {}, firstIdx[17] = {}, count[17] = {};
  uint8_t symbols[256] = {};
  bool present = false;
  bool build(const int sizes[16]) {
    int code = 0, idx = 0;
    for (int l = 1; l <= 16; l++) {
      firstCode[l] = code, firstIdx[l] = idx, count[l] = sizes[l - 1];
      code += sizes[l - 1];
      if (code > (1 << l)) return false;
      idx += sizes[l - 1];
      code <<= 1;
    }
    present = true;
    return idx <= 256;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dcPred = 0;
  int x = 0, y = 0, w2 = 0, h2 = 0;  // pixel size, padded size (whole MCUs)
  std::vector<uint8_t> data;          // w2 x h2 samples
  std::vector<int16_t> coeff;         // progressive: (w2 / 8) x (h2 / 8) blocks of 64
};

struct Jpeg {
  const uint8_t* p;
  const uint8_t* end;
  std::string& err;
  // entropy-coded segment reader
  uint32_t bitBuf = 0;
  int bitCount = 0;
  int marker = -1;  // a marker met while filling the bit buffer
  bool noMore = false;

  Huffman dc[4], ac[4];
  uint16_t dequant[4][64] = {};
  Component comp[4];
  int nComp = 0, width = 0, height = 0, hMax = 1, vMax = 1, mcuX = 0, mcuY = 0;
  bool progressive = false, jfif = false;
  int adobeTransform = -1, rgbIds = 0;
  int scanN = 0, order[4] = {}, specStart = 0, specEnd = 63, succHigh = 0, succLow = 0, eobRun = 0;
  int restartInterval = 0, todo = 0;

  Jpeg(const uint8_t* d, size_t n, std::string& e) : p(d), end(d + n), err(e) {}
  bool fail(const char* m) {
    err = std::string("JPEG: ") + m;
    return false;
  }
  int get8() { return p < end ? *p++ : 0; }
  int get16() {
    const int a = get8();
    return (a << 8) | get8();
  }
  bool eof() const { return p >= end; }

  void fill() {
    do {
      const int b = noMore ? 0 : get8();
      if (b == 0xff) {
        int c = get8();
        while (c == 0xff) c = get8();  // fill bytes
        if (c != 0) {
          marker = c;
          noMore = true;
          return;
        }
      }
      bitBuf |= uint32_t(b) << (24 - bitCount);
      bitCount += 8;
    } while (bitCount <= 24);
  }
  int getBits(int n) {  // unsigned; zeros once the stream ran out (as stb does)
    if (n == 0) return 0;
    if (bitCount < n) fill();
    if (bitCount < n) return 0;
    const int v = int(bitBuf >> (32 - n));
    bitBuf <<= n;
    bitCount -= n;
    return v;
  }
  int getBit() { return getBits(1); }
  int receiveExtend(int n) {  // JPEG RECEIVE + EXTEND
    if (bitCount < n) fill();
    if (bitCount < n) return 0;
    const int v = getBits(n);
    return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
  }
  int decodeHuff(const Huffman& h) {
    if (bitCount < 16) fill();
    int code = 0;
    for (int l = 1; l <= 16; l++) {
      if (bitCount < l) return -1;
      code = int(bitBuf >> (32 - l));
      if (h.count[l] && code - h.firstCode[l] < h.count[l] && code >= h.firstCode[l]) {
        bitBuf <<= l;
        bitCount -= l;
        return h.symbols[h.firstIdx[l] + code - h.firstCode[l]];
      }
    }
    return -1;
  }
  void resetEntropy() {
    bitBuf = 0, bitCount = 0, noMore = false, marker = -1;
    for (Component& c : comp) c.dcPred = 0;
    todo = restartInterval ? restartInterval : 0x7fffffff;
    eobRun = 0;
  }

  // ---- markers -------------------------------------------------------------------------------------
  int nextMarker() {
    if (marker != -1) {
      const int m = marker;
      marker = -1;
      return m;
    }
    int x = get8();
    if (x != 0xff) return -1;
    while (x == 0xff) x = get8();
    return x;
  }
  bool processMarker(int m) {
    if (m == -1) return fail("expected marker");
    if (m == 0xdd) {
      if (get16() != 4) return fail("bad DRI length");
      restartInterval = get16();
      return true;
    }
    if (m == 0xdb) {
      int L = get16() - 2;
      while (L > 0) {
        const int q = get8(), prec = q >> 4, t = q & 15;
        if (prec > 1 || t > 3) return fail("bad DQT");
        for (int i = 0; i < 64; i++) dequant[t][kZigzag[i]] = uint16_t(prec ? get16() : get8());
        L -= prec ? 129 : 65;
      }
      return L == 0 ? true : fail("bad DQT length");
    }
    if (m == 0xc4) {
      int L = get16() - 2;
      while (L > 0) {
        const int q = get8(), tc = q >> 4, th = q & 15;
        if (tc > 1 || th > 3) return fail("bad DHT header");
        int sizes[16], n = 0;
        for (int i = 0; i < 16; i++) n += sizes[i] = get8();
        if (n > 256) return fail("bad DHT header");
        Huffman& h = tc ? ac[th] : dc[th];
        if (!h.build(sizes)) return fail("bad code lengths");
        for (int i = 0; i < n; i++) h.symbols[i] = uint8_t(get8());
        L -= 17 + n;
      }
      return L == 0 ? true : fail("bad DHT length");
    }
    if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {
      int L = get16();
      if (L < 2) return fail("bad APP / COM length");
      L -= 2;
      if (m == 0xe0 && L >= 5) {
        static const char tag[5] = {'J', 'F', 'I', 'F', 0};
        bool ok = true;
        for (int i = 0; i < 5; i++) ok = (get8() == uint8_t(tag[i])) && ok;
        L -= 5;
        if (ok) jfif = true;
      } else if (m == 0xee && L >= 12) {
        static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
        bool ok = true;
        for (int i = 0; i < 6; i++) ok = (get8() == uint8_t(tag[i])) && ok;
        L -= 6;
        if (ok) {
          get8(), get16(), get16();
          adobeTransform = get8();
          L -= 6;
        }
      }
      if (L > end - p) return fail("truncated segment");
      p += L;
      return true;
    }
    return fail("unknown marker");
  }
  bool frameHeader() {
    const int Lf = get16();
    if (Lf < 11) return fail("bad SOF length");
    if (get8() != 8) return fail("only 8-bit samples are supported");
    height = get16(), width = get16();
    if (!height || !width) return fail("zero image size");
    nComp = get8();
    if (nComp == 4) return fail("4-component (CMYK / YCCK) images are not supported");
    if (nComp != 3 && nComp != 1) return fail("bad component count");
    if (Lf != 8 + 3 * nComp) return fail("bad SOF length");
    rgbIds = 0;
    for (int i = 0; i < nComp; i++) {
      Component& c = comp[i];
      c.id = get8();
      if (nComp == 3 && c.id == "RGB"[i]) rgbIds++;
      const int q = get8();
      c.h = q >> 4, c.v = q & 15, c.tq = get8();
      if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) return fail("bad sampling factors");
    }
    hMax = vMax = 1;
    for (int i = 0; i < nComp; i++) hMax = std::max(hMax, comp[i].h), vMax = std::max(vMax, comp[i].v);
    for (int i = 0; i < nComp; i++)
      if (hMax % comp[i].h || vMax % comp[i].v) return fail("bad sampling factors");
    mcuX = (width + hMax * 8 - 1) / (hMax * 8), mcuY = (height + vMax * 8 - 1) / (vMax * 8);
    if (uint64_t(mcuX) * mcuY * hMax * vMax > (1u << 22)) return fail("image too large");
    for (int i = 0; i < nComp; i++) {
      Component& c = comp[i];
      c.x = (width * c.h + hMax - 1) / hMax, c.y = (height * c.v + vMax - 1) / vMax;
      c.w2 = mcuX * c.h * 8, c.h2 = mcuY * c.v * 8;
      c.data.assign(size_t(c.w2) * c.h2, 0);
      if (progressive) c.coeff.assign(size_t(c.w2) * c.h2, 0);
    }
    return true;
  }
  bool scanHeader() {
    const int Ls = get16();
    scanN = get8();
    if (scanN < 1 || scanN > 4 || scanN > nComp) return fail("bad SOS component count");
    if (Ls != 6 + 2 * scanN) return fail("bad SOS length");
    for (int i = 0; i < scanN; i++) {
      const int id = get8(), q = get8();
      int which = 0;
      while (which < nComp && comp[which].id != id) which++;
      if (which == nComp) return fail("SOS names an unknown component");
      comp[which].hd = q >> 4, comp[which].ha = q & 15;
      if (comp[which].hd > 3 || comp[which].ha > 3) return fail("bad Huffman table index");
      order[i] = which;
    }
    specStart = get8(), specEnd = get8();
    const int aa = get8();
    succHigh = aa >> 4, succLow = aa & 15;
    if (progressive) {
      if (specStart > 63 || specEnd > 63 || specStart > specEnd || succHigh > 13 || succLow > 13) return fail("bad SOS");
    } else {
      if (specStart != 0 || succHigh != 0 || succLow != 0) return fail("bad SOS");
      specEnd = 63;
    }
    return true;
  }

  // ---- blocks --------------------------------------------------------------------------------------
  bool decodeBlock(int16_t* d, Component& c) {  // baseline: dequantised coefficients in natural order
    const Huffman &hd = dc[c.hd], &ha = ac[c.ha];
    const uint16_t* dq = dequant[c.tq];
    const int t = decodeHuff(hd);
    if (t < 0 || t > 15) return fail("bad Huffman code");
    memset(d, 0, 64 * sizeof(int16_t));
    const int diff = t ? receiveExtend(t) : 0;
    c.dcPred += diff;
    d[0] = int16_t(c.dcPred * dq[0]);
    int k = 1;
    do {
      const int rs = decodeHuff(ha);
      if (rs < 0) return fail("bad Huffman code");
      const int s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (rs != 0xf0) break;
        k += 16;
      } else {
        k += r;
        if (k > 63) return fail("bad Huffman code");
        const int zig = kZigzag[k++];
        d[zig] = int16_t(receiveExtend(s) * dq[zig]);
      }
    } while (k < 64);
    return true;
  }
  bool decodeBlockProgDc(int16_t* d, Component& c) {
    if (specEnd != 0) return fail("can't merge dc and ac");
    if (succHigh == 0) {
      memset(d, 0, 64 * sizeof(int16_t));
      const int t = decodeHuff(dc[c.hd]);
      if (t < 0 || t > 15) return fail("bad Huffman code");
      const int diff = t ? receiveExtend(t) : 0;
      c.dcPred += diff;
      d[0] = int16_t(c.dcPred * (1 << succLow));
    } else if (getBit()) {
      d[0] = int16_t(d[0] + (1 << succLow));
    }
    return true;
  }
  bool decodeBlockProgAc(int16_t* d, Component& c) {
    if (specStart == 0) return fail("can't merge dc and ac");
    const Huffman& ha = ac[c.ha];
    if (succHigh == 0) {
      if (eobRun) {
        eobRun--;
        return true;
      }
      int k = specStart;
      do {
        const int rs = decodeHuff(ha);
        if (rs < 0) return fail("bad Huffman code");
        const int s = rs & 15, r = rs >> 4;
        if (s == 0) {
          if (r < 15) {
            eobRun = 1 << r;
            if (r) eobRun += getBits(r);
            eobRun--;
            break;
          }
          k += 16;
        } else {
          k += r;
          if (k > 63) return fail("bad Huffman code");
          const int zig = kZigzag[k++];
          d[zig] = int16_t(receiveExtend(s) * (1 << succLow));
        }
      } while (k <= specEnd);
    } else {
      const int16_t bit = int16_t(1 << succLow);
      auto refine = [&](int16_t& v) {
        if (getBit() && (v & bit) == 0) v = int16_t(v > 0 ? v + bit : v - bit);
      };
      if (eobRun) {
        eobRun--;
        for (int k = specStart; k <= specEnd; k++) {
          int16_t& v = d[kZigzag[k]];
          if (v != 0) refine(v);
        }
      } else {
        int k = specStart;
        do {
          const int rs = decodeHuff(ha);
          if (rs < 0) return fail("bad Huffman code");
          int s = rs & 15, r = rs >> 4;
          if (s == 0) {
            if (r < 15) {
              eobRun = (1 << r) - 1;
              if (r) eobRun += getBits(r);
              r = 64;  // forces the end of the block
            }
          } else {
            if (s != 1) return fail("bad Huffman code");
            s = getBit() ? bit : -bit;
          }
          while (k <= specEnd) {
            int16_t& v = d[kZigzag[k++]];
            if (v != 0) {
              refine(v);
            } else {
              if (r == 0) {
                v = int16_t(s);
                break;
              }
              r--;
            }
          }
        } while (k <= specEnd);
      }
    }
    return true;
  }

  // ---- inverse DCT: stb_image's stbi__idct_block (derived from the IJG's jidctint, DCT_ISLOW) --------------
  static int f2f(double x) { return int(x * 4096 + 0.5); }
  struct Idct1D {
    int x0, x1, x2, x3, t0, t1, t2, t3;
  };
  static Idct1D idct1D(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
    Idct1D o;
    int p2 = s2, p3 = s6;
    int p1 = (p2 + p3) * f2f(0.5411961f);
    int t2 = p1 + p3 * f2f(-1.847759065f);
    int t3 = p1 + p2 * f2f(0.765366865f);
    p2 = s0, p3 = s4;
    int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
    o.x0 = t0 + t3, o.x3 = t0 - t3, o.x1 = t1 + t2, o.x2 = t1 - t2;
    t0 = s7, t1 = s5, t2 = s3, t3 = s1;
    p3 = t0 + t2;
    int p4 = t1 + t3;
    p1 = t0 + t3, p2 = t1 + t2;
    const int p5 = (p3 + p4) * f2f(1.175875602f);
    t0 = t0 * f2f(0.298631336f), t1 = t1 * f2f(2.053119869f), t2 = t2 * f2f(3.072711026f), t3 = t3 * f2f(1.501321110f);
    p1 = p5 + p1 * f2f(-0.899976223f), p2 = p5 + p2 * f2f(-2.562915447f);
    p3 = p3 * f2f(-1.961570560f), p4 = p4 * f2f(-0.390180644f);
    o.t3 = t3 + p1 + p4, o.t2 = t2 + p2 + p3, o.t1 = t1 + p2 + p4, o.t0 = t0 + p1 + p3;
    return o;
  }
  static uint8_t clamp8(int x) { return uint8_t(x < 0 ? 0 : x > 255 ? 255 : x); }
  static void idctBlock(uint8_t* out, int stride, const int16_t* d) {
    int val[64];
    for (int i = 0; i < 8; i++) {
      int* v = val + i;
      const int16_t* c = d + i;
      if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
        const int dcterm = c[0] * 4;
        v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
      } else {
        Idct1D r = idct1D(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
        r.x0 += 512, r.x1 += 512, r.x2 += 512, r.x3 += 512;  // 12 fractional bits down to 2
        v[0] = (r.x0 + r.t3) >> 10, v[56] = (r.x0 - r.t3) >> 10;
        v[8] = (r.x1 + r.t2) >> 10, v[48] = (r.x1 - r.t2) >> 10;
        v[16] = (r.x2 + r.t1) >> 10, v[40] = (r.x2 - r.t1) >> 10;
        v[24] = (r.x3 + r.t0) >> 10, v[32] = (r.x3 - r.t0) >> 10;
      }
    }
    for (int i = 0; i < 8; i++) {
      const int* v = val + 8 * i;
      uint8_t* o = out + size_t(stride) * i;
      Idct1D r = idct1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
      const int bias = 65536 + (128 << 17);  // rounding of the 17 bits removed + the level shift
      r.x0 += bias, r.x1 += bias, r.x2 += bias, r.x3 += bias;
      o[0] = clamp8((r.x0 + r.t3) >> 17), o[7] = clamp8((r.x0 - r.t3) >> 17);
      o[1] = clamp8((r.x1 + r.t2) >> 17), o[6] = clamp8((r.x1 - r.t2) >> 17);
      o[2] = clamp8((r.x2 + r.t1) >> 17), o[5] = clamp8((r.x2 - r.t1) >> 17);
      o[3] = clamp8((r.x3 + r.t0) >> 17), o[4] = clamp8((r.x3 - r.t0) >> 17);
    }
  }

  // ---- scans ---------------------------------------------------------------------------------------
  // Returns false on error; true also when the entropy-coded data ends early (stb keeps what it has).
  bool restartCheck(bool& stop) {
    stop = false;
    if (--todo <= 0) {
      if (bitCount < 24) fill();
      if (!(marker >= 0xd0 && marker <= 0xd7)) {
        stop = true;
        return true;
      }
      resetEntropy();
    }
    return true;
  }
  bool entropyCodedData() {
    resetEntropy();
    int16_t block[64];
    bool stop;
    if (scanN == 1) {
      Component& c = comp[order[0]];
      const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
      for (int j = 0; j < bh; j++)
        for (int i = 0; i < bw; i++) {
          if (!progressive) {
            if (!decodeBlock(block, c)) return false;
            idctBlock(&c.data[size_t(c.w2) * j * 8 + i * 8], c.w2, block);
          } else {
            int16_t* d = &c.coeff[64 * (size_t(i) + size_t(j) * (c.w2 / 8))];
            if (!(specStart == 0 ? decodeBlockProgDc(d, c) : decodeBlockProgAc(d, c))) return false;
          }
          if (!restartCheck(stop)) return false;
          if (stop) return true;
        }
      return true;
    }
    for (int j = 0; j < mcuY; j++)
      for (int i = 0; i < mcuX; i++) {
        for (int k = 0; k < scanN; k++) {
          Component& c = comp[order[k]];
          for (int y = 0; y < c.v; y++)
            for (int x = 0; x < c.h; x++) {
              const int bx = i * c.h + x, by = j * c.v + y;
              if (!progressive) {
                if (!decodeBlock(block, c)) return false;
                idctBlock(&c.data[size_t(c.w2) * by * 8 + bx * 8], c.w2, block);
              } else {
                if (!decodeBlockProgDc(&c.coeff[64 * (size_t(bx) + size_t(by) * (c.w2 / 8))], c)) return false;
              }
            }
        }
        if (!restartCheck(stop)) return false;
        if (stop) return true;
      }
    return true;
  }
  void finishProgressive() {
    for (int n = 0; n < nComp; n++) {
      Component& c = comp[n];
      const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
      for (int j = 0; j < bh; j++)
        for (int i = 0; i < bw; i++) {
          int16_t* d = &c.coeff[64 * (size_t(i) + size_t(j) * (c.w2 / 8))];
          const uint16_t* dq = dequant[c.tq];
          for (int k = 0; k < 64; k++) d[k] = int16_t(d[k] * dq[k]);
          idctBlock(&c.data[size_t(c.w2) * j * 8 + i * 8], c.w2, d);
        }
    }
  }
  int skipJunk() {  // stb: look for the next marker after the entropy-coded data
    while (!eof()) {
      int x = get8();
      while (x == 0xff) {
        if (eof()) return -1;
        x = get8();
        if (x != 0x00 && x != 0xff) return x;
      }
    }
    return -1;
  }
  bool decode() {
    if (nextMarker() != 0xd8) return fail("no SOI");
    int m = nextMarker();
    while (!(m == 0xc0 || m == 0xc1 || m == 0xc2)) {
      if (m == 0xc9 || m == 0xca || m == 0xcb) return fail("arithmetic coding is not supported");
      if (!processMarker(m)) return false;
      m = nextMarker();
      while (m == -1) {
        if (eof()) return fail("no SOF");
        m = nextMarker();
      }
    }
    progressive = m == 0xc2;
    if (!frameHeader()) return false;
    m = nextMarker();
    while (m != 0xd9) {
      if (m == 0xda) {
        if (!scanHeader() || !entropyCodedData()) return false;
        if (marker == -1) marker = skipJunk();
        m = nextMarker();
        if (m >= 0xd0 && m <= 0xd7) m = nextMarker();
      } else if (m == 0xdc) {
        const int Ld = get16(), NL = get16();
        if (Ld != 4 || NL != height) return fail("bad DNL");
        m = nextMarker();
      } else {
        // stb returns what it has at the first marker it cannot process (end of data included) — without the
        // progressive finishing pass
        if (!processMarker(m)) {
          err.clear();
          return true;
        }
        m = nextMarker();
      }
    }
    if (progressive) finishProgressive();
    return true;
  }
};

// chroma upsampling rows (stb_image: resample_row_1 / v_2 / h_2 / hv_2 / generic)
const uint8_t* resampleRow(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs) {
  if (hs == 1 && vs == 1) return nearRow;
  if (hs == 1 && vs == 2) {
    for (int i = 0; i < w; i++) out[i] = uint8_t((3 * nearRow[i] + farRow[i] + 2) >> 2);
    return out;
  }
  if (hs == 2 && vs == 1) {
    const uint8_t* in = nearRow;
    if (w == 1) {
      out[0] = out[1] = in[0];
      return out;
    }
    out[0] = in[0];
    out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; i++) {
      const int n = 3 * in[i] + 2;
      out[i * 2] = uint8_t((n + in[i - 1]) >> 2);
      out[i * 2 + 1] = uint8_t((n + in[i + 1]) >> 2);
    }
    out[i * 2] = uint8_t((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
    return out;
  }
  if (hs == 2 && vs == 2) {
    if (w == 1) {
      out[0] = out[1] = uint8_t((3 * nearRow[0] + farRow[0] + 2) >> 2);
      return out;
    }
    int t1 = 3 * nearRow[0] + farRow[0];
    out[0] = uint8_t((t1 + 2) >> 2);
    for (int i = 1; i < w; i++) {
      const int t0 = t1;
      t1 = 3 * nearRow[i] + farRow[i];
      out[i * 2 - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
      out[i * 2] = uint8_t((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = uint8_t((t1 + 2) >> 2);
    return out;
  }
  for (int i = 0; i < w; i++)
    for (int j = 0; j < hs; j++) out[i * hs + j] = nearRow[i];
  return out;
}

bool decodeJpeg(const uint8_t* data, size_t len, int& w, int& h, std::vector<uint8_t>& rgba, std::string& err) {
  Jpeg z(data, len, err);
  if (!z.decode()) return false;
  w = z.width, h = z.height;
  const bool isRgb = z.nComp == 3 && (z.rgbIds == 3 || (z.adobeTransform == 0 && !z.jfif));
  struct Resample {
    int hs, vs, ystep, wLores, ypos;
    const uint8_t *line0, *line1;
    std::vector<uint8_t> buf;
  } rs[3];
  for (int k = 0; k < z.nComp; k++) {
    Resample& r = rs[k];
    r.hs = z.hMax / z.comp[k].h, r.vs = z.vMax / z.comp[k].v;
    r.ystep = r.vs >> 1;
    r.wLores = (w + r.hs - 1) / r.hs;
    r.ypos = 0;
    r.line0 = r.line1 = z.comp[k].data.data();
    r.buf.assign(size_t(w) + 3 + 8, 0);
  }
  rgba.resize(size_t(w) * h * 4);
  const int fr = (int(1.40200f * 4096.0f + 0.5f)) << 8, fg1 = (int(0.71414f * 4096.0f + 0.5f)) << 8,
            fg2 = (int(0.34414f * 4096.0f + 0.5f)) << 8, fb = (int(1.77200f * 4096.0f + 0.5f)) << 8;
  for (int j = 0; j < h; j++) {
    const uint8_t* co[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < z.nComp; k++) {
      Resample& r = rs[k];
      const bool yBot = r.ystep >= (r.vs >> 1);
      co[k] = resampleRow(r.buf.data(), yBot ? r.line1 : r.line0, yBot ? r.line0 : r.line1, r.wLores, r.hs, r.vs);
      if (++r.ystep >= r.vs) {
        r.ystep = 0;
        r.line0 = r.line1;
        if (++r.ypos < z.comp[k].y) r.line1 += z.comp[k].w2;
      }
    }
    uint8_t* out = &rgba[size_t(j) * w * 4];
    for (int i = 0; i < w; i++, out += 4) {
      if (z.nComp == 1) {
        out[0] = out[1] = out[2] = co[0][i];
      } else if (isRgb) {
        out[0] = co[0][i], out[1] = co[1][i], out[2] = co[2][i];
      } else {
        // stbi__YCbCr_to_RGB_row: 20-bit fixed point, the Cb term of green truncated to its high 16 bits
        const int yFixed = (co[0][i] << 20) + (1 << 19), cr = co[2][i] - 128, cb = co[1][i] - 128;
        int r = yFixed + cr * fr;
        int g = yFixed + (cr * -fg1) + int(uint32_t(cb * -fg2) & 0xffff0000u);
        int b = yFixed + cb * fb;
        r >>= 20, g >>= 20, b >>= 20;
        out[0] = Jpeg::clamp8(r), out[1] = Jpeg::clamp8(g), out[2] = Jpeg::clamp8(b);
      }
      out[3] = 255;
    }
  }
  return true;
}

}  // namespace

// stbi_load_from_memory(data, len, &w, &h, nullptr, 4)
bool decodeImageRGBA8(const uint8_t* data, size_t len, int& w, int& h, std::vector<uint8_t>& rgba, std::string& err) {
  static const uint8_t pngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (len >= 8 && memcmp(data, pngSig, 8) == 0) return decodePng(data, len, w, h, rgba, err);
  if (len >= 3 && data[0] == 0xff && data[1] == 0xd8) return decodeJpeg(data, len, w, h, rgba, err);
  err = "unsupported image format (PNG and JPEG are decoded)";
  return false;
}

// stbi_loadf on a Radiance .hdr file, req_comp = 4, then the first three channels (loadTextureHDR, texture.cpp:21-35)
bool decodeRadianceHdr(const uint8_t* data, size_t len, int& w, int& h, std::vector<float>& rgb, std::string& err) {
  const uint8_t *p = data, *end = data + len;
  auto token = [&]() {
    std::string t;
    while (p < end && *p != '\n') {
      if (t.size() < 1022) t.push_back(char(*p));
      p++;
    }
    if (p < end) p++;
    return t;
  };
  const std::string magic = token();
  if (magic != "#?RADIANCE" && magic != "#?RGBE") return err = "not a Radiance HDR file", false;
  bool valid = false;
  for (;;) {
    const std::string t = token();
    if (t.empty()) break;
    if (t == "FORMAT=32-bit_rle_rgbe") valid = true;
    if (p >= end) break;
  }
  if (!valid) return err = "unsupported HDR format", false;
  const std::string dims = token();
  if (dims.compare(0, 3, "-Y ") != 0) return err = "unsupported HDR data layout", false;
  char* q = nullptr;
  h = int(strtol(dims.c_str() + 3, &q, 10));
  while (*q == ' ') q++;
  if (strncmp(q, "+X ", 3) != 0) return err = "unsupported HDR data layout", false;
  w = int(strtol(q + 3, nullptr, 10));
  if (w <= 0 || h <= 0 || w > (1 << 24) || h > (1 << 24) || uint64_t(w) * h > (1ull << 28)) return err = "bad HDR size", false;
  rgb.assign(size_t(w) * h * 3, 0.0f);
  auto convert = [&](size_t pixel, const uint8_t* rgbe) {  // stbi__hdr_convert: mantissa * 2^(e - 136)
    if (rgbe[3] != 0) {
      const float f1 = float(std::ldexp(1.0f, int(rgbe[3]) - (128 + 8)));
      rgb[pixel * 3] = rgbe[0] * f1, rgb[pixel * 3 + 1] = rgbe[1] * f1, rgb[pixel * 3 + 2] = rgbe[2] * f1;
    }
  };
  auto need = [&](size_t n) { return size_t(end - p) >= n; };
  bool flat = w < 8 || w >= 32768;
  size_t flatFrom = 0;
  if (!flat) {
    std::vector<uint8_t> scan(size_t(w) * 4);
    for (int j = 0; j < h && !flat; j++) {
      if (!need(4)) return err = "truncated HDR", false;
      const int c1 = p[0], c2 = p[1], l0 = p[2];
      if (c1 != 2 || c2 != 2 || (l0 & 0x80)) {
        // not run-length encoded: stb restarts as a flat file at pixel 0 with these four bytes
        flat = true;
        flatFrom = 0;
        break;
      }
      const int lenRow = (l0 << 8) | p[3];
      p += 4;
      if (lenRow != w) return err = "invalid decoded scanline length in HDR", false;
      for (int k = 0; k < 4; k++) {
        int i = 0;
        while (i < w) {
          if (!need(1)) return err = "truncated HDR", false;
          int count = *p++;
          if (count > 128) {
            count -= 128;
            if (count == 0 || count > w - i || !need(1)) return err = "bad RLE data in HDR", false;
            const uint8_t v = *p++;
            for (int z = 0; z < count; z++) scan[size_t(i++) * 4 + k] = v;
          } else {
            if (count == 0 || count > w - i || !need(size_t(count))) return err = "bad RLE data in HDR", false;
            for (int z = 0; z < count; z++) scan[size_t(i++) * 4 + k] = *p++;
          }
        }
      }
      for (int i = 0; i < w; i++) convert(size_t(j) * w + i, &scan[size_t(i) * 4]);
    }
  }
  if (flat) {
    for (size_t i = flatFrom; i < size_t(w) * h; i++) {
      if (!need(4)) return err = "truncated HDR", false;
      convert(i, p);
      p += 4;
    }
  }
  return true;
}

}  // namespace yartb
