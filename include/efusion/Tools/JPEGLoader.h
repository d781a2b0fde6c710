// JPEGLoader — decodes the JPEG colour payload of a .klg frame, for programs written against the reference's
// Tools/JPEGLoader.h:32-91 (same class name and readData signature). The reference calls libjpeg; no JPEG library is in
// scope here, so this is a self-contained baseline decoder whose OUTPUT IS BIT-IDENTICAL to libjpeg / libjpeg-turbo with
// their default decompression settings (what jpeg_read_header + jpeg_start_decompress give the reference): JDCT_ISLOW
// integer IDCT (jidctint.c: 13-bit constants, 2-bit pass-1 scaling), "fancy" triangle-filter chroma upsampling for 2x1 and
// 2x2 subsampled components (jdsample.c h2v1 / h2v2), and the 16-bit fixed-point YCbCr -> RGB tables (jdcolor.c).
// tests/test_jpeg_loader.py compares it byte for byte with libjpeg-turbo decodes (tests/golden/jpeg_*.npz).
//
// Supported: baseline / extended-sequential Huffman JPEG (SOF0 / SOF1), 8-bit, 1 or 3 components in one interleaved scan,
// luma sampling 1x1, 2x1 or 2x2 with 1x1 chroma, restart intervals. That covers what the reference's Logger writes
// (OpenCV cvEncodeImage: baseline 4:2:0). Progressive or arithmetic-coded streams throw std::runtime_error.
//
// As the reference (JPEGLoader.h:73-82), the decoder's R,G,B triplets are stored with channels 0 and 2 swapped: logs hold
// JPEG-compressed BGR images, so the swap yields RGB for ElasticFusion.
#ifndef EFUSION_B200_JPEGLOADER_H_
#define EFUSION_B200_JPEGLOADER_H_

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

class JPEGLoader {
 public:
  JPEGLoader() {}

  // src: numBytes of JPEG data; data: width * height * 3 bytes, rows top to bottom
  void readData(uint8_t* src, const int numBytes, uint8_t* data) {
    int w = 0, h = 0;
    decode(src, (size_t)numBytes, rgb_, w, h);
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i) {
      data[i * 3 + 0] = rgb_[i * 3 + 2];
      data[i * 3 + 1] = rgb_[i * 3 + 1];
      data[i * 3 + 2] = rgb_[i * 3 + 0];
    }
  }

  // decodes to libjpeg's JCS_RGB order; returns the image size
  static void decode(const uint8_t* p, size_t n, std::vector<uint8_t>& out, int& width, int& height) {
    Dec d;
    d.p = p;
    d.end = p + n;
    d.run(out);
    width = d.W;
    height = d.H;
  }

 private:
  std::vector<uint8_t> rgb_;

  struct Huff {
    // canonical Huffman table: codes of length l occupy [mincode[l], maxcode[l]] ; valptr[l] indexes vals
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool set = false;
  };
  struct Comp {
    int id = 0, hs = 1, vs = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;      // blocks per row / column (padded to whole MCUs)
    int dw = 0, dh = 0;      // downsampled_width / height: the real samples
    int pred = 0;
    std::vector<uint8_t> plane;  // bw*8 x bh*8 samples
  };
  struct Dec {
    const uint8_t *p, *end;
    int W = 0, H = 0, ncomp = 0, restart = 0;
    uint16_t qt[4][64];
    bool qt_set[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Comp comp[3];
    uint32_t bitbuf = 0;
    int bitcnt = 0;

    [[noreturn]] static void fail(const char* what) { throw std::runtime_error(std::string("JPEGLoader: ") + what); }
    int u8() {
      if (p >= end) fail("truncated stream");
      return *p++;
    }
    int u16() {
      const int a = u8();
      return (a << 8) | u8();
    }

    void run(std::vector<uint8_t>& out) {
      if (u16() != 0xFFD8) fail("not a JPEG stream (no SOI)");
      bool have_sof = false;
      while (true) {
        int m = u8();
        if (m != 0xFF) fail("marker expected");
        while ((m = u8()) == 0xFF) {
        }
        if (m == 0xD9) fail("no scan before EOI");
        const int len = u16() - 2;
        if (len < 0 || p + len > end) fail("bad segment length");
        const uint8_t* seg_end = p + len;
        if (m == 0xDB) {
          while (p < seg_end) {
            const int pq = u8();
            const int t = pq & 15;
            if (t > 3) fail("bad quantisation table id");
            for (int i = 0; i < 64; ++i) qt[t][zigzag(i)] = (uint16_t)((pq >> 4) ? u16() : u8());
            qt_set[t] = true;
          }
        } else if (m == 0xC4) {
          while (p < seg_end) {
            const int tc = u8();
            const int t = tc & 15;
            if (t > 3) fail("bad Huffman table id");
            Huff& hf = (tc >> 4) ? ac[t] : dc[t];
            int counts[17], total = 0;
            for (int l = 1; l <= 16; ++l) total += (counts[l] = u8());
            if (total > 256) fail("bad Huffman table");
            for (int i = 0; i < total; ++i) hf.vals[i] = (uint8_t)u8();
            int code = 0, k = 0;
            for (int l = 1; l <= 16; ++l) {
              hf.valptr[l] = k;
              hf.mincode[l] = code;
              code += counts[l];
              k += counts[l];
              hf.maxcode[l] = counts[l] ? code - 1 : -1;
              code <<= 1;
            }
            hf.maxcode[17] = 0x7fffffff;
            hf.set = true;
          }
        } else if (m == 0xC0 || m == 0xC1) {
          if (u8() != 8) fail("only 8-bit samples are supported");
          H = u16();
          W = u16();
          ncomp = u8();
          if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3)) fail("unsupported frame header");
          for (int c = 0; c < ncomp; ++c) {
            comp[c].id = u8();
            const int hv = u8();
            comp[c].hs = hv >> 4;
            comp[c].vs = hv & 15;
            comp[c].tq = u8();
            if (comp[c].tq > 3) fail("bad quantisation table selector");
          }
          have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
          fail("progressive / lossless / arithmetic JPEG is not supported (baseline only)");
        } else if (m == 0xDD) {
          restart = u16();
        } else if (m == 0xDA) {
          if (!have_sof) fail("scan before frame header");
          const int ns = u8();
          if (ns != ncomp) fail("only single-scan (interleaved) JPEG is supported");
          for (int i = 0; i < ns; ++i) {
            const int cid = u8();
            const int tt = u8();
            int c = 0;
            while (c < ncomp && comp[c].id != cid) ++c;
            if (c == ncomp) fail("scan references an unknown component");
            comp[c].td = tt >> 4;
            comp[c].ta = tt & 15;
          }
          p += 3;  // Ss, Se, Ah/Al
          scan();
          finish(out);
          return;
        }
        p = seg_end;
      }
    }

    static int zigzag(int i) {
      static const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
      return z[i];
    }

    // ---- entropy-coded segment
    void fill() {
      while (bitcnt <= 24) {
        int b = 0;
        if (p < end) {
          b = *p;
          if (b == 0xFF) {
            if (p + 1 < end && p[1] == 0x00) {
              p += 2;
            } else {
              b = 0;  // a marker: feed zeros, leave the pointer on it
            }
          } else {
            ++p;
          }
        }
        bitbuf |= (uint32_t)b << (24 - bitcnt);
        bitcnt += 8;
      }
    }
    int bits(int n) {
      if (n == 0) return 0;
      if (bitcnt < n) fill();
      const int v = (int)(bitbuf >> (32 - n));
      bitbuf <<= n;
      bitcnt -= n;
      return v;
    }
    int decode_sym(const Huff& hf) {
      if (!hf.set) fail("missing Huffman table");
      int code = 0;
      for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | bits(1);
        if (hf.maxcode[l] >= 0 && code <= hf.maxcode[l] && code >= hf.mincode[l]) return hf.vals[hf.valptr[l] + code - hf.mincode[l]];
      }
      fail("corrupt Huffman code");
    }
    static int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

    void scan() {
      int hmax = 1, vmax = 1;
      for (int c = 0; c < ncomp; ++c) {
        if (comp[c].hs < 1 || comp[c].vs < 1 || comp[c].hs > 2 || comp[c].vs > 2) fail("unsupported sampling factor");
        hmax = comp[c].hs > hmax ? comp[c].hs : hmax;
        vmax = comp[c].vs > vmax ? comp[c].vs : vmax;
        if (!qt_set[comp[c].tq]) fail("missing quantisation table");
      }
      if (ncomp == 1) comp[0].hs = comp[0].vs = hmax = vmax = 1;  // a single-component scan is never interleaved
      if (ncomp == 3 && (comp[1].hs != 1 || comp[1].vs != 1 || comp[2].hs != 1 || comp[2].vs != 1 || (comp[0].hs == 1 && comp[0].vs == 2)))
        fail("unsupported chroma subsampling (4:4:4, 4:2:2 and 4:2:0 are)");
      const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
      for (int c = 0; c < ncomp; ++c) {
        Comp& k = comp[c];
        k.bw = mcux * k.hs;
        k.bh = mcuy * k.vs;
        k.dw = (W * k.hs + hmax - 1) / hmax;
        k.dh = (H * k.vs + vmax - 1) / vmax;
        k.plane.assign((size_t)k.bw * 8 * k.bh * 8, 0);
        k.pred = 0;
      }
      bitbuf = 0;
      bitcnt = 0;
      int todo = restart, rst = 0;
      for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
          if (restart && todo == 0) {
            bitbuf = 0;
            bitcnt = 0;
            if (p + 1 < end && p[0] == 0xFF && p[1] == (0xD0 | rst)) p += 2;
            rst = (rst + 1) & 7;
            for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
            todo = restart;
          }
          for (int c = 0; c < ncomp; ++c) {
            Comp& k = comp[c];
            for (int by = 0; by < k.vs; ++by)
              for (int bx = 0; bx < k.hs; ++bx) {
                int blk[64];
                std::memset(blk, 0, sizeof(blk));
                const int t = decode_sym(dc[k.td]);
                k.pred += t ? extend(bits(t), t) : 0;
                blk[0] = k.pred * qt[k.tq][0];
                for (int i = 1; i < 64;) {
                  const int rs = decode_sym(ac[k.ta]);
                  const int r = rs >> 4, s = rs & 15;
                  if (s == 0) {
                    if (r != 15) break;
                    i += 16;
                    continue;
                  }
                  i += r;
                  if (i > 63) fail("corrupt coefficient run");
                  const int zz = zigzag(i);
                  blk[zz] = extend(bits(s), s) * qt[k.tq][zz];
                  ++i;
                }
                idct_islow(blk, k.plane.data() + ((size_t)(my * k.vs + by) * 8) * (k.bw * 8) + (size_t)(mx * k.hs + bx) * 8, k.bw * 8);
              }
          }
          if (restart) --todo;
        }
    }

    // ---- jidctint.c (JDCT_ISLOW): CONST_BITS 13, PASS1_BITS 2
    static inline int descale(long x, int n) { return (int)((x + (1L << (n - 1))) >> n); }
    static inline uint8_t clamp(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
    static void idct_1d(const long in[8], long out[8], int shift) {
      const long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137,
                 F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
      long z2 = in[2], z3 = in[6];
      long z1 = (z2 + z3) * F0_541;
      long tmp2 = z1 + z3 * (-F1_847);
      long tmp3 = z1 + z2 * F0_765;
      z2 = in[0];
      z3 = in[4];
      long tmp0 = (z2 + z3) << 13;
      long tmp1 = (z2 - z3) << 13;
      const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
      tmp0 = in[7];
      tmp1 = in[5];
      tmp2 = in[3];
      tmp3 = in[1];
      z1 = tmp0 + tmp3;
      z2 = tmp1 + tmp2;
      z3 = tmp0 + tmp2;
      long z4 = tmp1 + tmp3;
      const long z5 = (z3 + z4) * F1_175;
      tmp0 *= F0_298;
      tmp1 *= F2_053;
      tmp2 *= F3_072;
      tmp3 *= F1_501;
      z1 *= -F0_899;
      z2 *= -F2_562;
      z3 *= -F1_961;
      z4 *= -F0_390;
      z3 += z5;
      z4 += z5;
      tmp0 += z1 + z3;
      tmp1 += z2 + z4;
      tmp2 += z2 + z3;
      tmp3 += z1 + z4;
      out[0] = descale(tmp10 + tmp3, shift);
      out[7] = descale(tmp10 - tmp3, shift);
      out[1] = descale(tmp11 + tmp2, shift);
      out[6] = descale(tmp11 - tmp2, shift);
      out[2] = descale(tmp12 + tmp1, shift);
      out[5] = descale(tmp12 - tmp1, shift);
      out[3] = descale(tmp13 + tmp0, shift);
      out[4] = descale(tmp13 - tmp0, shift);
    }
    static void idct_islow(const int* blk, uint8_t* dst, int stride) {
      long ws[64];
      for (int c = 0; c < 8; ++c) {  // pass 1: columns, results scaled up by 2^PASS1_BITS
        long in[8], o[8];
        for (int r = 0; r < 8; ++r) in[r] = blk[r * 8 + c];
        idct_1d(in, o, 13 - 2);
        for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
      }
      for (int r = 0; r < 8; ++r) {  // pass 2: rows, remove the scaling and 1/8, level shift by 128
        long o[8];
        idct_1d(ws + r * 8, o, 13 + 2 + 3);
        for (int c = 0; c < 8; ++c) dst[(size_t)r * stride + c] = clamp((int)o[c] + 128);
      }
    }

    // ---- jdsample.c fancy upsampling + jdcolor.c colour conversion
    void finish(std::vector<uint8_t>& out) {
      out.assign((size_t)W * H * 3, 0);
      if (ncomp == 1) {
        const Comp& y = comp[0];
        for (int r = 0; r < H; ++r)
          for (int c = 0; c < W; ++c) {
            const uint8_t v = y.plane[(size_t)r * y.bw * 8 + c];
            uint8_t* o = &out[((size_t)r * W + c) * 3];
            o[0] = o[1] = o[2] = v;
          }
        return;
      }
      const int hs = comp[0].hs, vs = comp[0].vs;
      std::vector<uint8_t> cb((size_t)W * H + 2 * W + 4), cr((size_t)W * H + 2 * W + 4);
      upsample(comp[1], hs, vs, cb.data());
      upsample(comp[2], hs, vs, cr.data());
      int crr[256], cbb[256];
      long crg[256], cbg[256];
      for (int i = 0; i < 256; ++i) {
        const long x = i - 128;
        crr[i] = (int)((91881L * x + 32768L) >> 16);   // FIX(1.40200)
        cbb[i] = (int)((116130L * x + 32768L) >> 16);  // FIX(1.77200)
        crg[i] = -46802L * x;                          // FIX(0.71414)
        cbg[i] = -22554L * x + 32768L;                 // FIX(0.34414) + ONE_HALF
      }
      const Comp& y = comp[0];
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          const int yy = y.plane[(size_t)r * y.bw * 8 + c];
          const int b = cb[(size_t)r * W + c], q = cr[(size_t)r * W + c];
          uint8_t* o = &out[((size_t)r * W + c) * 3];
          o[0] = clamp(yy + crr[q]);
          o[1] = clamp(yy + (int)((cbg[b] + crg[q]) >> 16));
          o[2] = clamp(yy + cbb[b]);
        }
    }

    // chroma plane (dw x dh real samples) -> W x H. 1x1: copy; 2x1: h2v1_fancy_upsample; 2x2: h2v2_fancy_upsample
    // (the rows above the first / below the last real row are replicas of it, as jdmainct.c's context rows are)
    void upsample(const Comp& k, int hs, int vs, uint8_t* dst) const {
      const int sw = k.bw * 8;
      const int dw = k.dw, dh = k.dh;
      if (hs == 1 && vs == 1) {
        for (int r = 0; r < H; ++r) std::memcpy(dst + (size_t)r * W, &k.plane[(size_t)r * sw], (size_t)W);
        return;
      }
      std::vector<uint8_t> row((size_t)dw * 2 + 4);
      if (hs == 2 && vs == 1) {
        for (int r = 0; r < H; ++r) {
          const uint8_t* in = &k.plane[(size_t)r * sw];
          h2v1(in, dw, row.data());
          std::memcpy(dst + (size_t)r * W, row.data(), (size_t)W);
        }
        return;
      }
      for (int r = 0; r < H; ++r) {  // 2x2
        const int ir = r >> 1;
        int other = (r & 1) ? ir + 1 : ir - 1;  // the nearer neighbouring input row
        other = other < 0 ? 0 : (other >= dh ? dh - 1 : other);
        const uint8_t* in0 = &k.plane[(size_t)ir * sw];
        const uint8_t* in1 = &k.plane[(size_t)other * sw];
        uint8_t* o = row.data();
        if (dw == 1) {
          const int s = in0[0] * 3 + in1[0];
          o[0] = (uint8_t)((s * 4 + 8) >> 4);
          o[1] = (uint8_t)((s * 4 + 7) >> 4);
        } else {
          int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
          o[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
          o[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
          lastcol = thiscol;
          thiscol = nextcol;
          for (int c = 1; c < dw - 1; ++c) {
            nextcol = in0[c + 1] * 3 + in1[c + 1];
            o[2 * c] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
            o[2 * c + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
            lastcol = thiscol;
            thiscol = nextcol;
          }
          o[2 * (dw - 1)] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
          o[2 * (dw - 1) + 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
        }
        std::memcpy(dst + (size_t)r * W, row.data(), (size_t)W);
      }
    }
    static void h2v1(const uint8_t* in, int dw, uint8_t* o) {
      if (dw == 1) {
        o[0] = o[1] = in[0];
        return;
      }
      o[0] = in[0];
      o[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
      for (int c = 1; c < dw - 1; ++c) {
        const int v = in[c] * 3;
        o[2 * c] = (uint8_t)((v + in[c - 1] + 1) >> 2);
        o[2 * c + 1] = (uint8_t)((v + in[c + 1] + 2) >> 2);
      }
      o[2 * (dw - 1)] = (uint8_t)((in[dw - 1] * 3 + in[dw - 2] + 1) >> 2);
      o[2 * (dw - 1) + 1] = in[dw - 1];
    }
  };
};

#endif  // EFUSION_B200_JPEGLOADER_H_
