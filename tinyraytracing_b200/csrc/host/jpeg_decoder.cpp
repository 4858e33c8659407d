// Baseline JPEG decoder for Material::readinMap — the role cv::imread plays in the reference (material.cpp:6).
//
// Written from the JPEG standard (ITU T.81) and reproduces, bit for bit, what the IJG / libjpeg-turbo decoder that
// OpenCV links produces with its default settings, because texel values feed Kd directly (pathTracing.cpp:24-25):
//   * Huffman-coded baseline sequential DCT, 8-bit samples, 1 or 3 components, sampling factors 1 or 2, restart markers
//   * inverse DCT: the "slow-but-accurate" integer transform (13-bit constants, 2 extra bits after the first pass)
//   * chroma up-sampling: the "fancy" triangle filters (3/4 nearer + 1/4 farther sample, alternating rounding)
//   * YCbCr -> RGB with 16-bit fixed-point tables, output in OpenCV's BGR order
// Progressive, arithmetic-coded, 12-bit and CMYK files are reported as unsupported (the cg22 textures are baseline).
#include "tinyrt.h"

#include <cstdio>
#include <cstring>

namespace trt
{
namespace
{
struct Huff
{
    // canonical code tables: for each code length the smallest code, the largest code (-1 if none) and the index of
    // its first symbol
    int mincode[17], maxcode[18], valptr[17];
    unsigned char vals[256];
    bool present = false;
};

struct Comp
{
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int wblocks = 0, hblocks = 0; // padded size in 8x8 blocks
    int dw = 0, dh = 0;           // real down-sampled size in samples
    int pred = 0;
    std::vector<unsigned char> plane; // wblocks*8 x hblocks*8 samples after the inverse DCT
};

struct Decoder
{
    const unsigned char *p, *end;
    std::string err;
    unsigned short qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Comp comp[3];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart = 0;
    // bit reader
    unsigned long long bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;

    bool fail(const std::string &m)
    {
        if (err.empty())
            err = m;
        return false;
    }
    int u8() { return p < end ? *p++ : 0; }
    int u16()
    {
        int a = u8();
        return (a << 8) | u8();
    }

    void fill()
    {
        while (bitcnt <= 56)
        {
            int c = 0;
            if (!hit_marker && p < end)
            {
                c = *p++;
                if (c == 0xFF)
                {
                    int c2 = p < end ? *p : 0xD9;
                    if (c2 == 0)
                        ++p; // stuffed zero
                    else
                    {
                        // a marker: stop consuming, feed zeros (as the IJG decoder does) until it is handled
                        --p;
                        hit_marker = true;
                        c = 0;
                    }
                }
            }
            bitbuf |= (unsigned long long)c << (56 - bitcnt);
            bitcnt += 8;
        }
    }
    int getbits(int n)
    {
        if (n == 0)
            return 0;
        if (bitcnt < n)
            fill();
        const int v = (int)(bitbuf >> (64 - n));
        bitbuf <<= n;
        bitcnt -= n;
        return v;
    }
    int decodeHuff(const Huff &h)
    {
        if (bitcnt < 16)
            fill();
        int code = 0;
        for (int l = 1; l <= 16; ++l)
        {
            code = (code << 1) | (int)((bitbuf >> 63) & 1);
            bitbuf <<= 1;
            --bitcnt;
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l])
                return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        fail("bad Huffman code");
        return 0;
    }
    static int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

    bool readDQT(int len)
    {
        while (len > 0)
        {
            const int pq = u8();
            const int prec = pq >> 4, t = pq & 15;
            if (t > 3)
                return fail("bad DQT table id");
            static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
            for (int i = 0; i < 64; ++i)
                qt[t][zz[i]] = (unsigned short)(prec ? u16() : u8());
            qt_present[t] = true;
            len -= 1 + 64 * (prec ? 2 : 1);
        }
        return true;
    }
    bool readDHT(int len)
    {
        while (len > 0)
        {
            const int tc = u8();
            const int cls = tc >> 4, t = tc & 15;
            if (t > 3 || cls > 1)
                return fail("bad DHT table id");
            int bits[17], total = 0;
            for (int l = 1; l <= 16; ++l)
                total += (bits[l] = u8());
            if (total > 256)
                return fail("bad DHT counts");
            Huff &h = cls ? ac[t] : dc[t];
            for (int i = 0; i < total; ++i)
                h.vals[i] = (unsigned char)u8();
            int code = 0, k = 0;
            for (int l = 1; l <= 16; ++l)
            {
                h.valptr[l] = k;
                h.mincode[l] = code;
                code += bits[l];
                k += bits[l];
                h.maxcode[l] = bits[l] ? code - 1 : -1;
                code <<= 1;
            }
            h.present = true;
            len -= 17 + total;
        }
        return true;
    }
    bool readSOF(int len)
    {
        if (u8() != 8)
            return fail("only 8-bit samples are supported");
        height = u16();
        width = u16();
        ncomp = u8();
        if (width <= 0 || height <= 0)
            return fail("empty image");
        if (ncomp != 1 && ncomp != 3)
            return fail("only grayscale and YCbCr JPEGs are supported");
        for (int i = 0; i < ncomp; ++i)
        {
            comp[i].id = u8();
            const int hv = u8();
            comp[i].h = hv >> 4, comp[i].v = hv & 15;
            comp[i].tq = u8();
            if (comp[i].h < 1 || comp[i].h > 2 || comp[i].v < 1 || comp[i].v > 2 || comp[i].tq > 3)
                return fail("unsupported sampling factors");
            hmax = std::max(hmax, comp[i].h), vmax = std::max(vmax, comp[i].v);
        }
        if (ncomp == 3 && (comp[0].h != hmax || comp[0].v != vmax || comp[1].h != comp[2].h || comp[1].v != comp[2].v))
            return fail("unsupported component layout");
        (void)len;
        return true;
    }

    // ---- inverse DCT: integer "islow" transform, dequantisation folded in ------------------------------------
    static unsigned char rangeLimit(int x) // the decoder's 10-bit wrapped range-limit table, centred on 128
    {
        x &= 1023;
        if (x < 128)
            return (unsigned char)(x + 128);
        if (x < 512)
            return 255;
        if (x < 896)
            return 0;
        return (unsigned char)(x - 896);
    }
    static void idct(const short *coef, const unsigned short *q, unsigned char *out, int stride)
    {
        constexpr int CB = 13, P1 = 2;
        constexpr long F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633,
                       F1_501 = 12299, F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
        auto descale = [](long x, int n) { return (x + (1L << (n - 1))) >> n; };
        int ws[64];
        for (int c = 0; c < 8; ++c)
        {
            const short *in = coef + c;
            const unsigned short *qq = q + c;
            int *w = ws + c;
            if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0)
            {
                const int dcv = (int)((long)in[0] * qq[0]) << P1;
                for (int r = 0; r < 8; ++r)
                    w[8 * r] = dcv;
                continue;
            }
            long z2 = (long)in[16] * qq[16], z3 = (long)in[48] * qq[48];
            long z1 = (z2 + z3) * F0_541;
            long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            z2 = (long)in[0] * qq[0], z3 = (long)in[32] * qq[32];
            long tmp0 = (z2 + z3) << CB, tmp1 = (z2 - z3) << CB;
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = (long)in[56] * qq[56], tmp1 = (long)in[40] * qq[40], tmp2 = (long)in[24] * qq[24], tmp3 = (long)in[8] * qq[8];
            z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298, tmp1 *= F2_053, tmp2 *= F3_072, tmp3 *= F1_501;
            z1 *= -F0_899, z2 *= -F2_562, z3 *= -F1_961, z4 *= -F0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            w[0] = (int)descale(tmp10 + tmp3, CB - P1), w[56] = (int)descale(tmp10 - tmp3, CB - P1);
            w[8] = (int)descale(tmp11 + tmp2, CB - P1), w[48] = (int)descale(tmp11 - tmp2, CB - P1);
            w[16] = (int)descale(tmp12 + tmp1, CB - P1), w[40] = (int)descale(tmp12 - tmp1, CB - P1);
            w[24] = (int)descale(tmp13 + tmp0, CB - P1), w[32] = (int)descale(tmp13 - tmp0, CB - P1);
        }
        for (int r = 0; r < 8; ++r)
        {
            const int *w = ws + 8 * r;
            unsigned char *o = out + (size_t)r * stride;
            if (w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0 && w[6] == 0 && w[7] == 0)
            {
                const unsigned char dcv = rangeLimit((int)descale((long)w[0], P1 + 3));
                for (int c = 0; c < 8; ++c)
                    o[c] = dcv;
                continue;
            }
            long z2 = w[2], z3 = w[6];
            long z1 = (z2 + z3) * F0_541;
            long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            long tmp0 = ((long)w[0] + (long)w[4]) << CB, tmp1 = ((long)w[0] - (long)w[4]) << CB;
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7], tmp1 = w[5], tmp2 = w[3], tmp3 = w[1];
            z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298, tmp1 *= F2_053, tmp2 *= F3_072, tmp3 *= F1_501;
            z1 *= -F0_899, z2 *= -F2_562, z3 *= -F1_961, z4 *= -F0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            constexpr int S = CB + P1 + 3;
            o[0] = rangeLimit((int)descale(tmp10 + tmp3, S)), o[7] = rangeLimit((int)descale(tmp10 - tmp3, S));
            o[1] = rangeLimit((int)descale(tmp11 + tmp2, S)), o[6] = rangeLimit((int)descale(tmp11 - tmp2, S));
            o[2] = rangeLimit((int)descale(tmp12 + tmp1, S)), o[5] = rangeLimit((int)descale(tmp12 - tmp1, S));
            o[3] = rangeLimit((int)descale(tmp13 + tmp0, S)), o[4] = rangeLimit((int)descale(tmp13 - tmp0, S));
        }
    }

    bool decodeBlock(Comp &c, int bx, int by)
    {
        short coef[64];
        std::memset(coef, 0, sizeof coef);
        static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        int s = decodeHuff(dc[c.td]);
        if (s > 11)
            return fail("bad DC category");
        int diff = s ? extend(getbits(s), s) : 0;
        c.pred += diff;
        coef[0] = (short)c.pred;
        for (int k = 1; k < 64;)
        {
            const int rs = decodeHuff(ac[c.ta]);
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0)
            {
                if (r != 15)
                    break; // end of block
                k += 16;
                continue;
            }
            k += r;
            if (k > 63)
                return fail("AC index out of range");
            coef[zz[k]] = (short)extend(getbits(sz), sz);
            ++k;
        }
        if (!err.empty())
            return false;
        idct(coef, qt[c.tq], c.plane.data() + ((size_t)by * 8) * (c.wblocks * 8) + (size_t)bx * 8, c.wblocks * 8);
        return true;
    }

    bool handleRestart(int expected)
    {
        // discard the remaining bits, find the RSTn marker
        bitbuf = 0, bitcnt = 0, hit_marker = false;
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7))
            ++p;
        if (p + 1 >= end)
            return fail("missing restart marker");
        if ((p[1] & 7) != (expected & 7))
            return fail("restart marker out of sequence");
        p += 2;
        for (int i = 0; i < ncomp; ++i)
            comp[i].pred = 0;
        return true;
    }

    bool readScan()
    {
        const int ns = u8();
        if (ns != ncomp)
            return fail("non-interleaved multi-scan files are not supported");
        for (int i = 0; i < ns; ++i)
        {
            const int id = u8(), t = u8();
            int ci = -1;
            for (int k = 0; k < ncomp; ++k)
                if (comp[k].id == id)
                    ci = k;
            if (ci != i)
                return fail("unexpected component order in scan");
            comp[ci].td = t >> 4, comp[ci].ta = t & 15;
            if (comp[ci].td > 3 || comp[ci].ta > 3 || !dc[comp[ci].td].present || !ac[comp[ci].ta].present ||
                !qt_present[comp[ci].tq])
                return fail("scan refers to a missing table");
        }
        const int ss = u8(), se = u8(), ahal = u8();
        if (ss != 0 || se != 63 || ahal != 0)
            return fail("progressive JPEG is not supported");
        const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < ncomp; ++i)
        {
            Comp &c = comp[i];
            if (ncomp == 1)
                c.h = c.v = 1; // a single-component scan is never interleaved: one block per MCU
            c.wblocks = mcux * c.h, c.hblocks = mcuy * c.v;
            c.dw = (width * c.h + hmax - 1) / hmax, c.dh = (height * c.v + vmax - 1) / vmax;
            c.plane.assign((size_t)c.wblocks * 8 * c.hblocks * 8, 0);
            c.pred = 0;
        }
        if (ncomp == 1)
        {
            hmax = vmax = 1;
            comp[0].dw = width, comp[0].dh = height;
            comp[0].wblocks = (width + 7) / 8, comp[0].hblocks = (height + 7) / 8;
            comp[0].plane.assign((size_t)comp[0].wblocks * 8 * comp[0].hblocks * 8, 0);
        }
        const int nx = ncomp == 1 ? comp[0].wblocks : mcux, ny = ncomp == 1 ? comp[0].hblocks : mcuy;
        int count = 0, rst = 0;
        for (int my = 0; my < ny; ++my)
            for (int mx = 0; mx < nx; ++mx)
            {
                if (restart && count == restart)
                {
                    if (!handleRestart(rst++))
                        return false;
                    count = 0;
                }
                for (int i = 0; i < ncomp; ++i)
                    for (int v = 0; v < comp[i].v; ++v)
                        for (int h = 0; h < comp[i].h; ++h)
                            if (!decodeBlock(comp[i], mx * comp[i].h + h, my * comp[i].v + v))
                                return false;
                ++count;
            }
        return true;
    }

    // ---- chroma up-sampling ("fancy" triangle filters) -----------------------------------------------------------
    // in: rows of the down-sampled plane (pointer per row, with the vertical replication at the image edges already
    // applied by the caller); out: full-resolution rows
    static void h2v1Row(const unsigned char *in, int n, unsigned char *out)
    {
        if (n == 1)
        {
            out[0] = out[1] = in[0];
            return;
        }
        out[0] = in[0];
        out[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
        for (int c = 1; c < n - 1; ++c)
        {
            const int v = in[c] * 3;
            out[2 * c] = (unsigned char)((v + in[c - 1] + 1) >> 2);
            out[2 * c + 1] = (unsigned char)((v + in[c + 1] + 2) >> 2);
        }
        out[2 * n - 2] = (unsigned char)((in[n - 1] * 3 + in[n - 2] + 1) >> 2);
        out[2 * n - 1] = in[n - 1];
    }
    static void h2v2Row(const unsigned char *near, const unsigned char *far, int n, unsigned char *out)
    {
        auto colsum = [&](int c) { return near[c] * 3 + far[c]; };
        if (n == 1)
        {
            const int t = colsum(0);
            out[0] = (unsigned char)((t * 4 + 8) >> 4);
            out[1] = (unsigned char)((t * 4 + 7) >> 4);
            return;
        }
        int last, cur = colsum(0), next = colsum(1);
        out[0] = (unsigned char)((cur * 4 + 8) >> 4);
        out[1] = (unsigned char)((cur * 3 + next + 7) >> 4);
        last = cur, cur = next;
        for (int c = 1; c < n - 1; ++c)
        {
            next = colsum(c + 1);
            out[2 * c] = (unsigned char)((cur * 3 + last + 8) >> 4);
            out[2 * c + 1] = (unsigned char)((cur * 3 + next + 7) >> 4);
            last = cur, cur = next;
        }
        out[2 * n - 2] = (unsigned char)((cur * 3 + last + 8) >> 4);
        out[2 * n - 1] = (unsigned char)((cur * 4 + 7) >> 4);
    }

    // full-resolution plane (width x height, row stride = 2 * dw rounded as needed) of component i
    std::vector<unsigned char> upsample(const Comp &c, int &stride) const
    {
        const int ps = c.wblocks * 8;
        if (c.h == hmax && c.v == vmax)
        {
            stride = ps;
            return c.plane;
        }
        const int hx = hmax / c.h, vx = vmax / c.v;
        stride = c.dw * hx;
        std::vector<unsigned char> out((size_t)stride * c.dh * vx);
        auto row = [&](int r) { return c.plane.data() + (size_t)std::min(std::max(r, 0), c.dh - 1) * ps; };
        // the triangle filters need more than two columns; narrower planes are replicated (as the decoder does)
        const bool fancyH = c.dw > 2;
        for (int r = 0; r < c.dh; ++r)
        {
            if (hx == 2 && vx == 1 && fancyH)
                h2v1Row(row(r), c.dw, out.data() + (size_t)r * stride);
            else if (hx == 2 && vx == 2 && fancyH)
            {
                h2v2Row(row(r), row(r - 1), c.dw, out.data() + (size_t)(2 * r) * stride);
                h2v2Row(row(r), row(r + 1), c.dw, out.data() + (size_t)(2 * r + 1) * stride);
            }
            else if (hx == 1 && vx == 2)
            {
                // vertical triangle filter: 3/4 this row + 1/4 the row above (bias 1) / below (bias 2)
                const unsigned char *a = row(r), *up = row(r - 1), *dn = row(r + 1);
                unsigned char *o0 = out.data() + (size_t)(2 * r) * stride, *o1 = o0 + stride;
                for (int x = 0; x < c.dw; ++x)
                {
                    o0[x] = (unsigned char)((a[x] * 3 + up[x] + 1) >> 2);
                    o1[x] = (unsigned char)((a[x] * 3 + dn[x] + 2) >> 2);
                }
            }
            else
            {
                for (int v = 0; v < vx; ++v)
                    for (int x = 0; x < c.dw; ++x)
                        for (int h = 0; h < hx; ++h)
                            out[(size_t)(vx * r + v) * stride + (size_t)hx * x + h] = row(r)[x];
            }
        }
        return out;
    }

    bool decode(Image &img)
    {
        if (u16() != 0xFFD8)
            return fail("not a JPEG file");
        bool sof = false;
        for (;;)
        {
            int c = u8();
            if (p >= end)
                return fail("unexpected end of file");
            if (c != 0xFF)
                continue;
            int m = u8();
            while (m == 0xFF)
                m = u8();
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01 || m == 0)
                continue;
            if (m == 0xD9)
                return fail("no image data");
            const int len = u16() - 2;
            if (len < 0 || p + len > end)
                return fail("bad segment length");
            const unsigned char *next = p + len;
            if (m == 0xDB)
            {
                if (!readDQT(len))
                    return false;
            }
            else if (m == 0xC4)
            {
                if (!readDHT(len))
                    return false;
            }
            else if (m == 0xC0 || m == 0xC1)
            {
                if (!readSOF(len))
                    return false;
                sof = true;
            }
            else if (m == 0xC2)
                return fail("progressive JPEG is not supported");
            else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)
                return fail("unsupported JPEG process");
            else if (m == 0xDD)
                restart = u16();
            else if (m == 0xDA)
            {
                if (!sof)
                    return fail("scan before frame header");
                if (!readScan())
                    return false;
                break;
            }
            p = next;
        }
        // colour conversion (16-bit fixed point, the decoder's tables) into BGR
        img.rows = height, img.cols = width;
        img.data = std::make_shared<std::vector<unsigned char>>((size_t)width * height * 3);
        unsigned char *dst = img.data->data();
        if (ncomp == 1)
        {
            const int ps = comp[0].wblocks * 8;
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x)
                {
                    const unsigned char g = comp[0].plane[(size_t)y * ps + x];
                    unsigned char *o = dst + ((size_t)y * width + x) * 3;
                    o[0] = o[1] = o[2] = g;
                }
            return true;
        }
        int sy, scb, scr;
        const std::vector<unsigned char> Y = upsample(comp[0], sy), Cb = upsample(comp[1], scb), Cr = upsample(comp[2], scr);
        auto fix = [](double x) { return (long)(x * 65536.0 + 0.5); };
        int crr[256], cbb[256];
        long crg[256], cbg[256];
        for (int i = 0; i < 256; ++i)
        {
            const long x = i - 128;
            crr[i] = (int)((fix(1.40200) * x + 32768) >> 16);
            cbb[i] = (int)((fix(1.77200) * x + 32768) >> 16);
            crg[i] = -fix(0.71414) * x;
            cbg[i] = -fix(0.34414) * x + 32768;
        }
        auto clamp = [](int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x)
            {
                const int yy = Y[(size_t)y * sy + x], cb = Cb[(size_t)y * scb + x], cr = Cr[(size_t)y * scr + x];
                unsigned char *o = dst + ((size_t)y * width + x) * 3;
                o[2] = clamp(yy + crr[cr]);
                o[1] = clamp(yy + (int)((cbg[cb] + crg[cr]) >> 16));
                o[0] = clamp(yy + cbb[cb]);
            }
        return true;
    }
};
} // namespace

bool decodeJpeg(const unsigned char *bytes, size_t n, Image &out, std::string &error)
{
    Decoder d;
    d.p = bytes, d.end = bytes + n;
    std::memset(d.qt, 0, sizeof d.qt);
    const bool ok = d.decode(out);
    if (!ok)
    {
        error = d.err.empty() ? "JPEG decode failed" : d.err;
        out = Image();
    }
    return ok;
}

bool decodeJpegFile(const std::string &path, Image &out, std::string &error)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f)
    {
        error = "cannot open " + path;
        return false;
    }
    std::vector<unsigned char> buf;
    unsigned char chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0)
        buf.insert(buf.end(), chunk, chunk + got);
    std::fclose(f);
    return decodeJpeg(buf.data(), buf.size(), out, error);
}
} // namespace trt
