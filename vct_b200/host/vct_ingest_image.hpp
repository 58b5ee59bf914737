// vct_ingest_image.hpp — texture ingest for the host side of libvct_b200: PNG and DDS files -> the packed
// mip chains vct_upload_texture takes.  Header-only C++17, no dependencies (own inflate, own PNG, own S3TC decode).
//
// What it stands in for in the reference:
//   * GLHelper::createTextureFromImage (src/Graphics/GLHelper.cpp:165-211): stbi_load(file, STBI_default) -> R8 / RGB8 /
//     RGBA8 storage with log2(max(w, h)) + 1 levels -> glGenerateTextureMipmap.  decode_png returns exactly the bytes and
//     the channel count stb_image 2.15 (reference ext/include/stb_image.h) returns for a PNG: colour type decides the
//     channels (grey 1, grey+alpha 2, RGB 3, RGBA 4, palette 3 or 4 with tRNS), a tRNS colour key interleaves an alpha
//     channel that stb then fails to report (quirk replicated, see decode_png), 1/2/4-bit grey is scaled by 255 / 85 / 17,
//     16-bit samples keep their high byte, Adam7 files are de-interlaced.
//     Two-channel images are decoded but — like the reference, which has no branch for them (:194-205) — never uploaded.
//     tests/test_ingest.py compares decode_png byte for byte with the reference's stb_image compiled in place
//     (oracle/_ref/stb_dump) on every PNG the reference ships and on synthetic files of every colour type and depth.
//   * ResourceLoader::loadDDS (src/ResourceLoader.h:26-108): DXT1 / DXT3 / DXT5 blocks with the file's own mip levels go
//     to glCompressedTexImage2D and are decoded by the GL implementation.  Here the blocks are decoded on the host to
//     RGB8 (DXT1, GL_COMPRESSED_RGB_S3TC_DXT1_EXT: no alpha) or RGBA8 with the arithmetic of EXT_texture_compression_s3tc
//     as software GL (Mesa) evaluates it: 5:6:5 end points widened by bit replication, thirds as (2a + b) / 3 truncated,
//     halves as (a + b) / 2 truncated, DXT5 alpha ramps as (k a0 + (7-k) a1) / 7 and (k a0 + (5-k) a1) / 5 truncated.
//   * glGenerateTextureMipmap: implementation-defined filter; this repository's canonical choice is the
//     2x2 box, round half up, odd sizes dropping the last row/column — build_mips here, scene.build_mips in the harness.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace vct {

struct Image {
    int width = 0, height = 0, channels = 0, levels = 0;
    std::vector<uint8_t> pixels;            // all levels packed back to back, level 0 first (vct_upload_texture layout)
    std::string error;                      // empty on success
};

namespace image_detail {

// ------------------------------------------------------------------------------------------ inflate (RFC 1951)
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint64_t buf = 0; int count = 0; size_t used = 0, total;
    BitReader(const uint8_t* b, size_t n) : p(b), end(b + n), total(n * 8) {}
    // past the end of the input the reader supplies zero bits; `exhausted()` then reports the overrun, so a truncated
    // stream ends in an error instead of reading out of bounds or looping for ever
    void fill() { while (count <= 56) { if (p < end) buf |= (uint64_t)*p++ << count; count += 8; } }
    uint32_t peek(int n) { if (count < n) fill(); return (uint32_t)(buf & ((1ull << n) - 1)); }
    void drop(int n) { buf >>= n; count -= n; used += (size_t)n; }
    uint32_t bits(int n) { if (n == 0) return 0; const uint32_t v = peek(n); drop(n); return v; }
    void align() { const int r = (int)((8 - (used & 7)) & 7); if (r) { peek(r); drop(r); } }   // stored blocks start at a byte boundary
    bool exhausted() const { return used > total; }
};

struct Huffman {
    std::vector<uint16_t> table; int maxlen = 0;               // entry = symbol << 4 | length, indexed by maxlen reversed bits
    bool build(const uint8_t* lens, int n) {
        int count[16] = {0};
        for (int i = 0; i < n; ++i) count[lens[i]]++;
        count[0] = 0;
        maxlen = 15; while (maxlen > 0 && count[maxlen] == 0) --maxlen;
        table.assign(maxlen ? (size_t)1 << maxlen : 1, 0);
        if (maxlen == 0) return true;                           // no codes: any use is an error (entry length 0)
        int next[16], code = 0, left = 1;
        for (int l = 1; l <= 15; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }   // over-subscribed
        for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next[l] = code; }
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (!l) continue;
            uint32_t c = (uint32_t)next[l]++, r = 0;
            for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1); c >>= 1; }
            for (uint32_t k = r; k < ((uint32_t)1 << maxlen); k += (uint32_t)1 << l) table[k] = (uint16_t)(s << 4 | l);
        }
        return true;
    }
    int decode(BitReader& br) const {
        const uint16_t e = table[br.peek(maxlen)];
        if ((e & 15) == 0) return -1;
        br.drop(e & 15);
        return e >> 4;
    }
};

// `max_out`: the caller knows how many bytes a valid stream produces; anything longer is an error (no decompression bombs)
inline bool inflate_raw(const uint8_t* src, size_t n, std::vector<uint8_t>& out, std::string& err, size_t max_out = ~(size_t)0) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    BitReader br(src, n);
    Huffman lit, dist;
    for (bool last = false; !last;) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {
            br.align();
            const uint32_t len = br.bits(16), nlen = br.bits(16);
            if ((len ^ 0xFFFFu) != nlen) { err = "inflate: stored block length check failed"; return false; }
            if (br.used + (size_t)len * 8 > br.total) { err = "inflate: truncated stored block"; return false; }
            if (out.size() + len > max_out) { err = "inflate: more data than the image needs"; return false; }
            for (uint32_t i = 0; i < len; ++i) out.push_back((uint8_t)br.bits(8));
            continue;
        }
        if (type == 3) { err = "inflate: reserved block type"; return false; }
        uint8_t lens[320];
        if (type == 1) {
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            lit.build(lens, 288);
            for (int i = 0; i < 30; ++i) lens[i] = 5;
            dist.build(lens, 30);
        } else {
            const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
            if (hlit > 286 || hdist > 30) { err = "inflate: bad code counts"; return false; }
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman clh;
            if (!clh.build(cl, 19)) { err = "inflate: bad code-length code"; return false; }
            int i = 0;
            while (i < hlit + hdist) {
                const int s = clh.decode(br);
                if (s < 0) { err = "inflate: bad code-length symbol"; return false; }
                if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                int rep, val = 0;
                if (s == 16) { if (i == 0) { err = "inflate: repeat without a previous length"; return false; } val = lens[i - 1]; rep = 3 + (int)br.bits(2); }
                else if (s == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (i + rep > hlit + hdist) { err = "inflate: code lengths overflow"; return false; }
                while (rep--) lens[i++] = (uint8_t)val;
            }
            if (!lit.build(lens, hlit) || !dist.build(lens + hlit, hdist)) { err = "inflate: over-subscribed code"; return false; }
        }
        for (;;) {
            const int s = lit.decode(br);
            if (s < 0) { err = "inflate: invalid literal/length code"; return false; }
            if (s < 256) {
                if (out.size() >= max_out) { err = "inflate: more data than the image needs"; return false; }
                out.push_back((uint8_t)s);
                if (br.exhausted()) { err = "inflate: truncated stream"; return false; }
                continue;
            }
            if (s == 256) break;
            if (s > 285) { err = "inflate: invalid length symbol"; return false; }
            const uint32_t len = len_base[s - 257] + br.bits(len_extra[s - 257]);
            const int d = dist.decode(br);
            if (d < 0 || d > 29) { err = "inflate: invalid distance code"; return false; }
            const size_t back = dist_base[d] + br.bits(dist_extra[d]);
            if (back > out.size()) { err = "inflate: distance reaches before the start of the output"; return false; }
            if (out.size() + len > max_out) { err = "inflate: more data than the image needs"; return false; }
            const size_t at = out.size();
            out.resize(at + len);
            for (uint32_t i = 0; i < len; ++i) out[at + i] = out[at + i - back];       // overlapping copies replicate
            if (br.exhausted()) { err = "inflate: truncated stream"; return false; }
        }
        if (br.exhausted()) { err = "inflate: truncated stream"; return false; }
    }
    return true;
}

// zlib wrapper (RFC 1950): 2-byte header checked like stb_image does (multiple of 31, method 8, no preset dictionary);
// the Adler-32 trailer is not verified (stb_image does not verify it either).
inline bool inflate_zlib(const uint8_t* src, size_t n, std::vector<uint8_t>& out, std::string& err, size_t max_out = ~(size_t)0) {
    if (n < 2) { err = "zlib: stream too short"; return false; }
    const int cmf = src[0], flg = src[1];
    if ((cmf * 256 + flg) % 31 != 0) { err = "zlib: bad header"; return false; }
    if (flg & 32) { err = "zlib: preset dictionary not allowed"; return false; }
    if ((cmf & 15) != 8) { err = "zlib: bad compression method"; return false; }
    return inflate_raw(src + 2, n - 2, out, err, max_out);
}

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Undo the per-row filters of one (sub-)image of `w` x `h` pixels with `bits` per pixel; `raw` advances past it.
// Result: h rows of ceil(w * bits / 8) bytes, still bit-packed.
inline bool unfilter(const uint8_t*& raw, const uint8_t* raw_end, int w, int h, int bits, std::vector<uint8_t>& rows, std::string& err) {
    const size_t stride = ((size_t)w * bits + 7) / 8, bpp = bits >= 8 ? (size_t)bits / 8 : 1;
    rows.assign(stride * h, 0);
    for (int y = 0; y < h; ++y) {
        if ((size_t)(raw_end - raw) < stride + 1) { err = "png: not enough pixel data"; return false; }
        const int f = *raw++;
        if (f > 4) { err = "png: invalid filter"; return false; }
        uint8_t* cur = &rows[stride * y];
        const uint8_t* up = y ? cur - stride : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = raw[i];
            switch (f) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
            }
            cur[i] = (uint8_t)v;
        }
        raw += stride;
    }
    return true;
}

}  // namespace image_detail

// One-level image (levels = 1) with stb_image's STBI_default channel count; see the header comment.
inline Image decode_png(const uint8_t* data, size_t size) {
    using namespace image_detail;
    Image im;
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) { im.error = "png: bad signature"; return im; }
    uint32_t w = 0, h = 0; int depth = 0, color = 0, interlace = 0;
    uint8_t palette[256][4]; int pal_len = 0; bool has_trns = false; uint16_t key[3] = {0, 0, 0};
    std::vector<uint8_t> idat;
    bool first = true, done = false;
    size_t at = 8;
    while (!done) {
        if (at + 8 > size) { im.error = "png: truncated chunk header"; return im; }
        const uint32_t len = be32(data + at), type = be32(data + at + 4);
        const uint8_t* body = data + at + 8;
        if ((size_t)len > size - at - 8) { im.error = "png: truncated chunk"; return im; }
        if (first && type != 0x49484452u) { im.error = "png: first chunk is not IHDR"; return im; }
        switch (type) {
            case 0x49484452u: {   // IHDR
                if (!first || len != 13) { im.error = "png: bad IHDR"; return im; }
                w = be32(body); h = be32(body + 4); depth = body[8]; color = body[9]; interlace = body[12];
                if (w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) { im.error = "png: bad image size"; return im; }
                if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { im.error = "png: unsupported bit depth"; return im; }
                if (color > 6 || color == 1 || color == 5 || (color == 3 && depth == 16)) { im.error = "png: bad colour type"; return im; }
                if (body[10] || body[11] || interlace > 1) { im.error = "png: bad compression, filter or interlace method"; return im; }
                break;
            }
            case 0x504C5445u: {   // PLTE
                if (len > 256 * 3 || len % 3) { im.error = "png: bad PLTE"; return im; }
                pal_len = (int)len / 3;
                for (int i = 0; i < pal_len; ++i) { palette[i][0] = body[3 * i]; palette[i][1] = body[3 * i + 1]; palette[i][2] = body[3 * i + 2]; palette[i][3] = 255; }
                break;
            }
            case 0x74524E53u: {   // tRNS
                if (!idat.empty()) { im.error = "png: tRNS after IDAT"; return im; }
                if (color == 3) {
                    if (pal_len == 0 || (int)len > pal_len) { im.error = "png: bad tRNS"; return im; }
                    for (uint32_t i = 0; i < len; ++i) palette[i][3] = body[i];
                } else {
                    const int n = (color & 2) ? 3 : 1;
                    if ((color & 4) || len != (uint32_t)n * 2) { im.error = "png: bad tRNS"; return im; }
                    for (int k = 0; k < n; ++k) key[k] = (uint16_t)(body[2 * k] << 8 | body[2 * k + 1]);
                }
                has_trns = true;
                break;
            }
            case 0x49444154u: idat.insert(idat.end(), body, body + len); break;   // IDAT
            case 0x49454E44u: done = true; break;                                // IEND
            default: if (!(type & 0x20000000u)) { im.error = "png: unknown critical chunk"; return im; } break;
        }
        first = false;
        at += 8 + (size_t)len + 4;   // CRC not verified (neither does stb_image)
    }
    if (w == 0 || idat.empty()) { im.error = "png: no image data"; return im; }
    if (color == 3 && pal_len == 0) { im.error = "png: palette image without PLTE"; return im; }
    const int file_ch = color == 3 ? 1 : ((color & 2) ? 3 : 1) + ((color & 4) ? 1 : 0);
    const int out_ch = color == 3 ? (has_trns ? 4 : 3) : file_ch + ((has_trns && !(color & 4)) ? 1 : 0);
    const int bits = file_ch * depth;
    // bytes a valid IDAT stream inflates to: one filter byte + the packed row, per row (per Adam7 pass when interlaced)
    size_t expected = 0;
    if (!interlace) expected = (size_t)h * (((size_t)w * bits + 7) / 8 + 1);
    else {
        static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int p = 0; p < 7; ++p) {
            const size_t pw = ((size_t)w + xs[p] - 1 - xo[p]) / xs[p], ph = ((size_t)h + ys[p] - 1 - yo[p]) / ys[p];
            if ((int)w > xo[p] && (int)h > yo[p]) expected += ph * ((pw * bits + 7) / 8 + 1);
        }
    }
    if (expected > ((size_t)1 << 31) || (size_t)w * h * out_ch > ((size_t)1 << 31)) { im.error = "png: image too large"; return im; }
    std::vector<uint8_t> raw;
    raw.reserve(expected);
    if (!inflate_zlib(idat.data(), idat.size(), raw, im.error, expected)) return im;
    if (raw.size() < expected) { im.error = "png: not enough pixel data"; return im; }

    im.width = (int)w; im.height = (int)h; im.channels = out_ch; im.levels = 1;
    im.pixels.assign((size_t)w * h * out_ch, 0);

    // one pixel of a de-filtered row -> out_ch bytes
    static const int scale[9] = {0, 255, 85, 0, 17, 0, 0, 0, 1};
    auto emit = [&](const uint8_t* row, uint32_t x, uint8_t* o) {
        uint16_t s[4] = {0, 0, 0, 0};
        if (depth == 16) for (int c = 0; c < file_ch; ++c) s[c] = (uint16_t)(row[(x * file_ch + c) * 2] << 8 | row[(x * file_ch + c) * 2 + 1]);
        else if (depth == 8) for (int c = 0; c < file_ch; ++c) s[c] = row[x * file_ch + c];
        else { const uint32_t bit = x * depth; s[0] = (uint16_t)((row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1)); }
        if (color == 3) {
            const uint8_t* p = s[0] < pal_len ? palette[s[0]] : palette[0];   // out-of-range index: stb reads the (zero) table; entry 0 is the defined fallback here
            static const uint8_t zero[4] = {0, 0, 0, 255};
            if (s[0] >= pal_len) p = zero;
            for (int c = 0; c < out_ch; ++c) o[c] = p[c];
            return;
        }
        bool keyed = has_trns && !(color & 4);
        if (keyed) for (int c = 0; c < file_ch; ++c) keyed = keyed && s[c] == key[c];
        for (int c = 0; c < file_ch; ++c) o[c] = depth == 16 ? (uint8_t)(s[c] >> 8) : depth == 8 ? (uint8_t)s[c] : (uint8_t)(s[c] * scale[depth]);
        if (out_ch > file_ch) o[file_ch] = keyed ? 0 : 255;
    };

    const uint8_t* rp = raw.data(); const uint8_t* rend = raw.data() + raw.size();
    std::vector<uint8_t> rows;
    if (!interlace) {
        if (!unfilter(rp, rend, (int)w, (int)h, bits, rows, im.error)) return im;
        const size_t stride = ((size_t)w * bits + 7) / 8;
        for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x) emit(&rows[stride * y], x, &im.pixels[((size_t)y * w + x) * out_ch]);
    } else {
        static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int p = 0; p < 7; ++p) {
            const int pw = ((int)w - xo[p] + xs[p] - 1) / xs[p], ph = ((int)h - yo[p] + ys[p] - 1) / ys[p];
            if (pw <= 0 || ph <= 0) continue;
            if (!unfilter(rp, rend, pw, ph, bits, rows, im.error)) return im;
            const size_t stride = ((size_t)pw * bits + 7) / 8;
            for (int y = 0; y < ph; ++y) for (int x = 0; x < pw; ++x)
                emit(&rows[stride * y], (uint32_t)x, &im.pixels[(((size_t)y * ys[p] + yo[p]) * w + (size_t)x * xs[p] + xo[p]) * out_ch]);
        }
    }
    if (color != 3 && out_ch > file_ch) {
        // stb_image 2.15 quirk, replicated because the reference uploads what stbi_load reports: with a tRNS colour key the
        // decoded buffer is interleaved with the added alpha channel, but the channel count handed back is the FILE's
        // (stbi__do_png reports img_n, not img_out_n).  GLHelper.cpp:194-205 therefore reads the first w*h*channels bytes of
        // the interleaved data as a 1- or 3-channel image.
        im.channels = file_ch;
        im.pixels.resize((size_t)w * h * file_ch);
    }
    return im;
}

// ------------------------------------------------------------------------------------------------ DDS (S3TC)
namespace image_detail {
inline void rgb565(uint16_t c, int out[3]) {
    const int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
    out[0] = (r << 3) | (r >> 2); out[1] = (g << 2) | (g >> 4); out[2] = (b << 3) | (b >> 2);
}
// colour part of a block; `always4`: DXT3/DXT5 blocks never use the 3-colour mode
inline void s3tc_colors(const uint8_t* b, bool always4, uint8_t px[16][4]) {
    const uint16_t c0 = (uint16_t)(b[0] | b[1] << 8), c1 = (uint16_t)(b[2] | b[3] << 8);
    int p[4][3];
    rgb565(c0, p[0]); rgb565(c1, p[1]);
    const bool four = always4 || c0 > c1;
    for (int k = 0; k < 3; ++k) {
        if (four) { p[2][k] = (2 * p[0][k] + p[1][k]) / 3; p[3][k] = (p[0][k] + 2 * p[1][k]) / 3; }
        else { p[2][k] = (p[0][k] + p[1][k]) / 2; p[3][k] = 0; }
    }
    const uint32_t idx = (uint32_t)b[4] | (uint32_t)b[5] << 8 | (uint32_t)b[6] << 16 | (uint32_t)b[7] << 24;
    for (int i = 0; i < 16; ++i) {
        const int s = (idx >> (2 * i)) & 3;
        px[i][0] = (uint8_t)p[s][0]; px[i][1] = (uint8_t)p[s][1]; px[i][2] = (uint8_t)p[s][2];
        px[i][3] = (!four && s == 3) ? 0 : 255;
    }
}
inline void dxt5_alpha(const uint8_t* b, uint8_t px[16][4]) {
    const int a0 = b[0], a1 = b[1];
    int a[8] = {a0, a1, 0, 0, 0, 0, 0, 0};
    if (a0 > a1) for (int k = 1; k < 7; ++k) a[k + 1] = ((7 - k) * a0 + k * a1) / 7;
    else { for (int k = 1; k < 5; ++k) a[k + 1] = ((5 - k) * a0 + k * a1) / 5; a[6] = 0; a[7] = 255; }
    uint64_t idx = 0;
    for (int i = 0; i < 6; ++i) idx |= (uint64_t)b[2 + i] << (8 * i);
    for (int i = 0; i < 16; ++i) px[i][3] = (uint8_t)a[(idx >> (3 * i)) & 7];
}
}  // namespace image_detail

// DXT1 -> 3 channels, DXT3 / DXT5 -> 4 channels; the file's own mip levels are kept (ResourceLoader.h:94-103).
// Level sizes follow the DDS convention max(1, size >> level); the reference's loop halves towards 0 and would hand GL
// a zero-sized level for non-square images (:95-102) — never reached by its assets, which are square.
inline Image decode_dds(const uint8_t* data, size_t size) {
    using namespace image_detail;
    Image im;
    if (size < 128 || std::memcmp(data, "DDS ", 4) != 0) { im.error = "dds: invalid filecode"; return im; }
    auto u32 = [&](size_t off) { return (uint32_t)data[off] | (uint32_t)data[off + 1] << 8 | (uint32_t)data[off + 2] << 16 | (uint32_t)data[off + 3] << 24; };
    const uint32_t height = u32(12), width = u32(16), mips = u32(28);
    const char* fourcc = (const char*)data + 84;
    int kind;
    if (!std::memcmp(fourcc, "DXT1", 4)) kind = 1; else if (!std::memcmp(fourcc, "DXT3", 4)) kind = 3; else if (!std::memcmp(fourcc, "DXT5", 4)) kind = 5;
    else { im.error = "dds: only DXT1, DXT3 and DXT5 are supported"; return im; }
    if (width == 0 || height == 0 || width > 32768 || height > 32768) { im.error = "dds: bad image size"; return im; }
    const size_t block = kind == 1 ? 8 : 16;
    const int ch = kind == 1 ? 3 : 4;
    im.width = (int)width; im.height = (int)height; im.channels = ch;
    size_t at = 128;
    const uint32_t n_levels = mips ? mips : 1;
    for (uint32_t level = 0; level < n_levels && level < 16; ++level) {
        const uint32_t w = width >> level ? width >> level : 1, h = height >> level ? height >> level : 1;
        if (level && (width >> level) == 0 && (height >> level) == 0) break;
        const size_t bw = (w + 3) / 4, bh = (h + 3) / 4;
        if (at + bw * bh * block > size) { if (level == 0) { im.error = "dds: truncated"; im.levels = 0; return im; } break; }
        const size_t base = im.pixels.size();
        im.pixels.resize(base + (size_t)w * h * ch);
        for (size_t by = 0; by < bh; ++by) for (size_t bx = 0; bx < bw; ++bx) {
            const uint8_t* b = data + at + (by * bw + bx) * block;
            uint8_t px[16][4];
            if (kind == 1) s3tc_colors(b, false, px);
            else {
                s3tc_colors(b + 8, true, px);
                if (kind == 3) for (int i = 0; i < 16; ++i) { const int a4 = (b[i >> 1] >> ((i & 1) * 4)) & 15; px[i][3] = (uint8_t)(a4 * 17); }
                else dxt5_alpha(b, px);
            }
            for (int i = 0; i < 16; ++i) {
                const size_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
                if (x >= w || y >= h) continue;
                std::memcpy(&im.pixels[base + (y * w + x) * ch], px[i], ch);
            }
        }
        at += bw * bh * block;
        im.levels++;
    }
    return im;
}

// Append the 2x2 box-filter mip chain (round half up; odd sizes drop the last row / column; one-pixel-wide levels
// average pairs) to a one-level image: the canonical stand-in for glGenerateTextureMipmap (GLHelper.cpp:205).
inline void build_mips(Image& im) {
    if (im.levels != 1 || im.width < 1 || im.height < 1) return;
    int w = im.width, h = im.height; const int ch = im.channels;
    size_t src = 0;
    while ((w > 1 || h > 1) && im.levels < 16) {
        const int nw = w > 1 ? w / 2 : 1, nh = h > 1 ? h / 2 : 1;
        const size_t dst = im.pixels.size();
        im.pixels.resize(dst + (size_t)nw * nh * ch);
        const uint8_t* s = &im.pixels[src]; uint8_t* d = &im.pixels[dst];
        for (int y = 0; y < nh; ++y) for (int x = 0; x < nw; ++x) for (int c = 0; c < ch; ++c) {
            if (w > 1 && h > 1) d[((size_t)y * nw + x) * ch + c] = (uint8_t)((s[((size_t)(2 * y) * w + 2 * x) * ch + c] + s[((size_t)(2 * y + 1) * w + 2 * x) * ch + c] + s[((size_t)(2 * y) * w + 2 * x + 1) * ch + c] + s[((size_t)(2 * y + 1) * w + 2 * x + 1) * ch + c] + 2) >> 2);
            else if (h > 1) d[(size_t)y * ch + c] = (uint8_t)((s[(size_t)(2 * y) * ch + c] + s[(size_t)(2 * y + 1) * ch + c] + 1) >> 1);
            else d[(size_t)x * ch + c] = (uint8_t)((s[(size_t)(2 * x) * ch + c] + s[(size_t)(2 * x + 1) * ch + c] + 1) >> 1);
        }
        src = dst; w = nw; h = nh; im.levels++;
    }
}

inline bool read_binary(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t buf[1 << 16]; size_t got;
    out.clear();
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + got);
    std::fclose(f);
    return true;
}

// GLHelper::createTextureFromImage: by extension "dds" -> loadDDS (file mips), anything else -> image decode + generated
// mips.  Only PNG is decoded here (every non-DDS texture of the reference's scenes is a PNG).
inline Image load_texture_file(const std::string& path, bool with_mips = true) {
    Image im;
    std::vector<uint8_t> bytes;
    if (!read_binary(path, bytes)) { im.error = "cannot open " + path; return im; }
    const size_t dot = path.find_last_of('.');
    std::string ext = dot == std::string::npos ? std::string() : path.substr(dot + 1);
    if (ext == "dds") return decode_dds(bytes.data(), bytes.size());                 // case-sensitive like getExtension (GLHelper.cpp)
    im = decode_png(bytes.data(), bytes.size());
    if (im.error.empty() && with_mips) build_mips(im);
    return im;
}

}  // namespace vct
