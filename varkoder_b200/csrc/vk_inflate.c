/* vk_inflate.c -- host feed of the image hot path: gzip -> bytes, one call per file, straight into the caller's
 * (page-locked) buffer.
 *
 * The reference reads <int>/clean_reads/<sample>.fq.gz through Python's gzip module (varKoder/commands/image.py:662-667)
 * and then L more times through reformat.sh; a single gzip member cannot be split, so the host feed here
 * (varkoder_b200/feed.py) inflates every sample ONCE on a worker thread, N samples on N threads.  With the GPU side at a
 * fraction of a millisecond per sample, the batch is bound by this decoder (SURVEY.md section 8f N1,
 * profiles/bench_feed_r01.txt).  This is a table-driven DEFLATE decoder (RFC 1951 / 1952) written for that job:
 *   - 64-bit bit buffer refilled with one unaligned 8-byte load,
 *   - 11-bit literal/length table and 8-bit distance table with second-level tables for the longer codes; an entry
 *     carries the literal or the base value, the number of extra bits and the code length, so one look-up decodes a symbol,
 *   - up to three literals per refill, matches copied eight bytes at a time,
 *   - output written in place into the destination (no intermediate bytes objects, no window copy),
 *   - CRC-32 of every member checked (carry-less multiplication, slicing-by-8 without PCLMUL) together with ISIZE.
 * A careful byte-wise loop handles the last bytes of input and output, so nothing is read or written out of bounds.
 * Plain C, no dependencies; built by csrc/Makefile into libvk_feed.so and bound by feed.py with ctypes (the call runs
 * without the GIL).  Results are checked byte for byte against zlib in tests/test_feed.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define VKF_OK 0
#define VKF_EFORMAT (-1)   /* not gzip / corrupt stream */
#define VKF_ESPACE (-2)    /* output buffer too small (out_len holds what was produced so far) */
#define VKF_ECRC (-3)      /* CRC-32 or ISIZE mismatch */
#define VKF_ETRUNC (-4)    /* input ends inside a member */

#define LT_BITS 11
#define DT_BITS 8
#define PT_BITS 7
#define LT_CAP 4096
#define DT_CAP 1024
#define PT_CAP 128

#define F_LIT 0x8000u
#define F_SUB 0x4000u
#define F_EOB 0x2000u

typedef struct {
    uint32_t lt[LT_CAP];
    uint32_t dt[DT_CAP];
    uint32_t pt[PT_CAP];
    uint8_t lens[288 + 32 + 138];
} tables_t;

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t kPrecodeOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void copy64(uint8_t* d, const uint8_t* s) { uint64_t v; memcpy(&v, s, 8); memcpy(d, &v, 8); }

static inline unsigned bitrev(unsigned c, unsigned n)          /* reverse the low n <= 16 bits */
{
    c = ((c & 0x5555u) << 1) | ((c >> 1) & 0x5555u);
    c = ((c & 0x3333u) << 2) | ((c >> 2) & 0x3333u);
    c = ((c & 0x0F0Fu) << 4) | ((c >> 4) & 0x0F0Fu);
    c = ((c & 0x00FFu) << 8) | ((c >> 8) & 0x00FFu);
    return c >> (16 - n);
}

/* entry = value << 16 | flags | extra_bits << 8 | bits_to_consume.  kind: 0 literal/length, 1 distance, 2 precode */
static inline uint32_t symbol_entry(int kind, unsigned s)
{
    if (kind == 2) return (uint32_t)s << 16;
    if (kind == 1) return s < 30 ? ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 8) : 0u;
    if (s < 256) return ((uint32_t)s << 16) | F_LIT;
    if (s == 256) return F_EOB;
    if (s < 286) return ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)kLenExtra[s - 257] << 8);
    return 0u;
}

/* canonical Huffman code -> look-up table indexed by the next table_bits input bits (codes arrive LSB first, so by the
 * bit-reversed code).  An incomplete code leaves invalid (zero) entries; an over-subscribed one is an error. */
static int build_table(uint32_t* table, unsigned table_bits, size_t cap, const uint8_t* lens, unsigned nsym, int kind)
{
    unsigned count[16] = {0}, next_code[16], code = 0;
    uint8_t sub_bits[1u << LT_BITS];
    const unsigned main_size = 1u << table_bits, mask = main_size - 1;
    for (unsigned s = 0; s < nsym; ++s) count[lens[s]]++;
    count[0] = 0;
    int left = 1;
    for (unsigned l = 1; l <= 15; ++l) {
        left = (left << 1) - (int)count[l];
        if (left < 0) return -1;
    }
    for (unsigned l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = code; }
    memset(table, 0, main_size * sizeof(uint32_t));
    /* first pass over the long codes: bits of the second-level table behind every main-table prefix */
    int any_long = 0;
    for (unsigned l = table_bits + 1; l <= 15; ++l) any_long |= count[l] != 0;
    size_t next = main_size;
    if (any_long) {
        unsigned nc[16];
        memcpy(nc, next_code, sizeof(nc));
        memset(sub_bits, 0, main_size);
        for (unsigned s = 0; s < nsym; ++s) {
            const unsigned l = lens[s];
            if (l == 0) continue;
            const unsigned c = nc[l]++;
            if (l <= table_bits) continue;
            const unsigned prefix = bitrev(c, l) & mask;
            if (l - table_bits > sub_bits[prefix]) sub_bits[prefix] = (uint8_t)(l - table_bits);
        }
        for (unsigned p = 0; p < main_size; ++p) {
            if (!sub_bits[p]) continue;
            const size_t n = (size_t)1 << sub_bits[p];
            if (next + n > cap) return -1;
            table[p] = ((uint32_t)next << 16) | F_SUB | ((uint32_t)sub_bits[p] << 8) | table_bits;
            memset(table + next, 0, n * sizeof(uint32_t));
            next += n;
        }
    }
    for (unsigned s = 0; s < nsym; ++s) {
        const unsigned l = lens[s];
        if (l == 0) continue;
        const unsigned rev = bitrev(next_code[l]++, l);
        const uint32_t e = symbol_entry(kind, s);
        if (l <= table_bits) {
            for (unsigned i = rev; i < main_size; i += 1u << l) table[i] = e | l;
        } else {
            const uint32_t ptr = table[rev & mask];
            const unsigned sb = (ptr >> 8) & 31u, sl = l - table_bits;
            uint32_t* sub = table + (ptr >> 16);
            for (unsigned i = rev >> table_bits; i < (1u << sb); i += 1u << sl) sub[i] = e | sl;
        }
    }
    return 0;
}

static uint32_t crc_tab[8][256];
static int crc_ready = 0;
static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
    __atomic_store_n(&crc_ready, 1, __ATOMIC_RELEASE);
}
#if defined(__x86_64__)
#include <immintrin.h>
/* CRC-32 by carry-less multiplication: fold 64 bytes per step, then 128 -> 64 -> 32 bits with a Barrett reduction
 * (the published folding constants for the reflected polynomial 0xEDB88320).  n >= 64, n % 16 == 0; crc is the
 * running, pre-inverted register.  Checked against zlib.crc32 in tests/test_feed.py. */
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(uint32_t crc, const uint8_t* p, size_t n)
{
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124ll);
    const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
    const __m128i lo32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
    __m128i t;
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    p += 64;
    n -= 64;
    while (n >= 64) {
#define FOLD(x, off) t = _mm_clmulepi64_si128(x, k1k2, 0x00); x = _mm_clmulepi64_si128(x, k1k2, 0x11); \
                     x = _mm_xor_si128(_mm_xor_si128(x, t), _mm_loadu_si128((const __m128i*)(p + off)))
        FOLD(x1, 0); FOLD(x2, 16); FOLD(x3, 32); FOLD(x4, 48);
#undef FOLD
        p += 64;
        n -= 64;
    }
#define MERGE(y) t = _mm_clmulepi64_si128(x1, k3k4, 0x00); x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11); \
                 x1 = _mm_xor_si128(_mm_xor_si128(x1, t), y)
    MERGE(x2); MERGE(x3); MERGE(x4);
    while (n >= 16) {
        const __m128i y = _mm_loadu_si128((const __m128i*)p);
        MERGE(y);
        p += 16;
        n -= 16;
    }
#undef MERGE
    /* 128 -> 64 bits */
    t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_clmulepi64_si128(_mm_and_si128(x1, lo32), k5, 0x00);
    x1 = _mm_xor_si128(x1, t);
    /* Barrett reduction to 32 bits */
    t = _mm_clmulepi64_si128(_mm_and_si128(x1, lo32), poly, 0x10);
    t = _mm_clmulepi64_si128(_mm_and_si128(t, lo32), poly, 0x00);
    x1 = _mm_xor_si128(x1, t);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
static int have_clmul(void)
{
    static int cached = -1;
    if (cached < 0) cached = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return cached;
}
#endif

static uint32_t crc32_bytes(uint32_t crc, const uint8_t* p, size_t n)
{
    if (!__atomic_load_n(&crc_ready, __ATOMIC_ACQUIRE)) crc_init();      /* idempotent: a race writes the same values */
    crc = ~crc;
#if defined(__x86_64__)
    if (n >= 64 && have_clmul()) {
        const size_t m = n & ~(size_t)15;
        crc = crc32_clmul(crc, p, m);
        p += m;
        n -= m;
    }
#endif
    while (n && ((uintptr_t)p & 7)) { crc = (crc >> 8) ^ crc_tab[0][(crc ^ *p++) & 0xFF]; --n; }
    while (n >= 8) {
        const uint64_t v = load64(p) ^ crc;
        crc = crc_tab[7][v & 0xFF] ^ crc_tab[6][(v >> 8) & 0xFF] ^ crc_tab[5][(v >> 16) & 0xFF] ^ crc_tab[4][(v >> 24) & 0xFF] ^
              crc_tab[3][(v >> 32) & 0xFF] ^ crc_tab[2][(v >> 40) & 0xFF] ^ crc_tab[1][(v >> 48) & 0xFF] ^ crc_tab[0][v >> 56];
        p += 8;
        n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ crc_tab[0][(crc ^ *p++) & 0xFF];
    return ~crc;
}

/* ---- bit reader ------------------------------------------------------------------------------------------------
 * bitbuf holds bitcnt accounted bits (LSB first); `in` is the next byte that has not been accounted.  The fast refill
 * may leave up to 7 unaccounted garbage bits above bitcnt; they are ORed again with the same values by the next refill. */
#define REFILL_FAST() do { bitbuf |= load64(in) << bitcnt; in += (63u - bitcnt) >> 3; bitcnt |= 56u; } while (0)
#define REFILL_SAFE() do { bitbuf &= bitcnt < 64 ? (((uint64_t)1 << bitcnt) - 1) : ~(uint64_t)0; \
                           while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; } } while (0)
#define BITS(n) ((uint32_t)bitbuf & ((1u << (n)) - 1u))
#define DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)

#define FN_NAME inflate_raw
#define OUT_T uint8_t
#define SYMBOLIC 0
#include "vk_inflate_body.inc"
#undef FN_NAME
#undef OUT_T
#undef SYMBOLIC
#define FN_NAME inflate_raw16
#define OUT_T uint16_t
#define SYMBOLIC 1
#include "vk_inflate_body.inc"
#undef FN_NAME
#undef OUT_T
#undef SYMBOLIC

/* gzip file (one or more members, RFC 1952) -> out.  *out_len = bytes produced; returns VKF_OK or a negative VKF_E*.
 * With VKF_ESPACE, *out_len is the number of bytes produced before the buffer ran out (grow it and call again). */
int vkf_gunzip(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, int verify_crc)
{
    tables_t T;
    size_t ip = 0, op = 0;
    int members = 0;
    *out_len = 0;
    while (ip < in_len) {
        if (members && in[ip] == 0) { ++ip; continue; }          /* zero padding after the last member (tar, bgzip tools) */
        if (in_len - ip < 18) return members ? VKF_OK : VKF_EFORMAT;
        const uint8_t* h = in + ip;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || (h[3] & 0xE0)) return VKF_EFORMAT;
        const unsigned flg = h[3];
        size_t p = ip + 10;
        if (flg & 4) {
            if (in_len - p < 2) return VKF_ETRUNC;
            const size_t xl = in[p] | (in[p + 1] << 8);
            p += 2;
            if (in_len - p < xl) return VKF_ETRUNC;
            p += xl;
        }
        for (int f = 8; f <= 16; f <<= 1) {                       /* FNAME, FCOMMENT: zero-terminated */
            if (!(flg & f)) continue;
            while (p < in_len && in[p]) ++p;
            if (p >= in_len) return VKF_ETRUNC;
            ++p;
        }
        if (flg & 2) p += 2;
        if (p >= in_len) return VKF_ETRUNC;
        size_t used = 0, made = 0;
        int fin = 0;
        const int rc = inflate_raw(&T, in + p, in + in_len, NULL, out + op, out + out_cap, &used, &made, &fin);
        if (rc != VKF_OK) { *out_len = op + made; return rc; }
        p += used;
        if (in_len - p < 8) { *out_len = op + made; return VKF_ETRUNC; }
        const uint32_t crc = in[p] | (in[p + 1] << 8) | (in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
        const uint32_t isize = in[p + 4] | (in[p + 5] << 8) | (in[p + 6] << 16) | ((uint32_t)in[p + 7] << 24);
        if (isize != (uint32_t)made) return VKF_ECRC;
        if (verify_crc && crc32_bytes(0, out + op, made) != crc) return VKF_ECRC;
        op += made;
        *out_len = op;
        ip = p + 8;
        ++members;
    }
    return members ? VKF_OK : VKF_EFORMAT;
}

uint32_t vkf_crc32(uint32_t crc, const uint8_t* p, size_t n) { return crc32_bytes(crc, p, n); }

/* ---- one gzip member on several threads ---------------------------------------------------------------------------
 * pigz (what the reference's clean_reads step writes, image.py:534-540) ends every 128 KiB chunk with an empty stored
 * block, so the compressed stream has byte-aligned block starts every few tens of KiB, recognisable by the bytes
 * 00 00 FF FF in front of them.  feed.py cuts the stream at such points, decodes the pieces concurrently -- the first
 * one as bytes, the others as 16-bit symbols with placeholders for the 32 KiB they cannot see (pigz primes every chunk
 * with the previous one) -- and resolves the placeholders piece by piece once the bytes in front are final.  A cut at
 * a 00 00 FF FF that was not a block boundary makes its piece fail or end in the wrong place, and the CRC-32 of the
 * whole member is checked at the end: any doubt sends the file to the serial decoder. */

/* length of the gzip header at in[0..n), or 0 */
size_t vkf_gzip_header_len(const uint8_t* in, size_t n)
{
    if (n < 18 || in[0] != 0x1f || in[1] != 0x8b || in[2] != 8 || (in[3] & 0xE0)) return 0;
    const unsigned flg = in[3];
    size_t p = 10;
    if (flg & 4) {
        if (n - p < 2) return 0;
        const size_t xl = in[p] | (in[p + 1] << 8);
        p += 2;
        if (n - p < xl) return 0;
        p += xl;
    }
    for (int f = 8; f <= 16; f <<= 1) {
        if (!(flg & f)) continue;
        while (p < n && in[p]) ++p;
        if (p >= n) return 0;
        ++p;
    }
    if (flg & 2) p += 2;
    return p < n ? p : 0;
}

/* offset just behind the first 00 00 FF FF at or after `from` (a candidate block start), or n when there is none */
size_t vkf_next_sync(const uint8_t* in, size_t n, size_t from)
{
    if (n < 4) return n;
    for (size_t i = from; i + 4 <= n; ++i) {
        if (in[i + 3] != 0xFF) continue;                 /* rare byte first */
        if (in[i + 2] == 0xFF && in[i + 1] == 0 && in[i] == 0) return i + 4;
    }
    return n;
}

/* One piece of a raw DEFLATE stream.  in[0..in_len): from a block start to the cut behind an empty stored block
 * (last = 0) or to the end of the file (last = 1: runs up to the final block).  symbolic = 0: out is uint8_t[out_cap];
 * symbolic = 1: out is uint16_t[out_cap].  Returns VKF_OK only if the piece ended where it had to. */
int vkf_inflate_piece(const uint8_t* in, size_t in_len, int last, int symbolic, void* out, size_t out_cap, size_t* out_len,
                      size_t* in_used)
{
    tables_t T;
    int fin = 0, rc;
    const uint8_t* stop = last ? NULL : in + in_len;
    if (symbolic) rc = inflate_raw16(&T, in, in + in_len, stop, (uint16_t*)out, (uint16_t*)out + out_cap, in_used, out_len, &fin);
    else rc = inflate_raw(&T, in, in + in_len, stop, (uint8_t*)out, (uint8_t*)out + out_cap, in_used, out_len, &fin);
    if (rc != VKF_OK) return rc;
    if (last ? !fin : (fin || *in_used != in_len)) return VKF_EFORMAT;
    return VKF_OK;
}

/* symbols -> bytes.  window = the 32768 bytes in front of the piece (window[32767] is the byte just before it).
 * Returns the number of placeholders that pointed in front of `valid_from` (window offsets below it hold no data:
 * the stream is younger than 32 KiB there) -- must be 0. */
size_t vkf_resolve16(const uint16_t* sym, size_t n, const uint8_t* window, size_t valid_from, uint8_t* out)
{
    size_t bad = 0, i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t a, b;
        memcpy(&a, sym + i, 8);
        memcpy(&b, sym + i + 4, 8);
        if (((a | b) & 0xFF00FF00FF00FF00ull) == 0) {           /* eight plain bytes */
            out[i] = (uint8_t)a; out[i + 1] = (uint8_t)(a >> 16); out[i + 2] = (uint8_t)(a >> 32); out[i + 3] = (uint8_t)(a >> 48);
            out[i + 4] = (uint8_t)b; out[i + 5] = (uint8_t)(b >> 16); out[i + 6] = (uint8_t)(b >> 32); out[i + 7] = (uint8_t)(b >> 48);
            continue;
        }
        for (size_t j = i; j < i + 8; ++j) {
            const unsigned v = sym[j];
            if (v & 0x8000u) { bad += (v & 0x7FFFu) < valid_from; out[j] = window[v & 0x7FFFu]; }
            else out[j] = (uint8_t)v;
        }
    }
    for (; i < n; ++i) {
        const unsigned v = sym[i];
        if (v & 0x8000u) { bad += (v & 0x7FFFu) < valid_from; out[i] = window[v & 0x7FFFu]; }
        else out[i] = (uint8_t)v;
    }
    return bad;
}

/* CRC-32 of A || B from crc(A), crc(B) and len(B): multiply crc(A) by x^(8 len) modulo the polynomial (GF(2) matrix
 * squaring), then add crc(B). */
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec)
{
    uint32_t sum = 0;
    for (; vec; vec >>= 1, ++mat)
        if (vec & 1u) sum ^= *mat;
    return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat)
{
    for (int i = 0; i < 32; ++i) sq[i] = gf2_times(mat, mat[i]);
}
uint32_t vkf_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    uint32_t even[32], odd[32];
    if (len2 == 0) return crc1;
    odd[0] = 0xEDB88320u;                       /* operator for one zero bit */
    for (int i = 1; i < 32; ++i) odd[i] = 1u << (i - 1);
    gf2_square(even, odd);                      /* two bits */
    gf2_square(odd, even);                      /* four bits */
    do {                                        /* first square gives the operator for one zero byte */
        gf2_square(even, odd);
        if (len2 & 1) crc1 = gf2_times(even, crc1);
        len2 >>= 1;
        if (!len2) break;
        gf2_square(odd, even);
        if (len2 & 1) crc1 = gf2_times(odd, crc1);
        len2 >>= 1;
    } while (len2);
    return crc1 ^ crc2;
}
int vkf_abi_version(void) { return 1; }
