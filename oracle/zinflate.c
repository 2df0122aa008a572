/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (see zdeflate.c header).
 *
 * CPU restatement of zlib 1.2.8's inflate() as AntiZ drives it: a zlib-wrapped
 * stream (inflateInit -> wbits 15, wrap 1; Z/inflate.c:159,180-239), decoded
 * until Z_STREAM_END, a data error, or the end of the available input
 * (reference call sites: ZlibWrapper.h:58-84 from main.cpp:208-238, and
 * doInflate main.cpp:461-486).  It reproduces
 *   - the accept/reject set (every "msg" of Z/inflate.c:661-1192, Z/inffast.c,
 *     and the code-table validity rules of Z/inftrees.c:100-139),
 *   - total_in exactly, including at an error or when input runs out
 *     (zlib pulls whole bytes only as needed, NEEDBITS/PULLBYTE Z/inflate.c:461-479,
 *     and inflate_fast gives unused bytes back, Z/inffast.c:309-313, so the bytes
 *     consumed are always ceil(bits_consumed / 8)),
 *   - total_in at the moment the first-call output buffer is full
 *     (LIT/MATCH/COPY leave when left == 0, Z/inflate.c:881-885,1137,1168),
 * because ZBuffSearcher's accept logic depends on all three (main.cpp:229-239).
 * The input is a list of segments so that the cross-chunk continuation
 * (refillInput, main.cpp:208) can be replayed on the same decoder state.
 *
 * Parity pin: tests/test_oracle_inflate.py compares it with oracle/_ref/libz128.so
 * on valid streams, every truncation, and bit-flipped streams; plus the
 * known-answer vectors restated from Z/test/infcover.c:399-411,583-659.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OI_END = 0, OI_NEED_INPUT = 1, OI_DATA_ERROR = 2, OI_NEED_DICT = 3, OI_OUT_FULL = 4 };

typedef struct {
    int32_t status;       /* OI_* */
    int32_t err;          /* error class id (see E_* below), 0 if none */
    uint64_t total_in;    /* bytes consumed */
    uint64_t total_out;   /* bytes produced */
    uint64_t in_at_outcap;/* total_in when output first needed room beyond first_out_cap (== total_in if never) */
    uint32_t adler;       /* adler32 of the output produced */
} oi_result;

enum { E_HEADER_CHECK = 1, E_METHOD, E_WINDOW, E_BLOCK_TYPE, E_STORED_LEN, E_TOO_MANY_SYMS, E_CODELEN_SET, E_REPEAT,
       E_NO_EOB, E_LITLEN_SET, E_DIST_SET, E_LITLEN_CODE, E_DIST_CODE, E_DIST_FAR, E_DATA_CHECK };

typedef struct { const uint8_t *p; uint64_t n; } oi_seg;

typedef struct {
    const oi_seg *seg; int nseg, cur; uint64_t off, cidx; /* cursor: byte cidx of the stream is seg[cur].p[off] */
    uint64_t bits;                                  /* total bits consumed */
    uint64_t avail_bits;                            /* total bits available */
    /* output */
    uint8_t *out; uint64_t out_cap, nout; uint8_t *ring; /* ring: 32 KiB window when out == NULL */
    uint64_t first_cap, in_at_cap; int cap_seen;
    uint32_t a, b; /* adler */
} dec_t;

typedef struct { uint16_t count[16]; uint16_t sym[288]; int maxlen, nsyms; } code_t;

static const uint16_t LBASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t  LEXT[29]  = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t DBASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t  DEXT[30]  = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t  CLORD[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

static inline uint64_t bytes_used(const dec_t *d) { return (d->bits + 7) >> 3; }

/* read n (<= 16) bits LSB-first; returns 0 if the input ends first (then everything available counts as consumed,
 * exactly like NEEDBITS/PULLBYTE draining `have` before `goto inf_leave`, Z/inflate.c:461-479) */
static int getbits(dec_t *d, int n, unsigned *v) {
    if (d->bits + (uint64_t)n > d->avail_bits) { d->bits = d->avail_bits; return 0; }
    unsigned r = 0;
    for (int i = 0; i < n; i++) {
        uint64_t bi = d->bits + (uint64_t)i, idx = bi >> 3;
        while (d->cidx < idx) { d->cidx++; d->off++; }
        while (d->off >= d->seg[d->cur].n) { d->off -= d->seg[d->cur].n; d->cur++; }
        r |= ((d->seg[d->cur].p[d->off] >> (bi & 7)) & 1u) << i;
    }
    d->bits += (uint64_t)n; *v = r; return 1;
}

static void emit(dec_t *d, unsigned byte) {
    if (d->out) { if (d->nout < d->out_cap) d->out[d->nout] = (uint8_t)byte; } else d->ring[d->nout & 32767] = (uint8_t)byte;
    d->a += byte; if (d->a >= 65521u) d->a -= 65521u; d->b += d->a; if (d->b >= 65521u) d->b -= 65521u;
    d->nout++;
}
static inline unsigned back(dec_t *d, uint64_t dist) { return d->out ? d->out[d->nout - dist] : d->ring[(d->nout - dist) & 32767]; }
static inline void note_cap(dec_t *d) { if (!d->cap_seen && d->nout >= d->first_cap) { d->cap_seen = 1; d->in_at_cap = bytes_used(d); } }

/* Build a canonical code.  Returns 0 ok, -1 over-subscribed/incomplete in a way zlib rejects (Z/inftrees.c:100-139). */
static int make_code(code_t *c, const uint8_t *lens, int n, int is_codelen_code) {
    memset(c->count, 0, sizeof c->count);
    for (int i = 0; i < n; i++) c->count[lens[i]]++;
    int max = 15; while (max >= 1 && c->count[max] == 0) max--;
    c->maxlen = max; c->nsyms = n;
    if (max == 0) return 0;                      /* no codes: table of invalid entries, not an error here */
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= c->count[l]; if (left < 0) return -1; }
    if (left > 0 && (is_codelen_code || max != 1)) return -1;
    uint16_t offs[16]; offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + c->count[l]);
    for (int i = 0; i < n; i++) if (lens[i]) c->sym[offs[lens[i]]++] = (uint16_t)i;
    return 0;
}
/* Decode one symbol.  1 = ok, 0 = out of input, -1 = invalid code (bits consumed as zlib would: 1 bit, see
 * the {op 64, bits 1} filler entries of Z/inftrees.c:118-125,290-296). */
static int decode(dec_t *d, const code_t *c, int *sym) {
    if (c->maxlen == 0) { unsigned t; if (!getbits(d, 1, &t)) return 0; return -1; }
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= c->maxlen; len++) {
        unsigned bit; if (!getbits(d, 1, &bit)) return 0;
        code |= (int)bit;
        int cnt = c->count[len];
        if (code - cnt < first) { *sym = c->sym[index + (code - first)]; return 1; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -1; /* only reachable for the incomplete single 1-bit code: '1' read, 1 bit consumed */
}

static code_t g_fix_l, g_fix_d; static int g_fix = 0;
static void fixed_codes(void) {
    if (g_fix) return;
    uint8_t l[288]; for (int i = 0; i < 288; i++) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    make_code(&g_fix_l, l, 288, 0);
    uint8_t dl[32]; for (int i = 0; i < 32; i++) dl[i] = 5;
    make_code(&g_fix_d, dl, 32, 0);
    g_fix = 1;
}

#define FAIL(E) do { r->status = OI_DATA_ERROR; r->err = (E); goto done; } while (0)
#define NEED(n, v) do { if (!getbits(&d, (n), &(v))) { r->status = OI_NEED_INPUT; goto done; } } while (0)

/*
 * out == NULL: output is discarded (a 32 KiB ring keeps the back-reference window).
 * out != NULL: output stored up to out_cap; producing more gives OI_OUT_FULL.
 * first_out_cap: size of the scanner's output buffer (main.cpp:228 ZOBuffSz); 0 = unlimited.
 */
int oracle_inflate_segs(const oi_seg *seg, int nseg, uint8_t *out, uint64_t out_cap, uint64_t first_out_cap, oi_result *r) {
    dec_t d; memset(&d, 0, sizeof d);
    uint8_t *ring = NULL;
    d.seg = seg; d.nseg = nseg; d.out = out; d.out_cap = out_cap; d.a = 1; d.b = 0;
    d.first_cap = first_out_cap ? first_out_cap : (uint64_t)-1;
    for (int i = 0; i < nseg; i++) d.avail_bits += seg[i].n * 8;
    if (!out) { ring = (uint8_t *)malloc(32768); d.ring = ring; }
    fixed_codes();
    memset(r, 0, sizeof *r);
    unsigned v, last = 0;
    code_t lc, dc, cc; uint8_t lens[320];

    /* zlib header: Z/inflate.c:636-679 */
    NEED(16, v);
    { unsigned cmf = v & 0xff, flg = v >> 8;
      if (((cmf << 8) + flg) % 31) FAIL(E_HEADER_CHECK);
      if ((cmf & 15) != 8) FAIL(E_METHOD);
      if ((cmf >> 4) + 8 > 15) FAIL(E_WINDOW);
      if (flg & 0x20) { NEED(16, v); NEED(16, v); r->status = OI_NEED_DICT; goto done; } /* DICTID, Z/inflate.c:808-817 */
    }
    while (!last) {
        unsigned type;
        NEED(3, v); last = v & 1; type = v >> 1;           /* Z/inflate.c:829-864 */
        if (type == 3) FAIL(E_BLOCK_TYPE);
        if (type == 0) {                                    /* stored: Z/inflate.c:866-901 */
            d.bits = (d.bits + 7) & ~(uint64_t)7;
            unsigned len, nlen; NEED(16, len); NEED(16, nlen);
            if (len != (nlen ^ 0xffff)) FAIL(E_STORED_LEN);
            while (len) {
                if (d.bits >= d.avail_bits) { r->status = OI_NEED_INPUT; goto done; }
                note_cap(&d);
                if (d.out && d.nout >= d.out_cap) { r->status = OI_OUT_FULL; goto done; }
                NEED(8, v); emit(&d, v); len--;
            }
            continue;
        }
        const code_t *L, *D;
        if (type == 1) { L = &g_fix_l; D = &g_fix_d; }
        else {                                              /* dynamic: Z/inflate.c:903-1016 */
            unsigned nlen, ndist, ncode;
            NEED(14, v); nlen = (v & 31) + 257; ndist = ((v >> 5) & 31) + 1; ncode = (v >> 10) + 4;
            if (nlen > 286 || ndist > 30) FAIL(E_TOO_MANY_SYMS);
            memset(lens, 0, 19);
            for (unsigned i = 0; i < ncode; i++) { NEED(3, v); lens[CLORD[i]] = (uint8_t)v; }
            if (make_code(&cc, lens, 19, 1)) FAIL(E_CODELEN_SET);
            unsigned have = 0;
            while (have < nlen + ndist) {
                int sym, rc;
                if (cc.maxlen == 0) { NEED(1, v); sym = 0; rc = 1; }   /* filler entries decode as val 0 here, Z/inflate.c:944-953 */
                else rc = decode(&d, &cc, &sym);
                if (rc == 0) { r->status = OI_NEED_INPUT; goto done; }
                if (rc < 0) sym = 0;                          /* unreachable for a complete code */
                if (sym < 16) { lens[have++] = (uint8_t)sym; continue; }
                unsigned rep, val = 0;
                if (sym == 16) { NEED(2, v); if (have == 0) FAIL(E_REPEAT); val = lens[have - 1]; rep = 3 + v; }
                else if (sym == 17) { NEED(3, v); rep = 3 + v; }
                else { NEED(7, v); rep = 11 + v; }
                if (have + rep > nlen + ndist) FAIL(E_REPEAT);
                while (rep--) lens[have++] = (uint8_t)val;
            }
            if (lens[256] == 0) FAIL(E_NO_EOB);
            if (make_code(&lc, lens, (int)nlen, 0)) FAIL(E_LITLEN_SET);
            if (make_code(&dc, lens + nlen, (int)ndist, 0)) FAIL(E_DIST_SET);
            L = &lc; D = &dc;
        }
        for (;;) {                                          /* Z/inflate.c:1018-1172, Z/inffast.c:120-307 */
            int sym, rc = decode(&d, L, &sym);
            if (rc == 0) { r->status = OI_NEED_INPUT; goto done; }
            if (rc < 0 || sym > 285) FAIL(E_LITLEN_CODE);
            if (sym < 256) {
                note_cap(&d);
                if (d.out && d.nout >= d.out_cap) { r->status = OI_OUT_FULL; goto done; }
                emit(&d, (unsigned)sym); continue;
            }
            if (sym == 256) break;
            unsigned len = LBASE[sym - 257], dist;
            if (LEXT[sym - 257]) { NEED(LEXT[sym - 257], v); len += v; }
            rc = decode(&d, D, &sym);
            if (rc == 0) { r->status = OI_NEED_INPUT; goto done; }
            if (rc < 0 || sym > 29) FAIL(E_DIST_CODE);
            dist = DBASE[sym];
            if (DEXT[sym]) { NEED(DEXT[sym], v); dist += v; }
            note_cap(&d);                                   /* MATCH leaves on left == 0 before the distance check */
            if (d.out && d.nout >= d.out_cap) { r->status = OI_OUT_FULL; goto done; }
            if (dist > d.nout) FAIL(E_DIST_FAR);
            while (len--) {
                note_cap(&d);
                if (d.out && d.nout >= d.out_cap) { r->status = OI_OUT_FULL; goto done; }
                emit(&d, back(&d, dist));
            }
        }
    }
    d.bits = (d.bits + 7) & ~(uint64_t)7;                   /* CHECK: Z/inflate.c:1174-1195 */
    { unsigned hi, lo; NEED(16, hi); NEED(16, lo);
      uint32_t want = ((hi & 0xff) << 24) | ((hi >> 8) << 16) | ((lo & 0xff) << 8) | (lo >> 8);
      if (want != ((d.b << 16) | d.a)) FAIL(E_DATA_CHECK);
      r->status = OI_END; }
done:
    r->total_in = bytes_used(&d); r->total_out = d.nout; r->adler = (d.b << 16) | d.a;
    r->in_at_outcap = d.cap_seen ? d.in_at_cap : r->total_in;
    free(ring);
    return r->status;
}

int oracle_inflate(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t first_out_cap, oi_result *r) {
    oi_seg s; s.p = in; s.n = n;
    return oracle_inflate_segs(&s, 1, out, out_cap, first_out_cap, r);
}
