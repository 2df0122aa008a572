// TEST PROGRAM (built and run by tests/test_gpu_zlibwrapper.py): a chunked zlib-stream scanner written against the PUBLIC interface of
// ZlibInflator only - the way the reference's scanner drives that class (reference: ZBuffSearcher::operator(), main.cpp:205-246, fed by
// searchInfile, main.cpp:392-420): operator() on every accepted header pair with an output buffer of `chunksize` bytes, continuePrev
// while the output buffer is full, refillInput with the next chunk (first byte = the previous chunk's last) when the input ran out,
// totalInputByte <= 16 means "not a stream".  It is compiled against antiz_b200/host/ZlibWrapper.h (the GPU-backed drop-in) and
// prints one line per stream found: "offset type compressed inflated".  Not part of the product.
#include "ZlibWrapper.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static int header_type(unsigned b0, unsigned b1) {   // the 24 accepted CMF/FLG pairs (main.cpp:168-203), closed form
    if ((b0 & 0x8f) != 0x08 || b0 < 0x28 || (b1 & 0x20) || ((b0 << 8) | b1) % 31) return -1;
    return 4 * ((int)(b0 >> 4) - 2) + (int)(b1 >> 6);
}

struct Scanner {
    ZlibInflator inf;
    std::vector<uint8_t> out;
    bool need_more = false; uint64_t chunk_offset = 0, pending_offset = 0; int type = -1;
    explicit Scanner(uint64_t s) : out(s) {}
    void drain() { while (inf.avail_out() == 0) inf.continuePrev(out.data(), (ZlibInflator::size_type)out.size()); }
    void chunk(uint8_t *buf, uint64_t len) {
        uint64_t i = 0; const uint64_t redlen = len - 1;
        if (need_more) {
            inf.refillInput(buf, (ZlibInflator::size_type)len);
            drain();
            if (inf.lastRetVal() == Z_STREAM_END) {
                printf("%llu %d %lu %lu\n", (unsigned long long)pending_offset, type, inf.totalInputByte(), inf.totalOutputByte());
                i = len - inf.avail_in();
            }
            need_more = inf.avail_in() == 0;
        }
        for (; i < redlen && !need_more; i++) {
            const int t = header_type(buf[i], buf[i + 1]);
            if (t < 0) continue;
            type = t;
            inf(out.data(), (ZlibInflator::size_type)out.size(), buf + i, (ZlibInflator::size_type)(len - i));
            if (inf.totalInputByte() <= 16) continue;
            drain();
            if (inf.lastRetVal() == Z_STREAM_END) {
                printf("%llu %d %lu %lu\n", (unsigned long long)(i + chunk_offset), type, inf.totalInputByte(), inf.totalOutputByte());
                i += inf.totalInputByte() - 1;
            } else if ((need_more = inf.avail_in() == 0)) pending_offset = i + chunk_offset;
        }
        chunk_offset += redlen;
    }
};

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const uint64_t S = strtoull(argv[2], nullptr, 10);
    FILE *f = fopen(argv[1], "rb");
    if (!f || S < 2) return 2;
    std::vector<uint8_t> buf(S);
    Scanner sc(S);
    size_t got = fread(buf.data(), 1, S, f);
    if (!got) return 0;
    uint8_t last = buf[got - 1];
    sc.chunk(buf.data(), got);
    bool eof = got < S;
    while (!eof) {
        buf[0] = last;
        got = fread(buf.data() + 1, 1, S - 1, f);
        eof = got < S - 1;
        last = buf[got];
        sc.chunk(buf.data(), got + 1);
    }
    fclose(f);
    return 0;
}
