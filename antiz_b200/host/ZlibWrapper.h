// GPU-backed stand-in for the reference's zlib RAII wrapper (reference: /root/reference/ZlibWrapper.h:5-100).
// Same class names, method set, argument meaning and return values (zlib's Z_* codes), so ZBuffSearcher-style
// code keeps compiling; the inflate itself runs in the K2 kernel through atz_inflate_stream().  The streaming
// calls (continuePrev / refillInput) are served by re-running the kernel over the input seen so far - this class
// is a compatibility seam, not the fast path (the fast path is atz_scan, which batches every candidate).
#ifndef ANTIZ_B200_ZLIBWRAPPER_H
#define ANTIZ_B200_ZLIBWRAPPER_H
#include "antiz_b200.h"
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef Z_OK
#define Z_OK 0
#define Z_STREAM_END 1
#define Z_NEED_DICT 2
#define Z_STREAM_ERROR (-2)
#define Z_DATA_ERROR (-3)
#define Z_MEM_ERROR (-4)
#define Z_BUF_ERROR (-5)
#define Z_SYNC_FLUSH 2
#define Z_FINISH 4
#endif

class BasicZlib {
  public:
    std::string errToString(int ret) {
        switch (ret) {
        case Z_OK: return "Z_OK";
        case Z_DATA_ERROR: return "Z_DATA_ERROR";
        case Z_NEED_DICT: return "Z_NEED_DICT";
        case Z_MEM_ERROR: return "Z_MEM_ERROR";
        case Z_BUF_ERROR: return "Z_BUF_ERROR";
        case Z_STREAM_ERROR: return "Z_STREAM_ERROR";
        default: return std::string{"Unknown Error: "} + std::to_string(ret);
        }
    }
};

class ZlibInflator : public BasicZlib {
  public:
    typedef unsigned int size_type;   // zlib's uInt
    typedef unsigned char *byteP;

    ZlibInflator() { open(); }
    explicit ZlibInflator(int WindowBits) {
        if (WindowBits != 15) throw std::runtime_error{"inflateInit2 failed"};   // the reference only ever uses 15 (ZlibWrapper.h:31)
        open();
    }
    ~ZlibInflator() { if (ctx_) atz_ctx_destroy(ctx_); }
    ZlibInflator(const ZlibInflator &) = delete;
    ZlibInflator &operator=(const ZlibInflator &) = delete;

    unsigned long totalInputByte() { return total_in_; }
    unsigned long totalOutputByte() { return total_out_; }
    size_type avail_out() { return avail_out_; }
    size_type avail_in() { return avail_in_; }

    int operator()(void *dest, size_type destlen, void *src, size_type srclen, int FlushType = Z_SYNC_FLUSH) {
        (void)FlushType;   // inflateReset + new buffers (ZlibWrapper.h:58-69 of the reference)
        in_.assign((const uint8_t *)src, (const uint8_t *)src + srclen);
        served_ = 0; total_in_ = 0; total_out_ = 0; finished_ = 0;
        out_ptr_ = (uint8_t *)dest; avail_out_ = destlen;
        return step();
    }
    int continuePrev(void *dest, size_type destlen, int FlushType = Z_SYNC_FLUSH) {
        (void)FlushType;
        out_ptr_ = (uint8_t *)dest; avail_out_ = destlen;
        return step();
    }
    int refillInput(void *src, size_type srcLen, int FlushType = Z_SYNC_FLUSH) {
        (void)FlushType;   // next_in is replaced: whatever was not consumed is dropped
        in_.resize(total_in_);
        in_.insert(in_.end(), (const uint8_t *)src, (const uint8_t *)src + srcLen);
        return step();
    }
    int lastRetVal() { return LastRet; }

  private:
    atz_ctx *ctx_ = nullptr;
    std::vector<uint8_t> in_, buf_;
    uint64_t served_ = 0, total_in_ = 0, total_out_ = 0;
    size_type avail_in_ = 0, avail_out_ = 0;
    uint8_t *out_ptr_ = nullptr;
    int LastRet = Z_OK, finished_ = 0;
    void open() { if (atz_ctx_create(0, &ctx_) != ATZ_OK) throw std::runtime_error{"inflate init Failed"}; }
    // zlib keeps decoder state between calls; here the kernel is re-run over the input seen so far and the
    // part of its output that has not been handed out yet is copied to next_out.
    int step() {
        if (finished_) { avail_in_ = (size_type)(in_.size() - total_in_); return LastRet = finished_; }   // DONE / BAD consume nothing
        uint64_t cap = served_ + avail_out_, olen = 0, used = 0;
        buf_.resize(cap + 1);
        int rc = atz_inflate_stream(ctx_, in_.data(), in_.size(), buf_.data(), cap, &olen, &used);
        if (olen > cap) olen = cap;
        uint64_t fresh = olen - served_;
        if (fresh) std::memcpy(out_ptr_, buf_.data() + served_, fresh);
        out_ptr_ += fresh; avail_out_ -= (size_type)fresh;
        bool progress = fresh != 0 || used != total_in_;
        served_ = olen; total_out_ = olen; total_in_ = used;
        avail_in_ = (size_type)(in_.size() - used);
        if (rc == ATZ_OK) return LastRet = finished_ = Z_STREAM_END;
        if (rc == ATZ_E_DATA) return LastRet = finished_ = Z_DATA_ERROR;
        if (rc == ATZ_E_SMALL || rc == ATZ_E_TRUNCATED) return LastRet = progress ? Z_OK : Z_BUF_ERROR;
        throw errToString(Z_STREAM_ERROR);
    }
};
#endif
