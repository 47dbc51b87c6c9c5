// fdeflate_b200.hpp -- C++17 host-side mirror of the reference crate's public API over the C ABI
// (include/fdeflate_b200.h).  Header-only.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the host side that a Rust
// user would get from the shim in INTEGRATION.md is mirrored here in C++: same names, argument
// meaning and error behaviour as image-rs/fdeflate 0.4.0-dev (src/lib.rs:29-36):
//
//   fdeflate::decompress_to_vec            src/decompress.rs:1079   (throws DecompressionError)
//   fdeflate::decompress_to_vec_bounded    src/decompress.rs:1111   (throws BoundedDecompressionError)
//   fdeflate::compress_to_vec_ultra_fast   src/compress/mod.rs:313
//   fdeflate::UltraFastCompressor<W>       src/compress/ultrafast.rs:9-181   (W: anything with write(ptr, n))
//   fdeflate::Compressor<W>(w, 0, zlib)    src/compress/mod.rs:69-215, level 0 only ("stored")
//   fdeflate::Batch                        the batch entry points this project adds
//
// All compute happens on the GPU behind the C ABI; there is no CPU fallback (Context construction
// throws without a CUDA device).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fdeflate_b200.h"

namespace fdeflate {

// src/decompress.rs:13-48, declaration order; values equal the C ABI status codes
enum class DecompressionErrorKind : int32_t {
    BadZlibHeader = 1, InsufficientInput, InvalidBlockType, InvalidUncompressedBlockLength, InvalidHlit,
    InvalidHdist, InvalidCodeLengthRepeat, BadCodeLengthHuffmanTree, BadLiteralLengthHuffmanTree,
    BadDistanceHuffmanTree, InvalidLiteralLengthCode, InvalidDistanceCode, InputStartsWithRun,
    DistanceTooFarBack, WrongChecksum, ExtraInput
};

struct DecompressionError : std::runtime_error {
    DecompressionErrorKind kind;
    explicit DecompressionError(int32_t code)
        : std::runtime_error("DecompressionError(" + std::to_string(code) + ")"),
          kind(static_cast<DecompressionErrorKind>(code)) {}
};

// src/decompress.rs:1090-1102
struct BoundedDecompressionError : std::runtime_error {
    bool output_too_large;
    std::vector<uint8_t> partial_output;  // OutputTooLarge { partial_output }
    int32_t inner;                        // DecompressionError { inner }
    BoundedDecompressionError(std::vector<uint8_t> partial)
        : std::runtime_error("OutputTooLarge"), output_too_large(true), partial_output(std::move(partial)), inner(0) {}
    explicit BoundedDecompressionError(int32_t code)
        : std::runtime_error("DecompressionError"), output_too_large(false), inner(code) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = fdb_create(device, &h_);
        if (rc != 0 || !h_) throw std::runtime_error("fdb_create failed (" + std::to_string(rc) + "): no CUDA device; there is no CPU fallback");
    }
    ~Context() { fdb_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    fdb_ctx* handle() const { return h_; }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + fdb_last_error(h_));
    }
    // long streams by many warps each, for device-pointer calls (host-buffer calls decide by themselves)
    void set_split_large(bool on) { check(fdb_set_split_large(h_, on ? 1 : 0), "fdb_set_split_large"); }
    void set_split_scratch(size_t bytes) { check(fdb_set_split_scratch(h_, bytes), "fdb_set_split_scratch"); }
    void set_split_threshold(size_t inflate_stream_bytes, size_t deflate_input_bytes) {
        check(fdb_set_split_threshold(h_, inflate_stream_bytes, deflate_input_bytes), "fdb_set_split_threshold");
    }

private:
    fdb_ctx* h_ = nullptr;
};

// Several GPUs of one box behind one handle (fdb_multi_*): batches shard by stream, no collective.
class DeviceSet {
public:
    explicit DeviceSet(const std::vector<int>& devices) {
        int rc = fdb_multi_create(devices.data(), (int)devices.size(), &m_);
        if (rc != 0 || !m_) throw std::runtime_error("fdb_multi_create failed (code " + std::to_string(rc) + "): no usable CUDA device");
    }
    ~DeviceSet() { fdb_multi_destroy(m_); }
    DeviceSet(const DeviceSet&) = delete;
    DeviceSet& operator=(const DeviceSet&) = delete;
    fdb_multi* handle() const { return m_; }
    int device_count() const { return fdb_multi_device_count(m_); }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + fdb_multi_last_error(m_));
    }

private:
    fdb_multi* m_ = nullptr;
};

// ---- batch entry points (new surface) ---------------------------------------------------------
struct InflateResult {
    std::vector<int32_t> status;
    std::vector<std::vector<uint8_t>> output;
    std::vector<uint64_t> consumed;
};

class Batch {
public:
    explicit Batch(Context& ctx) : ctx_(&ctx) {}
    explicit Batch(DeviceSet& set) : set_(&set) {}  // the same calls, sharded over the GPUs of the set

    InflateResult inflate(const std::vector<std::vector<uint8_t>>& streams, const std::vector<uint64_t>& out_caps,
                          uint32_t flags = 0) {
        const size_t n = streams.size();
        std::vector<uint64_t> in_off(n), in_len(n), out_off(n), out_len(n), consumed(n);
        std::vector<int32_t> status(n);
        uint64_t ip = 0, op = 0;
        for (size_t i = 0; i < n; i++) {
            in_off[i] = ip;
            in_len[i] = streams[i].size();
            ip = (ip + in_len[i] + 15) & ~uint64_t(15);
            out_off[i] = op;
            op = (op + out_caps[i] + 15) & ~uint64_t(15);
        }
        std::vector<uint8_t> in(ip + 16), out(op + 16);
        for (size_t i = 0; i < n; i++)
            if (in_len[i]) std::memcpy(in.data() + in_off[i], streams[i].data(), in_len[i]);
        if (set_)
            set_->check(fdb_multi_inflate_batch(set_->handle(), in.data(), in_off.data(), in_len.data(), out.data(), out_off.data(),
                                                out_caps.data(), out_len.data(), consumed.data(), status.data(), n, flags),
                        "fdb_multi_inflate_batch");
        else
            ctx_->check(fdb_inflate_batch(ctx_->handle(), in.data(), in_off.data(), in_len.data(), out.data(), out_off.data(),
                                          out_caps.data(), out_len.data(), consumed.data(), status.data(), n, flags),
                        "fdb_inflate_batch");
        InflateResult r;
        r.status = status;
        r.consumed = consumed;
        r.output.resize(n);
        for (size_t i = 0; i < n; i++) r.output[i].assign(out.begin() + out_off[i], out.begin() + out_off[i] + out_len[i]);
        return r;
    }

    std::vector<std::vector<uint8_t>> deflate_ultra_fast(const std::vector<std::vector<uint8_t>>& inputs) {
        return deflate(inputs, false);
    }
    std::vector<std::vector<uint8_t>> deflate_stored(const std::vector<std::vector<uint8_t>>& inputs) {
        return deflate(inputs, true);
    }

private:
    std::vector<std::vector<uint8_t>> deflate(const std::vector<std::vector<uint8_t>>& inputs, bool stored) {
        const size_t n = inputs.size();
        std::vector<uint64_t> in_off(n), in_len(n), out_off(n), out_cap(n), out_len(n);
        std::vector<int32_t> status(n);
        uint64_t ip = 0, op = 0;
        for (size_t i = 0; i < n; i++) {
            in_off[i] = ip;
            in_len[i] = inputs[i].size();
            ip = (ip + in_len[i] + 15) & ~uint64_t(15);
            out_off[i] = op;
            out_cap[i] = stored ? fdb_deflate_stored_bound(in_len[i]) : fdb_deflate_ultrafast_bound(in_len[i]);
            op += out_cap[i];
        }
        std::vector<uint8_t> in(ip + 16), out(op + 16);
        for (size_t i = 0; i < n; i++)
            if (in_len[i]) std::memcpy(in.data() + in_off[i], inputs[i].data(), in_len[i]);
        if (set_) {
            int rc = stored ? fdb_multi_deflate_stored_batch(set_->handle(), in.data(), in_off.data(), in_len.data(), out.data(),
                                                             out_off.data(), out_cap.data(), out_len.data(), status.data(), n)
                            : fdb_multi_deflate_ultrafast_batch(set_->handle(), in.data(), in_off.data(), in_len.data(), out.data(),
                                                                out_off.data(), out_cap.data(), out_len.data(), status.data(), n);
            set_->check(rc, "fdb_multi_deflate_*_batch");
        } else {
            int rc = stored ? fdb_deflate_stored_batch(ctx_->handle(), in.data(), in_off.data(), in_len.data(), out.data(),
                                                       out_off.data(), out_cap.data(), out_len.data(), status.data(), n)
                            : fdb_deflate_ultrafast_batch(ctx_->handle(), in.data(), in_off.data(), in_len.data(), out.data(),
                                                          out_off.data(), out_cap.data(), out_len.data(), status.data(), n);
            ctx_->check(rc, "fdb_deflate_*_batch");
        }
        std::vector<std::vector<uint8_t>> res(n);
        for (size_t i = 0; i < n; i++) {
            if (status[i] != FDB_OK) throw std::runtime_error("deflate status " + std::to_string(status[i]));
            res[i].assign(out.begin() + out_off[i], out.begin() + out_off[i] + out_len[i]);
        }
        return res;
    }
    Context* ctx_ = nullptr;
    DeviceSet* set_ = nullptr;
};

// ---- single-stream API with the reference's names ------------------------------------------------
inline std::vector<uint8_t> decompress_to_vec_bounded(Context& ctx, const std::vector<uint8_t>& input, uint64_t maxlen) {
    Batch b(ctx);
    uint64_t cap = std::min<uint64_t>(maxlen, std::max<uint64_t>(1024, 4 * input.size()));
    for (;;) {
        InflateResult r = b.inflate({input}, {cap});
        if (r.status[0] == FDB_OK) return std::move(r.output[0]);
        if (r.status[0] == FDB_OUTPUT_TOO_LARGE) {
            if (cap >= maxlen) throw BoundedDecompressionError(std::move(r.output[0]));
            cap = std::min<uint64_t>(maxlen, cap * 4);  // the reference grows its Vec and carries on (:1132-1134)
            continue;
        }
        throw BoundedDecompressionError(r.status[0]);
    }
}

inline std::vector<uint8_t> decompress_to_vec(Context& ctx, const std::vector<uint8_t>& input) {
    try {
        return decompress_to_vec_bounded(ctx, input, uint64_t(1) << 62);
    } catch (const BoundedDecompressionError& e) {
        if (!e.output_too_large) throw DecompressionError(e.inner);
        throw;
    }
}

// src/decompress.rs:96-342: the streaming decoder with the read() contract of :158-184.  Its state machine lives on the
// device (fdb_stream_*): every call resumes at the token boundary the last one stopped at.  read() takes all of
// `input` (what cannot be parsed yet is kept by the context) and returns {input.size(), bytes written}; when the output
// is full, call again with more room and no input.
class Decompressor {
public:
    explicit Decompressor(Context& ctx) : ctx_(ctx) { ctx_.check(fdb_stream_open_batch(ctx_.handle(), &id_, 1), "fdb_stream_open_batch"); }
    ~Decompressor() { fdb_stream_close_batch(ctx_.handle(), &id_, 1); }
    Decompressor(const Decompressor&) = delete;
    Decompressor& operator=(const Decompressor&) = delete;
    void ignore_adler32() { flags_ |= FDB_FLAG_IGNORE_ADLER32; }
    bool is_done() const { return done_; }
    std::pair<size_t, size_t> read(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_len, size_t output_position) {
        if (done_) return {0, 0};  // :185-187
        if (output_position > output_len) throw std::out_of_range("output_position out of bounds");  // the reference panics (:189)
        const uint64_t in_off = 0, in_len = input_len, out_off = output_position, room = output_len - output_position;
        uint64_t produced = 0;
        int32_t status = 0;
        ctx_.check(fdb_stream_read_batch(ctx_.handle(), &id_, input, &in_off, &in_len, output, &out_off, &room, &produced, &status, 1, flags_),
                   "fdb_stream_read_batch");
        if (status > 0) throw DecompressionError(status);
        if (status == FDB_OK) done_ = true;
        return {input_len, (size_t)produced};
    }

private:
    Context& ctx_;
    uint32_t id_ = 0;
    uint32_t flags_ = 0;
    bool done_ = false;
};

inline std::vector<uint8_t> compress_to_vec_ultra_fast(Context& ctx, const std::vector<uint8_t>& input) {
    return Batch(ctx).deflate_ultra_fast({input})[0];
}

// src/compress/ultrafast.rs:9-181.  W needs `void write(const uint8_t*, size_t)`.
// The reference's bytes depend on how the input is cut into write_data calls: its zero-run counter and its 8-byte
// chunking restart with every call (:97-99).  Each call is therefore compressed as its own stream, all of them in ONE
// device batch at finish(), and their token bits are spliced on the host behind one header: byte for byte what the
// reference writes for the same call pattern (tests/cpp/test_cpp_api.cpp checks it against the oracle's
// new / write_data / finish).
namespace detail {
// bits [from, to) of `src` (LSB first) appended to `dst` at bit position `pos`
inline void append_bits(std::vector<uint8_t>& dst, uint64_t& pos, const std::vector<uint8_t>& src, uint64_t from, uint64_t to) {
    dst.resize((pos + (to - from) + 7) / 8 + 8, 0);
    for (uint64_t b = from; b < to;) {
        const uint64_t take = std::min<uint64_t>(to - b, 8 - (b & 7));  // bits of one source byte
        const uint32_t v = (uint32_t(src[b >> 3]) >> (b & 7)) & ((1u << take) - 1u);
        const uint32_t sh = uint32_t(pos & 7);
        dst[pos >> 3] |= uint8_t(v << sh);
        if (sh + take > 8) dst[(pos >> 3) + 1] |= uint8_t(v >> (8 - sh));
        pos += take;
        b += take;
    }
}
// RFC 1950 adler32, continued from `adler` (the checksum of a multi-call stream covers all calls in order)
inline uint32_t adler32(uint32_t adler, const uint8_t* p, size_t n) {
    uint32_t a = adler & 0xffffu, b = adler >> 16;
    while (n) {
        size_t k = std::min<size_t>(n, 5552);
        n -= k;
        while (k--) {
            a += *p++;
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
    }
    return (b << 16) | a;
}
}  // namespace detail

template <class W>
class UltraFastCompressor {
public:
    UltraFastCompressor(Context& ctx, W writer) : ctx_(ctx), w_(std::move(writer)) {}
    void write_data(const uint8_t* data, size_t n) { calls_.emplace_back(data, data + n); }
    W finish() {
        if (calls_.empty()) calls_.emplace_back();
        std::vector<std::vector<uint8_t>> z = Batch(ctx_).deflate_ultra_fast(calls_);
        if (z.size() == 1) {
            w_.write(z[0].data(), z[0].size());
            return std::move(w_);
        }
        const uint64_t header_bits = 53 * 8 + 5;  // ultrafast.rs:87-88
        std::vector<uint8_t> out;
        uint64_t pos = 0;
        detail::append_bits(out, pos, z[0], 0, header_bits);
        uint32_t adler = 1;
        for (size_t i = 0; i < z.size(); i++) {
            // the stream body ends with the 12-bit end-of-block code 0x8ff, whose top bit is the highest set bit
            // before the 4 checksum bytes (the padding behind it is zero)
            size_t last = z[i].size() - 5;
            while (z[i][last] == 0) last--;
            uint64_t end = 8 * uint64_t(last) + 8;
            for (uint8_t v = z[i][last]; !(v & 0x80); v = uint8_t(v << 1)) end--;
            detail::append_bits(out, pos, z[i], header_bits, end - 12);
            adler = detail::adler32(adler, calls_[i].data(), calls_[i].size());
        }
        const std::vector<uint8_t> eob = {0xff, 0x08};  // code 2303, 12 bits
        detail::append_bits(out, pos, eob, 0, 12);
        out.resize((pos + 7) / 8);
        for (int k = 3; k >= 0; k--) out.push_back(uint8_t(adler >> (8 * k)));
        w_.write(out.data(), out.size());
        return std::move(w_);
    }

private:
    Context& ctx_;
    W w_;
    std::vector<std::vector<uint8_t>> calls_;
};

// src/compress/mod.rs:47-215 restricted to level 0 ("stored"): block boundaries do not depend on the
// write_data call pattern, so calls are simply concatenated.
template <class W>
class Compressor {
public:
    Compressor(Context& ctx, W writer, uint8_t level, bool zlib) : ctx_(ctx), w_(std::move(writer)), zlib_(zlib) {
        if (level != 0) throw std::invalid_argument("only level 0 (stored) is on the accelerated path");
    }
    void write_data(const uint8_t* data, size_t n) { buf_.insert(buf_.end(), data, data + n); }
    W finish() {
        std::vector<uint8_t> z = Batch(ctx_).deflate_stored({buf_})[0];
        if (zlib_) w_.write(z.data(), z.size());
        else w_.write(z.data() + 2, z.size() - 6);
        return std::move(w_);
    }

private:
    Context& ctx_;
    W w_;
    bool zlib_;
    std::vector<uint8_t> buf_;
};

}  // namespace fdeflate
