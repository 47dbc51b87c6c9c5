// C++ harness for include/fdeflate_b200.hpp, linked against a library that implements the C ABI.
// The expected bytes come from the oracle (linked here as the CHECKER only).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fdeflate_b200.hpp"
extern "C" {
#include "fdeflate_oracle.h"
}

#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                   \
        }                                                               \
    } while (0)

struct VecWriter {
    std::vector<uint8_t> v;
    void write(const uint8_t* p, size_t n) { v.insert(v.end(), p, p + n); }
};

static std::vector<uint8_t> oracle_uf(const std::vector<uint8_t>& d) {
    std::vector<uint8_t> o(fdo_ultrafast_bound(d.size()));
    o.resize(fdo_compress_ultra_fast(d.data(), d.size(), o.data(), o.size()));
    return o;
}

int main() {
    fdeflate::Context ctx(0);
    std::vector<uint8_t> data(20000);
    uint32_t s = 12345;
    for (auto& b : data) {
        s = s * 1664525u + 1013904223u;
        b = (s >> 24) < 120 ? 0 : uint8_t((s >> 16) % 7);
    }
    // compress_to_vec_ultra_fast: byte-identical to the oracle, and it round-trips
    std::vector<uint8_t> z = fdeflate::compress_to_vec_ultra_fast(ctx, data);
    CHECK(z == oracle_uf(data));
    CHECK(fdeflate::decompress_to_vec(ctx, z) == data);
    // bounded: exact fit succeeds, one byte less is OutputTooLarge with the partial output
    CHECK(fdeflate::decompress_to_vec_bounded(ctx, z, data.size()) == data);
    try {
        fdeflate::decompress_to_vec_bounded(ctx, z, data.size() - 1);
        CHECK(false);
    } catch (const fdeflate::BoundedDecompressionError& e) {
        CHECK(e.output_too_large && e.partial_output.size() == data.size() - 1);
        CHECK(std::equal(e.partial_output.begin(), e.partial_output.end(), data.begin()));
    }
    // errors carry the reference's variant
    std::vector<uint8_t> bad = z;
    bad.back() ^= 1;
    try {
        fdeflate::decompress_to_vec(ctx, bad);
        CHECK(false);
    } catch (const fdeflate::DecompressionError& e) {
        CHECK(e.kind == fdeflate::DecompressionErrorKind::WrongChecksum);
    }
    std::vector<uint8_t> trunc(z.begin(), z.end() - 7);
    try {
        fdeflate::decompress_to_vec(ctx, trunc);
        CHECK(false);
    } catch (const fdeflate::DecompressionError& e) {
        CHECK(e.kind == fdeflate::DecompressionErrorKind::InsufficientInput);
    }
    // UltraFastCompressor / Compressor(level 0)
    fdeflate::UltraFastCompressor<VecWriter> uf(ctx, VecWriter{});
    uf.write_data(data.data(), data.size());
    CHECK(uf.finish().v == z);
    // several write_data calls: the reference's bytes depend on the call pattern (ultrafast.rs:97-99); compare with
    // the oracle's new / write_data / finish for a few patterns (cuts inside zero runs, empty calls, one-byte calls)
    {
        const size_t cut_sets[][6] = {{0, 1, 2, 3, 7000, 7001}, {8, 16, 4096, 4097, 12345, 19999}, {5, 5, 5, 20000, 20000, 20000},
                                      {3333, 6666, 9999, 13332, 16665, 19998}};
        for (const auto& cuts : cut_sets) {
            fdeflate::UltraFastCompressor<VecWriter> m(ctx, VecWriter{});
            std::vector<uint8_t> want(fdo_ultrafast_bound(data.size()) + 256);
            fdo_ultrafast* o = fdo_ultrafast_new(want.data(), want.size());
            size_t prev = 0;
            for (int k = 0; k <= 6; k++) {
                const size_t next = k < 6 ? cuts[k] : data.size();
                m.write_data(data.data() + prev, next - prev);
                fdo_ultrafast_write_data(o, data.data() + prev, next - prev);
                prev = next;
            }
            want.resize(fdo_ultrafast_finish(o));
            std::vector<uint8_t> got = m.finish().v;
            CHECK(got == want);
            CHECK(fdeflate::decompress_to_vec(ctx, got) == data);
        }
        fdeflate::UltraFastCompressor<VecWriter> none(ctx, VecWriter{});
        CHECK(none.finish().v == oracle_uf({}));
    }
    fdeflate::Compressor<VecWriter> st(ctx, VecWriter{}, 0, true);
    st.write_data(data.data(), 7000);
    st.write_data(data.data() + 7000, data.size() - 7000);
    std::vector<uint8_t> stored = st.finish().v;
    std::vector<uint8_t> want(fdo_stored_bound(data.size()));
    want.resize(fdo_compress_stored(data.data(), data.size(), want.data(), want.size()));
    CHECK(stored == want);
    CHECK(fdeflate::decompress_to_vec(ctx, stored) == data);
    // batch
    fdeflate::Batch b(ctx);
    auto r = b.inflate({z, stored, trunc}, {data.size(), data.size(), data.size()});
    CHECK(r.status[0] == FDB_OK && r.status[1] == FDB_OK && r.status[2] == FDB_INSUFFICIENT_INPUT);
    CHECK(r.output[0] == data && r.output[1] == data);
    // Decompressor::read fed in pieces of every size, small output rooms: same bytes as the whole-buffer call
    for (size_t step : {size_t(1), size_t(7), size_t(333), z.size()}) {
        fdeflate::Decompressor d(ctx);
        std::vector<uint8_t> got, room(257);
        size_t pos = 0;
        for (int guard = 0; !d.is_done() && guard < 200000; guard++) {
            const size_t take = std::min(step, z.size() - pos);
            auto rr = d.read(z.data() + pos, take, room.data(), room.size(), 0);
            pos += rr.first;
            got.insert(got.end(), room.begin(), room.begin() + rr.second);
        }
        CHECK(d.is_done() && got == data);
        CHECK(d.read(z.data(), 4, room.data(), room.size(), 0) == std::make_pair(size_t(0), size_t(0)));
    }
    {
        fdeflate::Decompressor d(ctx);
        std::vector<uint8_t> room(data.size());
        try {
            d.read(bad.data(), bad.size(), room.data(), room.size(), 0);
            CHECK(false);
        } catch (const fdeflate::DecompressionError& e) {
            CHECK(e.kind == fdeflate::DecompressionErrorKind::WrongChecksum);
        }
    }
    // a device set of two contexts (world = 2; on a one-GPU box both sit on device 0): same results as one context
    {
        fdeflate::DeviceSet set({0, 0});
        CHECK(set.device_count() == 2);
        fdeflate::Batch mb(set);
        std::vector<std::vector<uint8_t>> inputs;
        for (size_t k = 0; k < 9; k++) inputs.emplace_back(data.begin() + 100 * k, data.begin() + 100 * k + 1500 * (k + 1));
        inputs.push_back({});
        std::vector<std::vector<uint8_t>> zs = mb.deflate_ultra_fast(inputs);
        std::vector<uint64_t> caps;
        for (size_t k = 0; k < inputs.size(); k++) {
            CHECK(zs[k] == oracle_uf(inputs[k]));
            caps.push_back(inputs[k].size());
        }
        zs.push_back(trunc);
        caps.push_back(data.size());
        auto mr = mb.inflate(zs, caps);
        for (size_t k = 0; k < inputs.size(); k++) CHECK(mr.status[k] == FDB_OK && mr.output[k] == inputs[k]);
        CHECK(mr.status[inputs.size()] == FDB_INSUFFICIENT_INPUT);
        std::vector<uint32_t> owner(zs.size());
        CHECK(fdb_multi_last_partition(set.handle(), owner.data(), owner.size()) == 0);
        CHECK(*std::max_element(owner.begin(), owner.end()) == 1);
    }
    std::puts("cpp api ok");
    return 0;
}
