// capi.cu -- the C ABI of include/fdeflate_b200.h: context, kernel launches, host staging.
//
// Compiled by nvcc for sm_100a into libfdeflate_b200.so (the product).  The same file is also
// compiled by g++ -DFDB_EMUL against tests/emul/ for CPU-side logic tests of the kernels; that
// build is test infrastructure and is never loaded by the package.
#include "../../include/fdeflate_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <queue>
#include <thread>
#include <vector>

#include "simt.h"
#include "fdb_common.h"
#include "fdb_tables.h"
#include "adler.cuh"
#include "inflate_general.cuh"
#include "inflate_uf.cuh"
#include "inflate_uf2.cuh"
#include "deflate_uf.cuh"
#include "deflate_ufb.cuh"
#include "deflate_stored.cuh"
#include "synth.cuh"
#include "png_filter.cuh"
#include "deflate_png.cuh"
#include "crc32.cuh"

using namespace fdb;

// One pipeline lane of the host-buffer entry points: its own CUDA stream, work counters and fallback
// work list, so that chunks of one batch can overlap (H2D of chunk k+1 | kernels of chunk k | D2H of k-1).
// scratch of the span-by-span inflate path (FDB_FLAG_SPLIT_LARGE), grown on first use
struct fdb_split_scratch {
    K4Item* items = nullptr;
    size_t items_bytes = 0;
    uint32_t* per_stream = nullptr;  // item0 | nspans | need | done | failed, n words each
    size_t per_stream_bytes = 0;
    uint32_t* rec = nullptr;         // lane records the count pass leaves for the write pass (inflate_uf.cuh), 8 KiB per span
    size_t rec_bytes = 0;
};
static const uint32_t FDB_SPLIT_ITEMS = 1u << 20;  // spans per batch (64 GiB of compressed input); scratch is allocated on first use
// ... and of the segment-by-segment deflate path
struct fdb_dsplit_scratch {
    DfItem* items = nullptr;
    size_t items_bytes = 0;
    uint32_t* per_stream = nullptr;  // item0 | nseg | adler | work order, n words each, then 2 x DF_ORDER_CLASSES counters
    size_t per_stream_bytes = 0;
};

struct fdb_lane {
    cudaStream_t st = nullptr;
    uint32_t* d_counters = nullptr;  // same layout as fdb_ctx::d_counters
    uint32_t* d_worklist = nullptr;
    size_t worklist_cap = 0;
    fdb_split_scratch split;
    fdb_dsplit_scratch dsplit;
};
static const int FDB_LANES = 8;        // compute lanes created per context
static const int FDB_MAX_CHUNKS = 64;  // chunks per host-buffer call (one event pair each)

struct fdb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // lane 0 of the host-buffer entry points
    fdb_lane lanes[FDB_LANES];
    int n_lanes = 4;                // lanes in use (FDB_PIPELINE_LANES overrides; a tuning knob)
    int depth = 4;                  // chunks in flight ahead of the payload copies (FDB_PIPELINE_DEPTH)
    int png_fused = 0;              // PNG encode, filter modes 0..4: the filter runs inside the encoder (FDB_PNG_FUSED=1).  Opt-in:
                                    // it saves the filtered image's trip through device memory but measures 1.6-2.7x slower
                                    // than filter kernel + encoder (profiles/r03_png_fused_speed.txt)
    int direct_out = 0;             // ultra-fast deflate writes into pinned host output buffers itself (FDB_DIRECT_OUT=1;
                                    // measured slower than the payload copies on the bench step, so it is opt-in)
    cudaStream_t h2d_st = nullptr;  // every chunk's input travels on this stream, in order   (shared per device,
    cudaStream_t d2h_st = nullptr;  // every chunk's payload travels back on this one          see fdb_xfer)
    cudaEvent_t ev_in[FDB_MAX_CHUNKS] = {nullptr};   // chunk k's input is on the device
    cudaEvent_t ev_res[FDB_MAX_CHUNKS] = {nullptr};  // chunk k's kernels are done (its per-stream results are on the host)
    uint8_t* h_res = nullptr;       // pinned: per-stream results of the host-buffer entry points
    size_t h_res_cap = 0;
    int64_t last_general_host = -1; // fallback count of the last host-buffer inflate (-1: ask the device)
    int64_t last_split_host = -1;   // spans of the last host-buffer inflate (-1: ask the device)
    uint64_t chunk_bytes = 128ull << 20;  // slot span per pipeline chunk
    uint32_t* d_counters = nullptr; // [0] K4 next  [1] fallback count  [2] K3 next  [3] deflate next
                                    // [4] spans reserved  [5..7] next span of the count / scan / write pass
    uint32_t* d_worklist = nullptr;
    size_t worklist_cap = 0;
    fdb_split_scratch split;
    fdb_dsplit_scratch dsplit;
    int split_large = 0;            // device-pointer calls: long streams by many warps (fdb_set_split_large)
    uint64_t inflate_split_min = K4_SPLIT_MIN_BYTES;  // fdb_set_split_threshold
    uint64_t split_scratch = 256ull << 20;            // fdb_set_split_scratch: lane-record pool of a device-pointer inflate call
    uint64_t deflate_split_min = DF_SPLIT_MIN_BYTES;
    uint64_t deflate_auto_min = DF_AUTO_SPLIT_BYTES;  // host-buffer deflate: a chunk with an input this long takes the segment path
    UfEncTables* d_enc = nullptr;
    UfDecTables* d_dec = nullptr;
    CrcTables* d_crc = nullptr;
    // host-API staging (grow-only)
    uint8_t* d_in = nullptr;
    size_t d_in_cap = 0;
    uint8_t* d_out = nullptr;
    size_t d_out_cap = 0;
    uint8_t* d_mid = nullptr;    // filtered images between the two steps of fdb_png_{decode,encode}_batch
    size_t d_mid_cap = 0;
    uint64_t* d_meta = nullptr;  // in_off | in_len | out_off | out_cap | out_len | consumed | status(int32)
    size_t d_meta_cap = 0;
    uint64_t launches = 0;
    char err[512] = {0};
    // streaming decoders (fdb_stream_*): per-decoder device state + window buffer, host copy of the unparsed input tail
    struct stream_slot {
        bool open = false;
        K3StreamState* d_state = nullptr;
        uint8_t* d_buf = nullptr;
        size_t buf_cap = 0;
        uint64_t pos = 0, lo = 0;       // d_buf[lo .. pos) = the decoder's most recent output
        std::vector<uint8_t> tail;      // input bytes the decoder has not parsed yet
        uint32_t start_bit = 0;         // bits of tail[0] already consumed
        bool finished = false;
        int32_t final_status = 0;       // once finished: FDB_OK (stream complete) or the error
    };
    std::vector<stream_slot> streams;
    K3StreamJob* d_jobs = nullptr;
    size_t d_jobs_cap = 0;
};

// ---- optional timeline of the host pipeline (FDB_TRACE=<file>): one line per chunk with the device times
// (ms since the first traced call of the process) at which its H2D began / ended, its kernels ended and
// its payload D2H began / ended.  Diagnostic only; off unless the variable is set.
struct fdb_trace_chunk {
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};
static cudaEvent_t g_trace_base = nullptr;
static const char* trace_path() {
    static const char* p = getenv("FDB_TRACE");
    return (p && *p) ? p : nullptr;
}

// The copy streams are shared by every context on a device: copies of one direction then execute in the
// order the host threads queued them.  (Per-context copy streams are separate hardware channels, and a
// channel that always has a copy queued was seen to keep the copy engine until its whole batch was through,
// starving the other context for 20 ms.)
struct fdb_xfer {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    int refs = 0;
};
static std::mutex g_xfer_mu;
static fdb_xfer g_xfer[64];

static cudaError_t xfer_acquire(int device, cudaStream_t* h2d, cudaStream_t* d2h) {
    std::lock_guard<std::mutex> g(g_xfer_mu);
    fdb_xfer& x = g_xfer[device & 63];
    if (x.refs == 0) {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&x.h2d, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&x.d2h, cudaStreamNonBlocking)) != cudaSuccess) {
            cudaStreamDestroy(x.h2d);
            x.h2d = nullptr;
            return e;
        }
    }
    x.refs++;
    *h2d = x.h2d;
    *d2h = x.d2h;
    return cudaSuccess;
}
static void xfer_release(int device) {
    std::lock_guard<std::mutex> g(g_xfer_mu);
    fdb_xfer& x = g_xfer[device & 63];
    if (x.refs > 0 && --x.refs == 0) {
        cudaStreamSynchronize(x.h2d);
        cudaStreamSynchronize(x.d2h);
        cudaStreamDestroy(x.h2d);
        cudaStreamDestroy(x.d2h);
        x.h2d = x.d2h = nullptr;
    }
}

static int fail(fdb_ctx* c, const char* what, cudaError_t e) {
    if (c) snprintf(c->err, sizeof c->err, "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
    return e == cudaSuccess ? -1 : (int)e;
}
#define FDB_TRY(call)                                 \
    do {                                              \
        cudaError_t e_ = (call);                      \
        if (e_ != cudaSuccess) return fail(ctx, #call, e_); \
    } while (0)

extern "C" const char* fdb_version(void) {
#ifdef FDB_EMUL
    return "fdeflate_b200 0.1 (SIMT emulator build, tests only)";
#else
    return "fdeflate_b200 0.1 (CUDA sm_100a)";
#endif
}

extern "C" const char* fdb_last_error(const fdb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t fdb_launch_count(const fdb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int fdb_set_pipeline_chunk(fdb_ctx* ctx, size_t bytes) {
    if (!ctx) return -1;
    ctx->chunk_bytes = bytes ? bytes : (128ull << 20);
    return 0;
}

extern "C" int64_t fdb_last_general_count(fdb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return -1;
    if (ctx->last_general_host >= 0) return ctx->last_general_host;
    uint32_t v = 0;
    if (cudaStreamSynchronize((cudaStream_t)cuda_stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, ctx->d_counters + 1, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}

extern "C" int fdb_set_split_large(fdb_ctx* ctx, int on) {
    if (!ctx) return -1;
    ctx->split_large = on ? 1 : 0;
    return 0;
}

extern "C" int fdb_set_split_threshold(fdb_ctx* ctx, size_t inflate_bytes, size_t deflate_bytes) {
    if (!ctx) return -1;
    ctx->inflate_split_min = inflate_bytes ? inflate_bytes : K4_SPLIT_MIN_BYTES;
    ctx->deflate_split_min = deflate_bytes ? deflate_bytes : DF_SPLIT_MIN_BYTES;
    ctx->deflate_auto_min = deflate_bytes ? deflate_bytes : DF_AUTO_SPLIT_BYTES;
    return 0;
}

extern "C" int fdb_set_split_scratch(fdb_ctx* ctx, size_t bytes) {
    if (!ctx) return -1;
    ctx->split_scratch = bytes;
    return 0;
}

extern "C" int64_t fdb_last_split_spans(fdb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return -1;
    if (ctx->last_split_host >= 0) return ctx->last_split_host;
    uint32_t v = 0;
    if (cudaStreamSynchronize((cudaStream_t)cuda_stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, ctx->d_counters + 4, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)std::min<uint32_t>(v, FDB_SPLIT_ITEMS);
}

extern "C" int fdb_create(int device, fdb_ctx** out) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return e != cudaSuccess ? (int)e : -2;
    fdb_ctx* ctx = new (std::nothrow) fdb_ctx();
    if (!ctx) return -3;
    ctx->device = device;
    UfHostTables* ht = new (std::nothrow) UfHostTables();
    if (!ht || !build_uf_host_tables(*ht)) {
        delete ht;
        delete ctx;
        return -4;
    }
    UfEncTables enc;
    UfDecTables dec;
    memset(enc.lit, 0, sizeof enc.lit);
    for (int i = 0; i < 256; i++) {
        enc.lit[i].x = ht->code[i];
        enc.lit[i].y = ht->len[i];
    }
    memcpy(enc.tail_tok, ht->tail_tok, sizeof enc.tail_tok);
    memcpy(enc.header, ht->header, sizeof enc.header);
    memcpy(dec.wt, ht->wt, sizeof dec.wt);
    memcpy(dec.ct, ht->ct, sizeof dec.ct);
    memcpy(dec.bt, ht->bt, sizeof dec.bt);
    memcpy(dec.header, ht->header, sizeof dec.header);
    delete ht;
    auto bail = [&](cudaError_t err) {
        int r = fail(ctx, "fdb_create", err);
        fdb_destroy(ctx);
        return r;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e);
    if ((e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail(e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc((void**)&ctx->d_counters, 32 * sizeof(uint32_t))) != cudaSuccess) return bail(e);
    if ((e = xfer_acquire(device, &ctx->h2d_st, &ctx->d2h_st)) != cudaSuccess) return bail(e);
    for (int k = 0; k < FDB_MAX_CHUNKS; k++) {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_res[k], cudaEventDisableTiming)) != cudaSuccess) return bail(e);
    }
    if (const char* nl = getenv("FDB_PIPELINE_LANES")) {
        int v = atoi(nl);
        if (v >= 1 && v <= FDB_LANES) ctx->n_lanes = v;
    }
    if (const char* dq = getenv("FDB_DIRECT_OUT")) ctx->direct_out = atoi(dq) != 0;
    if (const char* pf = getenv("FDB_PNG_FUSED")) ctx->png_fused = atoi(pf) != 0;
    if (const char* nd = getenv("FDB_PIPELINE_DEPTH")) {
        int v = atoi(nd);
        if (v >= 1 && v <= FDB_MAX_CHUNKS) ctx->depth = v;
    }
    ctx->lanes[0].st = ctx->stream;
    for (int l = 0; l < ctx->n_lanes; l++) {
        if (l && (e = cudaStreamCreateWithFlags(&ctx->lanes[l].st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
        if ((e = cudaMalloc((void**)&ctx->lanes[l].d_counters, 32 * sizeof(uint32_t))) != cudaSuccess) return bail(e);
    }
    if ((e = cudaMalloc((void**)&ctx->d_enc, sizeof(UfEncTables))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc((void**)&ctx->d_dec, sizeof(UfDecTables))) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(ctx->d_enc, &enc, sizeof enc, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(ctx->d_dec, &dec, sizeof dec, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    {
        CrcTables crc;
        build_crc_tables(crc);
        if ((e = cudaMalloc((void**)&ctx->d_crc, sizeof(CrcTables))) != cudaSuccess) return bail(e);
        if ((e = cudaMemcpy(ctx->d_crc, &crc, sizeof crc, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    }
    if ((e = cudaFuncSetAttribute(inflate_uf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(inflate_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K3Smem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(inflate_uf_split_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(inflate_uf_split_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(inflate_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K3Smem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(deflate_stored_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StoredSmem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(deflate_ufb_png_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UbSmem))) != cudaSuccess)
        return bail(e);
    if ((e = cudaFuncSetAttribute(deflate_ufb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UbSmem))) != cudaSuccess)
        return bail(e);
    *out = ctx;
    return 0;
}

extern "C" void fdb_destroy(fdb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int l = 0; l < FDB_LANES; l++) {
        if (ctx->lanes[l].st) {
            cudaStreamSynchronize(ctx->lanes[l].st);
            cudaStreamDestroy(ctx->lanes[l].st);
        }
        cudaFree(ctx->lanes[l].d_counters);
        cudaFree(ctx->lanes[l].d_worklist);
        cudaFree(ctx->lanes[l].split.items);
        cudaFree(ctx->lanes[l].split.per_stream);
        cudaFree(ctx->lanes[l].split.rec);
        cudaFree(ctx->lanes[l].dsplit.items);
        cudaFree(ctx->lanes[l].dsplit.per_stream);
    }
    if (!ctx->lanes[0].st && ctx->stream) cudaStreamDestroy(ctx->stream);
    for (int k = 0; k < FDB_MAX_CHUNKS; k++) {
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
        if (ctx->ev_res[k]) cudaEventDestroy(ctx->ev_res[k]);
    }
    if (ctx->h2d_st) xfer_release(ctx->device);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_worklist);
    cudaFree(ctx->split.items);
    cudaFree(ctx->split.per_stream);
    cudaFree(ctx->split.rec);
    cudaFree(ctx->dsplit.items);
    cudaFree(ctx->dsplit.per_stream);
    cudaFree(ctx->d_enc);
    cudaFree(ctx->d_dec);
    cudaFree(ctx->d_crc);
    cudaFree(ctx->d_in);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_mid);
    cudaFree(ctx->d_meta);
    for (auto& sl : ctx->streams) {
        cudaFree(sl.d_state);
        cudaFree(sl.d_buf);
    }
    cudaFree(ctx->d_jobs);
    delete ctx;
}

static int grow(fdb_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    size_t ncap = std::max(need + need / 8, (size_t)1 << 20);
    ncap = (ncap + 255) & ~(size_t)255;
    if (*p) {
        for (int l = 0; l < FDB_LANES; l++)
            if (ctx->lanes[l].st) FDB_TRY(cudaStreamSynchronize(ctx->lanes[l].st));
        if (ctx->h2d_st) FDB_TRY(cudaStreamSynchronize(ctx->h2d_st));
        if (ctx->d2h_st) FDB_TRY(cudaStreamSynchronize(ctx->d2h_st));
        FDB_TRY(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    FDB_TRY(cudaMalloc(p, ncap));
    *cap = ncap;
    return 0;
}

// ---- inflate ----------------------------------------------------------------------------------
// K4 over the whole batch, then K3 over the streams K4 declined (or K3 alone with FDB_FLAG_GENERAL_ONLY)
// dense = the batch is one chunk of a pipeline: pack its streams onto as few SMs as they fill, so that the
// kernels of the chunks in flight (and of other contexts) run side by side instead of each one holding
// every SM with a warp or two.
static int launch_inflate(fdb_ctx* ctx, const InflateBatch& b, uint32_t* counters, uint32_t** worklist,
                          size_t* worklist_cap, cudaStream_t st, bool dense = false, fdb_split_scratch* ss = nullptr,
                          uint64_t in_total = 0 /* bytes of input in the batch if the host knows them */) {
    const size_t n = b.n;
    const bool split = ss && (b.flags & FDB_FLAG_SPLIT_LARGE) && !(b.flags & FDB_FLAG_GENERAL_ONLY);
    uint32_t rec_items = 0;
    if (split) {
        int r;
        // lane records: one span per 64 KiB of compressed data plus a partial one per split stream when the sizes are
        // known, the context's pool size otherwise; spans beyond the pool are counted again by the write pass
        uint64_t want_items = ctx->split_scratch / (K4_REC_WORDS * sizeof(uint32_t));
        if (in_total) {
            const uint64_t streams = std::min<uint64_t>(n, in_total / ctx->inflate_split_min + 1);
            want_items = in_total / (K4_SPAN_WORDS * 4) + streams;
        }
        want_items = std::min<uint64_t>(want_items, FDB_SPLIT_ITEMS);
        void* p = ss->rec;
        if (want_items && grow(ctx, &p, &ss->rec_bytes, (size_t)want_items * K4_REC_WORDS * sizeof(uint32_t)) == 0) {
            ss->rec = (uint32_t*)p;
            rec_items = (uint32_t)std::min<uint64_t>(ss->rec_bytes / (K4_REC_WORDS * sizeof(uint32_t)), FDB_SPLIT_ITEMS);
        } else {
            ss->rec = nullptr;  // (no memory for the pool: the write pass counts for itself)
            ss->rec_bytes = 0;
            cudaGetLastError();
        }
        p = ss->items;
        if ((r = grow(ctx, &p, &ss->items_bytes, (size_t)FDB_SPLIT_ITEMS * sizeof(K4Item)))) return r;
        ss->items = (K4Item*)p;
        p = ss->per_stream;
        r = grow(ctx, &p, &ss->per_stream_bytes, 5 * n * sizeof(uint32_t));
        ss->per_stream = (uint32_t*)p;
        if (r) return r;
    }
    if (n > *worklist_cap) {
        void* p = *worklist;
        size_t cap_bytes = *worklist_cap * sizeof(uint32_t);
        int r = grow(ctx, &p, &cap_bytes, n * sizeof(uint32_t));
        *worklist = (uint32_t*)p;
        *worklist_cap = cap_bytes / sizeof(uint32_t);
        if (r) return r;
    }
    FDB_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(uint32_t), st));
    const uint32_t sms = (uint32_t)std::max(ctx->sm_count, 1);
    const uint32_t* split_item0 = nullptr;
    if (split) {
        // long ultra-fast-format streams, span by span: plan, count, scan, write (inflate_uf.cuh)
        K4Split sp;
        sp.items = ss->items;
        sp.item_cap = FDB_SPLIT_ITEMS;
        sp.n_items = counters + 4;
        sp.item0 = ss->per_stream;
        sp.nspans = ss->per_stream + n;
        sp.need = ss->per_stream + 2 * n;
        sp.done = ss->per_stream + 3 * n;
        sp.failed = ss->per_stream + 4 * n;
        sp.next_count = counters + 5;
        sp.next_scan = counters + 6;
        sp.next_write = counters + 7;
        sp.min_bytes = ctx->inflate_split_min;
        sp.rec = rec_items ? ss->rec : nullptr;
        sp.rec_items = rec_items;
        split_item0 = sp.item0;
        FDB_LAUNCH(inflate_uf_plan_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0, st, b, (const UfDecTables*)ctx->d_dec, sp);
        FDB_LAUNCH(inflate_uf_split_count_kernel, dim3(sms), dim3(K4_WARPS * 32), sizeof(K4Smem), st, b,
                   (const UfDecTables*)ctx->d_dec, sp);
        FDB_LAUNCH(inflate_uf_split_scan_kernel, dim3((uint32_t)std::min<size_t>((n + 7) / 8, (size_t)sms * 4)), dim3(256), 0, st, b, sp,
                   *worklist, counters + 1);
        FDB_LAUNCH(inflate_uf_split_write_kernel, dim3(sms), dim3(K4_WARPS * 32), sizeof(K4Smem), st, b,
                   (const UfDecTables*)ctx->d_dec, sp, *worklist, counters + 1);
        ctx->launches += 4;
        FDB_TRY(cudaGetLastError());
    }
    if (!(b.flags & FDB_FLAG_GENERAL_ONLY)) {
        // one CTA per SM; with fewer streams than SMs the warps of n CTAs race for them, so a small
        // batch still spreads over the chip
        uint32_t grid = (uint32_t)std::min<size_t>(dense ? (n + K4_WARPS - 1) / K4_WARPS : n, (size_t)sms);
        FDB_LAUNCH(inflate_uf_kernel, dim3(grid), dim3(K4_WARPS * 32), sizeof(K4Smem), st, b, ctx->d_dec, counters + 0,
                   *worklist, counters + 1, split_item0);
        ctx->launches++;
        FDB_TRY(cudaGetLastError());
        uint32_t grid3 = (uint32_t)std::min<size_t>(n, (size_t)sms * 16);
        FDB_LAUNCH(inflate_general_kernel, dim3(grid3), dim3(32), sizeof(K3Smem), st, b, (const uint32_t*)*worklist,
                   (const uint32_t*)(counters + 1), counters + 2);
        ctx->launches++;
        FDB_TRY(cudaGetLastError());
    } else {
        uint32_t grid3 = (uint32_t)std::min<size_t>(n, (size_t)sms * 16);
        FDB_LAUNCH(inflate_general_kernel, dim3(grid3), dim3(32), sizeof(K3Smem), st, b, (const uint32_t*)nullptr,
                   (const uint32_t*)nullptr, counters + 2);
        ctx->launches++;
        FDB_TRY(cudaGetLastError());
    }
    return 0;
}

extern "C" int fdb_inflate_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                                        const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                                        const uint64_t* d_out_cap, uint64_t* d_out_len, uint64_t* d_consumed,
                                        int32_t* d_status, size_t n, uint32_t flags, void* cuda_stream) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !d_in_off || !d_in_len || !d_out_off || !d_out_cap || !d_out_len || !d_status)
        return fail(ctx, "fdb_inflate_batch_device", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    InflateBatch b;
    b.in_base = (const uint8_t*)d_in_base;
    b.in_off = d_in_off;
    b.in_len = d_in_len;
    b.out_base = (uint8_t*)d_out_base;
    b.out_off = d_out_off;
    b.out_cap = d_out_cap;
    b.out_len = d_out_len;
    b.consumed = d_consumed;
    b.status = d_status;
    b.n = (uint32_t)n;
    b.flags = flags | (ctx->split_large ? FDB_FLAG_SPLIT_LARGE : 0u);
    ctx->last_general_host = -1;
    ctx->last_split_host = -1;
    return launch_inflate(ctx, b, ctx->d_counters, &ctx->d_worklist, &ctx->worklist_cap, (cudaStream_t)cuda_stream, false,
                          &ctx->split);
}

// ---- deflate ----------------------------------------------------------------------------------
static int launch_deflate(fdb_ctx* ctx, int kind, const DeflateBatch& b, uint32_t* counter, cudaStream_t st,
                          bool dense = false, fdb_dsplit_scratch* ss = nullptr) {
    const size_t n = b.n;
    // counter = counters + 3; the segment path uses counters[8..11] of the same block
    FDB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    const uint32_t sms = (uint32_t)std::max(ctx->sm_count, 1);
    const uint32_t* split_item0 = nullptr;
    uint32_t* order = nullptr;
    if (kind == 0 && ss) {
        // long inputs, segment by segment: plan, count, scan, write (deflate_uf.cuh)
        int r;
        void* p = ss->items;
        if ((r = grow(ctx, &p, &ss->items_bytes, (size_t)FDB_SPLIT_ITEMS * sizeof(DfItem)))) return r;
        ss->items = (DfItem*)p;
        p = ss->per_stream;
        r = grow(ctx, &p, &ss->per_stream_bytes, (4 * n + 2 * DF_ORDER_CLASSES) * sizeof(uint32_t));
        ss->per_stream = (uint32_t*)p;
        if (r) return r;
        uint32_t* c8 = counter + 5;
        FDB_TRY(cudaMemsetAsync(c8, 0, 4 * sizeof(uint32_t), st));
        DfSplit sp;
        sp.items = ss->items;
        sp.item_cap = FDB_SPLIT_ITEMS;
        sp.n_items = c8;
        sp.item0 = ss->per_stream;
        sp.nseg = ss->per_stream + n;
        sp.adler = ss->per_stream + 2 * n;
        sp.next_count = c8 + 1;
        sp.next_scan = c8 + 2;
        sp.next_write = c8 + 3;
        sp.min_bytes = ctx->deflate_split_min;
        uint64_t* total = (uint64_t*)(counter + 15);  // counters[18..19]
        FDB_TRY(cudaMemsetAsync(total, 0, sizeof(uint64_t), st));
        FDB_LAUNCH(batch_total_kernel, dim3((uint32_t)std::min<size_t>((n + 255) / 256, 64)), dim3(256), 0, st, b.in_len, b.n, total);
        sp.total = total;
        sp.slots = sms * 32u;
        split_item0 = sp.item0;
        const uint32_t pgrid = sms * DEFLATE_MIN_CTAS;
        FDB_LAUNCH(deflate_uf_plan_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0, st, b, sp);
        FDB_LAUNCH(deflate_uf_split_count_kernel, dim3(pgrid), dim3(DEFLATE_WARPS * 32), 0, st, b, (const UfEncTables*)ctx->d_enc, sp);
        FDB_LAUNCH(deflate_uf_split_scan_kernel, dim3((uint32_t)std::min<size_t>((n + 7) / 8, (size_t)sms * 4)), dim3(256), 0, st, b, sp);
        FDB_LAUNCH(deflate_uf_split_write_kernel, dim3(pgrid), dim3(DEFLATE_WARPS * 32), 0, st, b, (const UfEncTables*)ctx->d_enc, sp);
        // the inputs that stay with one warp each are handed out longest first
        order = ss->per_stream + 3 * n;
        uint32_t* hist = ss->per_stream + 4 * n;
        FDB_TRY(cudaMemsetAsync(hist, 0, 2 * DF_ORDER_CLASSES * sizeof(uint32_t), st));
        const uint32_t ogrid = (uint32_t)std::min<size_t>((n + 255) / 256, 256);
        FDB_LAUNCH(order_hist_kernel, dim3(ogrid), dim3(256), 0, st, b.in_len, b.n, hist);
        FDB_LAUNCH(order_scatter_kernel, dim3(ogrid), dim3(256), 0, st, b.in_len, b.n, (const uint32_t*)hist, hist + DF_ORDER_CLASSES, order);
        ctx->launches += 7;
        FDB_TRY(cudaGetLastError());
    }
    if (kind == 0) {
        // persistent: every resident warp pulls streams from one counter, so SMs stay evenly loaded
        // even when the batch is not a multiple of the chip's warp slots
        // (FDB_DEFLATE_LANE16=1: the first-generation kernel, a lane per 16 bytes, which still encodes the segments)
        static const bool lane16 = [] { const char* e = getenv("FDB_DEFLATE_LANE16"); return e && atoi(e) != 0; }();
        if (lane16) {
            uint32_t grid = (uint32_t)std::min<size_t>(dense ? (n + DEFLATE_WARPS - 1) / DEFLATE_WARPS : n,
                                                       (size_t)sms * DEFLATE_MIN_CTAS);
            FDB_LAUNCH(deflate_uf_kernel, dim3(grid), dim3(DEFLATE_WARPS * 32), 0, st, b, (const UfEncTables*)ctx->d_enc,
                       counter, split_item0, (const uint32_t*)order);
        } else {
            uint32_t grid = (uint32_t)std::min<size_t>(dense ? (n + UB_WARPS - 1) / UB_WARPS : n, (size_t)sms * UB_MIN_CTAS);
            FDB_LAUNCH(deflate_ufb_kernel, dim3(grid), dim3(UB_WARPS * 32), sizeof(UbSmem), st, b, (const UfEncTables*)ctx->d_enc,
                       counter, split_item0, (const uint32_t*)order);
        }
    } else {
        uint32_t grid = (uint32_t)std::min<size_t>(n, (size_t)sms * 8);
        FDB_LAUNCH(deflate_stored_kernel, dim3(grid), dim3(STORED_THREADS), sizeof(StoredSmem), st, b, counter);
    }
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    return 0;
}

static int deflate_device(fdb_ctx* ctx, int kind, const void* d_in_base, const uint64_t* d_in_off,
                          const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                          const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_status, size_t n,
                          void* cuda_stream) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !d_in_off || !d_in_len || !d_out_off || !d_out_cap || !d_out_len || !d_status)
        return fail(ctx, "fdb_deflate_batch_device", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    DeflateBatch b;
    b.in_base = (const uint8_t*)d_in_base;
    b.in_off = d_in_off;
    b.in_len = d_in_len;
    b.out_base = (uint8_t*)d_out_base;
    b.out_off = d_out_off;
    b.out_cap = d_out_cap;
    b.out_len = d_out_len;
    b.status = d_status;
    b.n = (uint32_t)n;
    return launch_deflate(ctx, kind, b, ctx->d_counters + 3, (cudaStream_t)cuda_stream, false,
                          ctx->split_large ? &ctx->dsplit : nullptr);
}

extern "C" int fdb_deflate_ultrafast_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                                                  const uint64_t* d_in_len, void* d_out_base,
                                                  const uint64_t* d_out_off, const uint64_t* d_out_cap,
                                                  uint64_t* d_out_len, int32_t* d_status, size_t n,
                                                  void* cuda_stream) {
    return deflate_device(ctx, 0, d_in_base, d_in_off, d_in_len, d_out_base, d_out_off, d_out_cap, d_out_len, d_status,
                          n, cuda_stream);
}
extern "C" int fdb_deflate_stored_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                                               const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                                               const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_status,
                                               size_t n, void* cuda_stream) {
    return deflate_device(ctx, 1, d_in_base, d_in_off, d_in_len, d_out_base, d_out_off, d_out_cap, d_out_len, d_status,
                          n, cuda_stream);
}

extern "C" size_t fdb_deflate_ultrafast_bound(size_t n) {
    // header 53 bytes + 5 bits, <= 12 bits per input byte (run tokens are never longer than the
    // literals they replace), 12-bit EOB, pad, adler32; rounded up to 16 so slots stay aligned
    size_t bits = 53 * 8 + 5 + 12 * n + 12;
    return ((bits + 7) / 8 + 4 + 15) & ~(size_t)15;
}
extern "C" size_t fdb_deflate_stored_bound(size_t n) {
    return (2 + 5 * (n / 65535 + 1) + n + 4 + 15) & ~(size_t)15;
}

// ---- host-buffer entry points -----------------------------------------------------------------
struct Span {
    uint64_t in_span = 0, out_span = 0;
};
static Span spans(const uint64_t* in_off, const uint64_t* in_len, const uint64_t* out_off, const uint64_t* out_cap,
                  size_t n) {
    Span s;
    for (size_t i = 0; i < n; i++) {
        s.in_span = std::max(s.in_span, in_off[i] + in_len[i]);
        s.out_span = std::max(s.out_span, out_off[i] + out_cap[i]);
    }
    return s;
}

// If the n slots form an arithmetic progression (same stride, ascending) a single 2-D copy can move
// only the used prefix of every slot; otherwise the whole span is copied.
static bool uniform_stride(const uint64_t* off, size_t n, uint64_t* stride) {
    if (n < 2) return false;
    uint64_t s = off[1] - off[0];
    if (off[1] <= off[0]) return false;
    for (size_t i = 2; i < n; i++)
        if (off[i] - off[i - 1] != s || off[i] <= off[i - 1]) return false;
    *stride = s;
    return true;
}

// Copies the used part of rows [a, b) of a slot array between host and device.  ext[i] (clamped to cap[i] when
// given) is what row i holds.  Evenly spaced, mostly empty slots: one 2-D copy moves the widest row's width of every
// row BUT THE LAST (whose slot may end before that width does -- the caller's buffer ends with it), and a 1-D copy
// moves the last row's own extent.  Anything else: one copy per run of adjacent slots (up to 64 runs, which covers
// slots in any order), else one contiguous copy of the hull [min off, max off + ext).  Bytes between slots may be
// copied along; nothing outside the hull is touched.
// `exact` (device sets: the bytes between this call's slots may be another device's slots): a device-to-host copy
// writes nothing outside the slots themselves -- the 2-D copy only when every slot holds the widest row, the runs only
// over slots that touch, and as many runs as it takes.
static int copy_rows(fdb_ctx* ctx, uint8_t* dst_base, const uint8_t* src_base, const uint64_t* off, const uint64_t* ext,
                     const uint64_t* cap, size_t a, size_t b, bool uniform, uint64_t stride, cudaMemcpyKind kind,
                     cudaStream_t st, bool exact = false) {
    uint64_t lo = ~0ull, hi = 0, width = 0, min_cap = ~0ull;
    auto extent = [&](size_t i) { return cap ? std::min(ext[i], cap[i]) : ext[i]; };
    for (size_t i = a; i < b; i++) {
        if (cap) min_cap = std::min(min_cap, cap[i]);
        const uint64_t e = extent(i);
        if (!e) continue;
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + e);
        width = std::max(width, e);
    }
    if (hi <= lo) return 0;
    const size_t rows = b - a;
    const uint64_t span = hi - lo;
    const bool to_host = kind == cudaMemcpyDeviceToHost;
    if (exact && to_host) {
        if (uniform && rows > 2 && width <= stride && width <= min_cap && width * rows < span - span / 8) {
            FDB_TRY(cudaMemcpy2DAsync(dst_base + off[a], stride, src_base + off[a], stride, width, rows - 1, kind, st));
            const uint64_t e = extent(b - 1);
            if (e) FDB_TRY(cudaMemcpyAsync(dst_base + off[b - 1], src_base + off[b - 1], e, kind, st));
            return 0;
        }
        uint64_t p_lo = 0, p_hi = 0;
        for (size_t i = a; i < b; i++) {
            const uint64_t e = extent(i);
            if (!e) continue;
            if (p_hi > p_lo && off[i] == p_hi) {
                p_hi = off[i] + e;
                continue;
            }
            if (p_hi > p_lo) FDB_TRY(cudaMemcpyAsync(dst_base + p_lo, src_base + p_lo, p_hi - p_lo, kind, st));
            p_lo = off[i];
            p_hi = off[i] + e;
        }
        if (p_hi > p_lo) FDB_TRY(cudaMemcpyAsync(dst_base + p_lo, src_base + p_lo, p_hi - p_lo, kind, st));
        return 0;
    }
    if (uniform && rows > 2 && width <= stride && width * rows < span - span / 8) {
        FDB_TRY(cudaMemcpy2DAsync(dst_base + off[a], stride, src_base + off[a], stride, width, rows - 1, kind, st));
        const uint64_t e = extent(b - 1);
        if (e) FDB_TRY(cudaMemcpyAsync(dst_base + off[b - 1], src_base + off[b - 1], e, kind, st));
    } else {
        // runs of slots that follow each other (ascending, less than 256 bytes apart) travel in one copy each; if that
        // makes too many copies (widely and unevenly spaced slots) the whole hull goes in one
        struct Piece {
            uint64_t lo, hi;
        };
        Piece pieces[64];
        size_t np = 0;
        bool ok = true;
        for (size_t i = a; i < b && ok; i++) {
            const uint64_t e = extent(i);
            if (!e) continue;
            if (np && off[i] >= pieces[np - 1].hi && off[i] - pieces[np - 1].hi < 256)
                pieces[np - 1].hi = off[i] + e;
            else if (np == 64)
                ok = false;
            else
                pieces[np++] = {off[i], off[i] + e};
        }
        if (!ok) {
            FDB_TRY(cudaMemcpyAsync(dst_base + lo, src_base + lo, span, kind, st));
        } else {
            for (size_t k = 0; k < np; k++)
                FDB_TRY(cudaMemcpyAsync(dst_base + pieces[k].lo, src_base + pieces[k].lo, pieces[k].hi - pieces[k].lo, kind, st));
        }
    }
    return 0;
}

// Host-buffer batches run as a pipeline of chunks.  Every chunk's input is queued at once, in order, on
// one host-to-device stream (the staging buffers hold the whole batch, so nothing waits for a slot); each
// chunk's kernels run on one of n_lanes compute streams as soon as its input has landed, followed by a
// small copy of its per-stream results; the host picks the results up in order and they decide how much
// payload goes back on the device-to-host stream.  Both directions of the (full duplex) link therefore
// stay busy from the first chunk to the last, and a call costs about max(H2D, D2H) + one chunk.
static int png_launch(fdb_ctx* ctx, bool unfilter, const void* d_in_base, const uint64_t* d_in_off, void* d_out_base,
                      const uint64_t* d_out_off, const uint32_t* d_height, const uint32_t* d_stride, const uint32_t* d_bpp,
                      uint32_t mode, int32_t* d_status, size_t n, void* cuda_stream, uint32_t* counter = nullptr);

static int crc_launch(fdb_ctx* ctx, const void* d_base, const uint64_t* d_off, const uint64_t* d_len, uint32_t seed,
                      uint32_t* d_crc, size_t n, void* cuda_stream, uint32_t* counter);
static int launch_deflate_png(fdb_ctx* ctx, const DeflateBatch& b, const uint32_t* d_height, const uint32_t* d_stride,
                              const uint32_t* d_bpp, uint32_t mode, int32_t* d_fstatus, uint32_t* counter, cudaStream_t st,
                              bool dense);

// PNG encode rides the same pipeline: the inputs are raw images, and between "input landed" and the deflate
// kernels a filter kernel writes the filtered images into a second device buffer, which is what gets compressed.
struct PngPre {
    const uint32_t *height, *stride, *bpp;
    uint32_t mode;
    int32_t* filter_status;  // [n] host array
    uint32_t* idat_crc;      // [n] host array or null: CRC-32 of "IDAT" + the compressed stream, computed on the device
};

static int host_batch(fdb_ctx* ctx, int kind, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                      uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                      uint64_t* consumed, int32_t* status, size_t n, uint32_t flags, const PngPre* png = nullptr,
                      bool exact = false) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !in_off || !in_len || !out_off || !out_cap || !out_len || !status)
        return fail(ctx, "fdb_*_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    Span sp = spans(in_off, in_len, out_off, out_cap, n);
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, sp.in_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, sp.out_span + 64))) return r;
    const size_t meta_words = 8 * n;
    if ((r = grow(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, meta_words * sizeof(uint64_t)))) return r;
    uint64_t* m = ctx->d_meta;
    uint64_t *d_in_off = m, *d_in_len = m + n, *d_out_off = m + 2 * n, *d_out_cap = m + 3 * n;
    // PNG encode: filtered_off | filtered_len | height, stride, bpp (u32) | filter status (i32)
    uint64_t *d_filt_off = m + 4 * n, *d_filt_len = m + 5 * n;
    uint32_t* d_geo = (uint32_t*)(m + 6 * n);
    int32_t* d_fst = (int32_t*)(d_geo + 3 * n);
    std::vector<uint64_t> filt;
    uint64_t max_filtered = 0;
    bool png_fused = false;
    if (png) {
        filt.resize(2 * n);
        uint64_t span = 0;
        for (size_t i = 0; i < n; i++) {
            const uint64_t f = (uint64_t)png->height[i] * (1ull + png->stride[i]);
            filt[i] = span;
            filt[n + i] = f;
            max_filtered = std::max(max_filtered, f);
            span += (f + 15) & ~15ull;
        }
        // (nothing is staged when the encoder computes the filtered bytes itself, see `fused` below)
        png_fused = ctx->png_fused && kind == 1 && png->mode <= 4 && max_filtered <= 0xffffffffull && max_filtered < ctx->deflate_auto_min;
        if (!png_fused && (r = grow(ctx, (void**)&ctx->d_mid, &ctx->d_mid_cap, span + 64))) return r;
    }

    // chunking needs slots laid out in ascending order (what every packer produces); anything else is one chunk
    bool ascending = true;
    for (size_t i = 1; i < n && ascending; i++)
        ascending = in_off[i] >= in_off[i - 1] + in_len[i - 1] && out_off[i] >= out_off[i - 1] + out_cap[i - 1];
    size_t nchunk = 1;
    if (ascending)
        nchunk = (size_t)std::min<uint64_t>(std::min<uint64_t>((sp.in_span + sp.out_span) / ctx->chunk_bytes, FDB_MAX_CHUNKS), n);
    if (nchunk < 1) nchunk = 1;
    const size_t per = (n + nchunk - 1) / nchunk;
    nchunk = (n + per - 1) / per;

    std::vector<fdb_trace_chunk> trace;
    if (trace_path()) {
        if (!g_trace_base) {
            FDB_TRY(cudaEventCreate(&g_trace_base));
            FDB_TRY(cudaEventRecord(g_trace_base, ctx->h2d_st));
            FDB_TRY(cudaEventSynchronize(g_trace_base));
        }
        trace.resize(nchunk);
        for (auto& tc : trace)
            for (auto& e : tc.ev) FDB_TRY(cudaEventCreate(&e));
    }
    auto mark = [&](size_t k, int which, cudaStream_t st) {
        if (!trace.empty()) cudaEventRecord(trace[k].ev[which], st);
    };

    // The per-stream results live in pinned host memory and the kernels write them there directly (mapped
    // memory: a few bytes per stream over PCIe), so no copy stands between "kernels done" and the host
    // knowing how much payload to fetch.
    const size_t res_bytes = n * (8 + 8 + 4 + 4) + nchunk * 8 + 64;
    if (res_bytes > ctx->h_res_cap) {
        if (ctx->h_res) FDB_TRY(cudaFreeHost(ctx->h_res));
        ctx->h_res = nullptr;
        ctx->h_res_cap = 0;
        FDB_TRY(cudaMallocHost((void**)&ctx->h_res, res_bytes + res_bytes / 4));
        ctx->h_res_cap = res_bytes + res_bytes / 4;
    }
    uint64_t* h_out_len = (uint64_t*)ctx->h_res;
    uint64_t* h_consumed = h_out_len + n;
    int32_t* h_status = (int32_t*)(h_consumed + n);
    uint32_t* h_general = (uint32_t*)(h_status + n);
    uint32_t* h_split = h_general + nchunk;
    uint32_t* h_crc = h_split + nchunk;  // PNG encode: chunk CRCs, written by the device like the other results
    memset(h_split, 0, nchunk * sizeof(uint32_t));

    uint64_t in_stride = 0, out_stride = 0;
    const bool in_uniform = uniform_stride(in_off, n, &in_stride), out_uniform = uniform_stride(out_off, n, &out_stride);

    // stream descriptors: once, ahead of every chunk's input on the same stream
    const int L = ctx->n_lanes;
    cudaStream_t hs = ctx->h2d_st, ds = ctx->d2h_st;
    FDB_TRY(cudaMemcpyAsync(d_in_off, in_off, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_in_len, in_len, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_out_off, out_off, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_out_cap, out_cap, n * 8, cudaMemcpyHostToDevice, hs));
    if (png) {
        FDB_TRY(cudaMemcpyAsync(d_filt_off, filt.data(), 2 * n * 8, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaMemcpyAsync(d_geo, png->height, n * 4, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaMemcpyAsync(d_geo + n, png->stride, n * 4, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaMemcpyAsync(d_geo + 2 * n, png->bpp, n * 4, cudaMemcpyHostToDevice, hs));
    }
    const bool dense = nchunk > 1;

    // Ultra-fast deflate into a PINNED host buffer: the kernel stores its output words straight into the caller's
    // slots (mapped host memory, coalesced 128-byte stores over PCIe) -- exactly the bytes each stream produced, where
    // the 2-D payload copy moves the longest stream's width for every row (a quarter more on the bench tiles), and no
    // payload stage behind the kernels at all.  The kernel never reads its output back.  Chunks that take the segment
    // path (long inputs) and everything else use the device buffer and the copy.  OPT-IN (FDB_DIRECT_OUT=1): on the bench
    // step the stores of a chunk's few SMs over PCIe are slower than the copy engine, over-fetch included (e2e 51-55
    // against 57.5 GB/s, profiles/r03_e2e_direct_out.txt).
    uint8_t* direct_base = nullptr;
    if (kind == 1 && !png && ctx->direct_out && sp.out_span) {
        cudaPointerAttributes at0, at1;
        if (cudaPointerGetAttributes(&at0, out_base) == cudaSuccess && at0.type == cudaMemoryTypeHost && at0.devicePointer &&
            cudaPointerGetAttributes(&at1, out_base + sp.out_span - 1) == cudaSuccess && at1.type == cudaMemoryTypeHost &&
            (uint8_t*)at1.devicePointer == (uint8_t*)at0.devicePointer + (sp.out_span - 1))
            direct_base = (uint8_t*)at0.devicePointer;
        else
            cudaGetLastError();  // (older runtimes report unregistered memory as an error)
    }
    std::vector<uint8_t> direct_chunk(nchunk, 0);

    auto issue = [&](size_t k) -> int {  // input, kernels and results of chunk k
        const size_t a = k * per, b = std::min(n, a + per);
        fdb_lane& ln = ctx->lanes[k % L];
        uint64_t max_in = 0, sum_in = 0;
        for (size_t i = a; i < b; i++) {
            max_in = std::max(max_in, in_len[i]);
            sum_in += in_len[i];
        }
        mark(k, 0, hs);
        int rr = copy_rows(ctx, ctx->d_in, in_base, in_off, in_len, nullptr, a, b, in_uniform, in_stride, cudaMemcpyHostToDevice, hs);
        if (rr) return rr;
        mark(k, 1, hs);
        FDB_TRY(cudaEventRecord(ctx->ev_in[k], hs));
        FDB_TRY(cudaStreamWaitEvent(ln.st, ctx->ev_in[k], 0));
        if (kind == 0) {
            InflateBatch ib;
            ib.in_base = ctx->d_in;
            ib.in_off = d_in_off + a;
            ib.in_len = d_in_len + a;
            ib.out_base = ctx->d_out;
            ib.out_off = d_out_off + a;
            ib.out_cap = d_out_cap + a;
            ib.out_len = h_out_len + a;
            ib.consumed = h_consumed + a;
            ib.status = h_status + a;
            ib.n = (uint32_t)(b - a);
            ib.flags = flags;
            if (max_in >= ctx->inflate_split_min) ib.flags |= FDB_FLAG_SPLIT_LARGE;  // long streams: many warps each
            ib.general_out = h_general + k;
            ib.split_out = h_split + k;
            if ((rr = launch_inflate(ctx, ib, ln.d_counters, &ln.d_worklist, &ln.worklist_cap, ln.st, dense, &ln.split, kind == 0 ? sum_in : 0)))
                return rr;
        } else {
            DeflateBatch db;
            db.in_base = ctx->d_in;
            db.in_off = d_in_off + a;
            db.in_len = d_in_len + a;
            // PNG encode with one filter type for all rows: the encoder computes the filtered bytes itself from the raw
            // image (deflate_png.cuh) -- no filtered image in device memory.  The adaptive filter (mode 5), images whose
            // filtered size does not fit 32 bits and chunks that take the segment path keep the filter kernel.
            const bool fused = png_fused;
            if (png && !fused) {
                if ((rr = png_launch(ctx, false, ctx->d_in, d_in_off + a, ctx->d_mid, d_filt_off + a, d_geo + a, d_geo + n + a,
                                     d_geo + 2 * n + a, png->mode, d_fst + a, b - a, ln.st, ln.d_counters + 12)))
                    return rr;
                FDB_TRY(cudaMemcpyAsync(png->filter_status + a, d_fst + a, (b - a) * 4, cudaMemcpyDeviceToHost, ln.st));
                db.in_base = ctx->d_mid;
                db.in_off = d_filt_off + a;
                db.in_len = d_filt_len + a;
            }
            db.out_base = ctx->d_out;
            db.out_off = d_out_off + a;
            db.out_cap = d_out_cap + a;
            db.out_len = h_out_len + a;
            db.status = h_status + a;
            db.n = (uint32_t)(b - a);
            uint64_t max_len = png ? max_filtered : 0;
            for (size_t i = a; i < b; i++) max_len = std::max(max_len, in_len[i]);
            if (direct_base && max_len < ctx->deflate_auto_min) {
                db.out_base = direct_base;
                direct_chunk[k] = 1;
            }
            if (fused) {
                if ((rr = launch_deflate_png(ctx, db, d_geo + a, d_geo + n + a, d_geo + 2 * n + a, png->mode, d_fst + a,
                                             ln.d_counters + 3, ln.st, dense)))
                    return rr;
                FDB_TRY(cudaMemcpyAsync(png->filter_status + a, d_fst + a, (b - a) * 4, cudaMemcpyDeviceToHost, ln.st));
            } else if ((rr = launch_deflate(ctx, kind - 1, db, ln.d_counters + 3, ln.st, dense,
                                            max_len >= ctx->deflate_auto_min ? &ln.dsplit : nullptr)))
                return rr;
            // the CRC of each IDAT chunk ("IDAT" + stream): the stream lengths are read where the deflate kernel
            // wrote them (mapped host memory)
            if (png && png->idat_crc &&
                (rr = crc_launch(ctx, ctx->d_out, d_out_off + a, h_out_len + a, 0x35af061eu /* crc32("IDAT") */, h_crc + a, b - a,
                                 ln.st, ln.d_counters + 13)))
                return rr;
        }
        mark(k, 2, ln.st);
        FDB_TRY(cudaEventRecord(ctx->ev_res[k], ln.st));
        return 0;
    };
    auto finish = [&](size_t k) -> int {  // the results (already on the host) decide how much payload comes back
        const size_t a = k * per, b = std::min(n, a + per);
        mark(k, 3, ds);
        if (direct_chunk[k]) {  // the kernel has written the caller's slots itself
            mark(k, 4, ds);
            return 0;
        }
        int rr = copy_rows(ctx, out_base, ctx->d_out, out_off, h_out_len, out_cap, a, b, out_uniform, out_stride,
                           cudaMemcpyDeviceToHost, ds, exact);
        mark(k, 4, ds);
        return rr;
    };
    // The host keeps at most `depth` chunks between "input queued" and "payload queued": enough to cover the
    // kernels' latency, few enough that another context's copies are never stuck behind a long queue of ours
    // (copies of one direction execute in the order they were queued, whatever their stream).
    const size_t depth = (size_t)ctx->depth;
    size_t issued = 0, fin = 0;  // chunks [0, fin) have their payload on the way back
    while (fin < nchunk) {
        while (fin < issued && cudaEventQuery(ctx->ev_res[fin]) == cudaSuccess)
            if ((r = finish(fin++))) return r;
        if (issued < nchunk && issued < fin + depth) {
            if ((r = issue(issued++))) return r;
            continue;
        }
        if (fin < issued) {
            FDB_TRY(cudaEventSynchronize(ctx->ev_res[fin]));
            if ((r = finish(fin++))) return r;
        }
    }
    FDB_TRY(cudaStreamSynchronize(ds));
    FDB_TRY(cudaStreamSynchronize(hs));
    for (int l = 0; l < L; l++) FDB_TRY(cudaStreamSynchronize(ctx->lanes[l].st));
    if (!trace.empty()) {
        if (FILE* f = fopen(trace_path(), "a")) {
            for (size_t k = 0; k < nchunk; k++) {
                float t[5] = {0, 0, 0, 0, 0};
                for (int j = 0; j < 5; j++) cudaEventElapsedTime(&t[j], g_trace_base, trace[k].ev[j]);
                fprintf(f, "ctx %p kind %d chunk %zu/%zu lane %zu streams %zu h2d %.3f %.3f kern_end %.3f d2h %.3f %.3f\n",
                        (void*)ctx, kind, k, nchunk, k % (size_t)L, std::min(n, (k + 1) * per) - k * per, t[0], t[1], t[2],
                        t[3], t[4]);
            }
            fclose(f);
        }
        for (auto& tc : trace)
            for (auto& e : tc.ev) cudaEventDestroy(e);
    }
    memcpy(out_len, h_out_len, n * 8);
    memcpy(status, h_status, n * 4);
    if (png && png->idat_crc) memcpy(png->idat_crc, h_crc, n * 4);
    if (kind == 0) {
        if (consumed) memcpy(consumed, h_consumed, n * 8);
        int64_t g = 0;
        for (size_t k = 0; k < nchunk; k++) g += h_general[k];
        ctx->last_general_host = (flags & FDB_FLAG_GENERAL_ONLY) ? 0 : g;
        int64_t sp_total = 0;
        for (size_t k = 0; k < nchunk; k++) sp_total += h_split[k];
        ctx->last_split_host = sp_total;
    }
    return 0;
}

extern "C" int fdb_inflate_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                 uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                 uint64_t* out_len, uint64_t* consumed, int32_t* status, size_t n, uint32_t flags) {
    return host_batch(ctx, 0, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, consumed, status, n, flags);
}
extern "C" int fdb_deflate_ultrafast_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off,
                                           const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off,
                                           const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n) {
    return host_batch(ctx, 1, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, nullptr, status, n, 0);
}
extern "C" int fdb_deflate_stored_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off,
                                        const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off,
                                        const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n) {
    return host_batch(ctx, 2, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, nullptr, status, n, 0);
}

// ---- PNG row filters --------------------------------------------------------------------------
static int png_launch(fdb_ctx* ctx, bool unfilter, const void* d_in_base, const uint64_t* d_in_off, void* d_out_base,
                      const uint64_t* d_out_off, const uint32_t* d_height, const uint32_t* d_stride, const uint32_t* d_bpp,
                      uint32_t mode, int32_t* d_status, size_t n, void* cuda_stream, uint32_t* counter) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !d_in_off || !d_out_off || !d_height || !d_stride || !d_bpp || !d_status)
        return fail(ctx, "fdb_png_*_batch_device", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    PngBatch b;
    b.in_base = (const uint8_t*)d_in_base;
    b.in_off = d_in_off;
    b.out_base = (uint8_t*)d_out_base;
    b.out_off = d_out_off;
    b.height = d_height;
    b.stride = d_stride;
    b.bpp = d_bpp;
    b.status = d_status;
    b.n = (uint32_t)n;
    b.mode = mode;
    if (!counter) counter = ctx->d_counters + 12;
    FDB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    const uint32_t sms = (uint32_t)std::max(ctx->sm_count, 1);
    if (unfilter) {
        // a warp per image, persistent
        uint32_t grid = (uint32_t)std::min<size_t>((n + PNG_UNFILTER_WARPS - 1) / PNG_UNFILTER_WARPS, (size_t)sms * 6);
        FDB_LAUNCH(png_unfilter_kernel, dim3(grid), dim3(PNG_UNFILTER_WARPS * 32), 0, st, b, counter);
    } else {
        uint32_t grid = (uint32_t)std::min<size_t>(n, (size_t)sms * 4);
        FDB_LAUNCH(png_filter_kernel, dim3(grid), dim3(PNG_FILTER_WARPS * 32), 0, st, b, counter);
    }
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    return 0;
}
extern "C" int fdb_png_unfilter_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off, void* d_out_base,
                                             const uint64_t* d_out_off, const uint32_t* d_height, const uint32_t* d_stride,
                                             const uint32_t* d_bpp, int32_t* d_status, size_t n, void* cuda_stream) {
    return png_launch(ctx, true, d_in_base, d_in_off, d_out_base, d_out_off, d_height, d_stride, d_bpp, 0, d_status, n,
                      cuda_stream);
}
extern "C" int fdb_png_filter_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off, void* d_out_base,
                                           const uint64_t* d_out_off, const uint32_t* d_height, const uint32_t* d_stride,
                                           const uint32_t* d_bpp, uint32_t mode, int32_t* d_status, size_t n,
                                           void* cuda_stream) {
    return png_launch(ctx, false, d_in_base, d_in_off, d_out_base, d_out_off, d_height, d_stride, d_bpp, mode, d_status, n,
                      cuda_stream);
}

// filter fused into the encoder (deflate_png.cuh)
static int launch_deflate_png(fdb_ctx* ctx, const DeflateBatch& b, const uint32_t* d_height, const uint32_t* d_stride,
                              const uint32_t* d_bpp, uint32_t mode, int32_t* d_fstatus, uint32_t* counter, cudaStream_t st,
                              bool dense) {
    const uint32_t sms = (uint32_t)std::max(ctx->sm_count, 1);
    const size_t n = b.n;
    const uint32_t grid = (uint32_t)std::min<size_t>(dense ? (n + UB_WARPS - 1) / UB_WARPS : n, (size_t)sms * UB_MIN_CTAS);
    FDB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    FDB_LAUNCH(deflate_ufb_png_kernel, dim3(grid), dim3(UB_WARPS * 32), sizeof(UbSmem), st, b, d_height, d_stride, d_bpp, mode,
               d_fstatus, (const UfEncTables*)ctx->d_enc, counter);
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    return 0;
}
extern "C" int fdb_png_encode_batch_device(fdb_ctx* ctx, const void* d_raw_base, const uint64_t* d_raw_off,
                                           const uint32_t* d_height, const uint32_t* d_stride, const uint32_t* d_bpp,
                                           uint32_t mode, void* d_out_base, const uint64_t* d_out_off,
                                           const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_filter_status,
                                           int32_t* d_status, size_t n, void* cuda_stream) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !d_raw_off || !d_height || !d_stride || !d_bpp || !d_out_off || !d_out_cap || !d_out_len ||
        !d_filter_status || !d_status)
        return fail(ctx, "fdb_png_encode_batch_device", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    DeflateBatch b;
    b.in_base = (const uint8_t*)d_raw_base;
    b.in_off = d_raw_off;
    b.in_len = nullptr;  // (the kernel derives the lengths from the geometry)
    b.out_base = (uint8_t*)d_out_base;
    b.out_off = d_out_off;
    b.out_cap = d_out_cap;
    b.out_len = d_out_len;
    b.status = d_status;
    b.n = (uint32_t)n;
    return launch_deflate_png(ctx, b, d_height, d_stride, d_bpp, mode, d_filter_status, ctx->d_counters + 3,
                              (cudaStream_t)cuda_stream, false);
}

// host-buffer variants: stage through the context's device buffers on its first lane (one copy up, the kernel,
// one copy back; these calls are not pipelined)
static int png_host(fdb_ctx* ctx, bool unfilter, const uint8_t* in_base, const uint64_t* in_off, uint8_t* out_base,
                    const uint64_t* out_off, const uint32_t* height, const uint32_t* stride, const uint32_t* bpp, uint32_t mode,
                    int32_t* status, size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !in_off || !out_off || !height || !stride || !bpp || !status)
        return fail(ctx, "fdb_png_*_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    uint64_t in_span = 0, out_span = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t filtered = (uint64_t)height[i] * (1ull + stride[i]), raw = (uint64_t)height[i] * stride[i];
        in_span = std::max(in_span, in_off[i] + (unfilter ? filtered : raw));
        out_span = std::max(out_span, out_off[i] + (unfilter ? raw : filtered));
    }
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, in_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, out_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, 5 * n * sizeof(uint64_t)))) return r;
    cudaStream_t st = ctx->lanes[0].st;
    uint64_t* d_in_off = ctx->d_meta;
    uint64_t* d_out_off = d_in_off + n;
    uint32_t* d_h = (uint32_t*)(d_out_off + n);
    uint32_t* d_s = d_h + n;
    uint32_t* d_b = d_s + n;
    int32_t* d_status = (int32_t*)(d_b + n);
    FDB_TRY(cudaMemcpyAsync(d_in_off, in_off, n * 8, cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(d_out_off, out_off, n * 8, cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(d_h, height, n * 4, cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(d_s, stride, n * 4, cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(d_b, bpp, n * 4, cudaMemcpyHostToDevice, st));
    if (in_span) FDB_TRY(cudaMemcpyAsync(ctx->d_in, in_base, in_span, cudaMemcpyHostToDevice, st));
    if ((r = png_launch(ctx, unfilter, ctx->d_in, d_in_off, ctx->d_out, d_out_off, d_h, d_s, d_b, mode, d_status, n, st))) return r;
    if (out_span) FDB_TRY(cudaMemcpyAsync(out_base, ctx->d_out, out_span, cudaMemcpyDeviceToHost, st));
    FDB_TRY(cudaMemcpyAsync(status, d_status, n * 4, cudaMemcpyDeviceToHost, st));
    FDB_TRY(cudaStreamSynchronize(st));
    return 0;
}
extern "C" int fdb_png_unfilter_batch(fdb_ctx* ctx, const uint8_t* filtered_base, const uint64_t* filtered_off,
                                      uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* height,
                                      const uint32_t* stride, const uint32_t* bpp, int32_t* status, size_t n) {
    return png_host(ctx, true, filtered_base, filtered_off, raw_base, raw_off, height, stride, bpp, 0, status, n);
}
extern "C" int fdb_png_filter_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, uint8_t* filtered_base,
                                    const uint64_t* filtered_off, const uint32_t* height, const uint32_t* stride,
                                    const uint32_t* bpp, uint32_t mode, int32_t* status, size_t n) {
    return png_host(ctx, false, raw_base, raw_off, filtered_base, filtered_off, height, stride, bpp, mode, status, n);
}

// ---- PNG image data: zlib stream <-> raw pixels, the intermediate filtered image never leaves the device ----
// decode: IDAT payloads up, inflate (fast path / spans / general kernel as the streams require) into a device
// buffer of filtered images, unfilter into a second one, raw pixels back.  status = the inflate status if it is
// not Ok, else the unfilter status; a stream that inflates to anything but height * (1 + stride) bytes is
// reported as InsufficientInput / OutputTooLarge by the inflate step (slot = exactly that size).
extern "C" int fdb_png_decode_batch(fdb_ctx* ctx, const uint8_t* idat_base, const uint64_t* idat_off, const uint64_t* idat_len,
                                    uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* height, const uint32_t* stride,
                                    const uint32_t* bpp, int32_t* status, size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !idat_off || !idat_len || !raw_off || !height || !stride || !bpp || !status)
        return fail(ctx, "fdb_png_decode_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    std::vector<uint64_t> m(2 * n);  // filtered_off | filtered_cap
    uint64_t in_span = 0, raw_span = 0, filt_span = 0;
    bool ascending = true;
    for (size_t i = 0; i < n; i++) {
        const uint64_t filtered = (uint64_t)height[i] * (1ull + stride[i]);
        in_span = std::max(in_span, idat_off[i] + idat_len[i]);
        raw_span = std::max(raw_span, raw_off[i] + (uint64_t)height[i] * stride[i]);
        if (i && (idat_off[i] < idat_off[i - 1] + idat_len[i - 1] ||
                  raw_off[i] < raw_off[i - 1] + (uint64_t)height[i - 1] * stride[i - 1]))
            ascending = false;
        m[i] = filt_span;
        m[n + i] = filtered;
        filt_span += (filtered + 15) & ~15ull;
    }
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, in_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, raw_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_mid, &ctx->d_mid_cap, filt_span + 64))) return r;
    // idat_off | idat_len | filt_off | filt_cap | raw_off | consumed | h,s,b (u32)
    if ((r = grow(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, 8 * n * sizeof(uint64_t)))) return r;
    // per-image results, written by the kernels straight into pinned host memory (as in host_batch)
    const size_t res_bytes = n * (8 + 4 + 4) + 64;
    if (res_bytes > ctx->h_res_cap) {
        if (ctx->h_res) FDB_TRY(cudaFreeHost(ctx->h_res));
        ctx->h_res = nullptr;
        ctx->h_res_cap = 0;
        FDB_TRY(cudaMallocHost((void**)&ctx->h_res, res_bytes + res_bytes / 4));
        ctx->h_res_cap = res_bytes + res_bytes / 4;
    }
    uint64_t* h_out_len = (uint64_t*)ctx->h_res;
    int32_t* h_st1 = (int32_t*)(h_out_len + n);
    int32_t* h_st2 = h_st1 + n;
    cudaStream_t hs = ctx->h2d_st, ds = ctx->d2h_st;
    uint64_t* d = ctx->d_meta;
    uint64_t *d_idat_off = d, *d_idat_len = d + n, *d_filt_off = d + 2 * n, *d_filt_cap = d + 3 * n, *d_raw_off = d + 4 * n,
             *d_consumed = d + 5 * n;
    uint32_t* d_h = (uint32_t*)(d + 6 * n);
    uint32_t *d_s = d_h + n, *d_b = d_s + n;
    FDB_TRY(cudaMemcpyAsync(d_idat_off, idat_off, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_idat_len, idat_len, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_filt_off, m.data(), 2 * n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_raw_off, raw_off, n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_h, height, n * 4, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_s, stride, n * 4, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_b, bpp, n * 4, cudaMemcpyHostToDevice, hs));
    // chunks of streams: streams up | inflate + unfilter | pixels back overlap.  Nothing the host does depends on
    // a chunk's results (the pixel ranges are known from the geometry), so every chunk is queued at once.
    size_t nchunk = 1;
    if (ascending) nchunk = (size_t)std::min<uint64_t>(std::min<uint64_t>((in_span + raw_span) / ctx->chunk_bytes, FDB_MAX_CHUNKS), n);
    if (nchunk < 1) nchunk = 1;
    const size_t per = (n + nchunk - 1) / nchunk;
    nchunk = (n + per - 1) / per;
    const int L = ctx->n_lanes;
    ctx->last_general_host = -1;
    ctx->last_split_host = -1;
    for (size_t k = 0; k < nchunk; k++) {
        const size_t a = k * per, b = std::min(n, a + per);
        fdb_lane& ln = ctx->lanes[k % L];
        uint64_t max_in = 0, sum_in = 0;
        for (size_t i = a; i < b; i++) {
            max_in = std::max(max_in, idat_len[i]);
            sum_in += idat_len[i];
        }
        const uint64_t in_lo = nchunk == 1 ? 0 : idat_off[a], in_hi = nchunk == 1 ? in_span : idat_off[b - 1] + idat_len[b - 1];
        if (in_hi > in_lo) FDB_TRY(cudaMemcpyAsync(ctx->d_in + in_lo, idat_base + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaEventRecord(ctx->ev_in[k], hs));
        FDB_TRY(cudaStreamWaitEvent(ln.st, ctx->ev_in[k], 0));
        InflateBatch ib;
        ib.in_base = ctx->d_in;
        ib.in_off = d_idat_off + a;
        ib.in_len = d_idat_len + a;
        ib.out_base = ctx->d_mid;
        ib.out_off = d_filt_off + a;
        ib.out_cap = d_filt_cap + a;
        ib.out_len = h_out_len + a;
        ib.consumed = d_consumed + a;
        ib.status = h_st1 + a;
        ib.n = (uint32_t)(b - a);
        ib.flags = max_in >= ctx->inflate_split_min ? FDB_FLAG_SPLIT_LARGE : 0u;
        if ((r = launch_inflate(ctx, ib, ln.d_counters, &ln.d_worklist, &ln.worklist_cap, ln.st, nchunk > 1, &ln.split, sum_in))) return r;
        if ((r = png_launch(ctx, true, ctx->d_mid, d_filt_off + a, ctx->d_out, d_raw_off + a, d_h + a, d_s + a, d_b + a, 0,
                            h_st2 + a, b - a, ln.st, ln.d_counters + 12)))
            return r;
        FDB_TRY(cudaEventRecord(ctx->ev_res[k], ln.st));
        FDB_TRY(cudaStreamWaitEvent(ds, ctx->ev_res[k], 0));
        const uint64_t out_lo = nchunk == 1 ? 0 : raw_off[a];
        const uint64_t out_hi = nchunk == 1 ? raw_span : raw_off[b - 1] + (uint64_t)height[b - 1] * stride[b - 1];
        if (out_hi > out_lo) FDB_TRY(cudaMemcpyAsync(raw_base + out_lo, ctx->d_out + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, ds));
    }
    FDB_TRY(cudaStreamSynchronize(ds));
    FDB_TRY(cudaStreamSynchronize(hs));
    for (int l = 0; l < L; l++) FDB_TRY(cudaStreamSynchronize(ctx->lanes[l].st));
    for (size_t i = 0; i < n; i++) {
        int32_t s1 = h_st1[i];
        if (s1 == ST_OK && h_out_len[i] != m[n + i]) s1 = ST_INSUFFICIENT_INPUT;  // the stream ended before the image was complete
        status[i] = s1 != ST_OK ? s1 : h_st2[i];
    }
    return 0;
}

// encode: raw pixels up, filter, ultra-fast deflate, zlib streams back (slots of fdb_deflate_ultrafast_bound), on
// the pipeline of the host-buffer deflate call.
extern "C" int fdb_png_encode_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* height,
                                    const uint32_t* stride, const uint32_t* bpp, uint32_t mode, uint8_t* out_base,
                                    const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status,
                                    size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !raw_off || !height || !stride || !bpp || !out_off || !out_cap || !out_len || !status)
        return fail(ctx, "fdb_png_encode_batch", cudaSuccess);
    std::vector<uint64_t> raw_len(n);
    std::vector<int32_t> fst(n, 0);
    for (size_t i = 0; i < n; i++) raw_len[i] = (uint64_t)height[i] * stride[i];
    PngPre pre = {height, stride, bpp, mode, fst.data(), nullptr};
    int r = host_batch(ctx, 1, raw_base, raw_off, raw_len.data(), out_base, out_off, out_cap, out_len, nullptr, status, n, 0, &pre);
    if (r) return r;
    for (size_t i = 0; i < n; i++)
        if (fst[i] != ST_OK) {
            status[i] = fst[i];
            out_len[i] = 0;
        }
    return 0;
}

// raw pixels -> PNG files (signature, IHDR, one IDAT chunk, IEND) written straight into the caller's file slots:
// the deflate kernel writes the zlib stream at its place inside the file, the IDAT CRC comes from the device, the 57
// bytes around it (signature, IHDR chunk, IDAT length / type / CRC, IEND chunk) are written by the host.
static uint32_t host_crc32(const uint8_t* p, size_t n) {
    static CrcTables* t = [] {
        CrcTables* c = new CrcTables();
        build_crc_tables(*c);
        return c;
    }();
    uint32_t r = 0xffffffffu;
    for (size_t i = 0; i < n; i++) r = (r >> 8) ^ t->t[0][(r ^ p[i]) & 0xffu];
    return ~r;
}
static void put_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24);
    p[1] = (uint8_t)(v >> 16);
    p[2] = (uint8_t)(v >> 8);
    p[3] = (uint8_t)v;
}
static const size_t PNG_PRE = 8 + 25 + 8;   // signature + IHDR chunk + IDAT length and type
static const size_t PNG_POST = 4 + 12;      // IDAT CRC + IEND chunk

static bool png_geometry(uint32_t width, uint32_t height, uint32_t depth, uint32_t color, uint32_t* stride, uint32_t* bpp) {
    uint32_t channels = color == 0 ? 1 : color == 2 ? 3 : color == 4 ? 2 : color == 6 ? 4 : 0;  // (no palette images: no PLTE here)
    if (!channels || !width || !height || (depth != 8 && depth != 16)) return false;
    if ((uint64_t)width * channels * depth / 8 > 0x7fffffffull) return false;
    *bpp = channels * depth / 8;
    *stride = width * *bpp;
    return true;
}
extern "C" size_t fdb_png_file_bound(uint32_t width, uint32_t height, uint32_t bit_depth, uint32_t color_type) {
    uint32_t stride = 0, bpp = 0;
    if (!png_geometry(width, height, bit_depth, color_type, &stride, &bpp)) return 0;
    return PNG_PRE + fdb_deflate_ultrafast_bound((size_t)height * (1 + (size_t)stride)) + PNG_POST;
}
extern "C" int fdb_png_encode_files_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* width,
                                          const uint32_t* height, const uint32_t* bit_depth, const uint32_t* color_type,
                                          uint32_t mode, uint8_t* file_base, const uint64_t* file_off, const uint64_t* file_cap,
                                          uint64_t* file_len, int32_t* status, size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !raw_base || !raw_off || !width || !height || !bit_depth || !color_type || !file_base || !file_off ||
        !file_cap || !file_len || !status)
        return fail(ctx, "fdb_png_encode_files_batch", cudaSuccess);
    // Only images with a valid geometry and a slot that can hold the framing go to the device (compacted, in
    // their original order); the others get their status here and no byte of their slot is written.
    std::vector<size_t> idx;
    idx.reserve(n);
    std::vector<uint32_t> stride(n), bpp(n);
    for (size_t i = 0; i < n; i++) {
        file_len[i] = 0;
        if (!png_geometry(width[i], height[i], bit_depth[i], color_type[i], &stride[i], &bpp[i]))
            status[i] = ST_PNG_BAD_GEOMETRY;
        else if (file_cap[i] <= PNG_PRE + PNG_POST)
            status[i] = ST_OUTPUT_BUFFER_TOO_SMALL;
        else {
            status[i] = ST_OK;
            idx.push_back(i);
        }
    }
    const size_t m = idx.size();
    if (m == 0) return 0;
    std::vector<uint32_t> c_stride(m), c_bpp(m), c_h(m), crc(m);
    std::vector<uint64_t> c_raw_off(m), raw_len(m), z_off(m), z_cap(m), z_len(m);
    std::vector<int32_t> fst(m, 0), zst(m, 0);
    for (size_t k = 0; k < m; k++) {
        const size_t i = idx[k];
        c_stride[k] = stride[i];
        c_bpp[k] = bpp[i];
        c_h[k] = height[i];
        c_raw_off[k] = raw_off[i];
        raw_len[k] = (uint64_t)height[i] * stride[i];
        z_off[k] = file_off[i] + PNG_PRE;
        z_cap[k] = file_cap[i] - PNG_PRE - PNG_POST;
    }
    PngPre pre = {c_h.data(), c_stride.data(), c_bpp.data(), mode, fst.data(), crc.data()};
    int r = host_batch(ctx, 1, raw_base, c_raw_off.data(), raw_len.data(), file_base, z_off.data(), z_cap.data(), z_len.data(), nullptr,
                       zst.data(), m, 0, &pre);
    if (r) return r;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    for (size_t k = 0; k < m; k++) {
        const size_t i = idx[k];
        status[i] = fst[k] != ST_OK ? fst[k] : zst[k];
        if (status[i] != ST_OK) continue;
        uint8_t* f = file_base + file_off[i];
        memcpy(f, sig, 8);
        put_be32(f + 8, 13);
        memcpy(f + 12, "IHDR", 4);
        put_be32(f + 16, width[i]);
        put_be32(f + 20, height[i]);
        f[24] = (uint8_t)bit_depth[i];
        f[25] = (uint8_t)color_type[i];
        f[26] = f[27] = f[28] = 0;
        put_be32(f + 29, host_crc32(f + 12, 17));
        put_be32(f + 33, (uint32_t)z_len[k]);
        memcpy(f + 37, "IDAT", 4);
        uint8_t* t = f + PNG_PRE + z_len[k];
        put_be32(t, crc[k]);
        put_be32(t + 4, 0);
        memcpy(t + 8, "IEND", 4);
        put_be32(t + 12, 0xae426082u);  // crc32("IEND")
        file_len[i] = PNG_PRE + z_len[k] + PNG_POST;
    }
    return 0;
}

// ---- CRC-32 of a batch of byte ranges (PNG chunk CRCs) ------------------------------------------------
static int crc_launch(fdb_ctx* ctx, const void* d_base, const uint64_t* d_off, const uint64_t* d_len, uint32_t seed,
                      uint32_t* d_crc, size_t n, void* cuda_stream, uint32_t* counter);
extern "C" int fdb_crc32_batch_device(fdb_ctx* ctx, const void* d_base, const uint64_t* d_off, const uint64_t* d_len,
                                      uint32_t seed, uint32_t* d_crc, size_t n, void* cuda_stream) {
    return crc_launch(ctx, d_base, d_off, d_len, seed, d_crc, n, cuda_stream, nullptr);
}
static int crc_launch(fdb_ctx* ctx, const void* d_base, const uint64_t* d_off, const uint64_t* d_len, uint32_t seed,
                      uint32_t* d_crc, size_t n, void* cuda_stream, uint32_t* counter) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !d_off || !d_len || !d_crc) return fail(ctx, "fdb_crc32_batch_device", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CrcBatch b;
    b.base = (const uint8_t*)d_base;
    b.off = d_off;
    b.len = d_len;
    b.crc = d_crc;
    b.n = (uint32_t)n;
    b.seed = seed;
    if (!counter) counter = ctx->d_counters + 13;
    FDB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    const uint32_t sms = (uint32_t)std::max(ctx->sm_count, 1);
    uint32_t grid = (uint32_t)std::min<size_t>((n + CRC_WARPS - 1) / CRC_WARPS, (size_t)sms * 4);
    FDB_LAUNCH(crc32_kernel, dim3(grid), dim3(CRC_WARPS * 32), 0, st, b, (const CrcTables*)ctx->d_crc, counter);
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    return 0;
}
extern "C" int fdb_crc32_batch(fdb_ctx* ctx, const uint8_t* base, const uint64_t* off, const uint64_t* len, uint32_t seed,
                               uint32_t* crc, size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !off || !len || !crc) return fail(ctx, "fdb_crc32_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    uint64_t span = 0;
    for (size_t i = 0; i < n; i++) span = std::max(span, off[i] + len[i]);
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, 3 * n * sizeof(uint64_t)))) return r;
    cudaStream_t st = ctx->lanes[0].st;
    uint64_t* d_off = ctx->d_meta;
    uint64_t* d_len = d_off + n;
    uint32_t* d_crc = (uint32_t*)(d_len + n);
    FDB_TRY(cudaMemcpyAsync(d_off, off, n * 8, cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(d_len, len, n * 8, cudaMemcpyHostToDevice, st));
    if (span) FDB_TRY(cudaMemcpyAsync(ctx->d_in, base, span, cudaMemcpyHostToDevice, st));
    if ((r = fdb_crc32_batch_device(ctx, ctx->d_in, d_off, d_len, seed, d_crc, n, st))) return r;
    FDB_TRY(cudaMemcpyAsync(crc, d_crc, n * 4, cudaMemcpyDeviceToHost, st));
    FDB_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ---- PNG files: container walk on the host (C++), everything per byte on the device -------------------
// PNG (Third Edition) 5.2-5.3 (signature, chunk layout), 11.2.2 (IHDR), 5.6 (chunk ordering: IHDR first, IDAT chunks
// consecutive, IEND last).  Only the 8 header bytes of each chunk are read here; payload bytes are touched on the
// device only (CRC-32, gather of the IDAT payloads, inflate, unfilter).
struct PngChunkRef {
    uint64_t off;  // of the chunk's type field, relative to the batch base (the CRC covers type + payload)
    uint32_t len;  // payload length
    uint32_t crc;  // stored CRC
    uint32_t file;
    bool idat;
};
struct PngFileInfo {
    uint32_t width = 0, height = 0, depth = 0, color = 0, bpp = 0, stride = 0;
    int32_t status = ST_PNG_BAD_FILE;
    size_t first_chunk = 0, n_chunks = 0, n_idat = 0;
    uint64_t idat_bytes = 0;
};
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static void png_walk(const uint8_t* base, uint64_t off, uint64_t n, uint32_t file, PngFileInfo& fi, std::vector<PngChunkRef>* chunks) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    const uint8_t* f = base + off;
    fi.status = ST_PNG_BAD_FILE;
    if (chunks) fi.first_chunk = chunks->size();
    if (n < 8 + 25 || memcmp(f, sig, 8) != 0) return;
    uint64_t pos = 8;
    bool have_ihdr = false, seen_idat = false, idat_closed = false, seen_end = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(f + pos);
        const uint8_t* type = f + pos + 4;
        if ((uint64_t)len + 12 > n - pos) return;  // truncated chunk
        const bool is_idat = memcmp(type, "IDAT", 4) == 0;
        if (!have_ihdr) {
            if (memcmp(type, "IHDR", 4) != 0 || len != 13) return;
            const uint8_t* d = f + pos + 8;
            fi.width = be32(d);
            fi.height = be32(d + 4);
            fi.depth = d[8];
            fi.color = d[9];
            const uint32_t comp = d[10], filt = d[11], interlace = d[12];
            uint32_t channels = 0;
            bool depth_ok = false;
            switch (fi.color) {  // PNG table 11.1
                case 0: channels = 1; depth_ok = fi.depth == 1 || fi.depth == 2 || fi.depth == 4 || fi.depth == 8 || fi.depth == 16; break;
                case 2: channels = 3; depth_ok = fi.depth == 8 || fi.depth == 16; break;
                case 3: channels = 1; depth_ok = fi.depth == 1 || fi.depth == 2 || fi.depth == 4 || fi.depth == 8; break;
                case 4: channels = 2; depth_ok = fi.depth == 8 || fi.depth == 16; break;
                case 6: channels = 4; depth_ok = fi.depth == 8 || fi.depth == 16; break;
                default: break;
            }
            if (!channels || !depth_ok || !fi.width || !fi.height || comp || filt || interlace > 1) return;
            if ((uint64_t)fi.width * channels * fi.depth > 0x7fffffffull * 8) return;
            fi.bpp = std::max(1u, channels * fi.depth / 8);
            fi.stride = (uint32_t)(((uint64_t)fi.width * channels * fi.depth + 7) / 8);
            have_ihdr = true;
            if (interlace) {  // a valid file, but Adam7 passes are not decoded here
                fi.status = ST_PNG_UNSUPPORTED;
                return;
            }
        } else if (is_idat) {
            if (idat_closed) return;  // IDAT chunks must be consecutive
            seen_idat = true;
            fi.n_idat++;
            fi.idat_bytes += len;
        } else {
            if (seen_idat) idat_closed = true;
            if (memcmp(type, "IEND", 4) == 0) seen_end = true;
        }
        if (chunks) {
            chunks->push_back({off + pos + 4, len, be32(f + pos + 8 + len), file, is_idat});
            fi.n_chunks++;
        }
        pos += 12 + (uint64_t)len;
        if (seen_end) break;
    }
    if (have_ihdr && seen_idat && seen_end) fi.status = ST_OK;
}

extern "C" int fdb_png_probe_batch(const uint8_t* file_base, const uint64_t* file_off, const uint64_t* file_len, uint32_t* width,
                                   uint32_t* height, uint32_t* bit_depth, uint32_t* color_type, uint32_t* stride,
                                   int32_t* status, size_t n) {
    if (n && (!file_base || !file_off || !file_len || !status)) return -1;
    for (size_t i = 0; i < n; i++) {
        PngFileInfo fi;
        png_walk(file_base, file_off[i], file_len[i], (uint32_t)i, fi, nullptr);
        if (width) width[i] = fi.width;
        if (height) height[i] = fi.height;
        if (bit_depth) bit_depth[i] = fi.depth;
        if (color_type) color_type[i] = fi.color;
        if (stride) stride[i] = fi.stride;
        status[i] = fi.status;
    }
    return 0;
}

// files -> raw pixels.  Slot i holds raw_cap[i] bytes; a file whose IHDR asks for more (height * stride, see
// fdb_png_probe_batch) gets status OutputTooLarge before anything is sized from it, so a hostile header can neither
// overrun a slot nor make the batch's device buffers unallocatable.
extern "C" int fdb_png_decode_files_batch(fdb_ctx* ctx, const uint8_t* file_base, const uint64_t* file_off,
                                          const uint64_t* file_len, uint8_t* raw_base, const uint64_t* raw_off,
                                          const uint64_t* raw_cap, int32_t* status, size_t n) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (n > 0xffffffffull || !file_base || !file_off || !file_len || !raw_base || !raw_off || !raw_cap || !status)
        return fail(ctx, "fdb_png_decode_files_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    std::vector<PngFileInfo> fi(n);
    std::vector<PngChunkRef> chunks;
    chunks.reserve(4 * n);
    uint64_t lo = ~0ull, hi = 0;
    for (size_t i = 0; i < n; i++) {
        png_walk(file_base, file_off[i], file_len[i], (uint32_t)i, fi[i], &chunks);
        if (fi[i].status == ST_OK && (uint64_t)fi[i].height * fi[i].stride > raw_cap[i]) fi[i].status = ST_OUTPUT_TOO_LARGE;
        lo = std::min(lo, file_off[i]);
        hi = std::max(hi, file_off[i] + file_len[i]);
    }
    if (hi <= lo) lo = hi = 0;
    // streams: a file with one IDAT chunk is inflated where it lies; several chunks are gathered into one stream
    std::vector<uint64_t> m(6 * n);  // idat_off | idat_len | filt_off | filt_cap | raw_off | (spare)
    std::vector<uint32_t> geo(3 * n);
    std::vector<GatherItem> items;
    std::vector<uint64_t> crc_off, crc_len;
    std::vector<size_t> crc_first(n + 1, 0), item_first(n + 1, 0);  // per file: where its entries start in the two lists
    uint64_t gather_span = 0, filt_span = 0, raw_span = 0;
    const uint64_t in_span = hi - lo;  // device copy of the files: d_in[0 .. in_span), gathered streams behind it
    for (size_t i = 0; i < n; i++) {
        const PngFileInfo& f = fi[i];
        const bool ok = f.status == ST_OK;
        const uint64_t filtered = ok ? (uint64_t)f.height * (1ull + f.stride) : 0;
        geo[i] = ok ? f.height : 0;
        geo[n + i] = ok ? f.stride : 0;
        geo[2 * n + i] = ok ? f.bpp : 1;
        m[2 * n + i] = filt_span;
        m[3 * n + i] = filtered;
        filt_span += (filtered + 15) & ~15ull;
        m[4 * n + i] = raw_off[i];
        if (ok) raw_span = std::max(raw_span, raw_off[i] + (uint64_t)f.height * f.stride);
        m[i] = 0;
        m[n + i] = 0;
        crc_first[i] = crc_off.size();
        item_first[i] = items.size();
        if (!ok) continue;
        uint64_t dst = ((in_span + 15) & ~15ull) + gather_span;
        bool first = true;
        for (size_t c = f.first_chunk; c < f.first_chunk + f.n_chunks; c++) {
            const PngChunkRef& ch = chunks[c];
            crc_off.push_back(ch.off - lo);
            crc_len.push_back((uint64_t)ch.len + 4);
            if (!ch.idat) continue;
            if (f.n_idat == 1) {
                m[i] = ch.off + 4 - lo;
                m[n + i] = ch.len;
            } else {
                if (first) m[i] = dst;
                first = false;
                if (ch.len) items.push_back({ch.off + 4 - lo, dst, ch.len});
                dst += ch.len;
                m[n + i] += ch.len;
            }
        }
        if (f.n_idat != 1) gather_span += (f.idat_bytes + 15) & ~15ull;
    }
    const size_t n_crc = crc_off.size();
    crc_first[n] = n_crc;
    item_first[n] = items.size();
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, ((in_span + 15) & ~15ull) + gather_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, raw_span + 64))) return r;
    if ((r = grow(ctx, (void**)&ctx->d_mid, &ctx->d_mid_cap, filt_span + 64))) return r;
    const size_t meta_bytes = 7 * n * 8 + 3 * n * 4 + 2 * n * 4 + n_crc * (8 + 8 + 4) + items.size() * sizeof(GatherItem) + 64;
    if ((r = grow(ctx, (void**)&ctx->d_meta, &ctx->d_meta_cap, meta_bytes))) return r;
    cudaStream_t hs = ctx->h2d_st, ds = ctx->d2h_st;
    uint64_t* d = ctx->d_meta;
    uint64_t *d_idat_off = d, *d_idat_len = d + n, *d_filt_off = d + 2 * n, *d_filt_cap = d + 3 * n, *d_raw_off = d + 4 * n,
             *d_out_len = d + 5 * n, *d_consumed = d + 6 * n;
    uint64_t* d_crc_off = d + 7 * n;
    uint64_t* d_crc_len = d_crc_off + n_crc;
    GatherItem* d_items = (GatherItem*)(d_crc_len + n_crc);
    uint32_t* d_geo = (uint32_t*)(d_items + items.size());
    int32_t* d_st1 = (int32_t*)(d_geo + 3 * n);
    int32_t* d_st2 = d_st1 + n;
    uint32_t* d_crc = (uint32_t*)(d_st2 + n);
    FDB_TRY(cudaMemcpyAsync(d, m.data(), 5 * n * 8, cudaMemcpyHostToDevice, hs));
    FDB_TRY(cudaMemcpyAsync(d_geo, geo.data(), 3 * n * 4, cudaMemcpyHostToDevice, hs));
    if (n_crc) {
        FDB_TRY(cudaMemcpyAsync(d_crc_off, crc_off.data(), n_crc * 8, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaMemcpyAsync(d_crc_len, crc_len.data(), n_crc * 8, cudaMemcpyHostToDevice, hs));
    }
    if (!items.empty()) FDB_TRY(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(GatherItem), cudaMemcpyHostToDevice, hs));
    // chunks of files (files and pixel slots in ascending order, what every packer produces; else one chunk):
    // files up | CRCs, gather, inflate, unfilter | pixels back overlap, every chunk queued at once
    bool ascending = true;
    for (size_t i = 1; i < n && ascending; i++)
        ascending = file_off[i] >= file_off[i - 1] + file_len[i - 1] &&
                    raw_off[i] >= raw_off[i - 1] + (fi[i - 1].status == ST_OK ? (uint64_t)fi[i - 1].height * fi[i - 1].stride : 0);
    size_t nchunk = 1;
    if (ascending) nchunk = (size_t)std::min<uint64_t>(std::min<uint64_t>((in_span + raw_span) / ctx->chunk_bytes, FDB_MAX_CHUNKS), n);
    if (nchunk < 1) nchunk = 1;
    const size_t per = (n + nchunk - 1) / nchunk;
    nchunk = (n + per - 1) / per;
    const int L = ctx->n_lanes;
    ctx->last_general_host = -1;
    ctx->last_split_host = -1;
    for (size_t kc = 0; kc < nchunk; kc++) {
        const size_t a = kc * per, b = std::min(n, a + per);
        fdb_lane& ln = ctx->lanes[kc % L];
        const uint64_t f_lo = nchunk == 1 ? lo : file_off[a], f_hi = nchunk == 1 ? hi : file_off[b - 1] + file_len[b - 1];
        if (f_hi > f_lo) FDB_TRY(cudaMemcpyAsync(ctx->d_in + (f_lo - lo), file_base + f_lo, f_hi - f_lo, cudaMemcpyHostToDevice, hs));
        FDB_TRY(cudaEventRecord(ctx->ev_in[kc], hs));
        FDB_TRY(cudaStreamWaitEvent(ln.st, ctx->ev_in[kc], 0));
        const size_t c0 = crc_first[a], c1 = crc_first[b], g0 = item_first[a], g1 = item_first[b];
        if (c1 > c0 && (r = crc_launch(ctx, ctx->d_in, d_crc_off + c0, d_crc_len + c0, 0, d_crc + c0, c1 - c0, ln.st, ln.d_counters + 13)))
            return r;
        if (g1 > g0) {
            uint32_t* counter = ln.d_counters + 14;
            FDB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), ln.st));
            const uint32_t grid = (uint32_t)std::min<size_t>((g1 - g0 + 7) / 8, (size_t)std::max(ctx->sm_count, 1) * 8);
            FDB_LAUNCH(gather_kernel, dim3(grid), dim3(256), 0, ln.st, (const uint8_t*)ctx->d_in, ctx->d_in,
                       (const GatherItem*)(d_items + g0), (uint32_t)(g1 - g0), counter);
            ctx->launches++;
            FDB_TRY(cudaGetLastError());
        }
        InflateBatch ib;
        ib.in_base = ctx->d_in;
        ib.in_off = d_idat_off + a;
        ib.in_len = d_idat_len + a;
        ib.out_base = ctx->d_mid;
        ib.out_off = d_filt_off + a;
        ib.out_cap = d_filt_cap + a;
        ib.out_len = d_out_len + a;
        ib.consumed = d_consumed + a;
        ib.status = d_st1 + a;
        ib.n = (uint32_t)(b - a);
        uint64_t max_in = 0, sum_in = 0;
        for (size_t i = a; i < b; i++) {
            max_in = std::max(max_in, m[n + i]);
            sum_in += m[n + i];
        }
        ib.flags = max_in >= ctx->inflate_split_min ? FDB_FLAG_SPLIT_LARGE : 0u;
        if ((r = launch_inflate(ctx, ib, ln.d_counters, &ln.d_worklist, &ln.worklist_cap, ln.st, nchunk > 1, &ln.split, sum_in))) return r;
        if ((r = png_launch(ctx, true, ctx->d_mid, d_filt_off + a, ctx->d_out, d_raw_off + a, d_geo + a, d_geo + n + a,
                            d_geo + 2 * n + a, 0, d_st2 + a, b - a, ln.st, ln.d_counters + 12)))
            return r;
        FDB_TRY(cudaEventRecord(ctx->ev_res[kc], ln.st));
        FDB_TRY(cudaStreamWaitEvent(ds, ctx->ev_res[kc], 0));
        // pixels back (images that follow each other with less than 16 bytes of padding travel in one copy, padding included)
        for (size_t i = a; i < b;) {
            if (fi[i].status != ST_OK) {
                i++;
                continue;
            }
            const uint64_t begin = raw_off[i];
            uint64_t end = begin + (uint64_t)fi[i].height * fi[i].stride;
            size_t j = i + 1;
            while (j < b && fi[j].status == ST_OK && raw_off[j] >= end && raw_off[j] - end < 16) {
                end = raw_off[j] + (uint64_t)fi[j].height * fi[j].stride;
                j++;
            }
            FDB_TRY(cudaMemcpyAsync(raw_base + begin, ctx->d_out + begin, end - begin, cudaMemcpyDeviceToHost, ds));
            i = j;
        }
    }
    std::vector<int32_t> st12(2 * n);
    std::vector<uint64_t> olen(n);
    std::vector<uint32_t> crc(n_crc);
    for (int l = 0; l < L; l++) FDB_TRY(cudaStreamSynchronize(ctx->lanes[l].st));
    FDB_TRY(cudaMemcpyAsync(st12.data(), d_st1, 2 * n * 4, cudaMemcpyDeviceToHost, ds));
    FDB_TRY(cudaMemcpyAsync(olen.data(), d_out_len, n * 8, cudaMemcpyDeviceToHost, ds));
    if (n_crc) FDB_TRY(cudaMemcpyAsync(crc.data(), d_crc, n_crc * 4, cudaMemcpyDeviceToHost, ds));
    FDB_TRY(cudaStreamSynchronize(ds));
    FDB_TRY(cudaStreamSynchronize(hs));
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        if (fi[i].status != ST_OK) {
            status[i] = fi[i].status;
            continue;
        }
        bool crc_ok = true;
        for (size_t c = fi[i].first_chunk; c < fi[i].first_chunk + fi[i].n_chunks; c++, k++) crc_ok = crc_ok && crc[k] == chunks[c].crc;
        int32_t s1 = st12[i];
        if (s1 == ST_OK && olen[i] != m[3 * n + i]) s1 = ST_INSUFFICIENT_INPUT;
        status[i] = !crc_ok ? ST_PNG_BAD_CRC : s1 != ST_OK ? s1 : st12[n + i];
    }
    return 0;
}

// ---- synthetic tiles --------------------------------------------------------------------------
extern "C" size_t fdb_synth_tile_bytes(uint32_t width, uint32_t height) {
    return (size_t)height * (1u + 4u * (size_t)width);
}
extern "C" int fdb_synth_tiles_host(uint8_t* out, uint64_t first_tile, uint64_t n_tiles, uint32_t width,
                                    uint32_t height, uint64_t seed) {
    if (!out || !width || !height) return -1;
    const size_t tb = fdb_synth_tile_bytes(width, height);
    for (uint64_t t = 0; t < n_tiles; t++) {
        TileParams p = tile_params(seed, first_tile + t, width, height);
        for (uint32_t y = 0; y < height; y++)
            for (uint32_t x = 0; x < width; x++) synth_pixel(out + t * tb, p, x, y, width);
    }
    return 0;
}
extern "C" int fdb_synth_tiles_device(fdb_ctx* ctx, void* d_out, uint64_t first_tile, uint64_t n_tiles, uint32_t width,
                                      uint32_t height, uint64_t seed, void* cuda_stream) {
    if (!ctx || !d_out || !width || !height) return -1;
    if (n_tiles == 0) return 0;
    FDB_TRY(cudaSetDevice(ctx->device));
    uint64_t total = n_tiles * (uint64_t)width * height;
    uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)std::max(ctx->sm_count, 1) * 16);
    FDB_LAUNCH(synth_tiles_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)cuda_stream, (uint8_t*)d_out, first_tile,
               n_tiles, width, height, seed);
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    return 0;
}

// ---- streaming decoders ----------------------------------------------------------------------------------------
// Decompressor::read (src/decompress.rs:158-219) for many in-flight decoders at once: every decoder keeps its state on
// the device (inflate_general.cuh: K3StreamState), one launch advances all decoders of a call.
static const uint64_t STREAM_WINDOW = 32768;  // matches reach back 32 KiB at most (RFC 1951)

extern "C" int fdb_stream_open_batch(fdb_ctx* ctx, uint32_t* ids, size_t n) {
    if (!ctx || (n && !ids)) return -1;
    FDB_TRY(cudaSetDevice(ctx->device));
    size_t scan = 0;
    for (size_t k = 0; k < n; k++) {
        while (scan < ctx->streams.size() && ctx->streams[scan].open) scan++;
        if (scan == ctx->streams.size()) ctx->streams.emplace_back();
        fdb_ctx::stream_slot& sl = ctx->streams[scan];
        if (!sl.d_state) FDB_TRY(cudaMalloc((void**)&sl.d_state, sizeof(K3StreamState)));
        FDB_TRY(cudaMemsetAsync(sl.d_state, 0, sizeof(K3StreamState), ctx->stream));
        const uint32_t one = 1;  // adler32 starts at 1
        FDB_TRY(cudaMemcpyAsync(&sl.d_state->adler_a, &one, sizeof one, cudaMemcpyHostToDevice, ctx->stream));
        FDB_TRY(cudaStreamSynchronize(ctx->stream));
        sl.open = true;
        sl.pos = sl.lo = 0;
        sl.tail.clear();
        sl.start_bit = 0;
        sl.finished = false;
        sl.final_status = 0;
        ids[k] = (uint32_t)scan;
    }
    return 0;
}

extern "C" int fdb_stream_close_batch(fdb_ctx* ctx, const uint32_t* ids, size_t n) {
    if (!ctx || (n && !ids)) return -1;
    for (size_t k = 0; k < n; k++) {
        if (ids[k] >= ctx->streams.size() || !ctx->streams[ids[k]].open) return fail(ctx, "fdb_stream_close_batch: no such decoder", cudaSuccess);
        fdb_ctx::stream_slot& sl = ctx->streams[ids[k]];
        sl.open = false;
        sl.tail.clear();
        sl.tail.shrink_to_fit();
        if (sl.buf_cap > (1u << 20)) {  // keep small window buffers for the next decoder in this slot
            cudaFree(sl.d_buf);
            sl.d_buf = nullptr;
            sl.buf_cap = 0;
        }
    }
    return 0;
}

extern "C" int fdb_stream_read_batch(fdb_ctx* ctx, const uint32_t* ids, const uint8_t* in_base, const uint64_t* in_off,
                                     const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_room,
                                     uint64_t* produced, int32_t* status, size_t n, uint32_t flags) {
    if (!ctx) return -1;
    if (n == 0) return 0;
    if (!ids || !in_off || !in_len || !out_off || !out_room || !produced || !status || n > 0xffffffffull)
        return fail(ctx, "fdb_stream_read_batch", cudaSuccess);
    FDB_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<K3StreamJob> jobs;
    std::vector<size_t> owner;  // job -> index in the call
    jobs.reserve(n);
    uint64_t stage_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        if (ids[i] >= ctx->streams.size() || !ctx->streams[ids[i]].open) return fail(ctx, "fdb_stream_read_batch: no such decoder", cudaSuccess);
        for (size_t j = 0; j < i; j++)
            if (ids[j] == ids[i]) return fail(ctx, "fdb_stream_read_batch: a decoder appears twice in one call", cudaSuccess);
        fdb_ctx::stream_slot& sl = ctx->streams[ids[i]];
        produced[i] = 0;
        if (sl.finished) {  // :185-187 (Done) -- or the error the stream ended with
            status[i] = sl.final_status;
            continue;
        }
        if (in_len[i]) sl.tail.insert(sl.tail.end(), in_base + in_off[i], in_base + in_off[i] + in_len[i]);
        // room for the new output behind the window
        const uint64_t room = out_room[i];
        if (sl.pos + room + 64 > sl.buf_cap) {
            const uint64_t keep = std::min<uint64_t>(sl.pos - sl.lo, STREAM_WINDOW);
            if (sl.pos > keep) {  // drop everything but the window
                FDB_LAUNCH(stream_compact_kernel, dim3(1), dim3(1024), 0, st, sl.d_buf, sl.pos, keep);
                ctx->launches++;
                FDB_TRY(cudaGetLastError());
                sl.pos = keep;
                sl.lo = 0;
            }
            if (sl.pos + room + 64 > sl.buf_cap) {
                const size_t ncap = (size_t)(sl.pos + room + 64 + (64u << 10));
                uint8_t* nb = nullptr;
                FDB_TRY(cudaMalloc((void**)&nb, ncap));
                if (sl.pos) FDB_TRY(cudaMemcpyAsync(nb, sl.d_buf, sl.pos, cudaMemcpyDeviceToDevice, st));
                FDB_TRY(cudaStreamSynchronize(st));
                cudaFree(sl.d_buf);
                sl.d_buf = nb;
                sl.buf_cap = ncap;
            }
        }
        K3StreamJob jb;
        memset(&jb, 0, sizeof jb);
        jb.state = sl.d_state;
        jb.in = nullptr;  // set below, once the staging buffer exists
        jb.in_len = sl.tail.size();
        jb.start_bit = sl.start_bit;
        jb.flags = flags;
        jb.buf = sl.d_buf;
        jb.pos = sl.pos;
        jb.lo = sl.lo;
        jb.room = room;
        jobs.push_back(jb);
        owner.push_back(i);
        stage_bytes += (sl.tail.size() + 15 + 16) & ~(uint64_t)15;
    }
    if (jobs.empty()) return 0;
    int r;
    if ((r = grow(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, stage_bytes + 64))) return r;
    {
        void* p = ctx->d_jobs;
        size_t cap_bytes = ctx->d_jobs_cap * sizeof(K3StreamJob);
        if ((r = grow(ctx, &p, &cap_bytes, jobs.size() * sizeof(K3StreamJob)))) return r;
        ctx->d_jobs = (K3StreamJob*)p;
        ctx->d_jobs_cap = cap_bytes / sizeof(K3StreamJob);
    }
    std::vector<uint8_t> stage((size_t)stage_bytes + 64, 0);
    uint64_t so = 0;
    for (size_t k = 0; k < jobs.size(); k++) {
        const fdb_ctx::stream_slot& sl = ctx->streams[ids[owner[k]]];
        if (!sl.tail.empty()) memcpy(stage.data() + so, sl.tail.data(), sl.tail.size());
        jobs[k].in = ctx->d_in + so;
        so += (sl.tail.size() + 15 + 16) & ~(uint64_t)15;
    }
    FDB_TRY(cudaMemcpyAsync(ctx->d_in, stage.data(), stage.size(), cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemcpyAsync(ctx->d_jobs, jobs.data(), jobs.size() * sizeof(K3StreamJob), cudaMemcpyHostToDevice, st));
    FDB_TRY(cudaMemsetAsync(ctx->d_counters + 2, 0, sizeof(uint32_t), st));
    const uint32_t grid = (uint32_t)std::min<size_t>(jobs.size(), (size_t)std::max(ctx->sm_count, 1) * 8);
    FDB_LAUNCH(inflate_stream_kernel, dim3(grid), dim3(32), sizeof(K3Smem), st, ctx->d_jobs, (uint32_t)jobs.size(), ctx->d_counters + 2);
    ctx->launches++;
    FDB_TRY(cudaGetLastError());
    FDB_TRY(cudaMemcpyAsync(jobs.data(), ctx->d_jobs, jobs.size() * sizeof(K3StreamJob), cudaMemcpyDeviceToHost, st));
    FDB_TRY(cudaStreamSynchronize(st));
    for (size_t k = 0; k < jobs.size(); k++) {
        const size_t i = owner[k];
        fdb_ctx::stream_slot& sl = ctx->streams[ids[i]];
        const K3StreamJob& jb = jobs[k];
        if (jb.produced) FDB_TRY(cudaMemcpyAsync(out_base + out_off[i], sl.d_buf + sl.pos, jb.produced, cudaMemcpyDeviceToHost, st));
        produced[i] = jb.produced;
        sl.pos += jb.produced;
        const uint64_t whole = std::min<uint64_t>(jb.consumed_bits >> 3, sl.tail.size());
        sl.tail.erase(sl.tail.begin(), sl.tail.begin() + (ptrdiff_t)whole);
        sl.start_bit = (uint32_t)(jb.consumed_bits & 7);
        status[i] = jb.status;
        if (jb.status >= 0) {  // complete, or an error: the decoder takes no more input (bytes behind the checksum are ignored)
            sl.finished = true;
            sl.final_status = jb.status;
            sl.tail.clear();
        }
    }
    FDB_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ---- several GPUs behind one handle --------------------------------------------------------------------------
// Streams are independent: the batch is partitioned by stream, byte-balanced, and every device runs the
// ordinary host-buffer call on its share from its own host thread.  No collective, no peer access; the only shared
// resource is the host link.
struct fdb_multi {
    std::vector<fdb_ctx*> ctx;
    std::vector<uint32_t> owner;  // device index of every stream of the last call
    char err[600] = {0};
};

extern "C" int fdb_multi_create(const int* devices, int n_devices, fdb_multi** out) {
    if (!out) return -1;
    *out = nullptr;
    if (!devices || n_devices <= 0 || n_devices > 64) return -1;
    fdb_multi* m = new (std::nothrow) fdb_multi();
    if (!m) return -3;
    for (int k = 0; k < n_devices; k++) {
        fdb_ctx* c = nullptr;
        int r = fdb_create(devices[k], &c);
        if (r != 0) {
            for (fdb_ctx* x : m->ctx) fdb_destroy(x);
            delete m;
            return r;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return 0;
}
extern "C" void fdb_multi_destroy(fdb_multi* m) {
    if (!m) return;
    for (fdb_ctx* c : m->ctx) fdb_destroy(c);
    delete m;
}
extern "C" int fdb_multi_device_count(const fdb_multi* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" const char* fdb_multi_last_error(const fdb_multi* m) { return m ? m->err : "null handle"; }
extern "C" int fdb_multi_last_partition(const fdb_multi* m, uint32_t* owner, size_t n) {
    if (!m || !owner || n != m->owner.size()) return -1;
    memcpy(owner, m->owner.data(), n * sizeof(uint32_t));
    return 0;
}

// longest-processing-time greedy (the partition of fdeflate_b200/shard.py): streams by cost descending, ties by index,
// each to the least loaded device (first minimum).  Every device's list comes out in ascending stream order, so a
// packer's ascending slot layout stays ascending per device and the pipeline can still cut it into chunks.
static void partition_lpt(const std::vector<uint64_t>& cost, size_t world, std::vector<uint32_t>& owner) {
    const size_t n = cost.size();
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost[a] > cost[b]; });
    typedef std::pair<uint64_t, uint32_t> Load;  // (bytes, device): the smallest pair is the first minimum
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
    for (size_t d = 0; d < world; d++) pq.push(Load(0, (uint32_t)d));
    owner.assign(n, 0);
    for (uint32_t i : order) {
        Load l = pq.top();
        pq.pop();
        owner[i] = l.second;
        l.first += std::max<uint64_t>(cost[i], 1);
        pq.push(l);
    }
}

static int multi_batch(fdb_multi* m, int kind, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                       uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint64_t* consumed,
                       int32_t* status, size_t n, uint32_t flags) {
    if (!m || m->ctx.empty()) return -1;
    m->err[0] = 0;
    if (n == 0) {
        m->owner.clear();
        return 0;
    }
    if (!in_off || !in_len || !out_off || !out_cap || !out_len || !status) {
        snprintf(m->err, sizeof m->err, "fdb_multi_*: invalid argument");
        return -1;
    }
    const size_t world = m->ctx.size();
    std::vector<uint64_t> cost(n);
    for (size_t i = 0; i < n; i++) cost[i] = in_len[i] + (kind == 0 ? out_cap[i] : 0);
    partition_lpt(cost, world, m->owner);
    // When cutting the batch into `world` CONTIGUOUS runs of streams balances (nearly) as well -- it does whenever there
    // are many streams per device -- that partition is used instead: a packer's slots then stay one contiguous range
    // per device, which the pipeline moves in large copies.  Otherwise the LPT shares interleave in the caller's
    // buffers, and each device copies back exactly its own slots.
    bool exact = true;
    {
        uint64_t total = 0;
        for (uint64_t c : cost) total += std::max<uint64_t>(c, 1);
        std::vector<uint64_t> load(world, 0);
        for (size_t i = 0; i < n; i++) load[m->owner[i]] += std::max<uint64_t>(cost[i], 1);
        const uint64_t lpt_max = *std::max_element(load.begin(), load.end());
        std::vector<uint32_t> run(n);
        std::vector<uint64_t> rload(world, 0);
        uint64_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            const uint64_t c = std::max<uint64_t>(cost[i], 1);
            size_t d = (size_t)(((acc + c / 2) * world) / total);  // the device whose range holds the stream's midpoint
            if (d >= world) d = world - 1;
            run[i] = (uint32_t)d;
            rload[d] += c;
            acc += c;
        }
        bool ascending = true;
        for (size_t i = 1; i < n && ascending; i++)
            ascending = in_off[i] >= in_off[i - 1] + in_len[i - 1] && out_off[i] >= out_off[i - 1] + out_cap[i - 1];
        const uint64_t run_max = *std::max_element(rload.begin(), rload.end());
        if (ascending && run_max <= lpt_max + lpt_max / 32) {
            m->owner = run;
            exact = false;
        }
    }
    struct Share {
        std::vector<uint32_t> idx;
        std::vector<uint64_t> in_off, in_len, out_off, out_cap, out_len, consumed;
        std::vector<int32_t> status;
        int rc = 0;
    };
    std::vector<Share> sh(world);
    for (size_t i = 0; i < n; i++) sh[m->owner[i]].idx.push_back((uint32_t)i);
    auto run = [&](size_t d) {
        Share& s = sh[d];
        const size_t k = s.idx.size();
        if (k == 0) return;
        s.in_off.resize(k), s.in_len.resize(k), s.out_off.resize(k), s.out_cap.resize(k), s.out_len.assign(k, 0);
        s.consumed.assign(k, 0), s.status.assign(k, 0);
        for (size_t j = 0; j < k; j++) {
            const uint32_t i = s.idx[j];
            s.in_off[j] = in_off[i], s.in_len[j] = in_len[i], s.out_off[j] = out_off[i], s.out_cap[j] = out_cap[i];
        }
        s.rc = host_batch(m->ctx[d], kind, in_base, s.in_off.data(), s.in_len.data(), out_base, s.out_off.data(), s.out_cap.data(),
                          s.out_len.data(), kind == 0 ? s.consumed.data() : nullptr, s.status.data(), k, flags, nullptr, exact);
    };
#ifdef FDB_EMUL
    for (size_t d = 0; d < world; d++) run(d);  // (the test-only emulator is single-threaded)
#else
    {
        std::vector<std::thread> th;
        for (size_t d = 1; d < world; d++) th.emplace_back(run, d);
        run(0);
        for (auto& t : th) t.join();
    }
#endif
    int rc = 0;
    for (size_t d = 0; d < world; d++) {
        const Share& s = sh[d];
        if (s.rc != 0 && rc == 0) {
            rc = s.rc;
            snprintf(m->err, sizeof m->err, "device %zu: %s", d, fdb_last_error(m->ctx[d]));
        }
        for (size_t j = 0; j < s.idx.size(); j++) {
            const uint32_t i = s.idx[j];
            out_len[i] = s.out_len[j];
            status[i] = s.status[j];
            if (consumed) consumed[i] = s.consumed[j];
        }
    }
    return rc;
}

extern "C" int fdb_multi_inflate_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                       uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                       uint64_t* consumed, int32_t* status, size_t n, uint32_t flags) {
    return multi_batch(m, 0, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, consumed, status, n, flags);
}
extern "C" int fdb_multi_deflate_ultrafast_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off,
                                                 const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off,
                                                 const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n) {
    return multi_batch(m, 1, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, nullptr, status, n, 0);
}
extern "C" int fdb_multi_deflate_stored_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                              uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                              int32_t* status, size_t n) {
    return multi_batch(m, 2, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, nullptr, status, n, 0);
}
