// comm.cu -- the one exchange step of the hot path: the cross-rank reduction of `stat` partials.
//
//   reference semantics: src/tools/stat.rs:167-223 merge_final_from_pair over ALL records of the file in file order:
//   u64 sums of the 11 counters, minima of the two start columns, and the order-dependent f32 sum of inv_size.
//   Ranks hold contiguous byte ranges of the file in rank order, so "file order" = (rank, local order).
//
// Steps (wga_stat_allreduce):
//   1. pair dictionary: ranks all-gather the 128-bit hashes of their Pair keys (16 B per pair) and derive the same
//      global pair list (first-seen over ranks 0..N-1 = file order); a rank whose key SET equals rank 0's (count +
//      order-independent digests) is skipped without touching its list.  The names of the pairs a rank introduces are
//      all-gathered once (variable length, padded).
//   2. ONE all-reduce(sum, u64) of the dense [n_pairs x 11] counter table and one all-reduce(min, u64) of
//      [n_pairs x 2] starts  (ncclAllReduce over NVLink / NVSwitch).
//   3. inv_size: f32 addition is not associative, so the per-pair f32 partial travels the ranks in order
//      (ncclSend / ncclRecv, n_pairs x 4 B per hop); rank r continues the sum with its own '-' records in file order.
//      The last rank's vector is the reference's value and is broadcast back.
// The host logic is written against a small transport interface: NCCL (device staging buffers on the ctx stream) in
// production, caller-provided callbacks (wga_transport) for hosts that own a communicator of another kind and for the
// CPU tests, which run it over gloo with world_size 2.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <exception>
#include <new>

#include "ctx.cuh"
#include "stat_acc.hpp"

namespace {

struct Transport {
    int rank = 0, world = 1;
    virtual ~Transport() {}
    // host scratch the transport can move without an extra staging copy (pinned for NCCL); valid until the next call
    // with the same slot
    virtual void *scratch(int slot, size_t bytes) = 0;
    virtual int allgather(const void *send, void *recv, uint64_t bytes_per_rank) = 0;
    virtual int allreduce_u64(uint64_t *buf, uint64_t n, int op /*0 sum, 1 min*/) = 0;
    // one ring hop: send to dst (or -1) and receive from src (or -1)
    virtual int sendrecv(const void *send, uint64_t send_bytes, int dst, void *recv, uint64_t recv_bytes, int src) = 0;
    virtual const char *last_error() const { return ""; }
};

// ------------------------------------------------------------------ NCCL, resolved at run time
// dlopen instead of a link-time dependency: a process that already holds an NCCL (torch's bundled libnccl.so.2) must
// not get a second one, and the library must still load where no NCCL is installed (single-GPU hosts).
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi &nccl() {
    static NcclApi A;
    static bool tried = false;
    if (tried) return A;
    tried = true;
    A.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses, if any
    if (!A.h) A.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!A.h) A.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!A.h) return A;
    bool all = true;
#define WGA_SYM(field, name)                                   \
    do {                                                       \
        *(void **)(&A.field) = dlsym(A.h, name);               \
        all = all && A.field != nullptr;                       \
    } while (0)
    WGA_SYM(GetUniqueId, "ncclGetUniqueId");
    WGA_SYM(CommInitRank, "ncclCommInitRank");
    WGA_SYM(CommDestroy, "ncclCommDestroy");
    WGA_SYM(CommCount, "ncclCommCount");
    WGA_SYM(CommUserRank, "ncclCommUserRank");
    WGA_SYM(AllGather, "ncclAllGather");
    WGA_SYM(AllReduce, "ncclAllReduce");
    WGA_SYM(Send, "ncclSend");
    WGA_SYM(Recv, "ncclRecv");
    WGA_SYM(GroupStart, "ncclGroupStart");
    WGA_SYM(GroupEnd, "ncclGroupEnd");
    WGA_SYM(GetErrorString, "ncclGetErrorString");
#undef WGA_SYM
    A.ok = all;
    return A;
}

struct NcclTransport : Transport {
    wga_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    bool own_comm = false;
    std::string err;
    struct Pinned { void *p = nullptr; size_t cap = 0; } pin[6];
    void *d_a = nullptr, *d_b = nullptr;
    size_t cap_a = 0, cap_b = 0;

    ~NcclTransport() override {
        cudaSetDevice(ctx->device);
        for (auto &b : pin)
            if (b.p) cudaFreeHost(b.p);
        if (d_a) cudaFree(d_a);
        if (d_b) cudaFree(d_b);
        if (own_comm && comm) nccl().CommDestroy(comm);
    }
    const char *last_error() const override { return err.c_str(); }
    void *scratch(int slot, size_t bytes) override {
        Pinned &b = pin[slot];
        if (b.cap >= bytes && b.p) return b.p;
        if (b.p) cudaFreeHost(b.p);
        b.p = nullptr;
        b.cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&b.p, want, cudaHostAllocDefault) != cudaSuccess) { b.p = nullptr; err = "cudaHostAlloc failed"; return nullptr; }
        b.cap = want;
        return b.p;
    }
    bool dev(void **p, size_t *cap, size_t bytes) {
        if (*cap >= bytes && *p) return true;
        if (*p) { cudaStreamSynchronize(ctx->stream); cudaFree(*p); }
        *p = nullptr;
        *cap = 0;
        size_t want = ((bytes + bytes / 4 + 4096) + 255) & ~(size_t)255;
        if (cudaMalloc(p, want) != cudaSuccess) { *p = nullptr; err = "cudaMalloc failed for a collective staging buffer"; return false; }
        *cap = want;
        return true;
    }
    int fail(const char *what, ncclResult_t r) {
        err = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
        return WGA_E_NCCL;
    }
    int cufail(const char *what, cudaError_t e) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return WGA_E_CUDA;
    }
#define NCCL_TRY(call, what)                          \
    do {                                              \
        ncclResult_t r__ = (call);                    \
        if (r__ != ncclSuccess) return fail(what, r__); \
    } while (0)
#define CU_TRY(call, what)                            \
    do {                                              \
        cudaError_t e__ = (call);                     \
        if (e__ != cudaSuccess) return cufail(what, e__); \
    } while (0)
    int allgather(const void *send, void *recv, uint64_t bytes) override {
        if (bytes == 0) return WGA_OK;
        if (!dev(&d_a, &cap_a, bytes) || !dev(&d_b, &cap_b, bytes * world)) return WGA_E_NOMEM;
        CU_TRY(cudaMemcpyAsync(d_a, send, bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D");
        NCCL_TRY(nccl().AllGather(d_a, d_b, bytes, ncclUint8, comm, ctx->stream), "ncclAllGather");
        CU_TRY(cudaMemcpyAsync(recv, d_b, bytes * world, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sync");
        ctx->n_nccl++;
        return WGA_OK;
    }
    int allreduce_u64(uint64_t *buf, uint64_t n, int op) override {
        if (n == 0) return WGA_OK;
        if (!dev(&d_a, &cap_a, n * 8)) return WGA_E_NOMEM;
        CU_TRY(cudaMemcpyAsync(d_a, buf, n * 8, cudaMemcpyHostToDevice, ctx->stream), "H2D");
        NCCL_TRY(nccl().AllReduce(d_a, d_a, n, ncclUint64, op ? ncclMin : ncclSum, comm, ctx->stream), "ncclAllReduce");
        CU_TRY(cudaMemcpyAsync(buf, d_a, n * 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sync");
        ctx->n_nccl++;
        return WGA_OK;
    }
    int sendrecv(const void *send, uint64_t sb, int dst, void *recv, uint64_t rb, int src) override {
        const bool do_s = dst >= 0 && sb, do_r = src >= 0 && rb;
        if (!do_s && !do_r) return WGA_OK;
        if (do_s && !dev(&d_a, &cap_a, sb)) return WGA_E_NOMEM;
        if (do_r && !dev(&d_b, &cap_b, rb)) return WGA_E_NOMEM;
        if (do_s) CU_TRY(cudaMemcpyAsync(d_a, send, sb, cudaMemcpyHostToDevice, ctx->stream), "H2D");
        NCCL_TRY(nccl().GroupStart(), "ncclGroupStart");
        if (do_s) NCCL_TRY(nccl().Send(d_a, sb, ncclUint8, dst, comm, ctx->stream), "ncclSend");
        if (do_r) NCCL_TRY(nccl().Recv(d_b, rb, ncclUint8, src, comm, ctx->stream), "ncclRecv");
        NCCL_TRY(nccl().GroupEnd(), "ncclGroupEnd");
        if (do_r) CU_TRY(cudaMemcpyAsync(recv, d_b, rb, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sync");
        ctx->n_nccl++;
        return WGA_OK;
    }
#undef NCCL_TRY
#undef CU_TRY
};

struct CallbackTransport : Transport {
    wga_transport cb;
    std::vector<std::vector<char>> bufs{6};
    void *scratch(int slot, size_t bytes) override {
        if (bufs[slot].size() < bytes) bufs[slot].resize(bytes + bytes / 4 + 64);
        return bufs[slot].data();
    }
    int allgather(const void *send, void *recv, uint64_t bytes) override {
        return bytes ? cb.allgather(cb.user, send, recv, bytes) : WGA_OK;
    }
    int allreduce_u64(uint64_t *buf, uint64_t n, int op) override { return n ? cb.allreduce_u64(cb.user, buf, n, op) : WGA_OK; }
    int sendrecv(const void *send, uint64_t sb, int dst, void *recv, uint64_t rb, int src) override {
        return cb.sendrecv(cb.user, send, sb, dst, recv, rb, src);
    }
};

// second, independent 64-bit hash of a Pair key: with stat.cuh's pair hash it forms the 128-bit identity ranks agree on
inline uint64_t key2_of(const wga_stat_acc &a, const wga_stat_acc::Row &r) {
    uint64_t h = 0x9AE16A3B2F90404Full;
    const unsigned char *p = (const unsigned char *)a.names.data() + r.name_off;
    for (uint32_t i = 0; i < r.lr; ++i) h = (h + p[i] + 0x100) * 0x9FB21C651E98DF25ull;
    h = wga::mix64(h ^ (r.ref_size + 0x632BE59BD9B4E019ull));
    for (uint32_t i = 0; i < r.lq; ++i) h = (h + p[r.lr + i] + 0x200) * 0x9FB21C651E98DF25ull;
    return wga::mix64(h ^ (r.query_size * 0x9E3779B97F4A7C15ull));
}

struct GlobalIndex {  // (key, key2) -> global pair index: a 4-byte slot table over a flat key array (both stay cache resident)
    std::vector<uint32_t> slot;  // index + 1, 0 = empty
    std::vector<uint64_t> keys;  // two per global pair, in global order
    size_t mask = 0;
    void init(size_t n) {
        size_t cap = 1024;
        while (cap < 2 * n + 2) cap <<= 1;
        slot.assign(cap, 0u);
        keys.clear();
        keys.reserve(2 * n);
        mask = cap - 1;
    }
    // returns the index; *fresh = inserted now with index `next`
    uint32_t find_or_insert(uint64_t a, uint64_t b, uint32_t next, bool *fresh) {
        size_t s = a & mask;
        for (;;) {
            const uint32_t e = slot[s];
            if (!e) { slot[s] = next + 1; keys.push_back(a); keys.push_back(b); *fresh = true; return next; }
            if (keys[2 * (size_t)(e - 1)] == a && keys[2 * (size_t)(e - 1) + 1] == b) { *fresh = false; return e - 1; }
            s = (s + 1) & mask;
        }
    }
    int64_t find(uint64_t a, uint64_t b) const {
        size_t s = a & mask;
        for (;;) {
            const uint32_t e = slot[s];
            if (!e) return -1;
            if (keys[2 * (size_t)(e - 1)] == a && keys[2 * (size_t)(e - 1) + 1] == b) return (int64_t)e - 1;
            s = (s + 1) & mask;
        }
    }
};

int stat_allreduce_impl(Transport &T, wga_stat_acc *acc, std::string *err) {
    const int R = T.rank, W = T.world;
    static const bool dbg = getenv("WGA_DEBUG_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) {
        if (!dbg || R != 0) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[allreduce] %-14s %8.3f ms\n", w, std::chrono::duration<double, std::milli>(t - T0).count());
        T0 = t;
    };
    if (acc->each) { *err = "--each rows are concatenated in rank order, not reduced"; return WGA_E_INVALID_ARG; }
    for (const auto &r : acc->rows)
        if (r.inv_size != 0.f) { *err = "this accumulator already holds a reduced inv_size column"; return WGA_E_INVALID_ARG; }
    const uint64_t n = acc->rows.size();
    // ---- 1a. counts, set digests
    std::vector<uint64_t> my2(n);
    uint64_t dsum = 0, dxor = 0;
    {
        const unsigned nt = n >= 16384 ? 4u : 1u;  // hashing 10^5 name pairs is a millisecond per thread
        uint64_t ps[4] = {0, 0, 0, 0}, px[4] = {0, 0, 0, 0};
        auto part = [&](unsigned t) {
            uint64_t a = 0, b = 0;
            for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
                my2[i] = key2_of(*acc, acc->rows[i]);
                a += wga::mix64(acc->rows[i].key ^ my2[i]);
                b ^= wga::mix64(acc->rows[i].key + my2[i]);
            }
            ps[t] = a; px[t] = b;
        };
        if (nt == 1) part(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(part, t);
            for (auto &x : th) x.join();
        }
        for (unsigned t = 0; t < nt; ++t) { dsum += ps[t]; dxor ^= px[t]; }
    }
    uint64_t *hdr_s = (uint64_t *)T.scratch(0, 32), *hdr_r = (uint64_t *)T.scratch(1, 32 * (size_t)W);
    if (!hdr_s || !hdr_r) return WGA_E_NOMEM;
    hdr_s[0] = n; hdr_s[1] = dsum; hdr_s[2] = dxor; hdr_s[3] = 0;
    WGA_TRY(T.allgather(hdr_s, hdr_r, 32));
    lap("digests");
    std::vector<uint64_t> hdr(hdr_r, hdr_r + 4 * (size_t)W);
    uint64_t n_max = 0, n_tot = 0;
    for (int r = 0; r < W; ++r) { n_max = std::max(n_max, hdr[4 * r]); n_tot += hdr[4 * r]; }
    if (n_max == 0) return WGA_OK;  // nothing anywhere
    // ---- 1b. hash lists (16 B per pair, padded to the longest list)
    uint64_t *hl_s = (uint64_t *)T.scratch(0, n_max * 16), *hl_r = (uint64_t *)T.scratch(1, n_max * 16 * (size_t)W);
    if (!hl_s || !hl_r) return WGA_E_NOMEM;
    for (uint64_t i = 0; i < n; ++i) { hl_s[2 * i] = acc->rows[i].key; hl_s[2 * i + 1] = my2[i]; }
    memset(hl_s + 2 * n, 0, (n_max - n) * 16);
    WGA_TRY(T.allgather(hl_s, hl_r, n_max * 16));
    lap("hash lists");
    // ---- 1c. global order: first seen over ranks in order.  `uniform`: every rank holds rank 0's key set (an all-to-all
    //      file cut into ranges: every range meets every pair) -- the global list is rank 0's list, rank 0 maps to itself,
    //      and nobody needs anybody's names
    bool uniform = true;
    for (int r = 1; r < W; ++r) uniform = uniform && hdr[4 * r] == hdr[0] && hdr[4 * r + 1] == hdr[1] && hdr[4 * r + 2] == hdr[2];
    GlobalIndex G;
    uint32_t P = 0;
    std::vector<uint32_t> owner;       // global index -> rank that introduces it
    std::vector<uint32_t> mine_global;  // global indices this rank introduces, ascending
    std::vector<uint32_t> lmap(n);     // local row -> global index
    if (uniform && R == 0) {
        P = (uint32_t)n;
        for (uint64_t i = 0; i < n; ++i) lmap[i] = (uint32_t)i;
    } else {
        G.init(uniform ? hdr[0] : n_tot);
        owner.reserve(hdr[0] + 16);
        for (int r = 0; r < W; ++r) {
            const uint64_t nr = hdr[4 * r];
            if (r > 0 && nr == hdr[0] && hdr[4 * r + 1] == hdr[1] && hdr[4 * r + 2] == hdr[2]) continue;  // same key set as rank 0
            const uint64_t *hl = hl_r + (size_t)r * n_max * 2;
            for (uint64_t i = 0; i < nr; ++i) {
                bool fresh;
                G.find_or_insert(hl[2 * i], hl[2 * i + 1], P, &fresh);
                if (fresh) {
                    owner.push_back((uint32_t)r);
                    if (r == R) mine_global.push_back(P);
                    ++P;
                }
            }
        }
        for (uint64_t i = 0; i < n; ++i) {
            const int64_t g = G.find(acc->rows[i].key, my2[i]);
            if (g < 0) { *err = "internal: a local pair is missing from the global dictionary"; return WGA_E_INTERNAL; }
            lmap[i] = (uint32_t)g;
        }
    }
    lap("local map");
    lap("global order");
    // ---- 1d. names of the pairs each rank introduces: {u32 lr, u32 lq, u64 rs, u64 qs, u64 pair hash, bytes} in ascending
    //      global order
    std::vector<uint32_t> g2l(P, ~0u);
    for (uint64_t i = 0; i < n; ++i) g2l[lmap[i]] = (uint32_t)i;
    std::vector<uint64_t> blob_len(W, 0);
    uint64_t blob_max = 0;
    char *bl_r = nullptr;
    if (!uniform) {
        uint64_t my_blob = 0;
        for (uint32_t g : mine_global) { const auto &r = acc->rows[g2l[g]]; my_blob += 32 + r.lr + r.lq; }
        hdr_s = (uint64_t *)T.scratch(0, 8);
        hdr_r = (uint64_t *)T.scratch(1, 8 * (size_t)W);
        if (!hdr_s || !hdr_r) return WGA_E_NOMEM;
        hdr_s[0] = my_blob;
        WGA_TRY(T.allgather(hdr_s, hdr_r, 8));
        blob_len.assign(hdr_r, hdr_r + W);
        for (int r = 0; r < W; ++r) blob_max = std::max(blob_max, blob_len[r]);
        blob_max = (blob_max + 7) & ~7ull;
        char *bl_s = (char *)T.scratch(2, blob_max ? blob_max : 8);
        bl_r = (char *)T.scratch(3, (blob_max ? blob_max : 8) * (size_t)W);
        if (!bl_s || !bl_r) return WGA_E_NOMEM;
        char *p = bl_s;
        for (uint32_t g : mine_global) {
            const auto &r = acc->rows[g2l[g]];
            memcpy(p, &r.lr, 4); memcpy(p + 4, &r.lq, 4); memcpy(p + 8, &r.ref_size, 8); memcpy(p + 16, &r.query_size, 8);
            memcpy(p + 24, &r.key, 8);
            memcpy(p + 32, acc->names.data() + r.name_off, (size_t)r.lr + r.lq);
            p += 32 + r.lr + r.lq;
        }
        memset(p, 0, blob_max - my_blob);
        WGA_TRY(T.allgather(bl_s, bl_r, blob_max));
    }
    lap("names");
    // ---- 2. counters: one all-reduce(sum), one all-reduce(min).  Column 11 of the sum table carries, per pair, the integer
    //      sum of the '-' records' inv_size numerators (see 3.)
    constexpr size_t SC = 12;
    uint64_t *sums = (uint64_t *)T.scratch(4, (size_t)P * SC * 8), *mins = (uint64_t *)T.scratch(5, (size_t)P * 2 * 8);
    if (!sums || !mins) return WGA_E_NOMEM;
    if (n != P) {  // pairs this rank does not hold (n == P: lmap is a bijection, every entry is written below)
        memset(sums, 0, (size_t)P * SC * 8);
        memset(mins, 0xFF, (size_t)P * 16);
    }
    {
        const unsigned nt = n >= 16384 ? 4u : 1u;
        auto part = [&](unsigned t) {
            for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
                const uint64_t *v = acc->rows[i].v;
                uint64_t *d = sums + (size_t)lmap[i] * SC;
                memcpy(d, v, 88);
                d[11] = 0;
                mins[2 * (size_t)lmap[i]] = v[11];
                mins[2 * (size_t)lmap[i] + 1] = v[12];
            }
        };
        if (nt == 1) part(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(part, t);
            for (auto &x : th) x.join();
        }
    }
    {
        // integer sums are order-free: the record list is cut into one range per thread, each with its own compact table
        const size_t ni = acc->inv_pair.size();
        const unsigned nt = ni >= 200000 ? 4u : 1u;
        std::vector<std::vector<uint64_t>> ks(nt, std::vector<uint64_t>(P, 0));
        auto part = [&](unsigned t) {
            uint64_t *k = ks[t].data();
            for (size_t i = ni * t / nt; i < ni * (t + 1) / nt; ++i) k[lmap[acc->inv_pair[i]]] += acc->inv_num[i];
        };
        if (nt == 1) part(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(part, t);
            for (auto &x : th) x.join();
        }
        for (uint32_t g = 0; g < P; ++g) {
            uint64_t v = 0;
            for (unsigned t = 0; t < nt; ++t) v += ks[t][g];
            sums[(size_t)g * SC + 11] = v;
        }
    }
    WGA_TRY(T.allreduce_u64(sums, (uint64_t)P * SC, 0));
    WGA_TRY(T.allreduce_u64(mins, (uint64_t)P * 2, 1));
    lap("counters");
    // ---- 3. inv_size = the f32 sum, in file order, of (numerator as f32) / 2 over a pair's '-' records (common.rs:134-137).
    //      While a pair's numerators add up to less than 2^24 every term and every partial sum is exactly representable
    //      (k / 2 with k < 2^24), so the sum does not depend on the order: it IS (f32)total / 2, and the all-reduced integer
    //      column settles it.  Only the pairs beyond that bound need the order: their f32 partials (a compact array, one
    //      float per such pair) travel the ranks in file order, rank r continuing the sums with its own records -- a serial
    //      chain of W hops, skipped altogether when no pair needs it (every rank sees the same totals and takes the same
    //      branch).
    constexpr uint64_t kExact = 1ull << 24;
    float *fin = (float *)T.scratch(0, (size_t)P * 4);
    if (!fin) return WGA_E_NOMEM;
    std::vector<uint32_t> big_of;  // global pair -> index among the pairs that need the ring (~0u: settled)
    uint32_t n_big = 0;
    for (uint32_t g = 0; g < P; ++g) {
        const uint64_t k = sums[(size_t)g * SC + 11];
        fin[g] = (float)k / (float)2;
        if (k >= kExact) {
            if (big_of.empty()) big_of.assign(P, ~0u);
            big_of[g] = n_big++;
        }
    }
    if (n_big) {
        float *c = (float *)T.scratch(1, (size_t)n_big * 4), *c_all = (float *)T.scratch(2, (size_t)n_big * 4 * (size_t)W);
        if (!c || !c_all) return WGA_E_NOMEM;
        if (R == 0) memset(c, 0, (size_t)n_big * 4);
        else WGA_TRY(T.sendrecv(nullptr, 0, -1, c, (uint64_t)n_big * 4, R - 1));
        const size_t ni = acc->inv_pair.size();
        for (size_t i = 0; i < ni; ++i) {  // file order
            const uint32_t b = big_of[lmap[acc->inv_pair[i]]];
            if (b != ~0u) c[b] += (float)acc->inv_num[i] / (float)2;  // common.rs:134-137 with inv_event == 1
        }
        if (R + 1 < W) WGA_TRY(T.sendrecv(c, (uint64_t)n_big * 4, R + 1, nullptr, 0, -1));
        WGA_TRY(T.allgather(c, c_all, (uint64_t)n_big * 4));
        const float *last = c_all + (size_t)(W - 1) * n_big;
        for (uint32_t g = 0; g < P; ++g)
            if (big_of[g] != ~0u) fin[g] = last[big_of[g]];
    }
    lap("inv carry");
    // ---- rebuild the accumulator in global order: rows and names are laid out in bulk (no per-row vector growth)
    wga_stat_acc out;
    out.each = 0;
    out.rows.resize(P);
    {
        size_t name_bytes = 0;
        for (int r = 0; r < W; ++r) name_bytes += blob_len[r];  // upper bound: the blobs also hold 32 header bytes per pair
        out.names.resize(name_bytes);
    }
    std::vector<const char *> cur(W);
    for (int r = 0; r < W; ++r) cur[r] = uniform ? nullptr : bl_r + (size_t)r * blob_max;
    size_t noff = 0;
    if (uniform) out.names.resize(acc->names.size());
    char *nm = out.names.data();
    auto finish_row = [&](wga_stat_acc::Row &row, uint32_t g) {
        memcpy(row.v, sums + (size_t)g * SC, 88);
        row.v[11] = mins[2 * (size_t)g];
        row.v[12] = mins[2 * (size_t)g + 1];
        row.inv_size = fin[g];
    };
    if (uniform) {
        // every rank holds every pair: its own copy of the names and sizes.  Rows are independent once their name offsets
        // are known: ranges of the global order per thread, name bytes per range first, then the rows
        const unsigned nt = P >= 16384 ? 4u : 1u;
        std::vector<size_t> base(nt + 1, 0);
        auto count = [&](unsigned t) {
            size_t b = 0;
            for (uint32_t g = (uint32_t)((uint64_t)P * t / nt); g < (uint32_t)((uint64_t)P * (t + 1) / nt); ++g) {
                const wga_stat_acc::Row &l = acc->rows[g2l[g]];
                b += (size_t)l.lr + l.lq;
            }
            base[t + 1] = b;
        };
        auto fill = [&](unsigned t) {
            size_t off = base[t];
            for (uint32_t g = (uint32_t)((uint64_t)P * t / nt); g < (uint32_t)((uint64_t)P * (t + 1) / nt); ++g) {
                const wga_stat_acc::Row &l = acc->rows[g2l[g]];
                wga_stat_acc::Row &row = out.rows[g];
                row.lr = l.lr; row.lq = l.lq; row.ref_size = l.ref_size; row.query_size = l.query_size; row.key = l.key;
                const size_t nl = (size_t)row.lr + row.lq;
                row.name_off = off;
                memcpy(nm + off, acc->names.data() + l.name_off, nl);
                off += nl;
                finish_row(row, g);
            }
        };
        auto run = [&](auto &fn) {
            if (nt == 1) { fn(0); return; }
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
            for (auto &x : th) x.join();
        };
        run(count);
        for (unsigned t = 0; t < nt; ++t) base[t + 1] += base[t];
        run(fill);
        noff = base[nt];
    } else {
        for (uint32_t g = 0; g < P; ++g) {
            wga_stat_acc::Row &row = out.rows[g];
            const char *p = cur[owner[g]];
            memcpy(&row.lr, p, 4); memcpy(&row.lq, p + 4, 4); memcpy(&row.ref_size, p + 8, 8); memcpy(&row.query_size, p + 16, 8);
            // identity across ranks is the 128-bit (pair hash, key2) the dictionary was built on; the rebuilt accumulator's
            // own lookup table is made on demand from the rows' keys (find_or_add -> grow_table)
            memcpy(&row.key, p + 24, 8);
            const char *src = p + 32;
            cur[owner[g]] = p + 32 + (size_t)row.lr + row.lq;
            const size_t nl = (size_t)row.lr + row.lq;
            row.name_off = noff;
            memcpy(nm + noff, src, nl);
            noff += nl;
            finish_row(row, g);
        }
    }
    out.names.resize(noff);
    lap("rebuild");
    acc->names.swap(out.names);
    acc->rows.swap(out.rows);
    acc->table.clear();  // rebuilt on demand
    acc->inv_pair.clear();
    acc->inv_num.clear();
    return WGA_OK;
}

}  // namespace

struct wga_comm {
    Transport *t = nullptr;
    std::string err;
    ~wga_comm() { delete t; }
};

extern "C" {

int wga_nccl_unique_id(void *id128) {
    if (!id128) return WGA_E_INVALID_ARG;
    if (!nccl().ok) return WGA_E_NCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return WGA_E_NCCL;
    memcpy(id128, &id, sizeof id);
    return WGA_OK;
}

int wga_comm_init_nccl(wga_ctx *ctx, const void *id128, int rank, int world, wga_comm **out) {
    if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return WGA_E_INVALID_ARG;
    *out = nullptr;
    if (!nccl().ok) { ctx->set_error(WGA_E_NCCL, 0, "libnccl.so.2 not found (or too old)"); return WGA_E_NCCL; }
    cudaSetDevice(ctx->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    ncclResult_t r = nccl().CommInitRank(&c, world, id, rank);
    if (r != ncclSuccess) { ctx->set_error(WGA_E_NCCL, 0, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); return WGA_E_NCCL; }
    NcclTransport *t = new (std::nothrow) NcclTransport();
    wga_comm *cm = new (std::nothrow) wga_comm();
    if (!t || !cm) { delete t; delete cm; nccl().CommDestroy(c); return WGA_E_NOMEM; }
    t->ctx = ctx; t->comm = c; t->own_comm = true; t->rank = rank; t->world = world;
    cm->t = t;
    *out = cm;
    return WGA_OK;
}

int wga_comm_from_nccl(wga_ctx *ctx, void *nccl_comm, wga_comm **out) {
    if (!ctx || !nccl_comm || !out) return WGA_E_INVALID_ARG;
    *out = nullptr;
    if (!nccl().ok) { ctx->set_error(WGA_E_NCCL, 0, "libnccl.so.2 not found (or too old)"); return WGA_E_NCCL; }
    int rank = 0, world = 1;
    if (nccl().CommUserRank((ncclComm_t)nccl_comm, &rank) != ncclSuccess || nccl().CommCount((ncclComm_t)nccl_comm, &world) != ncclSuccess) {
        ctx->set_error(WGA_E_NCCL, 0, "not a valid ncclComm_t");
        return WGA_E_NCCL;
    }
    NcclTransport *t = new (std::nothrow) NcclTransport();
    wga_comm *cm = new (std::nothrow) wga_comm();
    if (!t || !cm) { delete t; delete cm; return WGA_E_NOMEM; }
    t->ctx = ctx; t->comm = (ncclComm_t)nccl_comm; t->own_comm = false; t->rank = rank; t->world = world;
    cm->t = t;
    *out = cm;
    return WGA_OK;
}

int wga_comm_from_transport(const wga_transport *tr, wga_comm **out) {
    if (!tr || !out || !tr->allgather || !tr->allreduce_u64 || !tr->sendrecv || tr->world < 1 || tr->rank < 0 || tr->rank >= tr->world)
        return WGA_E_INVALID_ARG;
    *out = nullptr;
    CallbackTransport *t = new (std::nothrow) CallbackTransport();
    wga_comm *cm = new (std::nothrow) wga_comm();
    if (!t || !cm) { delete t; delete cm; return WGA_E_NOMEM; }
    t->cb = *tr; t->rank = tr->rank; t->world = tr->world;
    cm->t = t;
    *out = cm;
    return WGA_OK;
}

void wga_comm_destroy(wga_comm *c) { delete c; }
int wga_comm_rank(const wga_comm *c) { return c && c->t ? c->t->rank : -1; }
int wga_comm_world(const wga_comm *c) { return c && c->t ? c->t->world : 0; }
const char *wga_comm_last_error(const wga_comm *c) { return c ? c->err.c_str() : "no communicator"; }

int wga_stat_allreduce(wga_comm *c, wga_stat_acc *acc) {
    if (!c || !c->t || !acc) return WGA_E_INVALID_ARG;
    c->err.clear();
    int rc;
    try {
        rc = stat_allreduce_impl(*c->t, acc, &c->err);
    } catch (const std::bad_alloc &) {
        c->err = "host allocation failed";
        return WGA_E_NOMEM;
    } catch (const std::exception &e) {
        c->err = std::string("internal: ") + e.what();
        return WGA_E_INTERNAL;
    }
    if (rc && c->err.empty()) c->err = c->t->last_error();
    return rc;
}

}  // extern "C"
