// stat_acc.hpp -- host accumulator behind wga_stat_acc_* : merges the per-slab pair tables coming
// back from the GPU, in file order, and prints the TSV of src/tools/stat.rs:107-126.
// merged mode : src/tools/stat.rs:167-223 merge_final_from_pair
// each mode   : src/tools/stat.rs:129-164 split_final
//
// Layout: rows are PODs, names live in one arena, the key -> row map is a flat open-addressing table keyed by the
// SAME 64-bit pair hash the device table uses (stat.cuh pair_key), so merging a 10^5-pair table coming back from the
// GPU is one probe + one memcmp per pair (no std::string, no node allocations).
#pragma once
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/wga_b200.h"
#include "hd_core.h"
#include "host_fmt.hpp"

namespace wga {
// host spelling of stat.cuh pair_key (FNV-1a over the names, splitmix finaliser over the sizes)
inline uint64_t host_pair_key(const char *rn, uint32_t lr, uint64_t rs, const char *qn, uint32_t lq, uint64_t qs) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < lr; ++i) h = (h ^ (unsigned char)rn[i]) * 0x100000001B3ull;
    h = mix64(h ^ rs);
    h ^= 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < lq; ++i) h = (h ^ (unsigned char)qn[i]) * 0x100000001B3ull;
    h = mix64(h ^ (qs * 0xD1342543DE82EF95ull));
    return h ? h : 1;
}
}  // namespace wga

struct wga_stat_acc {
    struct Row {
        uint64_t name_off = 0;  // names[name_off .. +lr) = ref_name, then lq bytes of query_name
        uint32_t lr = 0, lq = 0;
        uint64_t ref_size = 0, query_size = 0;
        uint64_t key = 0;      // pair hash (merged mode)
        uint64_t v[13] = {0};  // 11 counters (struct Cigar order), ref_start_min, query_start_min
        // each mode: the record's inv_size.  merged mode: the f32 sum of every '-' record that precedes this
        // accumulator's own list (set by the cross-rank ring carry of wga_stat_allreduce; 0 otherwise)
        float inv_size = 0.f;
    };
    int each = 0;
    std::vector<char> names;
    std::vector<Row> rows;            // first-seen order (merged) / record order (each)
    std::vector<uint32_t> table;      // merged mode: open addressing, row index + 1, 0 = empty
    std::vector<uint32_t> inv_pair;   // '-'-strand records in file order: row index
    std::vector<uint64_t> inv_num;    //   and (aligned + query_aligned)
    // the TSV: a raw buffer (std::string::resize would zero-fill 10^7 bytes that are overwritten right away)
    struct RawText {
        std::unique_ptr<char[]> p;
        size_t n = 0, cap = 0;
        void clear() { n = 0; }
        void resize_and_overwrite_compat(size_t m) {
            if (m > cap) { p.reset(new char[m]); cap = m; }
            n = m;
        }
        char *data() { return p.get(); }
        const char *data() const { return p.get(); }
        size_t size() const { return n; }
    } text;
    std::string blob;

    std::string_view ref_name(const Row &r) const { return std::string_view(names.data() + r.name_off, r.lr); }
    std::string_view query_name(const Row &r) const { return std::string_view(names.data() + r.name_off + r.lr, r.lq); }

    void grow_table(size_t want_rows) {
        size_t cap = table.size() ? table.size() : 1024;
        if (want_rows < rows.size()) want_rows = rows.size();
        while (cap < 2 * want_rows + 2) cap <<= 1;
        if (cap == table.size()) return;
        table.assign(cap, 0u);
        const size_t mask = cap - 1;
        for (uint32_t i = 0; i < rows.size(); ++i) {
            size_t s = rows[i].key & mask;
            while (table[s]) s = (s + 1) & mask;
            table[s] = i + 1;
        }
    }
    uint32_t push_row(const char *rn, uint32_t lr, uint64_t rs, const char *qn, uint32_t lq, uint64_t qs, uint64_t key) {
        Row r;
        r.name_off = names.size();
        r.lr = lr; r.lq = lq; r.ref_size = rs; r.query_size = qs; r.key = key;
        names.insert(names.end(), rn, rn + lr);
        names.insert(names.end(), qn, qn + lq);
        r.v[11] = rs;  // stat.rs:186,189: minima start at the sizes
        r.v[12] = qs;
        rows.push_back(r);
        return (uint32_t)rows.size() - 1;
    }
    // merged mode; key == 0: compute it here
    uint32_t find_or_add(const char *rn, uint32_t lr, uint64_t rs, const char *qn, uint32_t lq, uint64_t qs, uint64_t key = 0) {
        if (!key) key = wga::host_pair_key(rn, lr, rs, qn, lq, qs);
        if (table.size() < 2 * (rows.size() + 1) + 2) grow_table(rows.size() + 1);
        const size_t mask = table.size() - 1;
        size_t s = key & mask;
        for (;;) {
            const uint32_t e = table[s];
            if (!e) break;
            const Row &r = rows[e - 1];
            if (r.key == key && r.lr == lr && r.lq == lq && r.ref_size == rs && r.query_size == qs &&
                !memcmp(names.data() + r.name_off, rn, lr) && !memcmp(names.data() + r.name_off + lr, qn, lq))
                return e - 1;
            s = (s + 1) & mask;
        }
        const uint32_t id = push_row(rn, lr, rs, qn, lq, qs, key);
        table[s] = id + 1;
        return id;
    }
    // an EMPTY accumulator takes over a table of distinct pairs (names back to back: ref then query, per pair)
    void adopt_table(uint64_t n, const char *nm, uint64_t name_total, const uint32_t *nlen, const uint64_t *sizes /* rs, qs, key */,
                     const uint64_t *vals, int stride) {
        names.assign(nm, nm + name_total);
        rows.resize(n);
        uint64_t off = 0;
        for (uint64_t p = 0; p < n; ++p) {
            Row &r = rows[p];
            r.name_off = off; r.lr = nlen[2 * p]; r.lq = nlen[2 * p + 1];
            off += (uint64_t)r.lr + r.lq;
            r.ref_size = sizes[3 * p]; r.query_size = sizes[3 * p + 1]; r.key = sizes[3 * p + 2];
            const uint64_t *v = vals + p * stride;
            for (int i = 0; i < 11; ++i) r.v[i] = v[i];
            r.v[11] = v[11] < r.ref_size ? v[11] : r.ref_size;      // stat.rs:186,189: minima start at the sizes
            r.v[12] = v[12] < r.query_size ? v[12] : r.query_size;
        }
        table.clear();
        grow_table(n);
    }
    void add_pair_vals(uint32_t id, const uint64_t *v) {
        Row &r = rows[id];
        for (int i = 0; i < 11; ++i) r.v[i] += v[i];
        if (v[11] < r.v[11]) r.v[11] = v[11];
        if (v[12] < r.v[12]) r.v[12] = v[12];
    }

    // write_style_result (stat.rs:107-126): stable natord sort on ref_name, TSV with header
    void write() {
        static const bool dbg = getenv("WGA_DEBUG_TIMING") != nullptr;
        auto T0 = std::chrono::steady_clock::now();
        auto lap = [&](const char *w) {
            if (!dbg) return;
            auto t = std::chrono::steady_clock::now();
            fprintf(stderr, "[acc.write] %-10s %8.3f ms\n", w, std::chrono::duration<double, std::milli>(t - T0).count());
            T0 = t;
        };
        text.clear();
        const size_t n = rows.size();
        std::vector<float> inv(n, 0.f);
        for (size_t i = 0; i < n; ++i) inv[i] = rows[i].inv_size;
        if (!each) {
            // stat.inv_size += rec_stat.inv_size in record order (f32, order-dependent); a '+' record adds 0.0
            // (the order matters per PAIR only: with many '-' records a few threads each take the pairs of one residue class
            // and walk the whole list in file order)
            const size_t ni = inv_pair.size();
            auto add_class = [&](uint32_t t, uint32_t T) {
                for (size_t i = 0; i < ni; ++i) {
                    const uint32_t p = inv_pair[i];
                    if (T == 1 || (p & (T - 1)) == t) inv[p] += (float)inv_num[i] / (float)2;  // common.rs:134-137 with inv_event == 1
                }
            };
            if (ni < 200000) add_class(0, 1);
            else {
                std::vector<std::thread> th;
                for (uint32_t t = 0; t < 4; ++t) th.emplace_back(add_class, t, 4u);
                for (auto &x : th) x.join();
            }
        }
        // Rows with the same ref_name compare equal, so the stable sort only has to order the DISTINCT names:
        // group rows by ref_name (first-seen order), sort the groups, emit rows group by group.
        std::vector<uint32_t> grp_of(n), grp_first;
        {
            size_t cap = 64;
            while (cap < 2 * n + 2) cap <<= 1;
            std::vector<uint32_t> gt(cap, 0u);
            // the name hashes are independent (a few threads); only the probing below is sequential
            std::vector<uint64_t> hs(n);
            {
                const unsigned nt = n >= 16384 ? 4u : 1u;
                auto part = [&](unsigned t) {
                    for (size_t i = n * t / nt; i < n * (t + 1) / nt; ++i) {
                        uint64_t h = 0xCBF29CE484222325ull;
                        for (unsigned char c : ref_name(rows[i])) h = (h ^ c) * 0x100000001B3ull;
                        hs[i] = wga::mix64(h);
                    }
                };
                if (nt == 1) part(0);
                else {
                    std::vector<std::thread> th;
                    for (unsigned t = 0; t < nt; ++t) th.emplace_back(part, t);
                    for (auto &x : th) x.join();
                }
            }
            for (uint32_t i = 0; i < n; ++i) {
                const std::string_view nm = ref_name(rows[i]);
                size_t s = hs[i] & (cap - 1);
                for (;;) {
                    const uint32_t e = gt[s];
                    if (!e) { grp_first.push_back(i); gt[s] = (uint32_t)grp_first.size(); grp_of[i] = gt[s] - 1; break; }
                    if (ref_name(rows[grp_first[e - 1]]) == nm) { grp_of[i] = e - 1; break; }
                    s = (s + 1) & (cap - 1);
                }
            }
        }
        lap("group");
        const size_t ng = grp_first.size();
        std::vector<uint32_t> gorder(ng), grank(ng), gcount(ng + 1, 0u), order(n);
        for (uint32_t g = 0; g < ng; ++g) gorder[g] = g;
        std::stable_sort(gorder.begin(), gorder.end(), [&](uint32_t a, uint32_t b) {
            return wga::natord_cmp(ref_name(rows[grp_first[a]]), ref_name(rows[grp_first[b]])) < 0;
        });
        // groups that natord calls equal (e.g. names differing in whitespace only) must keep row order ACROSS groups:
        // give them one shared rank
        {
            uint32_t rank = 0;
            for (size_t k = 0; k < ng; ++k) {
                if (k && wga::natord_cmp(ref_name(rows[grp_first[gorder[k - 1]]]), ref_name(rows[grp_first[gorder[k]]])) != 0) ++rank;
                grank[gorder[k]] = rank;
            }
        }
        for (uint32_t i = 0; i < n; ++i) gcount[grank[grp_of[i]] + 1]++;
        for (size_t g = 0; g < ng; ++g) gcount[g + 1] += gcount[g];
        for (uint32_t i = 0; i < n; ++i) order[gcount[grank[grp_of[i]]]++] = i;

        lap("order");
        if (n == 0) return;
        static const char kHeader[] =
            "ref_name\tref_size\tref_start\tquery_name\tquery_size\tquery_start\taligned_size\t"
            "unaligned_size\tidentity\tsimilarity\tmatched\tmismatched\tins_event\tdel_event\t"
            "ins_size\tdel_size\tinv_event\tinv_size\tinv_ins_event\tinv_ins_size\t"
            "inv_del_event\tinv_del_size\n";
        // Rows are independent: ranges of the sorted order are formatted by a few host threads, each into its own buffer
        // sized for the worst case (19 integers <= 20 digits, 3 floats <= 24 bytes, 22 separators, the two names -- csv
        // quoting at most doubles them and adds two quotes), then laid end to end.
        struct Part { std::unique_ptr<char[]> p; size_t n = 0; };
        auto format_range = [&](size_t lo, size_t hi, Part &dst) {
            size_t cap = 16;
            for (size_t k = lo; k < hi; ++k) { const Row &r = rows[order[k]]; cap += 19 * 20 + 3 * 24 + 22 + 2 * ((size_t)r.lr + r.lq) + 4; }
            dst.p.reset(new char[cap]);  // uninitialised: only the bytes written are ever touched
            char *o = dst.p.get();
            auto name = [&](std::string_view f) {
                bool need = false;
                for (char c : f) need |= c == '\t' || c == '"' || c == '\r' || c == '\n';
                if (!need) { memcpy(o, f.data(), f.size()); o += f.size(); return; }
                *o++ = '"';
                for (char c : f) { if (c == '"') *o++ = '"'; *o++ = c; }
                *o++ = '"';
            };
            auto u = [&](uint64_t x) { *o++ = '\t'; o = std::to_chars(o, o + 20, x).ptr; };
            auto f = [&](float x) { *o++ = '\t'; o += wga::fmt_f32_to(o, x); };
            for (size_t k = lo; k < hi; ++k) {
                const uint32_t id = order[k];
                const Row &r = rows[id];
                const uint64_t *c = r.v;
                const uint64_t matched = c[0], mism = c[1];
                const uint64_t aligned = c[0] + c[1] + c[5] + c[9];          // common.rs:124-125
                const uint64_t unaligned = each ? 0 : r.ref_size - aligned;  // wraps like release Rust (stat.rs:216)
                const float identity = (float)matched / (float)aligned;
                const float similarity = (float)(matched + mism) / (float)aligned;
                name(ref_name(r));
                u(r.ref_size); u(c[11]);
                *o++ = '\t'; name(query_name(r));
                u(r.query_size); u(c[12]); u(aligned); u(unaligned);
                f(identity); f(similarity);
                u(matched); u(mism); u(c[2]); u(c[4]); u(c[3]); u(c[5]); u(c[10]);
                f(inv[id]);
                u(c[6]); u(c[7]); u(c[8]); u(c[9]);
                *o++ = '\n';
            }
            dst.n = (size_t)(o - dst.p.get());
        };
        unsigned nt = 1;
        if (n >= 8192) {
            nt = std::thread::hardware_concurrency();
            nt = nt == 0 ? 4 : (nt > 32 ? 32 : nt);
        }
        std::vector<Part> part(nt);
        if (nt == 1) {
            format_range(0, n, part[0]);
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(format_range, n * t / nt, n * (t + 1) / nt, std::ref(part[t]));
            for (auto &x : th) x.join();
        }
        lap("format");
        size_t total = sizeof kHeader - 1;
        std::vector<size_t> at(nt);
        for (unsigned t = 0; t < nt; ++t) { at[t] = total; total += part[t].n; }
        text.resize_and_overwrite_compat(total);
        memcpy(text.data(), kHeader, sizeof kHeader - 1);
        if (nt == 1) memcpy(text.data() + at[0], part[0].p.get(), part[0].n);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { memcpy(text.data() + at[t], part[t].p.get(), part[t].n); });
            for (auto &x : th) x.join();
        }
        lap("concat");
    }

    // flat blob: u64 each, u64 n_rows, rows{u32 lr,u32 lq,bytes,u64 rs,u64 qs,u64 v[13],f32 inv,u32 pad},
    //            u64 n_inv, {u32 pair}xn, {u64 num}xn
    void export_blob() {
        blob.clear();
        auto put = [&](const void *p, size_t n) { blob.append((const char *)p, n); };
        uint64_t e = (uint64_t)each, n = rows.size();
        put(&e, 8); put(&n, 8);
        for (const Row &r : rows) {
            put(&r.lr, 4); put(&r.lq, 4);
            put(names.data() + r.name_off, (size_t)r.lr + r.lq);
            put(&r.ref_size, 8); put(&r.query_size, 8); put(r.v, 13 * 8);
            uint32_t pad = 0;
            put(&r.inv_size, 4); put(&pad, 4);
        }
        uint64_t ni = inv_pair.size();
        put(&ni, 8);
        put(inv_pair.data(), ni * 4);
        put(inv_num.data(), ni * 8);
    }
    // Transactional: the blob is validated in full before anything is merged.  A merged-mode row's inv_size (a carried
    // f32 sum) is adopted only for rows this accumulator did not hold yet; f32 sums cannot be merged exactly otherwise.
    int import_blob(const char *p0, uint64_t len) {
        struct In { const char *rn; uint32_t lr, lq; uint64_t rs, qs, v[13]; float invs; };
        const char *p = p0, *end = p0 + len;
        auto get = [&](void *d, size_t n) -> bool {
            if ((size_t)(end - p) < n) return false;
            memcpy(d, p, n);
            p += n;
            return true;
        };
        uint64_t e, n;
        if (!get(&e, 8) || !get(&n, 8) || (int)e != each) return WGA_E_INVALID_ARG;
        if (n > (uint64_t)(end - p) / 136) return WGA_E_INVALID_ARG;  // a row is at least 8 + 16 + 104 + 8 bytes
        std::vector<In> in((size_t)n);
        for (uint64_t i = 0; i < n; ++i) {
            In &r = in[i];
            if (!get(&r.lr, 4) || !get(&r.lq, 4) || (uint64_t)(end - p) < (uint64_t)r.lr + r.lq) return WGA_E_INVALID_ARG;
            r.rn = p;
            p += (size_t)r.lr + r.lq;
            uint32_t pad;
            if (!get(&r.rs, 8) || !get(&r.qs, 8) || !get(r.v, 13 * 8) || !get(&r.invs, 4) || !get(&pad, 4)) return WGA_E_INVALID_ARG;
        }
        uint64_t ni;
        if (!get(&ni, 8) || ni > (uint64_t)(end - p) / 12) return WGA_E_INVALID_ARG;
        const char *pp = p, *pn = p + ni * 4;
        for (uint64_t i = 0; i < ni; ++i) {
            uint32_t a;
            memcpy(&a, pp + i * 4, 4);
            if (a >= n) return WGA_E_INVALID_ARG;
        }
        // ---- commit
        std::vector<uint32_t> map((size_t)n);
        for (uint64_t i = 0; i < n; ++i) {
            const In &r = in[i];
            if (each) {
                map[i] = push_row(r.rn, r.lr, r.rs, r.rn + r.lr, r.lq, r.qs, 0);
                Row &row = rows[map[i]];
                memcpy(row.v, r.v, sizeof r.v);
                row.inv_size = r.invs;
            } else {
                const size_t before = rows.size();
                map[i] = find_or_add(r.rn, r.lr, r.rs, r.rn + r.lr, r.lq, r.qs);
                add_pair_vals(map[i], r.v);
                if (rows.size() != before) rows[map[i]].inv_size = r.invs;
            }
        }
        inv_pair.reserve(inv_pair.size() + ni);
        inv_num.reserve(inv_num.size() + ni);
        for (uint64_t i = 0; i < ni; ++i) {
            uint32_t a;
            uint64_t b;
            memcpy(&a, pp + i * 4, 4);
            memcpy(&b, pn + i * 8, 8);
            inv_pair.push_back(map[a]);
            inv_num.push_back(b);
        }
        return WGA_OK;
    }
};
