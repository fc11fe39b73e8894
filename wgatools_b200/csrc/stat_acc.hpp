// stat_acc.hpp -- host accumulator behind wga_stat_acc_* : merges the per-slab pair tables coming
// back from the GPU, in file order, and prints the TSV of src/tools/stat.rs:107-126.
// merged mode : src/tools/stat.rs:167-223 merge_final_from_pair
// each mode   : src/tools/stat.rs:129-164 split_final
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/wga_b200.h"
#include "host_fmt.hpp"

struct wga_stat_acc {
    struct Row {
        std::string ref_name, query_name;
        uint64_t ref_size = 0, query_size = 0;
        uint64_t v[13] = {0};  // 11 counters (struct Cigar order), ref_start_min, query_start_min
        float inv_size = 0.f;  // each mode only
    };
    int each = 0;
    std::vector<Row> rows;                            // first-seen order (merged) / record order (each)
    std::unordered_map<std::string, uint32_t> index;  // merged mode: key -> row
    std::vector<uint32_t> inv_pair;                   // '-'-strand records in file order: row index
    std::vector<uint64_t> inv_num;                    //   and (aligned + query_aligned)
    std::string text, blob;

    static std::string key_of(const std::string &rn, uint64_t rs, const std::string &qn, uint64_t qs) {
        std::string k;
        k.reserve(rn.size() + qn.size() + 20);
        k += rn; k.push_back('\0'); k += qn; k.push_back('\0');
        k.append((const char *)&rs, 8);
        k.append((const char *)&qs, 8);
        return k;
    }
    // the key -> row map is only needed once a SECOND table is merged in (another slab, an imported blob); the first
    // table is bulk-loaded without it
    void ensure_index() {
        if (index.size() == rows.size()) return;
        index.clear();
        index.reserve(rows.size() * 2);
        for (uint32_t i = 0; i < rows.size(); ++i)
            index.emplace(key_of(rows[i].ref_name, rows[i].ref_size, rows[i].query_name, rows[i].query_size), i);
    }
    uint32_t add_new(const char *rn, size_t lr, uint64_t rs, const char *qn, size_t lq, uint64_t qs) {
        rows.emplace_back();
        Row &r = rows.back();
        r.ref_name.assign(rn, lr); r.query_name.assign(qn, lq); r.ref_size = rs; r.query_size = qs;
        r.v[11] = rs;  // stat.rs:186,189: minima start at the sizes
        r.v[12] = qs;
        return (uint32_t)rows.size() - 1;
    }
    uint32_t find_or_add(const std::string &rn, uint64_t rs, const std::string &qn, uint64_t qs) {
        ensure_index();
        std::string k = key_of(rn, rs, qn, qs);
        auto it = index.find(k);
        if (it != index.end()) return it->second;
        Row r;
        r.ref_name = rn; r.query_name = qn; r.ref_size = rs; r.query_size = qs;
        r.v[11] = rs;  // stat.rs:186,189: minima start at the sizes
        r.v[12] = qs;
        rows.push_back(std::move(r));
        uint32_t id = (uint32_t)rows.size() - 1;
        index.emplace(std::move(k), id);
        return id;
    }
    void add_pair_vals(uint32_t id, const uint64_t *v) {
        Row &r = rows[id];
        for (int i = 0; i < 11; ++i) r.v[i] += v[i];
        if (v[11] < r.v[11]) r.v[11] = v[11];
        if (v[12] < r.v[12]) r.v[12] = v[12];
    }

    void write() {
        text.clear();
        std::vector<float> inv(rows.size(), 0.f);
        if (!each) {
            // stat.inv_size += rec_stat.inv_size in record order (f32, order-dependent); a '+' record adds 0.0
            for (size_t i = 0; i < inv_pair.size(); ++i)
                inv[inv_pair[i]] += (float)inv_num[i] / (float)2;  // common.rs:134-137 with inv_event == 1
        } else {
            for (size_t i = 0; i < rows.size(); ++i) inv[i] = rows[i].inv_size;
        }
        std::vector<uint32_t> order(rows.size());
        for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return wga::natord_cmp(rows[a].ref_name, rows[b].ref_name) < 0;
        });
        if (!rows.empty())
            text += "ref_name\tref_size\tref_start\tquery_name\tquery_size\tquery_start\taligned_size\t"
                    "unaligned_size\tidentity\tsimilarity\tmatched\tmismatched\tins_event\tdel_event\t"
                    "ins_size\tdel_size\tinv_event\tinv_size\tinv_ins_event\tinv_ins_size\t"
                    "inv_del_event\tinv_del_size\n";
        for (uint32_t id : order) {
            const Row &r = rows[id];
            const uint64_t *c = r.v;
            uint64_t matched = c[0], mism = c[1];
            uint64_t aligned = c[0] + c[1] + c[5] + c[9];       // common.rs:124-125
            uint64_t unaligned = each ? 0 : r.ref_size - aligned;  // wraps like release Rust (stat.rs:216)
            float identity = (float)matched / (float)aligned;
            float similarity = (float)(matched + mism) / (float)aligned;
            wga::csv_field(text, r.ref_name);
            auto u = [&](uint64_t x) { text.push_back('\t'); text += std::to_string(x); };
            auto f = [&](float x) { text.push_back('\t'); text += wga::fmt_f32(x); };
            u(r.ref_size); u(c[11]);
            text.push_back('\t'); wga::csv_field(text, r.query_name);
            u(r.query_size); u(c[12]); u(aligned); u(unaligned);
            f(identity); f(similarity);
            u(matched); u(mism); u(c[2]); u(c[4]); u(c[3]); u(c[5]); u(c[10]);
            f(inv[id]);
            u(c[6]); u(c[7]); u(c[8]); u(c[9]);
            text.push_back('\n');
        }
    }

    // flat blob: u64 each, u64 n_rows, rows{u32 lr,u32 lq,bytes,u64 rs,u64 qs,u64 v[13],f32 inv,u32 pad},
    //            u64 n_inv, {u32 pair}xn, {u64 num}xn
    void export_blob() {
        blob.clear();
        auto put = [&](const void *p, size_t n) { blob.append((const char *)p, n); };
        uint64_t e = (uint64_t)each, n = rows.size();
        put(&e, 8); put(&n, 8);
        for (const Row &r : rows) {
            uint32_t lr = (uint32_t)r.ref_name.size(), lq = (uint32_t)r.query_name.size();
            put(&lr, 4); put(&lq, 4);
            put(r.ref_name.data(), lr); put(r.query_name.data(), lq);
            put(&r.ref_size, 8); put(&r.query_size, 8); put(r.v, 13 * 8);
            uint32_t pad = 0;
            put(&r.inv_size, 4); put(&pad, 4);
        }
        uint64_t ni = inv_pair.size();
        put(&ni, 8);
        put(inv_pair.data(), ni * 4);
        put(inv_num.data(), ni * 8);
    }
    int import_blob(const char *p, uint64_t len) {
        const char *end = p + len;
        auto get = [&](void *d, size_t n) -> bool {
            if ((size_t)(end - p) < n) return false;
            memcpy(d, p, n);
            p += n;
            return true;
        };
        uint64_t e, n;
        if (!get(&e, 8) || !get(&n, 8) || (int)e != each) return WGA_E_INVALID_ARG;
        std::vector<uint32_t> map(n);
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t lr, lq;
            if (!get(&lr, 4) || !get(&lq, 4) || (size_t)(end - p) < (size_t)lr + lq) return WGA_E_INVALID_ARG;
            std::string rn(p, lr), qn(p + lr, lq);
            p += lr + lq;
            uint64_t rs, qs, v[13];
            float invs;
            uint32_t pad;
            if (!get(&rs, 8) || !get(&qs, 8) || !get(v, 13 * 8) || !get(&invs, 4) || !get(&pad, 4))
                return WGA_E_INVALID_ARG;
            if (each) {
                Row r;
                r.ref_name = rn; r.query_name = qn; r.ref_size = rs; r.query_size = qs;
                memcpy(r.v, v, sizeof v);
                r.inv_size = invs;
                rows.push_back(std::move(r));
                map[i] = (uint32_t)rows.size() - 1;
            } else {
                map[i] = find_or_add(rn, rs, qn, qs);
                add_pair_vals(map[i], v);
            }
        }
        uint64_t ni;
        if (!get(&ni, 8) || (uint64_t)(end - p) < ni * 12) return WGA_E_INVALID_ARG;
        const char *pp = p, *pn = p + ni * 4;
        for (uint64_t i = 0; i < ni; ++i) {
            uint32_t a;
            uint64_t b;
            memcpy(&a, pp + i * 4, 4);
            memcpy(&b, pn + i * 8, 8);
            if (a >= n) return WGA_E_INVALID_ARG;
            inv_pair.push_back(map[a]);
            inv_num.push_back(b);
        }
        return WGA_OK;
    }
};
