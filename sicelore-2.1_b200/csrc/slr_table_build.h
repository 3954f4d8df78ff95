// slr_table_build.h — host-side construction of the digit-group bucket tables (see slr_table.cuh).
// Pure C++ (no CUDA calls): used by the C ABI (uploads the arrays) and by tests/host_sim (uses them in place).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "slr_table.cuh"

struct SlrTableHost {
    int bbits = 17;
    std::vector<uint8_t> slots;                // 4 tables back to back, (1 << bbits) buckets of 32 bytes each: 16 tag bytes, 16 pattern bytes
    std::vector<uint32_t> st_bucket;           // stash of all tables, sorted by (g << 24 | bucket)
    std::vector<uint16_t> st_slot;
    long long st_n[4] = {0, 0, 0, 0};          // stash entries per table (statistics)
    std::vector<uint32_t> ix_keys;
    std::vector<int32_t> ix_vals;
    uint32_t ix_mask = 0;
    std::vector<int32_t> rank;                 // empty = none
    long long n = 0;                           // number of input barcodes (index space of rank / counts)
    long long n_distinct = 0;
    long long n_ignored = 0;                   // entries with bits >= 32 set (can never equal a clean 16-nt window)
};

// bucket bits: ~6 keys per 16-slot bucket, clamped so that the tag fits 7 bits
inline int slr_choose_bbits(long long n)
{
    int b = 17;
    while (b < 22 && (6LL << b) < n) b++;
    return b;
}

inline void slr_build_table(const uint64_t *keys, const int32_t *rank, long long n, SlrTableHost &T, int force_bbits = 0)
{
    T.n = n;
    // ---- key -> first index map (duplicates keep the first index, like Map.put on a fresh map ignoring later puts) ----
    uint32_t cap = 16;
    while ((long long)cap < 2 * n + 2) cap <<= 1;
    T.ix_mask = cap - 1;
    T.ix_keys.assign(cap, 0);
    T.ix_vals.assign(cap, -1);
    std::vector<uint32_t> distinct;
    distinct.reserve((size_t)n);
    T.n_ignored = 0;
    for (long long i = 0; i < n; i++) {
        if (keys[i] >> 32) { T.n_ignored++; continue; }
        const uint32_t k = (uint32_t)keys[i];
        uint32_t h = slr_ix_hash(k) & T.ix_mask;
        bool dup = false;
        while (T.ix_vals[h] >= 0) {
            if (T.ix_keys[h] == k) { dup = true; break; }
            h = (h + 1) & T.ix_mask;
        }
        if (dup) continue;
        T.ix_keys[h] = k;
        T.ix_vals[h] = (int32_t)i;
        distinct.push_back(k);
    }
    T.n_distinct = (long long)distinct.size();
    if (rank) T.rank.assign(rank, rank + n); else T.rank.clear();

    T.bbits = force_bbits ? force_bbits : slr_choose_bbits(T.n_distinct);
    const int tb = 24 - T.bbits;
    const size_t nb = (size_t)1 << T.bbits;
    T.slots.assign(4 * nb * 32, 0);
    T.st_bucket.clear();
    T.st_slot.clear();
    // the four digit-group tables are independent: one host thread each (3 M keys: ~4 x faster than the serial loop)
    std::vector<std::pair<uint32_t, uint16_t>> stash[4];
    auto build_group = [&](int g) {
        std::vector<uint8_t> fill(nb, 0);
        for (uint32_t k : distinct) {
            const uint32_t m = slr_mix24(slr_key_rest(k, g));
            const uint32_t bucket = m >> tb, tag = m & ((1u << tb) - 1u);
            const uint16_t slot = (uint16_t)(0x8000u | (tag << 8) | slr_key_pat(k, g));
            if (fill[bucket] < 16) {
                uint8_t *b = &T.slots[((size_t)g * nb + bucket) * 32];
                b[fill[bucket]] = (uint8_t)(slot >> 8);
                b[16 + fill[bucket]] = (uint8_t)(slot & 0xFFu);
                fill[bucket]++;
            } else stash[g].emplace_back(((uint32_t)g << 24) | bucket, slot);
        }
        std::sort(stash[g].begin(), stash[g].end());
    };
    if (distinct.size() > (1u << 16)) {
        std::thread th[3];
        for (int g = 1; g < 4; g++) th[g - 1] = std::thread(build_group, g);
        build_group(0);
        for (int g = 1; g < 4; g++) th[g - 1].join();
    } else {
        for (int g = 0; g < 4; g++) build_group(g);
    }
    for (int g = 0; g < 4; g++) {
        T.st_n[g] = (long long)stash[g].size();
        for (auto &e : stash[g]) { T.st_bucket.push_back(e.first); T.st_slot.push_back(e.second); }
    }
}

// view of host arrays as the device struct (for the CPU simulation)
inline SlrTableDev slr_table_host_view(const SlrTableHost &T, unsigned long long *counts)
{
    SlrTableDev d;
    memset(&d, 0, sizeof(d));
    d.bk = reinterpret_cast<const uint4 *>(T.slots.data());
    d.st_bucket = T.st_bucket.data();
    d.st_slot = T.st_slot.data();
    d.st_total = (int)T.st_bucket.size();
    d.bbits = T.bbits;
    d.ix_keys = T.ix_keys.data();
    d.ix_vals = T.ix_vals.data();
    d.ix_mask = T.ix_mask;
    d.rank = T.rank.empty() ? nullptr : T.rank.data();
    d.counts = counts;
    d.n = T.n;
    return d;
}
