// guided_build.h — host-side builder of the candidate tables of the Illumina-guided search (guided_core.cuh layout).
// Shared by slr_api.cu (product) and tests/host_sim (CPU replay of the kernel).
#pragma once
#include <cstdint>
#include <vector>
#include "guided_core.cuh"

struct SlrGuidedSetsHost {
    std::vector<uint32_t> slots;
    std::vector<uint2> groups;
    uint2 all_set, empty_set;
};

// one bucketised table (capacity = power of two >= 2 * keys, >= 8 slots = one bucket) appended to `slots`; returns (first slot, meta).
// A key goes into the first free slot of its bucket, a full bucket spills into the next one.  Keys that do not fit 2 * seq_len bits can
// never equal a probe (clean sequences only match) and are left out.
inline uint2 slr_guided_add_set(std::vector<uint32_t> &slots, const uint64_t *keys, int64_t n, int seq_len)
{
    uint2 r;
    r.x = (uint32_t)slots.size();
    r.y = 0u;
    if (n <= 0) return r;
    uint32_t lg = 3;
    while ((1ull << lg) < 2ull * (uint64_t)n) lg++;
    const uint32_t cap = 1u << lg, nb_mask = (cap >> 3) - 1u;
    const size_t base = slots.size();                          // a multiple of 8: every table is
    slots.resize(base + cap, SLR_G_EMPTY);
    bool all_t = false, any = false;
    for (int64_t i = 0; i < n; i++) {
        if (seq_len < 32 && (keys[i] >> (2 * seq_len)) != 0ull) continue;
        const uint32_t key = (uint32_t)keys[i];
        any = true;
        if (key == SLR_G_EMPTY) { all_t = true; continue; }
        uint32_t bucket = slr_g_bucket_of(key, lg);
        bool done = false;
        while (!done) {
            uint32_t *q = &slots[base + 8u * bucket];
            for (int k = 0; k < 8 && !done; k++) {
                if (q[k] == key) done = true;                  // duplicate
                else if (q[k] == SLR_G_EMPTY) { q[k] = key; done = true; }
            }
            bucket = (bucket + 1u) & nb_mask;
        }
    }
    r.y = lg | (all_t ? 0x100u : 0u) | (any ? 0x200u : 0u);
    return r;
}

inline void slr_guided_build(const uint64_t *group_keys, const int64_t *group_offsets, int64_t n_groups, const uint64_t *all_keys,
                             int64_t n_all, const uint64_t *empty_keys, int64_t n_empty, int seq_len, SlrGuidedSetsHost &H)
{
    H.slots.clear();
    H.groups.resize((size_t)(n_groups > 0 ? n_groups : 0));
    for (int64_t g = 0; g < n_groups; g++)
        H.groups[(size_t)g] = slr_guided_add_set(H.slots, group_keys + group_offsets[g], group_offsets[g + 1] - group_offsets[g], seq_len);
    H.all_set = slr_guided_add_set(H.slots, all_keys, all_keys ? n_all : 0, seq_len);
    H.empty_set = slr_guided_add_set(H.slots, empty_keys, empty_keys ? n_empty : 0, seq_len);
    if (H.slots.empty()) H.slots.push_back(SLR_G_EMPTY);
}
