// slr_usedlist.cpp — host side of the hand-over between scanfastq's two passes: from the pass-1 counts and the collision tester's Matches
// (slr_bc_collide) to the used-barcode list, with ranks, that pass 2 searches (slr_bc_table_create).
//
// A few thousand barcodes, integer comparisons: host work in the reference too (it runs once, between the passes).  Restated from the bytecode:
//   UsedBarcodesListData.filterLowCounts / finalizeData   F!com/rw/nanoporereadscanner/analyzers/UsedCellBCListGenerator$UsedBarcodesListData.class
//                                                        (UsedCellBCListGenerator.java:L359-L363, L391-L392)
//   BarcodeDatasetColissionTester.generateColissionMergedBCmap   F!…/BarcodeDatasetColissionTester.class (…java:L158-L203; onSuccess L240-L241)
//   rank assignment                                       F!com/rw/nanoporereadscanner/WorkerReadscanner.class (WorkerReadscanner.java:L264-L270)
// The one order that decides results — the iteration of a java.util.HashMap<Long, Set<Long>> (a barcode that was itself removed removes
// nobody) — follows the JDK's table layout; pinned by tests/golden/ref_usedlist.npz (the reference's own class files run by the interpreter).
#include <algorithm>
#include <cstdint>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/sicelore_host.h"

extern "C" int slr_multi_fail(int code, const char *msg);      // slr_api.cu: sets the thread-local error message

namespace {

inline uint32_t long_bucket(uint64_t v, uint32_t cap)           // java.util.HashMap: spread(Long.hashCode(v)) & (cap - 1)
{
    uint32_t h = (uint32_t)(v ^ (v >> 32));
    return (h ^ (h >> 16)) & (cap - 1);
}

// Iteration order of a HashMap<Long, ?> filled with put() in the order of `keys` (distinct): bins by index, a bin in insertion order (a
// resize splits a bin without reordering it).  Growth: doubled when size exceeds 0.75 x capacity and — below 64 — when a bin receives its
// 9th entry (treeifyBin resizes instead).  *unpinned: a bin reached 9 entries at capacity >= 64 (a tree bin; its order is not reproduced).
std::vector<int64_t> hashmap_order(const std::vector<int64_t> &ids, const uint64_t *barcodes, bool *unpinned)
{
    uint32_t cap = 16;
    size_t size = 0;
    std::vector<uint32_t> cnt(cap, 0);
    auto recount = [&](size_t upto) {
        cnt.assign(cap, 0);
        for (size_t k = 0; k < upto; k++) cnt[long_bucket(barcodes[ids[k]], cap)]++;
    };
    for (size_t k = 0; k < ids.size(); k++) {
        uint32_t b = long_bucket(barcodes[ids[k]], cap);
        if (++cnt[b] >= 9) {
            if (cap < 64) { cap *= 2; recount(k + 1); }
            else *unpinned = true;
        }
        size++;
        if (size * 4 > (size_t)cap * 3) { cap *= 2; recount(k + 1); }
    }
    std::vector<int64_t> out(ids);
    std::stable_sort(out.begin(), out.end(), [&](int64_t a, int64_t b) { return long_bucket(barcodes[a], cap) < long_bucket(barcodes[b], cap); });
    return out;
}

}   // namespace

extern "C" {

int slr_bc_used_filter_low_counts(const int32_t *counts, int64_t n, int64_t record_count, uint8_t *keep_out)
{
    if (n < 0 || (n > 0 && (!counts || !keep_out))) return slr_multi_fail(SLR_E_INVALID, "slr_bc_used_filter_low_counts: NULL argument / n < 0");
    const float cutoff = 2.0f * (float)(int32_t)record_count / 5000000.0f;      // fconst_2 * i2f(recordCount.get()) / 5.0E6f
    for (int64_t i = 0; i < n; i++) keep_out[i] = (uint8_t)((float)counts[i] > cutoff && counts[i] > 1);
    return SLR_OK;
}

int slr_bc_used_merge_collisions(const uint64_t *barcodes, const int32_t *counts, const slr_collide_result *collide, int64_t n,
                                 int min_count_fold, int merge_ed, int cells_fold, uint8_t *keep_out, int32_t *rank_out, uint32_t *flags_out)
{
    if (n < 0 || (n > 0 && (!barcodes || !counts || !collide || !keep_out))) return slr_multi_fail(SLR_E_INVALID, "slr_bc_used_merge_collisions: NULL argument / n < 0");
    if (min_count_fold <= 0 || cells_fold <= 0) return slr_multi_fail(SLR_E_INVALID, "slr_bc_used_merge_collisions: minCountFold and cellsWithReadsnFoldBelowMaxToKeep must be > 0 (the Java divides by them)");
    if (flags_out) *flags_out = 0;
    try {
        std::unordered_map<uint64_t, int64_t> index;
        index.reserve((size_t)n * 2);
        for (int64_t i = 0; i < n; i++)
            if (!index.emplace(barcodes[i], i).second) return slr_multi_fail(SLR_E_INVALID, "slr_bc_used_merge_collisions: duplicate barcode");
        // L164-L183: the barcodes with a non-empty Matches (onSuccess keeps no other), by count descending — a stable sort over the entry set of a
        // ConcurrentHashMap the reference fills from many threads; equal counts keep input order here
        std::vector<int64_t> with;
        for (int64_t i = 0; i < n; i++)
            if (collide[i].valid & 3) with.push_back(i);
        std::stable_sort(with.begin(), with.end(), [&](int64_t a, int64_t b) { return counts[a] > counts[b]; });
        std::vector<int64_t> victim((size_t)n * 2, -1);        // per barcode: its much smaller colliders (at most one per ED level)
        for (int64_t i : with) {
            const int32_t cutoff = counts[i] / min_count_fold;
            for (int e = 0; e < 2; e++) {
                if (!(collide[i].valid >> e & 1) || e + 1 > merge_ed) continue;
                auto it = index.find(collide[i].bc[e]);
                if (it == index.end()) return slr_multi_fail(SLR_E_INVALID, "slr_bc_used_merge_collisions: a collision record names a barcode that is not in the list "
                                                                            "(records of another list?)");
                if (counts[it->second] < cutoff) victim[(size_t)i * 2 + (size_t)e] = it->second;
            }
        }
        bool unpinned = false;
        std::vector<int64_t> order = hashmap_order(with, barcodes, &unpinned);
        std::vector<uint8_t> alive((size_t)n, 1);
        for (int64_t i : order)                                // L186-L195: the filter runs element by element: a removed barcode removes nobody
            if (alive[(size_t)i])
                for (int e = 0; e < 2; e++)
                    if (victim[(size_t)i * 2 + (size_t)e] >= 0) alive[(size_t)victim[(size_t)i * 2 + (size_t)e]] = 0;
        int32_t best = 0;
        bool any = false;
        for (int64_t i = 0; i < n; i++)
            if (alive[(size_t)i]) { best = any ? std::max(best, counts[i]) : counts[i]; any = true; }
        if (!any) return slr_multi_fail(SLR_E_REFERENCE_THROWS, "java.util.NoSuchElementException at BarcodeDatasetColissionTester.java:L197 (empty barcode list)");
        const int32_t min_counts = best / cells_fold;          // L197-L198
        std::vector<int64_t> kept;
        for (int64_t i = 0; i < n; i++) {
            keep_out[i] = (uint8_t)(alive[(size_t)i] && counts[i] >= min_counts);
            if (keep_out[i]) kept.push_back(i);
        }
        bool ties = false;
        if (rank_out) {                                        // WorkerReadscanner.java:L264-L270: rank 1 = most reads
            std::stable_sort(kept.begin(), kept.end(), [&](int64_t a, int64_t b) { return counts[a] > counts[b]; });
            for (int64_t i = 0; i < n; i++) rank_out[i] = 0;
            for (size_t r = 0; r < kept.size(); r++) {
                rank_out[kept[r]] = (int32_t)(r + 1);
                if (r && counts[kept[r]] == counts[kept[r - 1]]) ties = true;
            }
        }
        if (flags_out) *flags_out = (unpinned ? SLR_UL_ORDER_UNPIN : 0u) | (ties ? SLR_UL_RANK_TIES : 0u);
        return SLR_OK;
    } catch (const std::bad_alloc &) {
        return slr_multi_fail(SLR_E_NOMEM, "slr_bc_used_merge_collisions: out of host memory");
    }
}

}   // extern "C"
