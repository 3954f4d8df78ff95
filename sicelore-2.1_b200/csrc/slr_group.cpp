// slr_group.cpp — host side of the clustering seam: the grouping that PRODUCES the (cell, region) jobs slr_umi_assign is fed with.
//
// The reference does this work in plain Java on its reader thread (no edit distances, O(n log n) over positions / keys), so it stays host
// code here too — C++ behind the C ABI for callers without a JVM; a JVM caller keeps its own ReadGrouper / UmiClustering.  Restated from the
// bytecode, quirks included (each one is pinned by the interpreter-run vectors of tests/golden/ref_grouper.npz and ref_jobs.npz):
//   ReadGrouper.groupSams            F!com/rw/umifinder/bamreaders/ReadGrouper.class (ReadGrouper.java:L82-L260)
//   ReadGrouper$Cluster              (L455-L667): lazily cached centre = Math.round((float) mean position), stale after an off-centre removal
//   ReadGrouper$ClusterList          (L675-L785): off-centre passes, merging of neighbouring clusters, size > 1 filter
//   UmiClustering.cluster            F!com/rw/umifinder/analyzers/clustering/UmiClustering.class (UmiClustering.java:L97-L118, L134-L143)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <new>
#include <unordered_set>
#include <vector>

#include "../../include/sicelore_host.h"

extern "C" int slr_multi_fail(int code, const char *msg);      // slr_api.cu: sets the thread-local error message

struct slr_grouper {
    int max_dist;
    long long next_id;
};

namespace {

struct NullCenter {};                                          // java.lang.NullPointerException at ReadGrouper.java:L173

int java_round_f32(double x)                                   // Math.round((float) x): nearest, ties toward +infinity (L615-L616)
{
    float f = (float)x;
    double r = std::floor((double)f + 0.5);
    if (r < -2147483648.0) return INT32_MIN;
    if (r > 2147483647.0) return INT32_MAX;
    return (int)r;
}

struct Work {                                                  // one groupSams call
    slr_grouper *g;
    std::vector<int64_t> pos;                                  // position-sorted (stable) reads that have a position
    std::vector<int64_t> idx;                                  // indexInList of each (its rank among the reads WITH a position, BAM order)
};

struct Cluster {
    Work *w;
    std::vector<int> list;                                     // members: indices into w->pos
    long long id;
    bool has_center = false, has_max = false;
    int center = 0;
    long long max_idx = 0;

    explicit Cluster(Work *w_) : w(w_), id(w_->g->next_id++) {}                                    // L520-L523
    Cluster(Work *w_, std::vector<int> members) : w(w_), list(std::move(members)), id(0)          // L529-L535: sorted by position, stable
    {
        std::stable_sort(list.begin(), list.end(), [&](int a, int b) { return w->pos[(size_t)a] < w->pos[(size_t)b]; });
        id = w->g->next_id++;
    }
    void touch() { has_center = has_max = false; }
    void add(int m) { touch(); list.push_back(m); }                                                // L557-L562
    void add_all(const std::vector<int> &ms) { list.insert(list.end(), ms.begin(), ms.end()); touch(); }      // L570-L573
    void remove_all(const std::vector<int> &ms)                                                    // L581-L584
    {
        std::unordered_set<int> drop(ms.begin(), ms.end());
        list.erase(std::remove_if(list.begin(), list.end(), [&](int m) { return drop.count(m) != 0; }), list.end());
        touch();
    }
    bool get_center(int *out)                                                                      // getCenter -> setCenter (L491, L614-L618)
    {
        if (!has_center && !list.empty()) {
            long long s = 0;
            for (int m : list) s += w->pos[(size_t)m];
            center = java_round_f32((double)s / (double)list.size());
            has_center = true;
        }
        if (out) *out = center;
        return has_center;
    }
    int center_or_throw() { int c; if (!get_center(&c)) throw NullCenter(); return c; }
    long long get_max_index()                                                                      // L592-L595
    {
        if (!has_max && !list.empty()) {
            max_idx = w->idx[(size_t)list[0]];
            for (int m : list) max_idx = std::max(max_idx, (long long)w->idx[(size_t)m]);
            has_max = true;
        }
        return max_idx;
    }
    // the BiFunction of L626-L638; left: pos < centre - max, right: pos > centre + max (L648-L659)
    std::unique_ptr<Cluster> remove_off_center(bool left)
    {
        auto pred = [&](int m) {
            int c = center_or_throw();                         // fills the cache; never empty here (the list is not)
            long long p = w->pos[(size_t)m], d = w->g->max_dist;
            return left ? p < (long long)c - d : p > (long long)c + d;
        };
        bool any = false;
        for (int m : list) any |= pred(m);                     // count(): the predicate runs on every member
        if (!any) return nullptr;
        touch();                                               // L630-L633
        std::vector<int> out, keep;
        for (int m : list) (pred(m) ? out : keep).push_back(m);      // L634: the centre is recomputed on the still complete list ...
        list.swap(keep);                                       // L635: ... and NOT cleared after the removal
        return std::unique_ptr<Cluster>(new Cluster(w, std::move(out)));
    }
};

typedef std::vector<std::unique_ptr<Cluster>> Owner;           // keeps every Cluster of the call alive; the lists below hold plain pointers

// ClusterList.sortAndRemoveEmpty (L703).  A single cluster is never compared: its centre cache is not filled.
void sort_nonempty(std::vector<Cluster *> &cl)
{
    cl.erase(std::remove_if(cl.begin(), cl.end(), [](Cluster *c) { return c->list.empty(); }), cl.end());
    if (cl.size() > 1) {
        for (Cluster *c : cl) c->get_center(nullptr);
        std::stable_sort(cl.begin(), cl.end(), [](Cluster *a, Cluster *b) { return a->center < b->center; });
    }
}

// ClusterList.refineClusters (L711-L785)
std::vector<Cluster *> refine(Work &w, Owner &own, std::vector<Cluster *> clusters)
{
    std::vector<Cluster *> outliers, current = clusters;
    while (!current.empty()) {                                 // L729-L731
        std::vector<Cluster *> nxt;
        for (Cluster *c : current)
            for (int side = 0; side < 2; side++) {
                std::unique_ptr<Cluster> o = c->remove_off_center(side == 0);
                if (o) { nxt.push_back(o.get()); own.push_back(std::move(o)); }
            }
        outliers.insert(outliers.end(), nxt.begin(), nxt.end());
        current.swap(nxt);
    }
    clusters.insert(clusters.end(), outliers.begin(), outliers.end());      // L734
    sort_nonempty(clusters);                                                // L752
    const long long md = w.g->max_dist;
    bool keep_merging = true;
    while (keep_merging) {                                     // L756-L781
        keep_merging = false;
        for (size_t i = 0; i + 1 < clusters.size(); i++) {
            Cluster *left = clusters[i], *right = clusters[i + 1];
            if (left->list.empty()) continue;
            int rc = right->center_or_throw(), lc = left->center_or_throw();
            if ((long long)rc - lc < 2 * md) {
                bool left_bigger = left->list.size() > right->list.size();
                Cluster *src = left_bigger ? right : left, *dst = left_bigger ? left : right;
                std::vector<int> move;
                for (int m : src->list)
                    if (std::llabs(w.pos[(size_t)m] - (long long)dst->center) <= md) move.push_back(m);
                if (!move.empty()) {
                    keep_merging = true;
                    dst->add_all(move);
                    src->remove_all(move);
                }
            }
        }
        clusters.erase(std::remove_if(clusters.begin(), clusters.end(), [](Cluster *c) { return c->list.empty(); }), clusters.end());
    }
    clusters.erase(std::remove_if(clusters.begin(), clusters.end(), [](Cluster *c) { return c->list.size() <= 1; }), clusters.end());      // L783
    return clusters;
}

// doClusteringOneStrand (L234-L260)
std::vector<Cluster *> cluster_one_strand(Work &w, Owner &own, const std::vector<int> &ind)
{
    std::vector<Cluster *> clusters;
    if (ind.size() <= 1) return clusters;
    const long long md = w.g->max_dist;
    auto fresh = [&]() { own.emplace_back(new Cluster(&w)); return own.back().get(); };
    Cluster *cur = fresh();
    if (w.pos[(size_t)ind[1]] - w.pos[(size_t)ind[0]] < md) cur->add(ind[0]);
    for (size_t i = 1; i < ind.size(); i++) {
        if (w.pos[(size_t)ind[i]] - w.pos[(size_t)ind[i - 1]] < md) cur->add(ind[i]);      // the read that opens a gap joins no run,
        else if (cur->list.size() > 2) { clusters.push_back(cur); cur = fresh(); }          // and a run of <= 2 reads is not closed at it
    }
    if (cur->list.size() > 2) clusters.push_back(cur);
    return refine(w, own, clusters);
}

}   // namespace

extern "C" {

int slr_grouper_create(int max_genome_distance, int64_t first_region_id, slr_grouper **out)
{
    if (!out) return slr_multi_fail(SLR_E_INVALID, "slr_grouper_create: out is NULL");
    if (max_genome_distance < 0) return slr_multi_fail(SLR_E_INVALID, "slr_grouper_create: max_genome_distance < 0");
    slr_grouper *g = new (std::nothrow) slr_grouper{max_genome_distance, (long long)first_region_id};
    if (!g) return slr_multi_fail(SLR_E_NOMEM, "slr_grouper_create: out of host memory");
    *out = g;
    return SLR_OK;
}

void slr_grouper_destroy(slr_grouper *g) { delete g; }

int64_t slr_grouper_next_region_id(const slr_grouper *g) { return g ? (int64_t)g->next_id : -1; }

int slr_grouper_group_sams(slr_grouper *g, const int32_t *position, const uint8_t *has_position, const int32_t *flags, int64_t n,
                           int keep_data_end, int64_t *region_io, int64_t *last_index_out)
{
    if (!g || !last_index_out || n < 0 || (n > 0 && (!position || !flags || !region_io)))
        return slr_multi_fail(SLR_E_INVALID, "slr_grouper_group_sams: NULL argument / n < 0");
    *last_index_out = -1;
    if (n == 0) return SLR_OK;                                 // L82-L83: an empty chunk is not handed on
    try {
        Work w;
        w.g = g;
        std::vector<int64_t> filt;                             // chunk index of every read with a position (L119-L123)
        for (int64_t i = 0; i < n; i++)
            if (!has_position || has_position[i]) filt.push_back(i);
        std::vector<int64_t> order(filt.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int64_t)k;
        std::stable_sort(order.begin(), order.end(),           // Arrays.parallelSort(Comparable[]) is stable (L128)
                         [&](int64_t a, int64_t b) { return position[filt[(size_t)a]] < position[filt[(size_t)b]]; });
        w.pos.resize(order.size());
        w.idx = order;
        std::vector<int> fwd, rev;
        for (size_t k = 0; k < order.size(); k++) {
            int64_t ci = filt[(size_t)order[k]];
            w.pos[k] = position[ci];
            ((flags[ci] & 16) ? rev : fwd).push_back((int)k);  // L129-L134
        }
        Owner own;
        std::vector<Cluster *> clusters = cluster_one_strand(w, own, fwd);
        std::vector<Cluster *> r2 = cluster_one_strand(w, own, rev);
        clusters.insert(clusters.end(), r2.begin(), r2.end());
        sort_nonempty(clusters);                               // L167
        int64_t last_index = n - 1;
        if (keep_data_end && !clusters.empty()) {              // L171-L186
            const long long most_right = w.pos.back();
            while (!clusters.empty()) {
                Cluster *c = clusters.back();
                if (!c->has_center) throw NullCenter();
                if ((long long)c->center <= most_right - 3ll * g->max_dist) break;
                clusters.pop_back();
            }
            if (!clusters.empty()) {
                last_index = clusters.back()->get_max_index();
                if (last_index < n / 3) last_index = n / 3;
            }
        }
        for (Cluster *c : clusters)                            // L189-L191
            for (int m : c->list) region_io[filt[(size_t)order[(size_t)m]]] = (int64_t)c->id;
        *last_index_out = last_index;
        return SLR_OK;
    } catch (const NullCenter &) {
        return slr_multi_fail(SLR_E_REFERENCE_THROWS, "java.lang.NullPointerException at ReadGrouper.java:L173: the only surviving cluster's centre cache "
                                                      "is empty (region numbers consumed, no read updated)");
    } catch (const std::bad_alloc &) {
        return slr_multi_fail(SLR_E_NOMEM, "slr_grouper_group_sams: out of host memory");
    }
}

int slr_group_jobs(const uint64_t *cell_bc, const int64_t *region, const uint8_t *valid, int64_t n, int min_size, int64_t ram_reserved,
                   int64_t *order_out, int64_t *job_offsets_out, int64_t *n_jobs_out)
{
    if (n < 0 || !n_jobs_out || !job_offsets_out || (n > 0 && (!cell_bc || !region || !order_out)))
        return slr_multi_fail(SLR_E_INVALID, "slr_group_jobs: NULL argument / n < 0");
    if (ram_reserved != 0 && ram_reserved < 300) return slr_multi_fail(SLR_E_INVALID, "slr_group_jobs: ram_reserved must be 0 (no split) or >= 300");
    try {
        std::vector<int64_t> o;
        for (int64_t i = 0; i < n; i++)
            if (!valid || valid[i]) o.push_back(i);
        std::stable_sort(o.begin(), o.end(), [&](int64_t a, int64_t b) {      // (barcode, region), reads of a group in input order
            return cell_bc[a] != cell_bc[b] ? cell_bc[a] < cell_bc[b] : region[a] < region[b];
        });
        const double root = ram_reserved ? std::sqrt((double)(ram_reserved / 300)) : 0.0;      // sqrt(MAX_SQUARE_NRECORDSPROCESSING), L59
        int64_t n_jobs = 0, w = 0;
        job_offsets_out[0] = 0;
        for (size_t a = 0; a < o.size();) {
            size_t b = a + 1;
            while (b < o.size() && cell_bc[o[b]] == cell_bc[o[a]] && region[o[b]] == region[o[a]]) b++;
            int64_t sz = (int64_t)(b - a);
            if (sz >= min_size) {                              // L135 (min_size = 2)
                int64_t part = sz;
                if (ram_reserved) {                            // L137-L140: nChunks = (int) ceil((float) n / sqrt(max)), parts of n / nChunks + 1
                    int n_chunks = (int)std::ceil((double)(float)sz / root);
                    part = n_chunks > 0 ? sz / n_chunks + 1 : sz + 1;
                }
                for (int64_t k = 0; k < sz; k += part) {
                    int64_t len = std::min(part, sz - k);
                    for (int64_t t = 0; t < len; t++) order_out[w++] = o[a + (size_t)(k + t)];
                    job_offsets_out[++n_jobs] = w;
                }
            }
            a = b;
        }
        *n_jobs_out = n_jobs;
        return SLR_OK;
    } catch (const std::bad_alloc &) {
        return slr_multi_fail(SLR_E_NOMEM, "slr_group_jobs: out of host memory");
    }
}

}   // extern "C"
