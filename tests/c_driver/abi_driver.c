/* TEST INFRASTRUCTURE: a plain-C caller of libsicelore_gpu.so (no Python, no torch) that replays recorded boundary buffers
 * through the C ABI and compares the records byte for byte with the recorded reference results — what the JNI glue does
 * from Java (java/sicelore_gpu_jni.c), minus the JVM.
 *   abi_driver <file>      file = "SLRB" | u32 kind (1 bc_assign, 2 umi_dist, 3 bc_collide, 4 bc_exact, 5 guided_match, 6 umi_cluster) | kind-specific payload
 * exit code 0 = identical, 1 = mismatch, 2 = bad file, 3 = no CUDA device (the library has no CPU fallback). */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sicelore_gpu.h"

static void *rd(FILE *f, size_t bytes)
{
    void *p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}
/* records compare field bytes only: the trailing padding of a record is not part of the contract */
static int rec_differ(const void *a, const void *b, int64_t n, size_t rec, size_t used)
{
    for (int64_t i = 0; i < n; i++)
        if (memcmp((const char *)a + i * rec, (const char *)b + i * rec, used)) return 1;
    return 0;
}
static int64_t rd64(FILE *f) { int64_t v; if (fread(&v, 8, 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); } return v; }

#define CHECK(call)                                                                    \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != SLR_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, slr_last_error());           \
            return rc_ == SLR_E_NODEVICE ? 3 : 1;                                      \
        }                                                                              \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: abi_driver <file>\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    char magic[4];
    uint32_t kind;
    if (!f || fread(magic, 1, 4, f) != 4 || memcmp(magic, "SLRB", 4) || fread(&kind, 4, 1, f) != 1) { fprintf(stderr, "bad file\n"); return 2; }
    if (slr_abi_version() != SLR_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    slr_ctx *ctx = NULL;
    CHECK(slr_ctx_create(0, 2, &ctx));
    int bad = 0;
    if (kind == 1 || kind == 3 || kind == 4) {
        const int64_t n_keys = rd64(f), has_rank = rd64(f);
        uint64_t *keys = rd(f, (size_t)n_keys * 8);
        int32_t *rank = has_rank ? rd(f, (size_t)n_keys * 4) : NULL;
        slr_bc_table *t = NULL;
        CHECK(slr_bc_table_create(ctx, keys, rank, n_keys, 16, &t));
        if (kind == 3) {
            const int64_t ed = rd64(f), n = rd64(f);
            uint64_t *q = rd(f, (size_t)n * 8);
            slr_collide_result *exp = rd(f, (size_t)n * sizeof(slr_collide_result)), *got = malloc((size_t)n * sizeof(slr_collide_result) + 1);
            CHECK(slr_bc_collide(ctx, t, (int)ed, q, n, got));
            bad = rec_differ(got, exp, n, sizeof(slr_collide_result), offsetof(slr_collide_result, pad));
            printf("bc_collide: %lld barcodes, ED %lld: %s\n", (long long)n, (long long)ed, bad ? "MISMATCH" : "OK");
        } else {
            const int64_t ed = rd64(f), pm = rd64(f), tp = rd64(f), n = rd64(f);
            uint8_t *slices = rd(f, (size_t)n * 32);
            int32_t *anchor = rd(f, (size_t)n * 4);
            slr_bc_result *exp = rd(f, (size_t)n * sizeof(slr_bc_result)), *got = malloc((size_t)n * sizeof(slr_bc_result) + 1);
            if (kind == 1) CHECK(slr_bc_assign(ctx, t, (int)ed, (int)pm, (int)tp, slices, 32, 32, NULL, anchor, n, got));
            else CHECK(slr_bc_exact(ctx, t, (int)tp, slices, 32, 32, NULL, anchor, n, got));
            bad = rec_differ(got, exp, n, sizeof(slr_bc_result), offsetof(slr_bc_result, flags) + sizeof(uint32_t));
            int64_t *counts = malloc((size_t)n_keys * 24 + 8), total = 0, assigned = 0;
            CHECK(slr_bc_counts_read(ctx, t, counts));
            for (int64_t i = 0; i < n_keys * 3; i++) total += counts[i];
            for (int64_t i = 0; i < n; i++) assigned += exp[i].flags & SLR_F_ASSIGNED;
            bad |= total != assigned;
            printf("%s: %lld reads, ED %lld, %lld assigned, counters %lld: %s\n", kind == 1 ? "bc_assign" : "bc_exact", (long long)n,
                   (long long)ed, (long long)assigned, (long long)total, bad ? "MISMATCH" : "OK");
        }
        slr_bc_table_destroy(t);
    } else if (kind == 2) {
        const int64_t umi_len = rd64(f), n_jobs = rd64(f), m = rd64(f), cells = rd64(f);
        uint8_t *umis = rd(f, (size_t)m * 16);
        int64_t *joff = rd(f, (size_t)(n_jobs + 1) * 8), *ooff = rd(f, (size_t)(n_jobs + 1) * 8);
        int32_t *exp = rd(f, (size_t)cells * 4), *got = malloc((size_t)cells * 4 + 4);
        CHECK(slr_umi_dist(ctx, umis, 16, (int)umi_len, joff, n_jobs, got, ooff));
        bad = memcmp(got, exp, (size_t)cells * 4) != 0;
        printf("umi_dist: %lld jobs, %lld reads, %lld cells: %s\n", (long long)n_jobs, (long long)m, (long long)cells, bad ? "MISMATCH" : "OK");
    } else if (kind == 6) {
        const int64_t umi_len = rd64(f), n_jobs = rd64(f), m = rd64(f), cells = rd64(f), ed = rd64(f);
        uint8_t *umis = rd(f, (size_t)m * 16);
        int64_t *joff = rd(f, (size_t)(n_jobs + 1) * 8), *ooff = rd(f, (size_t)(n_jobs + 1) * 8);
        int32_t *exp_m = rd(f, (size_t)cells * 4), *got_m = malloc((size_t)cells * 4 + 4);
        slr_umi_cluster_rec *exp = rd(f, (size_t)m * sizeof(slr_umi_cluster_rec)), *got = malloc((size_t)m * sizeof(slr_umi_cluster_rec) + 1);
        CHECK(slr_umi_cluster(ctx, umis, 16, (int)umi_len, joff, n_jobs, (int)ed, NULL, NULL, got_m, ooff, got));
        bad = memcmp(got_m, exp_m, (size_t)cells * 4) != 0 || memcmp(got, exp, (size_t)m * sizeof(slr_umi_cluster_rec)) != 0;
        memset(got, 0xff, (size_t)m * sizeof(slr_umi_cluster_rec));
        CHECK(slr_umi_cluster(ctx, umis, 16, (int)umi_len, joff, n_jobs, (int)ed, NULL, NULL, NULL, NULL, got));      /* matrices stay on the device */
        bad |= memcmp(got, exp, (size_t)m * sizeof(slr_umi_cluster_rec)) != 0;
        int64_t keys = 0;
        for (int64_t i = 0; i < m; i++) keys += exp[i].best_key >= 0;
        printf("umi_cluster: %lld jobs, %lld reads, ED %lld, %lld keys: %s\n", (long long)n_jobs, (long long)m, (long long)ed, (long long)keys,
               bad ? "MISMATCH" : "OK");
    } else if (kind == 7) {                              /* clustering + UMI assignment (S6): small jobs and jobs above 100 reads in one call */
        const int64_t umi_len = rd64(f), n_jobs = rd64(f), m = rd64(f);
        uint8_t *umis = rd(f, (size_t)m * 16);
        int64_t *joff = rd(f, (size_t)(n_jobs + 1) * 8);
        uint8_t *qv = rd(f, (size_t)n_jobs);
        slr_umi_assign_rec *exp = rd(f, (size_t)m * sizeof(slr_umi_assign_rec)), *got = malloc((size_t)m * sizeof(slr_umi_assign_rec) + 1);
        CHECK(slr_umi_assign(ctx, umis, 16, (int)umi_len, joff, n_jobs, NULL, qv, NULL, NULL, got));
        bad = memcmp(got, exp, (size_t)m * sizeof(slr_umi_assign_rec)) != 0;
        int64_t assigned = 0, deep = 0;
        for (int64_t i = 0; i < m; i++) { assigned += exp[i].flags & SLR_UA_ASSIGNED; deep += (exp[i].flags & SLR_UA_DEEP) != 0; }
        printf("umi_assign: %lld jobs, %lld reads (%lld in jobs above 100 reads), %lld assigned: %s\n", (long long)n_jobs, (long long)m, (long long)deep,
               (long long)assigned, bad ? "MISMATCH" : "OK");
    } else if (kind == 5) {
        const int64_t L = rd64(f), bc = rd64(f), pm = rd64(f), post_len = rd64(f), bailout = rd64(f), slice_len = rd64(f), raw_cap = rd64(f);
        const int64_t n_groups = rd64(f), n_keys = rd64(f), n_all = rd64(f), n_empty = rd64(f), n = rd64(f);
        uint64_t *gk = rd(f, (size_t)n_keys * 8);
        int64_t *go = rd(f, (size_t)(n_groups + 1) * 8);
        uint64_t *ak = rd(f, (size_t)n_all * 8), *ek = rd(f, (size_t)n_empty * 8);
        uint8_t *slices = rd(f, (size_t)n * 32);
        int32_t *anchor = rd(f, (size_t)n * 4), *gid = rd(f, (size_t)n * 4), *ed = rd(f, (size_t)n * 4);
        slr_guided_result *exp = rd(f, (size_t)n * sizeof(slr_guided_result)), *got = malloc((size_t)n * sizeof(slr_guided_result) + 1);
        slr_guided_hit *eraw = rd(f, (size_t)(n * raw_cap) * sizeof(slr_guided_hit)), *graw = malloc((size_t)(n * raw_cap) * sizeof(slr_guided_hit) + 1);
        slr_guided_sets *gs = NULL;
        CHECK(slr_guided_sets_create(ctx, gk, go, n_groups, bc ? ak : NULL, n_all, 3, bc ? ek : NULL, n_empty, 2, (int)bc, (int)L, &gs));
        CHECK(slr_guided_match(ctx, gs, (int)pm, (int)post_len, (int)bailout, slices, 32, (int)slice_len, anchor, gid, ed, n, got, graw, (int)raw_cap));
        bad = rec_differ(got, exp, n, sizeof(slr_guided_result), offsetof(slr_guided_result, pad));
        int64_t found = 0;
        for (int64_t i = 0; i < n; i++) {
            found += exp[i].n_distinct > 0;
            if (!exp[i].flags)                      /* the raw records of a read that throws are unspecified */
                bad |= memcmp(graw + i * raw_cap, eraw + i * raw_cap, (size_t)raw_cap * sizeof(slr_guided_hit)) != 0;
        }
        printf("guided_match: %lld reads, L %lld, %s flavour, %lld found: %s\n", (long long)n, (long long)L, bc ? "BC" : "UMI", (long long)found,
               bad ? "MISMATCH" : "OK");
        slr_guided_sets_destroy(gs);
    } else { fprintf(stderr, "unknown kind %u\n", kind); return 2; }
    slr_ctx_destroy(ctx);
    return bad ? 1 : 0;
}
