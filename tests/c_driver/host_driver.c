/* host_driver.c — a plain-C caller (no JVM, no Python, no GPU) of the library's host-side entry points: replays recorded inputs of
 * ReadGrouper.groupSams, the job former and the Needleman step through the C ABI and compares with the outputs the reference's own class
 * files produced (dumped from tests/golden/ref_grouper.npz, ref_jobs.npz, ref_needleman.npz by tests/test_abi.py).
 * File: "SLRH" u32 kind, then per kind (all little-endian):
 *   1 groupSams:  i64 n_calls; per call: i64 n, i32 max_dist, i64 id_before, i64 id_after, i32 keep, i32 thrown, i64 n_done,
 *                 i32 position[n], i32 flags[n], u8 has[n], i64 region_in[n], i64 region_out[n]
 *   2 jobs:       i64 n_cases; per case: i64 n, i64 ram, i64 n_jobs, u64 cell[n], i64 region[n], u8 valid[n], i64 offsets[n_jobs + 1], i64 order[offsets[n_jobs]]
 *   3 needleman:  i64 n_pairs; per pair: u64 template, u64 read, i32 len, i32 custom, i32 scores[7], i32 counts[4]
 *   4 used list:  i64 n_cases; per case: i64 n, i32 ed, i32 min_count_fold, i32 cells_fold, i64 record_count, u64 barcodes[n], i32 counts[n],
 *                 slr_collide_result collide[n], u8 kept[n], u8 count_filter_keep[n] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sicelore_host.h"

static FILE *f;
static void rd(void *p, size_t n) { if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
static int64_t rd64(void) { int64_t v; rd(&v, 8); return v; }
static int32_t rd32(void) { int32_t v; rd(&v, 4); return v; }
static void *buf(size_t n) { void *p = malloc(n ? n : 1); if (!p) exit(2); return p; }
#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, " (%s)\n", slr_last_error()); return 1; } while (0)

static int group_sams(void)
{
    int64_t calls = rd64(), reads = 0;
    for (int64_t c = 0; c < calls; c++) {
        int64_t n = rd64(); int32_t md = rd32(); int64_t id0 = rd64(), id1 = rd64(); int32_t keep = rd32(), thrown = rd32(); int64_t n_done = rd64();
        int32_t *pos = buf((size_t)n * 4), *fl = buf((size_t)n * 4); uint8_t *has = buf((size_t)n);
        int64_t *reg = buf((size_t)n * 8), *want = buf((size_t)n * 8);
        rd(pos, (size_t)n * 4); rd(fl, (size_t)n * 4); rd(has, (size_t)n); rd(reg, (size_t)n * 8); rd(want, (size_t)n * 8);
        slr_grouper *g; int64_t last = -7;
        if (slr_grouper_create(md, id0, &g)) FAIL("call %lld: create", (long long)c);
        int rc = slr_grouper_group_sams(g, pos, has, fl, n, keep, reg, &last);
        if (thrown ? rc != SLR_E_REFERENCE_THROWS : rc != SLR_OK) FAIL("call %lld: rc %d, reference threw %d", (long long)c, rc, thrown);
        if (slr_grouper_next_region_id(g) != id1) FAIL("call %lld: region counter %lld != %lld", (long long)c, (long long)slr_grouper_next_region_id(g), (long long)id1);
        if (!thrown && (last + 1 != n_done || memcmp(reg, want, (size_t)n * 8))) FAIL("call %lld: grouping differs (done %lld / %lld)", (long long)c, (long long)(last + 1), (long long)n_done);
        slr_grouper_destroy(g);
        free(pos); free(fl); free(has); free(reg); free(want);
        reads += n;
    }
    printf("groupSams: %lld calls, %lld reads OK\n", (long long)calls, (long long)reads);
    return 0;
}

static int jobs(void)
{
    int64_t cases = rd64(), total = 0;
    for (int64_t c = 0; c < cases; c++) {
        int64_t n = rd64(), ram = rd64(), nj = rd64();
        uint64_t *cell = buf((size_t)n * 8); int64_t *region = buf((size_t)n * 8); uint8_t *valid = buf((size_t)n);
        int64_t *woff = buf((size_t)(nj + 1) * 8);
        rd(cell, (size_t)n * 8); rd(region, (size_t)n * 8); rd(valid, (size_t)n); rd(woff, (size_t)(nj + 1) * 8);
        int64_t *word = buf((size_t)woff[nj] * 8);
        rd(word, (size_t)woff[nj] * 8);
        int64_t *order = buf((size_t)n * 8), *off = buf((size_t)(n + 1) * 8), got = -1;
        if (slr_group_jobs(cell, region, valid, n, 2, ram, order, off, &got)) FAIL("case %lld: slr_group_jobs", (long long)c);
        if (got != nj || memcmp(off, woff, (size_t)(nj + 1) * 8) || memcmp(order, word, (size_t)woff[nj] * 8)) FAIL("case %lld: jobs differ (%lld / %lld)", (long long)c, (long long)got, (long long)nj);
        free(cell); free(region); free(valid); free(woff); free(word); free(order); free(off);
        total += nj;
    }
    printf("job former: %lld cases, %lld jobs OK\n", (long long)cases, (long long)total);
    return 0;
}

static int needleman(void)
{
    int64_t pairs = rd64();
    for (int64_t i = 0; i < pairs; i++) {
        uint64_t t, r; rd(&t, 8); rd(&r, 8);
        int32_t len = rd32(), custom = rd32(); slr_needleman_scores sc; int32_t want[4], got[4];
        rd(&sc, sizeof sc); rd(want, sizeof want);
        if (slr_needleman_errors(t, r, len, custom ? &sc : NULL, got)) FAIL("pair %lld: slr_needleman_errors", (long long)i);
        if (memcmp(got, want, sizeof got)) FAIL("pair %lld: %d %d %d %d, expected %d %d %d %d", (long long)i, got[0], got[1], got[2], got[3], want[0], want[1], want[2], want[3]);
    }
    printf("needleman: %lld alignments OK\n", (long long)pairs);
    return 0;
}

static int used_list(void)
{
    int64_t cases = rd64(), total = 0;
    for (int64_t c = 0; c < cases; c++) {
        int64_t n = rd64(); int32_t ed = rd32(), fold = rd32(), cells = rd32(); int64_t rec = rd64();
        uint64_t *bc = buf((size_t)n * 8); int32_t *cnt = buf((size_t)n * 4); slr_collide_result *col = buf((size_t)n * sizeof *col);
        uint8_t *want = buf((size_t)n), *fwant = buf((size_t)n), *keep = buf((size_t)n), *fkeep = buf((size_t)n); int32_t *rank = buf((size_t)n * 4);
        rd(bc, (size_t)n * 8); rd(cnt, (size_t)n * 4); rd(col, (size_t)n * sizeof *col); rd(want, (size_t)n); rd(fwant, (size_t)n);
        uint32_t flags = 99;
        if (slr_bc_used_merge_collisions(bc, cnt, col, n, fold, ed, cells, keep, rank, &flags)) FAIL("case %lld: slr_bc_used_merge_collisions", (long long)c);
        if (memcmp(keep, want, (size_t)n) || (flags & SLR_UL_ORDER_UNPIN)) FAIL("case %lld: kept list differs", (long long)c);
        if (slr_bc_used_filter_low_counts(cnt, n, rec, fkeep) || memcmp(fkeep, fwant, (size_t)n)) FAIL("case %lld: count filter differs", (long long)c);
        free(bc); free(cnt); free(col); free(want); free(fwant); free(keep); free(fkeep); free(rank);
        total += n;
    }
    printf("used list: %lld lists, %lld barcodes OK\n", (long long)cases, (long long)total);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc != 2 || !(f = fopen(argv[1], "rb"))) { fprintf(stderr, "usage: host_driver file\n"); return 2; }
    char magic[4]; rd(magic, 4);
    if (memcmp(magic, "SLRH", 4)) { fprintf(stderr, "bad magic\n"); return 2; }
    int32_t kind = rd32();
    return kind == 1 ? group_sams() : kind == 2 ? jobs() : kind == 3 ? needleman() : kind == 4 ? used_list() : 2;
}
