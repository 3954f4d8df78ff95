// synth.cpp — deterministic synthetic inputs for the hot path (libslr_synth.so, plain C++, no CUDA).
//
// The reference's whitelists (Jar/737K-august-2016.txt, Jar/3M-february-2018.txt.gz) and its test reads
// (Data/fastq_pass) are missing from the mount (/root/reference/.MISSING_LARGE_BLOBS), so every workload of
// BASELINE.json is synthesised here.  Counter-based RNG: read i depends only on (seed, i), so
// any shard of a workload can be generated independently on any rank and is bit-identical everywhere.
//
// Read layout (stranded orientation, what FastqRecordExt.getStrandedSeq() returns, cf. the X= example in
// /root/reference/README.md:400):
//   3' kit:  ...cDNA  polyA  revcomp(UMI)  revcomp(BC)  revcomp(adapter)      adapter = CTACACGACGCTCTTCCGATCT (config.xml:111-113)
//   5' kit:  adapter  BC  UMI  TSO ...
// The boundary slice handed to slr_bc_assign is 32 bytes around the reported adapter end, anchor = 8.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

inline uint64_t splitmix64(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
struct Rng {
    uint64_t s;
    Rng(uint64_t seed, uint64_t idx) : s(seed * 0xD1342543DE82EF95ull + idx * 0x2545F4914F6CDD1Dull + 0x1234567ull) { splitmix64(s); }
    uint64_t next() { return splitmix64(s); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};
const char BASES[4] = {'A', 'G', 'C', 'T'};                  // 2-bit code order of the reference
inline char comp(char c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : 'N'; }
inline void decode16(uint64_t bc, char *out) { for (int i = 0; i < 16; i++) out[i] = BASES[(bc >> (2 * (15 - i))) & 3]; }

}  // namespace

extern "C" {

// n distinct pseudo-random 16-mers (2-bit packed): fmix32 is a bijection on 32 bits, so distinct by construction.
void slr_synth_whitelist(uint64_t *out, int64_t n, uint64_t seed)
{
    const uint32_t c = fmix32((uint32_t)seed ^ 0xA5A5A5A5u) | 1u;
    for (int64_t i = 0; i < n; i++) out[i] = fmix32((uint32_t)i * 0x9E3779B1u + c);
}

struct slr_synth_params {
    double p_sub, p_ins, p_del;     // per-base error rates (default 0.02 / 0.01 / 0.02)
    double p_random;                // fraction of reads with no true barcode (default 0.10)
    double p_n;                     // fraction of reads with an N inside the window (default 0.01)
    double jitter[5];               // P(reported adapter end - true end = -2..+2) (default .02 .08 .8 .08 .02)
    int64_t n_cells;                // barcodes are drawn uniformly from the first n_cells list entries (<=0: whole list)
    int three_prime;
};

void slr_synth_default_params(slr_synth_params *p)
{
    p->p_sub = 0.02; p->p_ins = 0.01; p->p_del = 0.02; p->p_random = 0.10; p->p_n = 0.01;
    const double j[5] = {0.02, 0.08, 0.80, 0.08, 0.02};
    memcpy(p->jitter, j, sizeof(j));
    p->n_cells = 0; p->three_prime = 1;
}

// reads [first, first+n): slices n x 32 bytes, anchor n (always 8), truth n (index into wl or -1), umi n x 12 (true UMI, ASCII)
void slr_synth_reads(const uint64_t *wl, int64_t n_wl, int64_t first, int64_t n, uint64_t seed, const slr_synth_params *prm,
                     uint8_t *slices, int32_t *anchor, int64_t *truth, uint8_t *umi_out)
{
    static const char ADAPTER[] = "CTACACGACGCTCTTCCGATCT";
    const int AL = 22;
    const int64_t ncell = (prm->n_cells > 0 && prm->n_cells < n_wl) ? prm->n_cells : n_wl;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        Rng g(seed, (uint64_t)(first + r));
        char tmpl[96];
        int tl = 0, key_pos;                                       // key_pos: template index of the first adapter base (3') / first BC base (5')
        char bc[16], umi[12];
        const bool random_read = g.uni() < prm->p_random;
        int64_t tidx = -1;
        if (!random_read && ncell > 0) {
            tidx = (int64_t)(g.next() % (uint64_t)ncell);
            decode16(wl[tidx], bc);
        } else {
            for (int i = 0; i < 16; i++) bc[i] = BASES[g.below(4)];
        }
        for (int i = 0; i < 12; i++) umi[i] = BASES[g.below(4)];
        if (umi_out) memcpy(umi_out + r * 12, umi, 12);
        if (prm->three_prime) {
            for (int i = 0; i < 10; i++) tmpl[tl++] = BASES[g.below(4)];                 // cDNA tail
            for (int i = 0; i < 14; i++) tmpl[tl++] = 'A';                               // polyA
            for (int i = 11; i >= 0; i--) tmpl[tl++] = comp(umi[i]);                     // revcomp(UMI)
            for (int i = 15; i >= 0; i--) tmpl[tl++] = comp(bc[i]);                      // revcomp(BC)
            key_pos = tl;
            for (int i = AL - 1; i >= AL - 14; i--) tmpl[tl++] = comp(ADAPTER[i]);       // revcomp(adapter), first 14 bases
        } else {
            for (int i = AL - 14; i < AL; i++) tmpl[tl++] = ADAPTER[i];                  // adapter tail
            key_pos = tl;
            for (int i = 0; i < 16; i++) tmpl[tl++] = bc[i];
            for (int i = 0; i < 12; i++) tmpl[tl++] = umi[i];
            for (int i = 0; i < 24; i++) tmpl[tl++] = BASES[g.below(4)];                 // TSO / cDNA
        }
        // sequencing errors
        char seq[224];
        int sl = 0, key_out = -1;
        for (int i = 0; i < tl; i++) {
            if (i == key_pos) key_out = sl;
            const double u = g.uni();
            if (u < prm->p_del) continue;
            char c = tmpl[i];
            if (u < prm->p_del + prm->p_sub) { char d; do d = BASES[g.below(4)]; while (d == c); c = d; }
            seq[sl++] = c;
            if (g.uni() < prm->p_ins) seq[sl++] = BASES[g.below(4)];
        }
        if (key_out < 0) key_out = sl;
        while (sl < 200) seq[sl++] = BASES[g.below(4)];
        // reported adapter end
        double u = g.uni();
        int jit = 2;
        for (int k = 0; k < 5; k++) { if (u < prm->jitter[k]) { jit = k - 2; break; } u -= prm->jitter[k]; }
        // 3': window = [ap-16, ap) with ap = index of the first adapter base; 5': window = [ap, ap+16) with ap = first BC base
        const int ws = prm->three_prime ? (key_out + jit - 16) : (key_out + jit);
        int s0 = ws - 8;
        if (s0 < 0) s0 = 0;
        uint8_t *dst = slices + r * 32;
        memcpy(dst, seq + s0, 32);
        anchor[r] = ws - s0;
        if (g.uni() < prm->p_n) dst[8 + g.below(16)] = 'N';
        if (truth) truth[r] = tidx;
    }
}

// UMI jobs: n_jobs (cell, region) groups; group sizes ~ geometric(mean) capped at cap (>= 1); inside a group reads
// come from true UMIs (1-3 reads each) with the same error model.  Each read = 16 bytes: umi_len + 2 4-bit codes
// (one flanking base either side of the predicted 12-nt window), the window start is off by -1/+1 with p_shift.
// Pass 1: sizes only (umis == NULL) fills job_offsets[n_jobs+1]; pass 2 fills umis (m x 16).
// first_job: the jobs generated are first_job .. first_job + n_jobs - 1 of the run's global job stream (job j is a pure function of
// (seed, j)), so that every rank of a sharded run — and the checker of a job that crosses a shard boundary — can produce any piece of it.
void slr_synth_umi_jobs_at(int64_t first_job, int64_t n_jobs, double mean, int64_t cap, uint64_t seed, double p_err, double p_shift,
                           int umi_len, int64_t *job_offsets, uint8_t *umis)
{
    static const uint8_t CODE[4] = {1, 2, 4, 8};
    if (!umis) {
        job_offsets[0] = 0;
        for (int64_t j = 0; j < n_jobs; j++) {
            Rng g(seed ^ 0x5151, (uint64_t)(first_job + j));
            int64_t n = 1;
            const double q = 1.0 - 1.0 / mean;
            while (n < cap && g.uni() < q) n++;
            job_offsets[j + 1] = job_offsets[j] + n;
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t j = 0; j < n_jobs; j++) {
        Rng g(seed, (uint64_t)(first_job + j));
        const int64_t r0 = job_offsets[j], n = job_offsets[j + 1] - r0;
        int64_t i = 0;
        while (i < n) {
            uint8_t t[20];
            for (int k = 0; k < umi_len + 6; k++) t[k] = CODE[g.below(4)];          // 3 flank + UMI + 3 flank
            int copies = 1 + (int)g.below(3);
            for (; copies > 0 && i < n; copies--, i++) {
                uint8_t s[40];
                int sl = 0;
                for (int k = 0; k < umi_len + 6; k++) {
                    const double u = g.uni();
                    if (u < p_err * 0.4) continue;                                   // deletion
                    uint8_t c = t[k];
                    if (u < p_err * 0.8) c = CODE[g.below(4)];                       // substitution (may be silent)
                    s[sl++] = c;
                    if (g.uni() < p_err * 0.2) s[sl++] = CODE[g.below(4)];           // insertion
                }
                while (sl < umi_len + 8) s[sl++] = CODE[g.below(4)];
                int start = 2;                                                       // window -1 starts one base before the UMI
                const double u = g.uni();
                if (u < p_shift) start = 1; else if (u < 2 * p_shift) start = 3;
                uint8_t *dst = umis + (r0 + i) * 16;
                memset(dst, 0, 16);
                memcpy(dst, s + start, (size_t)umi_len + 2);
                if (g.uni() < 0.005) dst[g.below((uint32_t)umi_len + 2)] = 15;      // N
            }
        }
    }
}

void slr_synth_umi_jobs(int64_t n_jobs, double mean, int64_t cap, uint64_t seed, double p_err, double p_shift, int umi_len,
                        int64_t *job_offsets, uint8_t *umis)
{
    slr_synth_umi_jobs_at(0, n_jobs, mean, cap, seed, p_err, p_shift, umi_len, job_offsets, umis);
}

}  // extern "C"
