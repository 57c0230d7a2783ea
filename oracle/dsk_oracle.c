/*
 * oracle/dsk_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
 *
 * CPU restatement of the native half of varKoder's image hot path
 * (/root/reference/varKoder/commands/image.py:629-806):
 *   - FASTQ framing and base counting as done by split_fastq (image.py:662-667),
 *   - read sub-sampling (the reference shells out to BBTools reformat.sh, image.py:582-596),
 *   - canonical k-mer counting (the reference shells out to GATB dsk 2.3.3, image.py:771-790),
 *   - the text dump of dsk2ascii (image.py:875-899).
 *
 * PARITY UNPINNED at the dsk / reformat.sh boundary: neither binary nor its source is in
 * /root/reference (conda pins: dsk=2.3.3, bbmap unpinned; conda_environments/linux.yml:10-11) and the
 * reference's own tests hold no expected counts (tests/03_test_installation.sh:88-90 checks exit codes).
 * The rules below restate dsk's published behaviour:
 *   D1 k-mers never span records;            D2 2-bit code (ascii>>1)&3 => A=0 C=1 T=2 G=3;
 *   D3 a window holding a non-ACGT byte is dropped;
 *   D4 canonical = min(forward, reverse complement) under the D2 code;
 *   D5 abundance = total occurrences, every k-mer with count >= 1 reported;
 *   D7 text = "KMER COUNT\n" per canonical k-mer;   D8 4-line records, '@'/'+' legal in qualities.
 * reformat.sh options restated: breaklength=500 (reads cut into consecutive <=500-base pieces),
 * iupacToN=t (any non-ACGT letter behaves as N).  The Java RNG of samplebasestarget cannot be
 * reproduced; the selection rule is this project's own seeded hash (vko_prio), see DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * Index convention of all histograms here: lexicographic, first base most significant, A=0 C=1 G=2 T=3.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int lex_code(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* splitmix64 finaliser over (seed, read index): the project's seeded per-read priority. */
uint64_t vko_prio(uint64_t seed, uint64_t read_index)
{
    uint64_t z = seed + (read_index + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* floor(target * 2^64 / nsites) for target < nsites; UINT64_MAX-saturated "everything" flag is the caller's job. */
uint64_t vko_threshold(uint64_t target, uint64_t nsites)
{
    unsigned __int128 q = ((unsigned __int128)target << 64) / nsites;
    return (uint64_t)q;
}

/*
 * FASTQ framing as split_fastq sees it (image.py:663-667): lines are split on '\n' only; line index
 * mod 4 == 1 is a sequence line; the reference adds len(line)-1 per sequence line, where len()
 * includes the trailing '\n' when there is one -- an unterminated final sequence line is therefore
 * under-counted by one (kept in *nsites_ref; *nsites_true has the real number of bytes).
 * starts/lens (nullable) receive byte offset and length (without '\n') of every sequence line.
 * Returns the number of sequence lines (records whose header line was terminated).
 */
int64_t vko_parse_fastq(const uint8_t* buf, int64_t n, int64_t* starts, int64_t* lens, int64_t cap,
                        int64_t* nsites_ref, int64_t* nsites_true, int64_t* n_lines)
{
    int64_t line = 0, line_start = 0, nreads = 0, ref = 0, tru = 0;
    for (int64_t i = 0; i <= n; ++i) {
        int at_end = (i == n);
        if (at_end && line_start == n) break;           /* no unterminated tail */
        if (at_end || buf[i] == '\n') {
            if ((line & 3) == 1) {
                int64_t len = i - line_start;
                if (starts && nreads < cap) { starts[nreads] = line_start; lens[nreads] = len; }
                ++nreads;
                tru += len;
                ref += at_end ? len - 1 : len;
            }
            ++line;
            line_start = i + 1;
        }
    }
    /* a header line terminated right at EOF opens an empty, unterminated sequence line that Python's
       line iterator never yields: it is not a read for the reference, and holds no k-mer for dsk. */
    if (nsites_ref) *nsites_ref = ref;
    if (nsites_true) *nsites_true = tru;
    if (n_lines) *n_lines = line;
    return nreads;
}

/*
 * dsk restatement: forward-strand histogram over the selected reads, then fold to canonical.
 * select: nullable per-read byte (non-zero = read is in the sub-sample).
 * breaklen: reformat.sh breaklength (0 = off): a read is cut into consecutive pieces of <= breaklen bases
 * before dsk sees it, so no k-mer spans a cut.
 * fwd: uint64[4^k], lexicographic index; ADDED to (caller zeroes).
 */
void vko_count_forward(const uint8_t* buf, const int64_t* starts, const int64_t* lens, int64_t n_reads,
                       int k, int breaklen, const uint8_t* select, uint64_t* fwd, int n_threads)
{
    const uint32_t nk = 1u << (2 * k);
    const uint32_t mask = nk - 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
    {
        uint64_t* h = (uint64_t*)calloc(nk, sizeof(uint64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
        for (int64_t r = 0; r < n_reads; ++r) {
            if (select && !select[r]) continue;
            const uint8_t* s = buf + starts[r];
            const int64_t len = lens[r];
            uint32_t kmer = 0;
            int run = 0;
            for (int64_t i = 0; i < len; ++i) {
                if (breaklen > 0 && i > 0 && (i % breaklen) == 0) run = 0;   /* new piece = new read (D1) */
                int c = lex_code(s[i]);
                if (c < 0) { run = 0; continue; }                             /* D3 */
                kmer = ((kmer << 2) | (uint32_t)c) & mask;
                if (++run >= k) ++h[kmer];
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (uint32_t i = 0; i < nk; ++i) fwd[i] += h[i];
        }
        free(h);
    }
}

/* reverse complement of a lexicographic k-mer index (A<->T, C<->G: code -> 3-code, order reversed) */
uint32_t vko_revcomp(uint32_t x, int k)
{
    uint32_t r = 0;
    for (int i = 0; i < k; ++i) { r = (r << 2) | (3u - (x & 3u)); x >>= 2; }
    return r;
}

/* canon_full[K] = canon_full[rc K] = abundance of the canonical class of K (D4, D5). */
void vko_fold_canonical(const uint64_t* fwd, int k, uint64_t* canon_full)
{
    const uint32_t nk = 1u << (2 * k);
    for (uint32_t x = 0; x < nk; ++x) {
        uint32_t rc = vko_revcomp(x, k);
        canon_full[x] = (rc == x) ? fwd[x] : fwd[x] + fwd[rc];
    }
}

/* dsk's own integer order (D2): A=0 C=1 T=2 G=3, first base most significant */
static uint32_t dsk_code_of_lex(uint32_t x, int k)
{
    static const uint32_t m[4] = {0, 1, 3, 2};   /* lex A,C,G,T -> dsk 0,1,3,2 */
    uint32_t r = 0;
    for (int i = k - 1; i >= 0; --i) r = (r << 2) | m[(x >> (2 * i)) & 3u];
    return r;
}

/*
 * dsk2ascii-format text (D7): one "KMER COUNT\n" line per canonical k-mer with abundance >= 1; the
 * representative printed is the smaller of K / rc(K) under dsk's code order (D4).
 * Returns bytes written (or needed, if > cap).
 */
int64_t vko_dsk2ascii(const uint64_t* canon_full, int k, char* out, int64_t cap)
{
    static const char L[4] = {'A', 'C', 'G', 'T'};
    const uint32_t nk = 1u << (2 * k);
    int64_t w = 0;
    char line[64];
    for (uint32_t x = 0; x < nk; ++x) {
        if (canon_full[x] == 0) continue;
        uint32_t rc = vko_revcomp(x, k);
        if (rc != x && dsk_code_of_lex(rc, k) < dsk_code_of_lex(x, k)) continue;   /* rc is the representative */
        for (int i = 0; i < k; ++i) line[i] = L[(x >> (2 * (k - 1 - i))) & 3u];
        int m = k + snprintf(line + k, sizeof(line) - k, " %llu\n", (unsigned long long)canon_full[x]);
        if (w + m <= cap) memcpy(out + w, line, m);
        w += m;
    }
    return w;
}

/* the whole native half in one call, for the CPU baseline: frame, select by seeded priority, count, fold.
 * levels: thresholds per level (UINT64_MAX with all_flag => every read); out: canon_full per level. */
int64_t vko_count_levels(const uint8_t* buf, int64_t n, int k, int breaklen, uint64_t seed, uint64_t read_index_base,
                         const uint64_t* thr, const uint8_t* take_all, int n_levels, uint64_t* canon_full_out,
                         int n_threads)
{
    int64_t cap = n / 4 + 16;
    int64_t* starts = (int64_t*)malloc(sizeof(int64_t) * cap);
    int64_t* lens = (int64_t*)malloc(sizeof(int64_t) * cap);
    int64_t nr = vko_parse_fastq(buf, n, starts, lens, cap, NULL, NULL, NULL);
    uint8_t* sel = (uint8_t*)malloc(nr > 0 ? nr : 1);
    const uint32_t nk = 1u << (2 * k);
    uint64_t* fwd = (uint64_t*)malloc(sizeof(uint64_t) * nk);
    for (int l = 0; l < n_levels; ++l) {
        for (int64_t r = 0; r < nr; ++r)
            sel[r] = take_all[l] ? 1 : (vko_prio(seed, read_index_base + (uint64_t)r) < thr[l]);
        memset(fwd, 0, sizeof(uint64_t) * nk);
        vko_count_forward(buf, starts, lens, nr, k, breaklen, sel, fwd, n_threads);
        vko_fold_canonical(fwd, k, canon_full_out + (size_t)l * nk);
    }
    free(fwd); free(sel); free(starts); free(lens);
    return nr;
}
