/*
 * Deterministic synthetic corpora for the benchmark configurations of BASELINE.json
 * (SURVEY.md §8d).  splitmix64 everywhere; byte 0x00 is never produced (the reference cuts its
 * input at the first NUL, bpe.c:555).
 *
 *   zipf_ascii : 50,000 words, length U[1,12], letters U[a-z], rank-frequency ~ r^-1.1,
 *                single 0x20 separators, truncated to `size`            (config "100 MB Zipf-word ASCII")
 *   zipf_bytes : 65,536 words, length U[2,10], bytes U[1,255], ~ r^-1.0, no separators
 *                (configs "1 GB byte-level", "10 GB encode", "8 GB"); the word list depends on
 *                `vocab_seed` only, so two corpora can share a vocabulary
 *
 * Built both as a CLI (gen_corpus <kind> <size> <seed> [vocab_seed] <out-file>) and, with
 * -DGEN_CORPUS_LIB, as a tiny shared library used by bench.py to fill a host buffer.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct
{
    uint32_t n_words;
    uint32_t *off; /* n_words + 1 */
    uint8_t *chars;
    uint64_t *cdf; /* cumulative integer weights */
} vocab_t;

static int vocab_build(vocab_t *v, uint32_t n_words, unsigned min_len, unsigned max_len, unsigned lo, unsigned hi,
                       double exponent, uint64_t seed)
{
    uint64_t s = seed;
    v->n_words = n_words;
    v->off = (uint32_t *)malloc(((size_t)n_words + 1) * sizeof(uint32_t));
    v->chars = (uint8_t *)malloc((size_t)n_words * max_len);
    v->cdf = (uint64_t *)malloc((size_t)n_words * sizeof(uint64_t));
    if (!v->off || !v->chars || !v->cdf)
        return -1;
    uint32_t o = 0;
    for (uint32_t w = 0; w < n_words; w++)
    {
        v->off[w] = o;
        unsigned len = min_len + (unsigned)(splitmix64(&s) % (max_len - min_len + 1));
        for (unsigned k = 0; k < len; k++)
            v->chars[o++] = (uint8_t)(lo + splitmix64(&s) % (hi - lo + 1));
    }
    v->off[n_words] = o;
    uint64_t acc = 0;
    for (uint32_t w = 0; w < n_words; w++)
    {
        /* integer weights so the sampler itself is exact integer arithmetic */
        double wt = 1e12 * pow((double)(w + 1), -exponent);
        acc += (uint64_t)wt + 1;
        v->cdf[w] = acc;
    }
    return 0;
}

static void vocab_free(vocab_t *v)
{
    free(v->off);
    free(v->chars);
    free(v->cdf);
}

static inline uint32_t vocab_sample(const vocab_t *v, uint64_t *s)
{
    uint64_t total = v->cdf[v->n_words - 1];
    uint64_t u = (uint64_t)(((unsigned __int128)splitmix64(s) * total) >> 64);
    uint32_t lo = 0, hi = v->n_words - 1;
    while (lo < hi)
    {
        uint32_t mid = lo + (hi - lo) / 2;
        if (v->cdf[mid] > u)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

/* kind 0 = zipf_ascii, 1 = zipf_bytes.  Returns 0 on success. */
int gen_corpus_fill(int kind, uint8_t *out, uint64_t size, uint64_t seed, uint64_t vocab_seed)
{
    vocab_t v;
    int rc;
    if (kind == 0)
        rc = vocab_build(&v, 50000, 1, 12, 'a', 'z', 1.1, vocab_seed);
    else
        rc = vocab_build(&v, 65536, 2, 10, 1, 255, 1.0, vocab_seed);
    if (rc)
        return -1;
    uint64_t s = seed ^ 0xD1B54A32D192ED03ull;
    uint64_t p = 0;
    while (p < size)
    {
        uint32_t w = vocab_sample(&v, &s);
        uint32_t len = v.off[w + 1] - v.off[w];
        for (uint32_t k = 0; k < len && p < size; k++)
            out[p++] = v.chars[v.off[w] + k];
        if (kind == 0 && p < size)
            out[p++] = ' ';
    }
    vocab_free(&v);
    return 0;
}

/* FNV-1a 64 of a buffer: lets bench lines and tests name the exact corpus they ran on */
uint64_t gen_corpus_fnv1a(const uint8_t *p, uint64_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < n; i++)
    {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

#ifndef GEN_CORPUS_LIB
int main(int argc, char **argv)
{
    if (argc < 5)
    {
        fprintf(stderr, "usage: %s zipf_ascii|zipf_bytes <size> <seed> [vocab_seed] <out>\n", argv[0]);
        return 2;
    }
    int kind = strcmp(argv[1], "zipf_ascii") ? 1 : 0;
    uint64_t size = strtoull(argv[2], NULL, 10), seed = strtoull(argv[3], NULL, 10);
    uint64_t vseed = (argc >= 6) ? strtoull(argv[4], NULL, 10) : (kind ? 65536 : 50000);
    const char *path = argv[argc - 1];
    uint8_t *buf = (uint8_t *)malloc(size ? size : 1);
    if (!buf || gen_corpus_fill(kind, buf, size, seed, vseed))
        return 1;
    FILE *f = fopen(path, "wb");
    if (!f)
    {
        perror(path);
        return 1;
    }
    fwrite(buf, 1, size, f);
    fclose(f);
    printf("%s size=%llu seed=%llu vocab_seed=%llu fnv1a=%016llx\n", argv[1], (unsigned long long)size,
           (unsigned long long)seed, (unsigned long long)vseed, (unsigned long long)gen_corpus_fnv1a(buf, size));
    free(buf);
    return 0;
}
#endif
