/*
 * TEST INFRASTRUCTURE — command-line front end of the CPU oracle (bpe_oracle.c).
 *   bpe_oracle_cli train  <input> <cap|0> <mode 0|1|2> <out-prefix>
 *   bpe_oracle_cli encode <input> <merges-file> <out-prefix>
 * Environment: BO_WORKERS, BO_CAND_FLOOR select the helper-thread / candidate-list variants of the FAST modes.
 * Output files use the reference's on-disk record (bpe.c:243-339): LE {u32 a,u32 b} per merge
 * from id 256, and LE u32 per token.
 */
#define _POSIX_C_SOURCE 200809L
#include "bpe_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static uint8_t *slurp(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f)
    {
        perror(path);
        return NULL;
    }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    rewind(f);
    uint8_t *buf = (uint8_t *)malloc((size_t)len + 1);
    if (buf && fread(buf, 1, (size_t)len, f) != (size_t)len)
    {
        free(buf);
        buf = NULL;
    }
    fclose(f);
    *n = (size_t)len;
    return buf;
}

static int dump(const char *prefix, const char *ext, const void *p, size_t bytes)
{
    char path[4096];
    snprintf(path, sizeof path, "%s.%s", prefix, ext);
    FILE *f = fopen(path, "wb");
    if (!f)
    {
        perror(path);
        return -1;
    }
    fwrite(p, 1, bytes, f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    /* BO_WORKERS=n: helper threads on every stream length; BO_CAND_FLOOR=c: candidate-list threshold (bpe_oracle.h) */
    if (getenv("BO_WORKERS"))
        bo_set_workers(atoi(getenv("BO_WORKERS")), 0);
    if (getenv("BO_CAND_FLOOR"))
        bo_set_candidate_floor((uint32_t)strtoul(getenv("BO_CAND_FLOOR"), NULL, 10));
    if (argc >= 6 && !strcmp(argv[1], "train"))
    {
        size_t n;
        uint8_t *buf = slurp(argv[2], &n);
        if (!buf)
            return 2;
        bo_pair_t *merges;
        uint32_t *ids;
        size_t nm, ni;
        bo_stats_t st;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int rc = bo_train(buf, n, strtoull(argv[3], NULL, 10), atoi(argv[4]), &merges, &nm, &ids, &ni, &st);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (rc)
        {
            if (rc == BO_ERR_SHORT)
                printf("Error: File contains less than 2 characters\n"); /* bpe.c:560 */
            fprintf(stderr, "bo_train rc=%d\n", rc);
            return 1;
        }
        dump(argv[5], "merges", merges, nm * sizeof(bo_pair_t));
        dump(argv[5], "ids", ids, ni * sizeof(uint32_t));
        double sec = (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
        printf("%zu %zu %.6f ties=%llu edges=%llu faithful=%llu census=%llu D=%llu bt0=%llu bt15=%llu\n", nm, ni, sec,
               (unsigned long long)st.same_bucket_ties, (unsigned long long)st.threshold_edges,
               (unsigned long long)st.faithful_iters, (unsigned long long)st.census_iters,
               (unsigned long long)st.final_distinct, (unsigned long long)st.thread_buckets[0],
               (unsigned long long)st.thread_buckets[15]);
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "encode"))
    {
        size_t n, mb;
        uint8_t *buf = slurp(argv[2], &n);
        uint8_t *mg = slurp(argv[3], &mb);
        if (!buf || !mg)
            return 2;
        uint32_t *ids;
        size_t ni;
        int rc = bo_encode(buf, n, (const bo_pair_t *)mg, mb / sizeof(bo_pair_t), &ids, &ni);
        if (rc)
            return 1;
        dump(argv[4], "ids", ids, ni * sizeof(uint32_t));
        printf("%zu\n", ni);
        return 0;
    }
    fprintf(stderr, "usage: %s train <input> <cap|0> <mode> <out> | encode <input> <merges> <out>\n", argv[0]);
    return 2;
}
