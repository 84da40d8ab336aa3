/*
 * TEST INFRASTRUCTURE — not part of the shipped engine.
 *
 * Harness around the UNMODIFIED reference sources (compiled where they lie under
 * $(REF_ROOT), never copied).  It text-includes the reference bpe.c after installing two
 * call-site hooks, so that
 *   - every merge the reference records (bpe.c:753 dyn_arr_set(pair_arr, next_symbol, ..))
 *     is streamed to "<out>.merges" as it happens (long runs can be inspected live), and
 *   - an optional merge cap can stop the loop: once `cap` merges are recorded, the argmax
 *     result handed back to bpe.c:745 has freq = 0, which takes the reference's own
 *     `max.freq <= 1` exit.  The loop body itself is untouched.
 *
 * usage: ref_harness <input-file> <cap|0> <out-prefix>
 *   writes <out>.merges  (LE {u32 a,u32 b} per merge, ids 256.. in order)
 *          <out>.ids     (LE u32 per final token)
 *   prints one line: n_merges n_tokens seconds
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

/* pull the reference's public headers in first (include guards keep the hooks below from
 * renaming the declarations) */
#include REF_BPE_H

static unsigned long long g_cap = 0, g_done = 0;
static FILE *g_merge_log = NULL;

static bool hook_dyn_arr_max(dyn_arr_t *arr, size_t s, size_t e, dyn_compare_t less, void *out)
{
    bool ok = dyn_arr_max(arr, s, e, less, out);
    if (ok && g_cap && g_done >= g_cap)
        ((pair_freq_t *)out)->freq = 0; /* -> reference takes its own freq<=1 exit */
    return ok;
}

static bool hook_dyn_arr_set(dyn_arr_t *arr, size_t index, const void *item)
{
    if (arr && item && arr->item_size == sizeof(pair_t) && index >= 256)
    {
        g_done++;
        if (g_merge_log)
        {
            fwrite(item, sizeof(pair_t), 1, g_merge_log);
            if ((g_done & 63) == 0)
                fflush(g_merge_log);
        }
    }
    return dyn_arr_set(arr, index, item);
}

#define dyn_arr_max hook_dyn_arr_max
#define dyn_arr_set hook_dyn_arr_set
#include REF_BPE_C
#undef dyn_arr_max
#undef dyn_arr_set

int main(int argc, char **argv)
{
    if (argc < 4)
    {
        fprintf(stderr, "usage: %s <input> <cap|0> <out-prefix>\n", argv[0]);
        return 2;
    }
    g_cap = strtoull(argv[2], NULL, 10);
    char path[4096];
    snprintf(path, sizeof path, "%s.merges", argv[3]);
    g_merge_log = fopen(path, "wb");
    if (!g_merge_log)
    {
        perror("fopen merges");
        return 2;
    }

    uint32_t *ids = NULL;
    size_t n_ids = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    dyn_arr_t *pairs = compress(argv[1], &ids, &n_ids);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fclose(g_merge_log);
    if (!pairs)
    {
        fprintf(stderr, "compress() returned NULL\n");
        return 1;
    }
    snprintf(path, sizeof path, "%s.ids", argv[3]);
    FILE *f = fopen(path, "wb");
    if (!f)
    {
        perror("fopen ids");
        return 2;
    }
    fwrite(ids, sizeof(uint32_t), n_ids, f);
    fclose(f);
    double sec = (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
    printf("%llu %zu %.6f\n", g_done, n_ids, sec);
    free(ids);
    dyn_arr_free(pairs);
    return 0;
}
