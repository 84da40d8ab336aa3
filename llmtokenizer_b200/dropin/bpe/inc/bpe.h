/* Drop-in for the reference's bpe/inc/bpe.h: every type and prototype of bpe/inc/bpe.h:14-37 is kept,
 * so the reference's main.c compiles and runs unchanged against this tree.  compress() runs its merge
 * loop (bpe.c:669-783) on the GPU through the C ABI of include/bpe_cuda.h; everything else stays on
 * the host.  The declarations after the "additive" line do not exist in the reference. */
#ifndef BPE_H
#define BPE_H

#include <errno.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../dyn_arr/inc/dyn_arr.h"
#include "../../hash_table/inc/hash_table.h"

typedef struct
{
    uint32_t a, b;
} pair_t;

typedef struct
{
    pair_t pair;
    uint32_t freq;
} pair_freq_t;

char *get_file(const char *path);
bool dump_pairs(const char *path, dyn_arr_t *pair_arr);
dyn_arr_t *read_pairs(const char *path);

void print_text(const uint32_t *text, int length);
void print_graph(dyn_arr_t *pair_arr, const char *png_name, bool add_ascii);

dyn_arr_t *compress(const char *path, uint32_t **encoding, size_t *len);
char *decompress(uint32_t *encoding, size_t len, dyn_arr_t *pair_arr);
void render_pairs(dyn_arr_t *pair_arr);
char *resolve_pair(uint32_t pair_index, dyn_arr_t *pair_arr, hash_table_t *memoization_table);

bool is_less(const void *a, const void *b);

/* ---- additive (not in the reference) ------------------------------------------------------- */
/* compress() with a merge cap (0 = to exhaustion, like compress) and a GPU count.  compress() itself
 * reads the same two knobs from the environment: BPE_MAX_MERGES, BPE_GPUS. */
dyn_arr_t *compress_n(const char *path, uint32_t **encoding, size_t *len, size_t max_merges, int n_gpus);
/* Apply a learned vocabulary (as returned by compress / read_pairs) to another file. */
uint32_t *bpe_encode_file(const char *path, dyn_arr_t *pair_arr, size_t *len, int n_gpus);

#endif /* BPE_H */
