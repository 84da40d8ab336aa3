/* Drop-in for the reference's dyn_arr/inc/dyn_arr.h (same type, same prototypes): the result container
 * compress() hands back (main.c:23 frees it with dyn_arr_free) and the `first maximum wins`
 * scan (dyn_arr.c:136-181).  Chunked array of fixed-size items, 256 items per node. */
#ifndef DYN_ARR_H
#define DYN_ARR_H

#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#define MAX_NODE_SIZE (1U << 8)

typedef struct
{
    size_t len;        /* nodes in the pointer table            */
    size_t last_index; /* highest index ever set                */
    size_t item_size;  /* bytes per item                        */
    void **nodes;      /* node pointers (NULL until first used) */
} dyn_arr_t;

typedef bool (*dyn_compare_t)(const void *a, const void *b);

dyn_arr_t *dyn_arr_create(size_t min_size, size_t item_size);
void dyn_arr_free(dyn_arr_t *dyn_arr);
bool dyn_arr_set(dyn_arr_t *dyn_arr, size_t index, const void *item);
bool dyn_arr_append(dyn_arr_t *dyn_arr, const void *item);
bool dyn_arr_get(dyn_arr_t *dyn_arr, size_t index, void *output);
bool dyn_arr_sort(dyn_arr_t *dyn_arr, size_t start_index, size_t end_index, dyn_compare_t compare);
bool dyn_arr_max(dyn_arr_t *dyn_arr, size_t start_index, size_t end_index, dyn_compare_t is_less, void *output);
bool dyn_arr_min(dyn_arr_t *dyn_arr, size_t start_index, size_t end_index, dyn_compare_t is_less, void *output);

#endif /* DYN_ARR_H */
