/* Drop-in for the reference's hash_table/inc/hash_table.h (same types, same prototypes).  On the hot
 * path this structure is gone (the GPU pair table replaces it; only its iteration ORDER survives, as
 * the tie-break contract); the host keeps it for decompress()'s memo table
 * (bpe.c:12-92) and for callers that use it directly. */
#ifndef HASH_TABLE_H
#define HASH_TABLE_H

#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct node
{
    void *key;
    void *value;
    bool is_free;
    struct node *next;
} node_t;

typedef bool (*hash_value_add)(const void *val_one, const void *val_two, const void *result);

typedef struct
{
    size_t num_of_buckets;
    size_t key_size;
    size_t value_size;
    node_t **buckets;
    node_t *free_nodes;
    size_t num_of_nodes;
} hash_table_t;

hash_table_t *hash_table_create(size_t num_of_buckets, size_t key_size, size_t value_size);
void hash_table_destroy(hash_table_t *table);
bool hash_table_insert(hash_table_t *table, const void *key, const void *value);
bool hash_table_delete(hash_table_t *table, const void *key);
bool hash_table_search(hash_table_t *table, const void *key, void *value);
bool hash_table_clear(hash_table_t *table);
hash_table_t *hash_table_merge(hash_table_t **hash_table_arr, size_t len, hash_value_add add_value, size_t key_size,
                               size_t value_size, size_t new_bucket_num);

#endif /* HASH_TABLE_H */
