/* Chained hash map behind the reference's hash_table API (hash_table/inc/hash_table.h:29-37).
 * Same observable rules as the reference: murmur3_32 with seed 0x9747b28c (hash_table.c:5,8-53),
 * new keys go to the head of their chain (:300-302), the table doubles at the top of an insert once
 * nodes >= 0.3 * buckets (:248-254), clear keeps the bucket count and recycles nodes (:310-338). */
#include "../inc/hash_table.h"

#include <string.h>

static uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static uint32_t murmur3(const void *key, size_t len)
{
    const uint8_t *p = (const uint8_t *)key;
    uint32_t h = 0x9747b28cu;
    size_t i = 0;
    for (; i + 4 <= len; i += 4)
    {
        uint32_t k;
        memcpy(&k, p + i, 4);
        k *= 0xcc9e2d51u;
        k = rotl(k, 15);
        k *= 0x1b873593u;
        h ^= k;
        h = rotl(h, 13);
        h = h * 5u + 0xe6546b64u;
    }
    uint32_t k = 0;
    switch (len & 3)
    {
    case 3:
        k ^= (uint32_t)p[i + 2] << 16; /* fall through */
    case 2:
        k ^= (uint32_t)p[i + 1] << 8; /* fall through */
    case 1:
        k ^= p[i];
        k *= 0xcc9e2d51u;
        k = rotl(k, 15);
        k *= 0x1b873593u;
        h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

hash_table_t *hash_table_create(size_t buckets, size_t key_size, size_t value_size)
{
    if (!buckets || !key_size || !value_size)
        return NULL;
    hash_table_t *t = (hash_table_t *)calloc(1, sizeof *t);
    if (!t)
        return NULL;
    t->buckets = (node_t **)calloc(buckets, sizeof(node_t *));
    if (!t->buckets)
    {
        free(t);
        return NULL;
    }
    t->num_of_buckets = buckets;
    t->key_size = key_size;
    t->value_size = value_size;
    return t;
}

static void free_chain(node_t *n)
{
    while (n)
    {
        node_t *nx = n->next;
        free(n->key);
        free(n->value);
        free(n);
        n = nx;
    }
}

void hash_table_destroy(hash_table_t *t)
{
    if (!t)
        return;
    for (size_t b = 0; b < t->num_of_buckets; b++)
        free_chain(t->buckets[b]);
    free_chain(t->free_nodes);
    free(t->buckets);
    free(t);
}

static bool grow(hash_table_t *t)
{
    const size_t nb = t->num_of_buckets * 2;
    node_t **nbk = (node_t **)calloc(nb, sizeof(node_t *));
    if (!nbk)
        return false;
    /* bucket 0.., head to tail, each node to the head of its new chain (reverses chains, as the reference does) */
    for (size_t b = 0; b < t->num_of_buckets; b++)
        for (node_t *n = t->buckets[b]; n;)
        {
            node_t *nx = n->next;
            const size_t i = murmur3(n->key, t->key_size) % nb;
            n->next = nbk[i];
            nbk[i] = n;
            n = nx;
        }
    free(t->buckets);
    t->buckets = nbk;
    t->num_of_buckets = nb;
    return true;
}

bool hash_table_insert(hash_table_t *t, const void *key, const void *value)
{
    if (!t || !key || !value)
        return false;
    if ((double)t->num_of_nodes >= 0.3 * (double)t->num_of_buckets && !grow(t))
        return false;
    const size_t i = murmur3(key, t->key_size) % t->num_of_buckets;
    for (node_t *n = t->buckets[i]; n; n = n->next)
        if (!memcmp(n->key, key, t->key_size))
        {
            memcpy(n->value, value, t->value_size);
            return true;
        }
    node_t *n = t->free_nodes;
    if (n)
        t->free_nodes = n->next;
    else
    {
        n = (node_t *)calloc(1, sizeof *n);
        if (!n)
            return false;
        n->key = malloc(t->key_size);
        n->value = malloc(t->value_size);
        if (!n->key || !n->value)
        {
            free(n->key);
            free(n->value);
            free(n);
            return false;
        }
    }
    memcpy(n->key, key, t->key_size);
    memcpy(n->value, value, t->value_size);
    n->is_free = false;
    n->next = t->buckets[i];
    t->buckets[i] = n;
    t->num_of_nodes++;
    return true;
}

bool hash_table_search(hash_table_t *t, const void *key, void *value)
{
    if (!t || !key)
        return false;
    for (node_t *n = t->buckets[murmur3(key, t->key_size) % t->num_of_buckets]; n; n = n->next)
        if (!memcmp(n->key, key, t->key_size))
        {
            if (value)
                memcpy(value, n->value, t->value_size);
            return true;
        }
    return false;
}

bool hash_table_delete(hash_table_t *t, const void *key)
{
    if (!t || !key)
        return false;
    node_t **pp = &t->buckets[murmur3(key, t->key_size) % t->num_of_buckets];
    for (; *pp; pp = &(*pp)->next)
        if (!memcmp((*pp)->key, key, t->key_size))
        {
            node_t *n = *pp;
            *pp = n->next;
            n->is_free = true;
            n->next = t->free_nodes;
            t->free_nodes = n;
            t->num_of_nodes--;
            return true;
        }
    return false;
}

bool hash_table_clear(hash_table_t *t)
{
    if (!t)
        return false;
    for (size_t b = 0; b < t->num_of_buckets; b++)
    {
        for (node_t *n = t->buckets[b]; n;)
        {
            node_t *nx = n->next;
            n->is_free = true;
            n->next = t->free_nodes;
            t->free_nodes = n;
            n = nx;
        }
        t->buckets[b] = NULL;
    }
    t->num_of_nodes = 0;
    return true;
}

hash_table_t *hash_table_merge(hash_table_t **arr, size_t len, hash_value_add add_value, size_t key_size, size_t value_size,
                               size_t new_bucket_num)
{
    if (!arr || !add_value)
        return NULL;
    hash_table_t *m = hash_table_create(new_bucket_num, key_size, value_size);
    void *cur = malloc(value_size), *sum = malloc(value_size);
    if (!m || !cur || !sum)
        goto fail;
    /* table 0.., bucket 0.., chain head to tail: this traversal is what fixes the reference's tie-break */
    for (size_t k = 0; k < len; k++)
    {
        const hash_table_t *t = arr[k];
        if (!t || t->key_size != key_size || t->value_size != value_size)
            goto fail;
        for (size_t b = 0; b < t->num_of_buckets; b++)
            for (const node_t *n = t->buckets[b]; n; n = n->next)
            {
                if (hash_table_search(m, n->key, cur))
                {
                    if (!add_value(cur, n->value, sum) || !hash_table_insert(m, n->key, sum))
                        goto fail;
                }
                else if (!hash_table_insert(m, n->key, n->value))
                    goto fail;
            }
    }
    free(cur);
    free(sum);
    return m;
fail:
    free(cur);
    free(sum);
    hash_table_destroy(m);
    return NULL;
}
