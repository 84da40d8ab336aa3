/* Host-side container behind the reference's dyn_arr API (dyn_arr/inc/dyn_arr.h:27-85).  Written
 * from the interface; none of the reference's quirks (log2 growth, the double free in its sort) are
 * kept, only the observable contract: last_index = highest index set, max/min keep the FIRST
 * extremal element (dyn_arr.c:163-174). */
#include "../inc/dyn_arr.h"

static bool ensure_node(dyn_arr_t *d, size_t node)
{
    if (node >= d->len)
    {
        size_t nl = d->len ? d->len : 1;
        while (nl <= node)
            nl *= 2;
        void **nn = (void **)realloc(d->nodes, nl * sizeof(void *));
        if (!nn)
            return false;
        for (size_t i = d->len; i < nl; i++)
            nn[i] = NULL;
        d->nodes = nn;
        d->len = nl;
    }
    if (!d->nodes[node])
    {
        d->nodes[node] = calloc(MAX_NODE_SIZE, d->item_size);
        if (!d->nodes[node])
            return false;
    }
    return true;
}

static void *slot(const dyn_arr_t *d, size_t index)
{
    const size_t node = index / MAX_NODE_SIZE;
    if (node >= d->len || !d->nodes[node])
        return NULL;
    return (char *)d->nodes[node] + (index % MAX_NODE_SIZE) * d->item_size;
}

dyn_arr_t *dyn_arr_create(size_t min_size, size_t item_size)
{
    if (!item_size)
        return NULL;
    dyn_arr_t *d = (dyn_arr_t *)calloc(1, sizeof *d);
    if (!d)
        return NULL;
    d->item_size = item_size;
    const size_t nodes = (min_size + MAX_NODE_SIZE - 1) / MAX_NODE_SIZE;
    for (size_t i = 0; i < nodes; i++)
        if (!ensure_node(d, i))
        {
            dyn_arr_free(d);
            return NULL;
        }
    return d;
}

void dyn_arr_free(dyn_arr_t *d)
{
    if (!d)
        return;
    for (size_t i = 0; i < d->len; i++)
        free(d->nodes[i]);
    free(d->nodes);
    free(d);
}

bool dyn_arr_set(dyn_arr_t *d, size_t index, const void *item)
{
    if (!d || !item || !ensure_node(d, index / MAX_NODE_SIZE))
        return false;
    memcpy(slot(d, index), item, d->item_size);
    if (index > d->last_index)
        d->last_index = index;
    return true;
}

bool dyn_arr_append(dyn_arr_t *d, const void *item)
{
    if (!d)
        return false;
    /* an array nothing was ever set in has last_index 0 and no item at 0 */
    const bool empty = d->last_index == 0 && (d->len == 0 || !d->nodes[0]);
    return dyn_arr_set(d, empty ? 0 : d->last_index + 1, item);
}

bool dyn_arr_get(dyn_arr_t *d, size_t index, void *output)
{
    if (!d || !output || index > d->last_index)
        return false;
    const void *p = slot(d, index);
    if (!p)
        return false;
    memcpy(output, p, d->item_size);
    return true;
}

static bool extremum(dyn_arr_t *d, size_t lo, size_t hi, dyn_compare_t is_less, void *out, bool want_max)
{
    if (!d || !is_less || !out || lo > hi || hi > d->last_index)
        return false;
    const void *best = slot(d, lo);
    if (!best)
        return false;
    for (size_t i = lo + 1; i <= hi; i++)
    {
        const void *p = slot(d, i);
        if (!p)
            return false;
        /* strict comparison: among equal elements the first one stays (bpe.c:4-10 relies on it) */
        if (want_max ? is_less(best, p) : is_less(p, best))
            best = p;
    }
    memcpy(out, best, d->item_size);
    return true;
}

bool dyn_arr_max(dyn_arr_t *d, size_t lo, size_t hi, dyn_compare_t is_less, void *out) { return extremum(d, lo, hi, is_less, out, true); }
bool dyn_arr_min(dyn_arr_t *d, size_t lo, size_t hi, dyn_compare_t is_less, void *out) { return extremum(d, lo, hi, is_less, out, false); }

bool dyn_arr_sort(dyn_arr_t *d, size_t lo, size_t hi, dyn_compare_t before)
{
    if (!d || !before || lo > hi || hi > d->last_index)
        return false;
    const size_t n = hi - lo + 1, sz = d->item_size;
    char *tmp = (char *)malloc(n * sz), *aux = (char *)malloc(n * sz);
    if (!tmp || !aux)
    {
        free(tmp);
        free(aux);
        return false;
    }
    for (size_t i = 0; i < n; i++)
        memcpy(tmp + i * sz, slot(d, lo + i), sz);
    /* bottom-up merge sort: stable, no recursion */
    for (size_t w = 1; w < n; w *= 2)
    {
        for (size_t s = 0; s < n; s += 2 * w)
        {
            size_t i = s, m = s + w < n ? s + w : n, j = m, e = s + 2 * w < n ? s + 2 * w : n, k = s;
            while (i < m && j < e)
                memcpy(aux + (k++) * sz, before(tmp + j * sz, tmp + i * sz) ? tmp + (j++) * sz : tmp + (i++) * sz, sz);
            while (i < m)
                memcpy(aux + (k++) * sz, tmp + (i++) * sz, sz);
            while (j < e)
                memcpy(aux + (k++) * sz, tmp + (j++) * sz, sz);
        }
        char *t = tmp;
        tmp = aux;
        aux = t;
    }
    for (size_t i = 0; i < n; i++)
        memcpy(slot(d, lo + i), tmp + i * sz, sz);
    free(tmp);
    free(aux);
    return true;
}
