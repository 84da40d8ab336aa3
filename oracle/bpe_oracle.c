/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference's BPE merge loop (see bpe_oracle.h).
 * Never linked into, or called from, the shipped engine.
 *
 * Citations are reference file:line (relative to /root/reference).
 *
 * Three ways of choosing the next merge are implemented and cross-checked by the tests:
 *   FAITHFUL  every iteration recounts the stream into 16 emulated chained "thread" tables
 *             (bpe.c:428-527), folds them into a fresh 65,536-bucket table in the reference's
 *             traversal order (hash_table.c:146-188), walks that table bucket by bucket, chain
 *             head to tail (bpe.c:705-728) and keeps the first maximum (dyn_arr.c:163-174,
 *             bpe.c:4-10).  This is the literal restatement.
 *   FAST      keeps pair counts incrementally and uses the closed form of that
 *             order: max frequency, then smallest `murmur3 % B(D)`.  When two or more maximal
 *             pairs share the winning bucket, or D sits exactly on a resize threshold, it runs
 *             the FAITHFUL selection for that one iteration.
 *   FAST_CF   as FAST, but those iterations are resolved by the closed-form chain-order ranking
 *             (rank of first sight + parity of later resizes) that the CUDA engine implements.
 *
 * Work split above 1,048,576 tokens (bpe.c:479-520) is a race in the reference (which thread
 * counts which 64 Ki chunk).  The canonical schedule used here — and by the engine — is the
 * legal schedule in which worker 0 takes every chunk, in order.  It never changes the bucket
 * order (which decides all but ~0.3 % of merges); it only fixes the chain order inside one
 * bucket and the workers' persistent bucket counts.
 */
#include "bpe_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define BO_THREADS 16u                           /* bpe.c:409 */
#define BO_CHUNK (64u * 1024u)                   /* bpe.c:423 */
#define BO_STATIC_LIMIT ((size_t)BO_CHUNK * BO_THREADS) /* bpe.c:449 */
#define BO_THREAD_BUCKETS 256u                   /* bpe.c:610 */
#define BO_MERGED_BUCKETS 65536u                 /* bpe.c:611 */
#define BO_SENT 0xFFFFFFFFu                      /* "no token here" in halo windows */

/* ------------------------------------------------------------------------------------------ */
/* Optional helper threads for the FAST modes' two whole-array loops (argmax scan of the count
 * map, rewrite + delta emission).  They only split the same loops over index ranges and combine
 * the partial results in range order, so every result is the one the sequential loops give
 * (tests/test_oracle.py checks that); default: off.  Used offline for the full-size fixtures
 * (tools/make_full_golden.py), where a 1 GB corpus costs ~1 s per merge on one core. */
#define BO_MAX_WORKERS 64
static int g_workers = 1;
static size_t g_workers_min_n = (size_t)1 << 22;

void bo_set_workers(int workers, size_t min_tokens)
{
    g_workers = workers < 1 ? 1 : (workers > BO_MAX_WORKERS ? BO_MAX_WORKERS : workers);
    g_workers_min_n = min_tokens;
}

typedef struct
{
    void (*fn)(void *ctx, int w, int nw);
    void *ctx;
    int w, nw;
} par_job_t;

static void *par_tramp(void *p)
{
    par_job_t *j = (par_job_t *)p;
    j->fn(j->ctx, j->w, j->nw);
    return NULL;
}

/* fork-join: fn(ctx, w, nw) for w = 0..nw-1, worker 0 on the calling thread */
static void par_run(void (*fn)(void *, int, int), void *ctx, int nw)
{
    pthread_t th[BO_MAX_WORKERS];
    par_job_t job[BO_MAX_WORKERS];
    int started[BO_MAX_WORKERS];
    for (int w = 1; w < nw; w++)
    {
        job[w].fn = fn;
        job[w].ctx = ctx;
        job[w].w = w;
        job[w].nw = nw;
        started[w] = (pthread_create(&th[w], NULL, par_tramp, &job[w]) == 0);
    }
    fn(ctx, 0, nw);
    for (int w = 1; w < nw; w++)
    {
        if (started[w])
            pthread_join(th[w], NULL);
        else
            fn(ctx, w, nw); /* no thread to be had: do the share here */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* hash_table.c:8-53 specialised to the 8-byte key {u32 a; u32 b} (two blocks, no tail)        */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t bo_murmur3_pair(uint32_t a, uint32_t b)
{
    uint32_t h = 0x9747b28cu; /* hash_table.c:5 */
    uint32_t k = a;
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5u + 0xe6546b64u;
    k = b;
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5u + 0xe6546b64u;
    h ^= 8u; /* key_size, hash_table.c:45 */
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

/* smallest node count at which `nodes >= 0.3 * buckets` holds (hash_table.c:6,248) */
static inline int resize_due(uint64_t nodes, uint64_t buckets) { return (double)nodes >= 0.3 * (double)buckets; }

static uint64_t resize_threshold(uint64_t buckets)
{
    uint64_t t = (uint64_t)(0.3 * (double)buckets);
    while (!resize_due(t, buckets))
        t++;
    while (t > 0 && resize_due(t - 1, buckets))
        t--;
    return t;
}

/* B(D) away from the exact-threshold edge: a doubling for bucket count B has happened iff some
 * insert call saw nodes >= thr(B), which is certain once D > thr(B) */
uint64_t bo_merged_buckets(uint64_t distinct)
{
    uint64_t b = BO_MERGED_BUCKETS;
    while (distinct > resize_threshold(b))
        b *= 2;
    return b;
}

static inline uint64_t pack_key(uint32_t a, uint32_t b) { return (uint64_t)a | ((uint64_t)b << 32); }

/* ------------------------------------------------------------------------------------------ */
/* emulated chained table: head insertion (hash_table.c:300-302), doubling checked at the top of
 * every insert (hash_table.c:248-254), resize re-inserts bucket 0.., head->tail, each at the
 * head of its new chain (hash_table.c:208-223), clear keeps the bucket count (hash_table.c:310-338) */
typedef struct
{
    uint64_t nb;
    uint64_t head_cap;
    int32_t *head;
    uint64_t *key;
    uint64_t *val;
    uint32_t *hsh;
    int32_t *next;
    size_t nn, cap;
} cht_t;

static int cht_set_buckets(cht_t *t, uint64_t nb)
{
    if (nb > t->head_cap)
    {
        int32_t *h = (int32_t *)realloc(t->head, nb * sizeof(int32_t));
        if (!h)
            return -1;
        t->head = h;
        t->head_cap = nb;
    }
    t->nb = nb;
    memset(t->head, 0xff, nb * sizeof(int32_t));
    t->nn = 0;
    return 0;
}

static void cht_release(cht_t *t)
{
    free(t->head);
    free(t->key);
    free(t->val);
    free(t->hsh);
    free(t->next);
    memset(t, 0, sizeof *t);
}

static int cht_resize(cht_t *t, uint64_t nnb)
{
    int32_t *nh = (int32_t *)malloc(nnb * sizeof(int32_t));
    if (!nh)
        return -1;
    memset(nh, 0xff, nnb * sizeof(int32_t));
    for (uint64_t i = 0; i < t->nb; i++)
    {
        int32_t c = t->head[i];
        while (c >= 0)
        {
            int32_t nx = t->next[c];
            uint64_t bk = (uint64_t)t->hsh[c] % nnb;
            t->next[c] = nh[bk];
            nh[bk] = c;
            c = nx;
        }
    }
    free(t->head);
    t->head = nh;
    t->head_cap = nnb;
    t->nb = nnb;
    return 0;
}

/* search + insert(old+add) as get_freq (bpe.c:465-470) and hash_table_merge (hash_table.c:156-183) do */
static int cht_add(cht_t *t, uint64_t key, uint32_t h, uint64_t add)
{
    if (resize_due(t->nn, t->nb))
        if (cht_resize(t, t->nb * 2))
            return -1;
    uint64_t bk = (uint64_t)h % t->nb;
    for (int32_t c = t->head[bk]; c >= 0; c = t->next[c])
        if (t->key[c] == key)
        {
            t->val[c] += add;
            return 0;
        }
    if (t->nn == t->cap)
    {
        size_t nc = t->cap ? t->cap * 2 : 1024;
        uint64_t *k = (uint64_t *)realloc(t->key, nc * sizeof(uint64_t));
        if (k)
            t->key = k;
        uint64_t *v = (uint64_t *)realloc(t->val, nc * sizeof(uint64_t));
        if (v)
            t->val = v;
        uint32_t *hh = (uint32_t *)realloc(t->hsh, nc * sizeof(uint32_t));
        if (hh)
            t->hsh = hh;
        int32_t *nx = (int32_t *)realloc(t->next, nc * sizeof(int32_t));
        if (nx)
            t->next = nx;
        if (!k || !v || !hh || !nx)
            return -1;
        t->cap = nc;
    }
    int32_t idx = (int32_t)t->nn++;
    t->key[idx] = key;
    t->val[idx] = add;
    t->hsh[idx] = h;
    t->next[idx] = t->head[bk];
    t->head[bk] = idx;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* incremental pair-count map (open addressing) used by the FAST modes                          */
#define PM_EMPTY UINT64_MAX
typedef struct
{
    uint64_t *key;
    uint32_t *cnt;
    uint32_t *hsh;
    uint64_t cap, used;
    uint64_t distinct; /* D: entries with cnt > 0 */
    /* argmax shortcut: the slots whose count has reached cand_T since the list was made (0: no
     * list).  Only an increment can carry a count over cand_T, and pmap_add offers those, so while
     * the largest count in the list is >= cand_T the list holds every maximal entry. */
    uint64_t *cand, *inlist;
    uint64_t cand_n, cand_cap;
    uint32_t cand_T;
} pmap_t;

static uint32_t g_cand_floor = 64; /* a list is only made while the maximum count is at least this; 0: never */

void bo_set_candidate_floor(uint32_t min_count) { g_cand_floor = min_count; }

static int pmap_init(pmap_t *m, uint64_t cap)
{
    m->cap = cap;
    m->used = 0;
    m->distinct = 0;
    m->cand = m->inlist = NULL;
    m->cand_n = m->cand_cap = 0;
    m->cand_T = 0;
    m->key = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->cnt = (uint32_t *)calloc(cap, sizeof(uint32_t));
    m->hsh = (uint32_t *)malloc(cap * sizeof(uint32_t));
    if (!m->key || !m->cnt || !m->hsh)
        return -1;
    for (uint64_t i = 0; i < cap; i++)
        m->key[i] = PM_EMPTY;
    return 0;
}

static void pmap_release(pmap_t *m)
{
    free(m->key);
    free(m->cnt);
    free(m->hsh);
    free(m->cand);
    free(m->inlist);
    memset(m, 0, sizeof *m);
}

static int cand_push(pmap_t *m, uint64_t s)
{
    if (m->inlist[s >> 6] >> (s & 63) & 1)
        return 0;
    m->inlist[s >> 6] |= 1ull << (s & 63);
    if (m->cand_n == m->cand_cap)
    {
        uint64_t nc = m->cand_cap ? m->cand_cap * 2 : 4096;
        uint64_t *c2 = (uint64_t *)realloc(m->cand, nc * sizeof(uint64_t));
        if (!c2)
            return -1;
        m->cand = c2;
        m->cand_cap = nc;
    }
    m->cand[m->cand_n++] = s;
    return 0;
}

/* (re)build the list for threshold T from the whole map */
static int cand_build(pmap_t *m, uint32_t T)
{
    if (!m->inlist)
        m->inlist = (uint64_t *)malloc((m->cap / 64 + 1) * sizeof(uint64_t));
    if (!m->inlist)
        return -1;
    memset(m->inlist, 0, (m->cap / 64 + 1) * sizeof(uint64_t));
    m->cand_n = 0;
    m->cand_T = 0;
    for (uint64_t s = 0; s < m->cap; s++)
        if (m->key[s] != PM_EMPTY && m->cnt[s] >= T)
            if (cand_push(m, s))
                return -1;
    m->cand_T = T;
    return 0;
}

static int64_t pmap_find(const pmap_t *m, uint64_t key, uint32_t h)
{
    uint64_t s = ((uint64_t)h * 0x9E3779B97F4A7C15ull >> 20) & (m->cap - 1);
    while (m->key[s] != PM_EMPTY)
    {
        if (m->key[s] == key)
            return (int64_t)s;
        s = (s + 1) & (m->cap - 1);
    }
    return -1;
}

static int pmap_grow(pmap_t *m);

static int64_t pmap_upsert(pmap_t *m, uint64_t key, uint32_t h)
{
    if ((m->used + 1) * 2 > m->cap)
        if (pmap_grow(m))
            return -1;
    uint64_t s = ((uint64_t)h * 0x9E3779B97F4A7C15ull >> 20) & (m->cap - 1);
    while (m->key[s] != PM_EMPTY)
    {
        if (m->key[s] == key)
            return (int64_t)s;
        s = (s + 1) & (m->cap - 1);
    }
    m->key[s] = key;
    m->hsh[s] = h;
    m->cnt[s] = 0;
    m->used++;
    return (int64_t)s;
}

static int pmap_grow(pmap_t *m)
{
    pmap_t n;
    uint64_t want = m->cap;
    while (m->distinct * 4 + 16 > want)
        want *= 2;
    if (want == m->cap && m->used * 2 + 2 > m->cap && m->distinct * 4 + 16 > m->cap / 2)
        want *= 2;
    if (pmap_init(&n, want))
        return -1;
    for (uint64_t i = 0; i < m->cap; i++)
        if (m->key[i] != PM_EMPTY && m->cnt[i])
        {
            uint64_t s = ((uint64_t)m->hsh[i] * 0x9E3779B97F4A7C15ull >> 20) & (n.cap - 1);
            while (n.key[s] != PM_EMPTY)
                s = (s + 1) & (n.cap - 1);
            n.key[s] = m->key[i];
            n.hsh[s] = m->hsh[i];
            n.cnt[s] = m->cnt[i];
            n.used++;
        }
    n.distinct = m->distinct;
    pmap_release(m);
    *m = n;
    return 0;
}

static int pmap_add(pmap_t *m, uint32_t a, uint32_t b, int64_t d)
{
    if (!d)
        return 0;
    uint64_t key = pack_key(a, b);
    uint32_t h = bo_murmur3_pair(a, b);
    int64_t s = pmap_upsert(m, key, h);
    if (s < 0)
        return -1;
    uint32_t before = m->cnt[s];
    uint32_t after = (uint32_t)((int64_t)before + d);
    m->cnt[s] = after;
    if (m->cand_T && before < m->cand_T && after >= m->cand_T && cand_push(m, (uint64_t)s))
        return -1;
    if (!before && after)
        m->distinct++;
    if (before && !after)
        m->distinct--;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* epoch-stamped hash set used by the census and the closed-form resolver                      */
typedef struct
{
    uint64_t *key;
    uint32_t *stamp;
    uint32_t *aux;
    uint64_t cap;
    uint32_t epoch;
} hset_t;

static int hset_reserve(hset_t *s, uint64_t items)
{
    uint64_t want = 1024;
    while (want < items * 2 + 2)
        want *= 2;
    if (want <= s->cap)
        return 0;
    free(s->key);
    free(s->stamp);
    free(s->aux);
    s->key = (uint64_t *)malloc(want * sizeof(uint64_t));
    s->stamp = (uint32_t *)calloc(want, sizeof(uint32_t));
    s->aux = (uint32_t *)malloc(want * sizeof(uint32_t));
    s->cap = want;
    s->epoch = 0;
    return (s->key && s->stamp && s->aux) ? 0 : -1;
}

static void hset_next_epoch(hset_t *s)
{
    if (++s->epoch == 0)
    {
        memset(s->stamp, 0, s->cap * sizeof(uint32_t));
        s->epoch = 1;
    }
}

/* returns slot; *fresh = 1 if the key was not present in this epoch */
static uint64_t hset_put(hset_t *s, uint64_t key, uint32_t h, int *fresh)
{
    uint64_t p = ((uint64_t)h * 0x9E3779B97F4A7C15ull >> 20) & (s->cap - 1);
    for (;;)
    {
        if (s->stamp[p] != s->epoch)
        {
            s->stamp[p] = s->epoch;
            s->key[p] = key;
            *fresh = 1;
            return p;
        }
        if (s->key[p] == key)
        {
            *fresh = 0;
            return p;
        }
        p = (p + 1) & (s->cap - 1);
    }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct
{
    uint64_t bt[BO_THREADS]; /* persistent bucket counts of the 16 worker tables (bpe.c:615, hash_table.c:310-338) */
    cht_t tt[BO_THREADS];
    cht_t merged;
    hset_t set_a, set_b;
    /* scratch for the closed-form resolver */
    uint64_t *ent_key;
    uint32_t *ent_hsh;
    uint32_t *ent_rank;
    uint8_t *ent_first;
    size_t ent_cap;
} emu_t;

static void emu_init(emu_t *e)
{
    memset(e, 0, sizeof *e);
    for (unsigned t = 0; t < BO_THREADS; t++)
        e->bt[t] = BO_THREAD_BUCKETS;
}

static void emu_release(emu_t *e)
{
    for (unsigned t = 0; t < BO_THREADS; t++)
        cht_release(&e->tt[t]);
    cht_release(&e->merged);
    free(e->set_a.key);
    free(e->set_a.stamp);
    free(e->set_a.aux);
    free(e->set_b.key);
    free(e->set_b.stamp);
    free(e->set_b.aux);
    free(e->ent_key);
    free(e->ent_hsh);
    free(e->ent_rank);
    free(e->ent_first);
}

/* which positions worker t counts: bpe.c:449-477 (static slices) / canonical schedule for
 * bpe.c:479-520 (worker 0 takes every chunk).  Pair positions are i with i+1 < n (bpe.c:462). */
static void slice_of(size_t n, unsigned t, size_t *start, size_t *end)
{
    if (n < BO_STATIC_LIMIT)
    {
        size_t per = n / BO_THREADS;
        size_t s = (size_t)t * per;
        size_t len = (t == BO_THREADS - 1) ? per + n % BO_THREADS : per;
        size_t e = s + len;
        if (e > n - 1)
            e = n - 1; /* i + 1 >= text_size -> break */
        if (s > e)
            s = e;
        *start = s;
        *end = e;
    }
    else
    {
        *start = 0;
        *end = (t == 0) ? n - 1 : 0;
    }
}

/* FAITHFUL selection for one iteration.  found=0 when the merged table is empty (bpe.c:730). */
static int faithful_select(emu_t *e, const uint32_t *text, size_t n, bo_pair_t *best, uint32_t *best_freq,
                           uint64_t *distinct, int *found)
{
    for (unsigned t = 0; t < BO_THREADS; t++)
    {
        if (cht_set_buckets(&e->tt[t], e->bt[t])) /* hash_table_clear keeps the grown bucket count */
            return -1;
        size_t s, en;
        slice_of(n, t, &s, &en);
        for (size_t i = s; i < en; i++)
            if (cht_add(&e->tt[t], pack_key(text[i], text[i + 1]), bo_murmur3_pair(text[i], text[i + 1]), 1))
                return -1;
        e->bt[t] = e->tt[t].nb;
    }
    if (cht_set_buckets(&e->merged, BO_MERGED_BUCKETS)) /* bpe.c:684 fresh table every iteration */
        return -1;
    for (unsigned t = 0; t < BO_THREADS; t++) /* hash_table.c:146-188 */
    {
        cht_t *tt = &e->tt[t];
        for (uint64_t bk = 0; bk < tt->nb; bk++)
            for (int32_t c = tt->head[bk]; c >= 0; c = tt->next[c])
                if (cht_add(&e->merged, tt->key[c], tt->hsh[c], tt->val[c]))
                    return -1;
    }
    *found = 0;
    *best_freq = 0;
    cht_t *m = &e->merged;
    for (uint64_t bk = 0; bk < m->nb; bk++) /* bpe.c:705-728 + dyn_arr.c:163-174 */
        for (int32_t c = m->head[bk]; c >= 0; c = m->next[c])
        {
            uint32_t f = (uint32_t)m->val[c]; /* bpe.c:717 truncation */
            if (!*found || *best_freq < f)    /* strict <: first maximum wins (bpe.c:9) */
            {
                *found = 1;
                *best_freq = f;
                best->a = (uint32_t)(m->key[c] & 0xFFFFFFFFu);
                best->b = (uint32_t)(m->key[c] >> 32);
            }
        }
    *distinct = m->nn;
    return 0;
}

/* resizes a table that starts the iteration with `b0` buckets goes through, given that it sees
 * `d` distinct keys and whether its very last insert call created the d-th key */
static uint64_t grown_buckets(uint64_t b0, uint64_t d, int last_call_is_new)
{
    uint64_t b = b0;
    for (;;)
    {
        uint64_t th = resize_threshold(b);
        if (d > th || (d == th && !last_call_is_new))
            b *= 2;
        else
            return b;
    }
}

/* number of doublings (starting at b0, ending at bfinal) that happen after the key of 1-based
 * first-sight rank r went in: a doubling for bucket count B fires on the first insert call after
 * thr(B) keys are present */
static unsigned doublings_after(uint64_t b0, uint64_t bfinal, uint64_t r)
{
    unsigned c = 0;
    for (uint64_t b = b0; b < bfinal; b *= 2)
        if (resize_threshold(b) >= r)
            c++;
    return c;
}

/* keeps bt[] up to date on iterations that do not run faithful_select */
static int census(emu_t *e, const uint32_t *text, size_t n, const pmap_t *pm, int *ran)
{
    *ran = 0;
    if (n < 2)
        return 0;
    if (n >= BO_STATIC_LIMIT)
    {
        /* worker 0 sees every pair: D keys; its last call creates a key iff the last pair occurs once */
        int64_t s = pmap_find(pm, pack_key(text[n - 2], text[n - 1]), bo_murmur3_pair(text[n - 2], text[n - 1]));
        int last_new = (s >= 0 && pm->cnt[s] == 1);
        e->bt[0] = grown_buckets(e->bt[0], pm->distinct, last_new);
        return 0;
    }
    int need = 0;
    for (unsigned t = 0; t < BO_THREADS; t++)
    {
        size_t s, en;
        slice_of(n, t, &s, &en);
        uint64_t most = en - s;
        if (most > pm->distinct)
            most = pm->distinct;
        if (most >= resize_threshold(e->bt[t]))
            need = 1;
    }
    if (!need)
        return 0;
    *ran = 1;
    if (hset_reserve(&e->set_a, n / BO_THREADS + BO_THREADS + 2))
        return -1;
    for (unsigned t = 0; t < BO_THREADS; t++)
    {
        size_t s, en;
        slice_of(n, t, &s, &en);
        if (en == s)
            continue;
        hset_next_epoch(&e->set_a);
        uint64_t d = 0;
        int fresh = 0;
        for (size_t i = s; i < en; i++)
        {
            hset_put(&e->set_a, pack_key(text[i], text[i + 1]), bo_murmur3_pair(text[i], text[i + 1]), &fresh);
            d += (uint64_t)fresh;
        }
        e->bt[t] = grown_buckets(e->bt[t], d, fresh);
    }
    return 0;
}

/* position of an entry inside one chain, as a sortable number: keys that went in while an even
 * number of doublings was still to come sit in front, youngest first; the others behind, oldest
 * first (each doubling reverses every chain, hash_table.c:208-223) */
static inline uint64_t chain_slot(unsigned doublings_to_come, uint64_t r)
{
    return (doublings_to_come & 1u) ? ((1ull << 40) | r) : ((1ull << 40) - 1 - r);
}

static int ent_reserve(emu_t *e, size_t n)
{
    if (n <= e->ent_cap)
        return 0;
    free(e->ent_key);
    free(e->ent_hsh);
    free(e->ent_rank);
    free(e->ent_first);
    e->ent_key = (uint64_t *)malloc(n * sizeof(uint64_t));
    e->ent_hsh = (uint32_t *)malloc(n * sizeof(uint32_t));
    e->ent_rank = (uint32_t *)malloc(n * sizeof(uint32_t));
    e->ent_first = (uint8_t *)malloc(n);
    e->ent_cap = n;
    return (e->ent_key && e->ent_hsh && e->ent_rank && e->ent_first) ? 0 : -1;
}

/* Closed-form selection (what the CUDA resolver computes).  Needs the exact counts in pm. */
static int closedform_select(emu_t *e, const uint32_t *text, size_t n, const pmap_t *pm, bo_pair_t *best,
                             uint32_t *best_freq)
{
    const size_t npairs = n - 1;
    if (ent_reserve(e, npairs) || hset_reserve(&e->set_a, npairs) || hset_reserve(&e->set_b, npairs))
        return -1;
    /* pass 1: per worker, distinct keys in first-sight order with their 1-based rank; whether the
     * key is new to the whole merge sequence (not seen by a lower worker) */
    size_t ent_begin[BO_THREADS + 1];
    uint64_t b0[BO_THREADS], b1[BO_THREADS], first_seen_cnt[BO_THREADS];
    size_t n_ent = 0;
    hset_next_epoch(&e->set_b); /* global: key -> seen by some lower (or this) worker */
    int last_nonempty = -1;
    for (unsigned t = 0; t < BO_THREADS; t++)
    {
        size_t s, en;
        slice_of(n, t, &s, &en);
        ent_begin[t] = n_ent;
        first_seen_cnt[t] = 0;
        b0[t] = e->bt[t];
        hset_next_epoch(&e->set_a);
        uint32_t r = 0;
        int fresh = 0;
        for (size_t i = s; i < en; i++)
        {
            uint64_t key = pack_key(text[i], text[i + 1]);
            uint32_t h = bo_murmur3_pair(text[i], text[i + 1]);
            hset_put(&e->set_a, key, h, &fresh);
            if (fresh)
            {
                int gfresh;
                hset_put(&e->set_b, key, h, &gfresh);
                e->ent_key[n_ent] = key;
                e->ent_hsh[n_ent] = h;
                e->ent_rank[n_ent] = ++r;
                e->ent_first[n_ent] = (uint8_t)gfresh;
                first_seen_cnt[t] += (uint64_t)gfresh;
                n_ent++;
            }
        }
        b1[t] = grown_buckets(b0[t], r, fresh);
        e->bt[t] = b1[t];
        if (r)
            last_nonempty = (int)t;
    }
    ent_begin[BO_THREADS] = n_ent;
    const uint64_t D = pm->distinct;

    /* is the very last insert call of the merge a key creation?  It is the last entry, in table
     * order, of the last non-empty worker table. */
    int last_call_new = 0;
    if (last_nonempty >= 0)
    {
        unsigned t = (unsigned)last_nonempty;
        uint64_t best_ord_b = 0, best_ord_c = 0;
        int have = 0;
        for (size_t k = ent_begin[t]; k < ent_begin[t + 1]; k++)
        {
            uint64_t ob = (uint64_t)e->ent_hsh[k] % b1[t];
            uint64_t oc = chain_slot(doublings_after(b0[t], b1[t], e->ent_rank[k]), e->ent_rank[k]);
            if (!have || ob > best_ord_b || (ob == best_ord_b && oc > best_ord_c))
            {
                have = 1;
                best_ord_b = ob;
                best_ord_c = oc;
                last_call_new = e->ent_first[k];
            }
        }
    }
    const uint64_t bm = grown_buckets(BO_MERGED_BUCKETS, D, last_call_new);

    /* max frequency and the winning bucket under bm */
    uint32_t fmax = 0;
    uint64_t wbucket = 0;
    for (uint64_t s = 0; s < pm->cap; s++)
        if (pm->key[s] != PM_EMPTY && pm->cnt[s])
        {
            uint64_t bk = (uint64_t)pm->hsh[s] % bm;
            if (pm->cnt[s] > fmax || (pm->cnt[s] == fmax && bk < wbucket))
            {
                fmax = pm->cnt[s];
                wbucket = bk;
            }
        }
    *best_freq = fmax;
    if (!fmax)
        return 0;

    /* rank every candidate (max frequency, winning bucket) in the merge sequence */
    uint64_t prefix_first[BO_THREADS + 1];
    prefix_first[0] = 0;
    for (unsigned t = 0; t < BO_THREADS; t++)
        prefix_first[t + 1] = prefix_first[t] + first_seen_cnt[t];
    int have_best = 0;
    uint64_t best_slot = 0;
    for (unsigned t = 0; t < BO_THREADS; t++)
        for (size_t k = ent_begin[t]; k < ent_begin[t + 1]; k++)
        {
            if (!e->ent_first[k] || (uint64_t)e->ent_hsh[k] % bm != wbucket)
                continue;
            int64_t ps = pmap_find(pm, e->ent_key[k], e->ent_hsh[k]);
            if (ps < 0 || pm->cnt[ps] != fmax)
                continue;
            /* entries of worker t, new to the sequence, that precede this one in table-t order */
            uint64_t ob = (uint64_t)e->ent_hsh[k] % b1[t];
            uint64_t oc = chain_slot(doublings_after(b0[t], b1[t], e->ent_rank[k]), e->ent_rank[k]);
            uint64_t before = 0;
            for (size_t j = ent_begin[t]; j < ent_begin[t + 1]; j++)
            {
                if (!e->ent_first[j] || j == k)
                    continue;
                uint64_t jb = (uint64_t)e->ent_hsh[j] % b1[t];
                if (jb < ob)
                    before++;
                else if (jb == ob)
                {
                    uint64_t jc = chain_slot(doublings_after(b0[t], b1[t], e->ent_rank[j]), e->ent_rank[j]);
                    if (jc < oc)
                        before++;
                }
            }
            uint64_t R = prefix_first[t] + before + 1;
            uint64_t slot = chain_slot(doublings_after(BO_MERGED_BUCKETS, bm, R), R);
            if (!have_best || slot < best_slot)
            {
                have_best = 1;
                best_slot = slot;
                best->a = (uint32_t)(e->ent_key[k] & 0xFFFFFFFFu);
                best->b = (uint32_t)(e->ent_key[k] >> 32);
            }
        }
    return have_best ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------ */
size_t bo_rewrite(const uint32_t *in, size_t n, uint32_t a, uint32_t b, uint32_t z, uint32_t *out)
{
    size_t m = 0;
    for (size_t i = 0; i < n; i++) /* bpe.c:760-772 */
    {
        if (i + 1 < n && in[i] == a && in[i + 1] == b)
        {
            out[m++] = z;
            i++;
        }
        else
            out[m++] = in[i];
    }
    return m;
}

/* Per-replacement bookkeeping shared by the incremental path and the sharded/tiled emulation.
 * w points at the match start inside a window where w[-2..3] are the OLD neighbours (BO_SENT
 * where the stream ends).  Every pair INSTANCE that disappears or appears is charged to exactly
 * one replacement: a replacement always owns its left neighbour pair, and owns
 * its right neighbour pair unless another replacement starts right behind it.  Instances of
 * (a,b) itself are not listed: that count is set to zero after the merge (A.5.1). */
static inline void match_deltas(const uint32_t *w, uint32_t a, uint32_t b, uint32_t z, int32_t *delta)
{
    const uint32_t x = w[-1], y = w[2];
    const int same = (a == b);
    const int prev_match = same ? (x == a) : (w[-2] == a && x == b);
    const int next_match = (y == a && w[3] == b);
    if (x != BO_SENT)
    {
        if (!(same && x == a))
            delta[(size_t)x * 4 + 0]++; /* (x,a) gone */
        delta[(size_t)(prev_match ? z : x) * 4 + 2]++; /* (x',z) new */
    }
    if (y != BO_SENT && !next_match)
    {
        if (!(same && y == a))
            delta[(size_t)y * 4 + 1]++; /* (b,y) gone */
        delta[(size_t)y * 4 + 3]++;     /* (z,y) new */
    }
}

static int apply_deltas(pmap_t *pm, uint32_t a, uint32_t b, uint32_t z, int32_t *delta)
{
    for (uint32_t t = 0; t <= z; t++)
    {
        int32_t *d = delta + (size_t)t * 4;
        if (d[0] && pmap_add(pm, t, a, -(int64_t)d[0]))
            return -1;
        if (d[1] && pmap_add(pm, b, t, -(int64_t)d[1]))
            return -1;
        if (d[2] && pmap_add(pm, t, z, d[2]))
            return -1;
        if (d[3] && pmap_add(pm, z, t, d[3]))
            return -1;
        d[0] = d[1] = d[2] = d[3] = 0;
    }
    int64_t s = pmap_find(pm, pack_key(a, b), bo_murmur3_pair(a, b));
    if (s >= 0 && pm->cnt[s])
    {
        pm->cnt[s] = 0;
        pm->distinct--;
    }
    return 0;
}

/* rewrite + delta emission over a stream that carries 4 sentinel slots on either side */
static size_t rewrite_with_deltas(const uint32_t *in, size_t n, uint32_t a, uint32_t b, uint32_t z, uint32_t *out,
                                  int32_t *delta)
{
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
    {
        if (i + 1 < n && in[i] == a && in[i + 1] == b)
        {
            match_deltas(in + i, a, b, z, delta);
            out[m++] = z;
            i++;
        }
        else
            out[m++] = in[i];
    }
    return m;
}

/* ---- the same two loops split over helper threads (bo_set_workers) ---------------------- */
/* first position a worker whose share starts at `bound` has to look at: `bound` itself, or the
 * one behind it when `bound` is the second token of a replacement that starts at bound - 1
 * (a != b: occurrences cannot overlap, so every occurrence is a replacement; a == b: replacements
 * pair up from the start of the run, bpe.c:760-772 scanning left to right) */
static size_t share_start(const uint32_t *in, size_t n, size_t bound, uint32_t a, uint32_t b)
{
    if (bound == 0 || bound >= n)
        return bound > n ? n : bound;
    if (a != b)
        return (in[bound - 1] == a && in[bound] == b) ? bound + 1 : bound;
    if (in[bound] != a)
        return bound;
    size_t r = bound;
    while (r > 0 && in[r - 1] == a)
        r--;
    return ((bound - r) & 1) ? bound + 1 : bound;
}

typedef struct
{
    const uint32_t *in;
    size_t n;
    uint32_t a, b, z;
    uint32_t *out;
    int32_t *delta0, *delta_rest; /* worker 0 / workers 1.. (delta_len each) */
    size_t delta_len;
    size_t start[BO_MAX_WORKERS + 1], cnt[BO_MAX_WORKERS], off[BO_MAX_WORKERS];
    int write;
} prw_t;

static void prw_worker(void *ctx, int w, int nw)
{
    prw_t *c = (prw_t *)ctx;
    (void)nw;
    const uint32_t *in = c->in;
    const size_t n = c->n, end = c->start[w + 1];
    const uint32_t a = c->a, b = c->b, z = c->z;
    if (!c->write)
    {
        size_t m = 0;
        for (size_t i = c->start[w]; i < end; i++, m++)
            if (i + 1 < n && in[i] == a && in[i + 1] == b)
                i++;
        c->cnt[w] = m;
        return;
    }
    uint32_t *out = c->out + c->off[w];
    int32_t *delta = w ? c->delta_rest + (size_t)(w - 1) * c->delta_len : c->delta0;
    size_t m = 0;
    for (size_t i = c->start[w]; i < end; i++)
    {
        if (i + 1 < n && in[i] == a && in[i + 1] == b)
        {
            match_deltas(in + i, a, b, z, delta);
            out[m++] = z;
            i++;
        }
        else
            out[m++] = in[i];
    }
}

static size_t rewrite_with_deltas_par(const uint32_t *in, size_t n, uint32_t a, uint32_t b, uint32_t z, uint32_t *out,
                                      int32_t *delta0, int32_t *delta_rest, size_t delta_len, int nw)
{
    prw_t c;
    memset(&c, 0, sizeof c);
    c.in = in;
    c.n = n;
    c.a = a;
    c.b = b;
    c.z = z;
    c.out = out;
    c.delta0 = delta0;
    c.delta_rest = delta_rest;
    c.delta_len = delta_len;
    for (int w = 0; w < nw; w++)
        c.start[w] = share_start(in, n, (size_t)((unsigned __int128)n * (unsigned)w / (unsigned)nw), a, b);
    c.start[nw] = n;
    c.write = 0;
    par_run(prw_worker, &c, nw);
    size_t total = 0;
    for (int w = 0; w < nw; w++)
    {
        c.off[w] = total;
        total += c.cnt[w];
    }
    c.write = 1;
    par_run(prw_worker, &c, nw);
    const size_t used = ((size_t)z + 1) * 4;
    for (int w = 1; w < nw; w++)
    {
        int32_t *d = delta_rest + (size_t)(w - 1) * delta_len;
        for (size_t k = 0; k < used; k++)
            if (d[k])
            {
                delta0[k] += d[k];
                d[k] = 0;
            }
    }
    return total;
}

typedef struct
{
    const pmap_t *pm;
    uint64_t bm;
    uint32_t freq[BO_MAX_WORKERS];
    uint64_t bucket[BO_MAX_WORKERS], mult[BO_MAX_WORKERS], key[BO_MAX_WORKERS];
} pam_t;

static void pam_worker(void *ctx, int w, int nw)
{
    pam_t *c = (pam_t *)ctx;
    const pmap_t *pm = c->pm;
    const uint64_t lo = (uint64_t)((unsigned __int128)pm->cap * (unsigned)w / (unsigned)nw);
    const uint64_t hi = (uint64_t)((unsigned __int128)pm->cap * (unsigned)(w + 1) / (unsigned)nw);
    uint32_t bf = 0;
    uint64_t bb = 0, mult = 0, key = 0;
    for (uint64_t s = lo; s < hi; s++)
        if (pm->key[s] != PM_EMPTY && pm->cnt[s])
        {
            uint64_t bk = (uint64_t)pm->hsh[s] % c->bm;
            if (pm->cnt[s] > bf || (pm->cnt[s] == bf && bk < bb))
            {
                bf = pm->cnt[s];
                bb = bk;
                key = pm->key[s];
                mult = 1;
            }
            else if (pm->cnt[s] == bf && bk == bb)
                mult++;
        }
    c->freq[w] = bf;
    c->bucket[w] = bb;
    c->mult[w] = mult;
    c->key[w] = key;
}

/* ------------------------------------------------------------------------------------------ */
int bo_train(const uint8_t *bytes, size_t n_in, uint64_t max_merges, int mode, bo_pair_t **merges_out,
             size_t *n_merges_out, uint32_t **tokens_out, size_t *n_tokens_out, bo_stats_t *stats)
{
    if (!bytes || !merges_out || !n_merges_out || !tokens_out || !n_tokens_out)
        return BO_ERR_ARG;
    size_t n = 0;
    while (n < n_in && bytes[n]) /* strlen: bpe.c:555 */
        n++;
    if (n < 2)
        return BO_ERR_SHORT;

    const size_t pad = 4;
    uint32_t *buf0 = (uint32_t *)malloc((n + 2 * pad) * sizeof(uint32_t));
    uint32_t *buf1 = (uint32_t *)malloc((n + 2 * pad) * sizeof(uint32_t));
    size_t mcap = 1024, nm = 0;
    bo_pair_t *merges = (bo_pair_t *)malloc(mcap * sizeof(bo_pair_t));
    int32_t *delta = NULL, *pdelta = NULL;
    size_t delta_cap = 0, pdelta_len = 0;
    emu_t emu;
    pmap_t pm;
    memset(&pm, 0, sizeof pm);
    emu_init(&emu);
    bo_stats_t st;
    memset(&st, 0, sizeof st);
    int rc = BO_ERR_NOMEM;
    if (!buf0 || !buf1 || !merges)
        goto done;
    for (size_t i = 0; i < pad; i++)
        buf0[i] = buf1[i] = BO_SENT;
    uint32_t *text = buf0 + pad, *temp = buf1 + pad;
    for (size_t i = 0; i < n; i++)
        text[i] = (uint32_t)bytes[i]; /* bpe.c:582 unsigned widening */

    const int fast = (mode != BO_MODE_FAITHFUL);
    if (fast)
    {
        if (pmap_init(&pm, 1u << 16))
            goto done;
        for (size_t i = 0; i + 1 < n; i++)
            if (pmap_add(&pm, text[i], text[i + 1], 1))
                goto done;
    }

    uint32_t next_symbol = 256; /* bpe.c:588 */
    for (;;)
    {
        bo_pair_t best = {0, 0};
        uint32_t best_freq = 0;
        uint64_t D = 0;
        int found = 0;
        if (!fast)
        {
            if (n < 2)
                break; /* no pairs: the merged table is empty (bpe.c:730) */
            if (faithful_select(&emu, text, n, &best, &best_freq, &D, &found))
                goto done;
            st.faithful_iters++;
            st.final_distinct = D;
            if (!found)
                break;
        }
        else
        {
            D = pm.distinct;
            st.final_distinct = D;
            if (!D)
                break;
            const uint64_t bm = bo_merged_buckets(D);
            uint64_t best_bucket = 0, mult = 0;
            const int nw = (g_workers > 1 && n >= g_workers_min_n && n >= 8u * (size_t)g_workers) ? g_workers : 1;
            int from_list = 0;
            if (pm.cand_T)
            {
                for (uint64_t k = 0; k < pm.cand_n; k++)
                {
                    const uint64_t s = pm.cand[k];
                    if (!pm.cnt[s])
                        continue;
                    uint64_t bk = (uint64_t)pm.hsh[s] % bm;
                    if (pm.cnt[s] > best_freq || (pm.cnt[s] == best_freq && bk < best_bucket))
                    {
                        best_freq = pm.cnt[s];
                        best_bucket = bk;
                        best.a = (uint32_t)(pm.key[s] & 0xFFFFFFFFu);
                        best.b = (uint32_t)(pm.key[s] >> 32);
                        mult = 1;
                    }
                    else if (pm.cnt[s] == best_freq && bk == best_bucket)
                        mult++;
                }
                if (best_freq >= pm.cand_T)
                    from_list = 1; /* every maximal entry is in the list */
                else
                {
                    pm.cand_T = 0; /* the maximum may sit below the threshold: look at the whole map */
                    best_freq = 0;
                    best_bucket = mult = 0;
                }
            }
            if (from_list)
                ;
            else if (nw > 1)
            {
                pam_t pa;
                pa.pm = &pm;
                pa.bm = bm;
                par_run(pam_worker, &pa, nw);
                for (int w = 0; w < nw; w++) /* shares in slot order: the first maximum stays the first */
                {
                    if (!pa.freq[w])
                        continue;
                    if (pa.freq[w] > best_freq || (pa.freq[w] == best_freq && pa.bucket[w] < best_bucket))
                    {
                        best_freq = pa.freq[w];
                        best_bucket = pa.bucket[w];
                        best.a = (uint32_t)(pa.key[w] & 0xFFFFFFFFu);
                        best.b = (uint32_t)(pa.key[w] >> 32);
                        mult = pa.mult[w];
                    }
                    else if (pa.freq[w] == best_freq && pa.bucket[w] == best_bucket)
                        mult += pa.mult[w];
                }
            }
            else
            for (uint64_t s = 0; s < pm.cap; s++)
                if (pm.key[s] != PM_EMPTY && pm.cnt[s])
                {
                    uint64_t bk = (uint64_t)pm.hsh[s] % bm;
                    if (pm.cnt[s] > best_freq || (pm.cnt[s] == best_freq && bk < best_bucket))
                    {
                        best_freq = pm.cnt[s];
                        best_bucket = bk;
                        best.a = (uint32_t)(pm.key[s] & 0xFFFFFFFFu);
                        best.b = (uint32_t)(pm.key[s] >> 32);
                        mult = 1;
                    }
                    else if (pm.cnt[s] == best_freq && bk == best_bucket)
                        mult++;
                }
            if (!from_list && g_cand_floor && best_freq >= g_cand_floor)
                if (cand_build(&pm, best_freq / 2 ? best_freq / 2 : 1))
                    goto done;
            int on_threshold = 0;
            for (uint64_t b = BO_MERGED_BUCKETS; resize_threshold(b) <= D; b *= 2)
                if (resize_threshold(b) == D)
                    on_threshold = 1;
            if (best_freq > 1 && (mult > 1 || on_threshold))
            {
                if (mult > 1)
                    st.same_bucket_ties++;
                if (on_threshold)
                    st.threshold_edges++;
                if (mode == BO_MODE_FAST)
                {
                    uint64_t d2 = 0;
                    uint32_t f2 = 0;
                    if (faithful_select(&emu, text, n, &best, &f2, &d2, &found))
                        goto done;
                    st.faithful_iters++;
                    if (!found || f2 != best_freq || d2 != D)
                    {
                        rc = -100; /* incremental counts diverged from a recount: oracle bug */
                        goto done;
                    }
                }
                else
                {
                    uint32_t f2 = 0;
                    if (closedform_select(&emu, text, n, &pm, &best, &f2) || f2 != best_freq)
                    {
                        rc = -101;
                        goto done;
                    }
                }
            }
            else
            {
                int ran = 0;
                if (census(&emu, text, n, &pm, &ran))
                    goto done;
                st.census_iters += (uint64_t)ran;
            }
        }
        if (best_freq <= 1) /* bpe.c:745 */
            break;
        if (max_merges && nm >= max_merges)
            break;
        if (nm == mcap)
        {
            mcap *= 2;
            bo_pair_t *m2 = (bo_pair_t *)realloc(merges, mcap * sizeof(bo_pair_t));
            if (!m2)
                goto done;
            merges = m2;
        }
        merges[nm++] = best; /* bpe.c:752-758 */

        size_t new_n;
        if (fast)
        {
            size_t need = ((size_t)next_symbol + 1) * 4;
            if (need > delta_cap)
            {
                size_t nc = delta_cap ? delta_cap : 4096;
                while (nc < need)
                    nc *= 2;
                int32_t *d2 = (int32_t *)realloc(delta, nc * sizeof(int32_t));
                if (!d2)
                    goto done;
                memset(d2 + delta_cap, 0, (nc - delta_cap) * sizeof(int32_t));
                delta = d2;
                delta_cap = nc;
            }
            for (size_t i = 0; i < pad; i++)
                text[n + i] = BO_SENT;
            const int nw = (g_workers > 1 && n >= g_workers_min_n && n >= 8u * (size_t)g_workers) ? g_workers : 1;
            if (nw > 1)
            {
                if (pdelta_len != delta_cap) /* the helpers' vectors: all zero between merges */
                {
                    free(pdelta);
                    pdelta = (int32_t *)calloc((size_t)(BO_MAX_WORKERS - 1) * delta_cap, sizeof(int32_t));
                    if (!pdelta)
                        goto done;
                    pdelta_len = delta_cap;
                }
                new_n = rewrite_with_deltas_par(text, n, best.a, best.b, next_symbol, temp, delta, pdelta, pdelta_len, nw);
            }
            else
                new_n = rewrite_with_deltas(text, n, best.a, best.b, next_symbol, temp, delta);
            if (apply_deltas(&pm, best.a, best.b, next_symbol, delta))
                goto done;
        }
        else
            new_n = bo_rewrite(text, n, best.a, best.b, next_symbol, temp);
        uint32_t *sw = text; /* bpe.c:774-779 */
        text = temp;
        temp = sw;
        n = new_n;
        next_symbol++;
    }

    st.n_merges = nm;
    st.n_tokens = n;
    for (unsigned t = 0; t < BO_THREADS; t++)
        st.thread_buckets[t] = emu.bt[t];
    {
        uint32_t *res = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        if (!res)
            goto done;
        memcpy(res, text, n * sizeof(uint32_t));
        *tokens_out = res;
        *n_tokens_out = n;
        *merges_out = merges;
        *n_merges_out = nm;
        merges = NULL;
        if (stats)
            *stats = st;
        rc = BO_OK;
    }
done:
    free(buf0);
    free(buf1);
    free(merges);
    free(delta);
    free(pdelta);
    emu_release(&emu);
    if (pm.key)
        pmap_release(&pm);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
int bo_encode(const uint8_t *bytes, size_t n_in, const bo_pair_t *merges, size_t n_merges, uint32_t **tokens_out,
              size_t *n_tokens_out)
{
    if (!bytes || (!merges && n_merges) || !tokens_out || !n_tokens_out)
        return BO_ERR_ARG;
    size_t n = 0;
    while (n < n_in && bytes[n])
        n++;
    uint32_t *text = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *temp = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint64_t *occ = (uint64_t *)calloc(256 + n_merges + 1, sizeof(uint64_t)); /* occurrences per id */
    if (!text || !temp || !occ)
    {
        free(text);
        free(temp);
        free(occ);
        return BO_ERR_NOMEM;
    }
    for (size_t i = 0; i < n; i++)
    {
        text[i] = (uint32_t)bytes[i];
        occ[bytes[i]]++;
    }
    for (size_t r = 0; r < n_merges; r++)
    {
        const uint32_t a = merges[r].a, b = merges[r].b, z = (uint32_t)(256 + r);
        if (a >= z || b >= z || !occ[a] || !occ[b])
            continue; /* a pass would change nothing */
        size_t m = bo_rewrite(text, n, a, b, z, temp);
        size_t repl = n - m;
        occ[a] -= repl;
        occ[b] -= repl;
        occ[z] = repl;
        uint32_t *sw = text;
        text = temp;
        temp = sw;
        n = m;
    }
    free(temp);
    free(occ);
    *tokens_out = text;
    *n_tokens_out = n;
    return BO_OK;
}

int bo_decode(const uint32_t *tokens, size_t n_tokens, const bo_pair_t *merges, size_t n_merges, uint8_t **bytes_out,
              size_t *n_bytes_out)
{
    if ((!tokens && n_tokens) || (!merges && n_merges) || !bytes_out || !n_bytes_out)
        return BO_ERR_ARG;
    const size_t V = 256 + n_merges;
    uint64_t *len = (uint64_t *)malloc(V * sizeof(uint64_t));
    if (!len)
        return BO_ERR_NOMEM;
    for (size_t i = 0; i < 256; i++)
        len[i] = 1;
    for (size_t r = 0; r < n_merges; r++)
    {
        if (merges[r].a >= 256 + r || merges[r].b >= 256 + r)
        {
            free(len);
            return BO_ERR_ARG;
        }
        len[256 + r] = len[merges[r].a] + len[merges[r].b];
    }
    uint64_t total = 0;
    for (size_t i = 0; i < n_tokens; i++)
    {
        if (tokens[i] >= V)
        {
            free(len);
            return BO_ERR_ARG;
        }
        total += len[tokens[i]];
    }
    uint8_t *out = (uint8_t *)malloc(total ? total : 1);
    uint32_t *stack = (uint32_t *)malloc((n_merges + 2) * sizeof(uint32_t));
    if (!out || !stack)
    {
        free(len);
        free(out);
        free(stack);
        return BO_ERR_NOMEM;
    }
    size_t w = 0;
    for (size_t i = 0; i < n_tokens; i++)
    {
        size_t sp = 0;
        stack[sp++] = tokens[i];
        while (sp)
        {
            uint32_t t = stack[--sp];
            if (t < 256)
                out[w++] = (uint8_t)t;
            else
            {
                stack[sp++] = merges[t - 256].b;
                stack[sp++] = merges[t - 256].a;
            }
        }
    }
    free(len);
    free(stack);
    *bytes_out = out;
    *n_bytes_out = w;
    return BO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* The multi-GPU / tiled formulation of one merge, emulated on the CPU.
 *
 * Each shard publishes a pair-independent edge record (first three tokens, last two tokens,
 * length, length of its trailing run of equal tokens).  From all records a shard derives the two
 * tokens in front of it, the three tokens behind it, and — for a == b — the parity of the run of
 * `a` that ends right in front of it.  Inside a shard the stream is cut into tiles; a tile sees
 * two tokens of halo in front and three behind, plus that parity carried tile to tile. */
typedef struct
{
    uint64_t len;
    uint32_t first[3];
    uint32_t last[2]; /* last[1] is the final token, last[0] the one before it */
    uint64_t trail_run;
} edge_t;

static void edge_of(const uint32_t *s, size_t len, edge_t *e)
{
    e->len = len;
    for (int k = 0; k < 3; k++)
        e->first[k] = ((size_t)k < len) ? s[k] : BO_SENT;
    e->last[1] = len >= 1 ? s[len - 1] : BO_SENT;
    e->last[0] = len >= 2 ? s[len - 2] : BO_SENT;
    uint64_t r = 0;
    while (r < len && s[len - 1 - r] == s[len - 1])
        r++;
    e->trail_run = r;
}

size_t bo_rewrite_sharded(const uint32_t *in, size_t n, uint32_t a, uint32_t b, uint32_t z, int n_shards, size_t tile,
                          uint32_t *out, int32_t *delta)
{
    if (n_shards < 1)
        n_shards = 1;
    if (tile < 1)
        tile = 1;
    const int same = (a == b);
    edge_t *edges = (edge_t *)malloc((size_t)n_shards * sizeof(edge_t));
    size_t *sh_begin = (size_t *)malloc(((size_t)n_shards + 1) * sizeof(size_t));
    uint32_t *win = (uint32_t *)malloc((tile + 8) * sizeof(uint32_t));
    size_t m = 0;
    if (!edges || !sh_begin || !win)
        goto out;
    for (int s = 0; s <= n_shards; s++)
        sh_begin[s] = (size_t)((unsigned __int128)n * (unsigned)s / (unsigned)n_shards);
    for (int s = 0; s < n_shards; s++)
        edge_of(in + sh_begin[s], sh_begin[s + 1] - sh_begin[s], &edges[s]);

    for (int s = 0; s < n_shards; s++)
    {
        const uint32_t *sh = in + sh_begin[s];
        const size_t len = sh_begin[s + 1] - sh_begin[s];
        /* halo in front: walk back over the records until two tokens are collected */
        uint32_t before[2] = {BO_SENT, BO_SENT};
        int got = 0;
        for (int q = s - 1; q >= 0 && got < 2; q--)
        {
            if (edges[q].len >= 1 && got < 2)
                before[1 - got++] = edges[q].last[1];
            if (edges[q].len >= 2 && got < 2)
                before[1 - got++] = edges[q].last[0];
        }
        /* halo behind: walk forward until three tokens are collected */
        uint32_t after[3] = {BO_SENT, BO_SENT, BO_SENT};
        got = 0;
        for (int q = s + 1; q < n_shards && got < 3; q++)
            for (int k = 0; k < 3 && got < 3; k++)
                if ((uint64_t)k < edges[q].len)
                    after[got++] = edges[q].first[k];
        /* parity of the run of `a` that ends right in front of this shard */
        unsigned carry = 0;
        if (same)
            for (int q = s - 1; q >= 0; q--)
            {
                if (!edges[q].len)
                    continue;
                if (edges[q].last[1] != a)
                    break;
                carry ^= (unsigned)(edges[q].trail_run & 1u);
                if (edges[q].trail_run != edges[q].len)
                    break;
            }
        for (size_t t0 = 0; t0 < len; t0 += tile)
        {
            const size_t tl = (len - t0 < tile) ? len - t0 : tile;
            /* window: win[2+k] = sh[t0+k], halos from the neighbours or the shard edges */
            for (int k = -2; k < (int)tl + 3; k++)
            {
                long long g = (long long)t0 + k;
                uint32_t v;
                if (g < 0)
                    v = before[2 + g];
                else if ((size_t)g < len)
                    v = sh[g];
                else
                    v = (g - (long long)len < 3) ? after[g - (long long)len] : BO_SENT;
                win[2 + k] = v;
            }
            const uint32_t *w = win + 2;
            /* odd[k] = parity of the number of consecutive `a` right in front of position k */
            unsigned par = carry;
            unsigned prev_match = 0; /* does a replacement start at position k-1 ? */
            if (same)
                prev_match = (w[-1] == a && w[0] == a && carry == 1);
            else
                prev_match = (w[-1] == a && w[0] == b);
            for (size_t k = 0; k < tl; k++)
            {
                unsigned is_match;
                if (same)
                    is_match = (w[k] == a && w[k + 1] == a && par == 0);
                else
                    is_match = (w[k] == a && w[k + 1] == b);
                if (!prev_match)
                {
                    if (is_match)
                    {
                        out[m++] = z;
                        match_deltas(w + k, a, b, z, delta);
                    }
                    else
                        out[m++] = w[k];
                }
                /* a dropped position never starts a replacement: for a != b it holds b, for
                 * a == b the parity rule already says so */
                prev_match = prev_match ? 0 : is_match;
                par = (w[k] == a) ? (par ^ 1u) : 0u;
            }
            carry = par;
        }
    }
out:
    free(edges);
    free(sh_begin);
    free(win);
    return m;
}

void bo_free(void *p) { free(p); }
