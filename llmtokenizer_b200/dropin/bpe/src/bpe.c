/* Host side of the drop-in: the reference's bpe/src/bpe.c API on top of the CUDA engine.
 *
 * compress()  (reference bpe.c:541-811)  reads the file exactly like the reference (get_file, then
 *             strlen semantics: everything after the first NUL is ignored, bpe.c:555), hands the bytes
 *             to bpe_cuda_train() and wraps the results in the containers main.c expects: a
 *             dyn_arr_t of pair_t created with dyn_arr_create(512, 8) whose entries 0..255 are {i,0}
 *             (bpe.c:589-608) and 256+k is the k-th merge (bpe.c:752-758), plus a malloc'd uint32_t
 *             stream the caller free()s (main.c:22).
 * The cold functions (decode, pair-file I/O, printing) are plain host C with the reference's
 * observable behaviour; they are not on the accelerated path.
 * There is no CPU merge loop in this file: without a GPU compress() fails (returns NULL). */
#include "../inc/bpe.h"

#include "../../../../include/bpe_cuda.h"

/* ---- small helpers -------------------------------------------------------------------------- */
static size_t env_size(const char *name, size_t dflt)
{
    const char *e = getenv(name);
    if (!e || !*e)
        return dflt;
    return (size_t)strtoull(e, NULL, 10);
}

bool is_less(const void *a, const void *b) /* bpe.c:4-10: strict <, so the first maximum wins */
{
    return ((const pair_freq_t *)a)->freq < ((const pair_freq_t *)b)->freq;
}

/* whole file, NUL-terminated; *size_out (optional) = bytes read */
static char *read_whole_file(const char *path, size_t *size_out)
{
    FILE *f = fopen(path, "r");
    if (!f)
    {
        perror("fopen"); /* bpe.c:133-137 */
        return NULL;
    }
    if (fseek(f, 0, SEEK_END) != 0)
    {
        perror("fseek");
        fclose(f);
        return NULL;
    }
    const long sz = ftell(f);
    if (sz < 0)
    {
        perror("ftell");
        fclose(f);
        return NULL;
    }
    rewind(f);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (!buf)
    {
        perror("malloc");
        fclose(f);
        return NULL;
    }
    const size_t got = fread(buf, 1, (size_t)sz, f);
    if (got != (size_t)sz && ferror(f))
    {
        perror("fread");
        free(buf);
        fclose(f);
        return NULL;
    }
    buf[got] = '\0';
    fclose(f);
    if (size_out)
        *size_out = got;
    return buf;
}

char *get_file(const char *path) /* bpe.c:130-180 */
{
    if (!path)
        return NULL;
    return read_whole_file(path, NULL);
}

static dyn_arr_t *new_vocabulary(void)
{
    dyn_arr_t *arr = dyn_arr_create(512, sizeof(pair_t)); /* bpe.c:589 */
    if (!arr)
        return NULL;
    for (uint32_t i = 0; i < 256; i++) /* bpe.c:598-608 */
    {
        const pair_t p = {i, 0};
        if (!dyn_arr_set(arr, i, &p))
        {
            dyn_arr_free(arr);
            return NULL;
        }
    }
    return arr;
}

/* ---- the accelerated path -------------------------------------------------------------------- */
dyn_arr_t *compress_n(const char *path, uint32_t **encoding, size_t *len, size_t max_merges, int n_gpus)
{
    if (!path || !encoding || !len) /* bpe.c:548-549 */
        return NULL;
    {
        /* the reference reports an unreadable file through perror (bpe.c:133-137) */
        FILE *f = fopen(path, "r");
        if (!f)
        {
            perror("fopen");
            return NULL;
        }
        fclose(f);
    }
    /* File ingest belongs to the engine (bpe.c:130-180 get_file, :555 strlen cut, :580-584 widen): the file is read in
     * pinned 32 MB pieces, each copied to the GPU, widened and counted while the next one is being read. */
    bpe_pair_t *merges = NULL;
    uint32_t *tokens = NULL;
    size_t n_merges = 0, n_tokens = 0;
    const int rc = bpe_cuda_train_file(path, (uint64_t)max_merges, n_gpus < 1 ? 1 : n_gpus, &merges, &n_merges, &tokens, &n_tokens,
                                       NULL);
    if (rc == BPE_CUDA_ERR_SHORT)
    {
        printf("Error: File contains less than 2 characters\n"); /* bpe.c:558-563 (stdout) */
        return NULL;
    }
    if (rc != BPE_CUDA_OK)
    {
        fprintf(stderr, "compress: %s\n", bpe_cuda_last_error());
        *encoding = NULL; /* bpe.c:841-842 */
        *len = 0;
        return NULL;
    }
    dyn_arr_t *arr = new_vocabulary();
    for (size_t k = 0; arr && k < n_merges; k++)
    {
        const pair_t p = {merges[k].a, merges[k].b};
        if (!dyn_arr_set(arr, 256 + k, &p)) /* bpe.c:752-758 */
        {
            dyn_arr_free(arr);
            arr = NULL;
        }
    }
    bpe_cuda_free(merges);
    if (!arr)
    {
        bpe_cuda_free(tokens);
        *encoding = NULL;
        *len = 0;
        return NULL;
    }
    *encoding = tokens; /* malloc-owned: main.c:22 free()s it */
    *len = n_tokens;
    return arr;
}

dyn_arr_t *compress(const char *path, uint32_t **encoding, size_t *len) /* bpe.c:541 */
{
    return compress_n(path, encoding, len, env_size("BPE_MAX_MERGES", 0), (int)env_size("BPE_GPUS", 1));
}

uint32_t *bpe_encode_file(const char *path, dyn_arr_t *pair_arr, size_t *len, int n_gpus)
{
    if (!path || !pair_arr || !len || pair_arr->last_index < 255)
        return NULL;
    const size_t n_merges = pair_arr->last_index - 255;
    bpe_pair_t *merges = (bpe_pair_t *)malloc((n_merges ? n_merges : 1) * sizeof *merges);
    uint32_t *tokens = NULL;
    size_t n_tokens = 0;
    int rc = merges ? BPE_CUDA_OK : BPE_CUDA_ERR_NOMEM;
    for (size_t k = 0; rc == BPE_CUDA_OK && k < n_merges; k++)
    {
        pair_t p;
        if (!dyn_arr_get(pair_arr, 256 + k, &p))
            rc = BPE_CUDA_ERR_ARG;
        merges[k].a = p.a;
        merges[k].b = p.b;
    }
    if (rc == BPE_CUDA_OK)
        rc = bpe_cuda_encode_file(path, merges, n_merges, n_gpus < 1 ? 1 : n_gpus, &tokens, &n_tokens, NULL);
    free(merges);
    if (rc != BPE_CUDA_OK)
    {
        fprintf(stderr, "bpe_encode_file: %s\n", bpe_cuda_last_error());
        return NULL;
    }
    *len = n_tokens;
    return tokens;
}

/* ---- output ---------------------------------------------------------------------------------- */
/* bpe.c:182-196 prints one printf per token; a GB-scale encoding makes that the slowest part of main().  Same
 * bytes on stdout, formatted into a 1 MB buffer and written in blocks. */
void print_text(const uint32_t *text, int length) /* bpe.c:182-196 */
{
    enum
    {
        CAP = 1 << 20
    };
    char *buf = (char *)malloc(CAP + 16);
    if (!buf)
    {
        for (int i = 0; i < length; i++)
        {
            if (text[i] < 32 || text[i] > 126)
                printf("[%u]", text[i]);
            else
                printf("%c", (char)text[i]);
        }
        printf("\n");
        return;
    }
    size_t w = 0;
    for (int i = 0; i < length; i++)
    {
        const uint32_t t = text[i];
        if (t >= 32 && t <= 126)
            buf[w++] = (char)t;
        else
        {
            char d[10];
            int nd = 0;
            uint32_t v = t;
            do
            {
                d[nd++] = (char)('0' + v % 10);
                v /= 10;
            } while (v);
            buf[w++] = '[';
            while (nd)
                buf[w++] = d[--nd];
            buf[w++] = ']';
        }
        if (w >= CAP)
        {
            fwrite(buf, 1, w, stdout);
            w = 0;
        }
    }
    buf[w++] = '\n';
    fwrite(buf, 1, w, stdout);
    free(buf);
}

/* ---- decode (bpe.c:12-128, 341-394): ids -> bytes through the vocabulary --------------------- */
/* byte length of every id's expansion (0 for ids that are not defined) */
static size_t *expansion_lengths(dyn_arr_t *pair_arr)
{
    const size_t nv = pair_arr->last_index + 1;
    size_t *L = (size_t *)calloc(nv, sizeof *L);
    if (!L)
        return NULL;
    for (size_t i = 0; i < nv; i++)
    {
        pair_t p;
        if (!dyn_arr_get(pair_arr, i, &p))
            continue;
        if (i < 256)
            L[i] = 1;
        else if (p.a < i && p.b < i) /* ids only refer to earlier ids */
            L[i] = L[p.a] + L[p.b];
    }
    return L;
}

/* Write the expansion of id to out[0 .. cap) (iteratively: an explicit, growing stack instead of the reference's
 * recursion).  The vocabulary is not trusted: an id may only refer to EARLIER ids (a pairs file can say anything), and
 * nothing is written beyond the length expansion_lengths() computed.  Returns the end of the output, NULL on a bad
 * vocabulary or when memory runs out. */
static char *expand_into(uint32_t id, dyn_arr_t *pair_arr, char *out, size_t cap)
{
    size_t scap = 64, sp = 0;
    uint32_t *stack = (uint32_t *)malloc(scap * sizeof *stack);
    char *const end = out + cap;
    if (!stack)
        return NULL;
    stack[sp++] = id;
    while (sp)
    {
        const uint32_t t = stack[--sp];
        if (t < 256)
        {
            if (out == end)
                goto bad;
            *out++ = (char)t;
            continue;
        }
        pair_t p;
        if (!dyn_arr_get(pair_arr, t, &p) || p.a >= t || p.b >= t)
            goto bad;
        if (sp + 2 > scap)
        {
            uint32_t *ns = (uint32_t *)realloc(stack, 2 * scap * sizeof *stack);
            if (!ns)
                goto bad;
            stack = ns;
            scap *= 2;
        }
        stack[sp++] = p.b; /* right first so the left part is emitted first */
        stack[sp++] = p.a;
    }
    free(stack);
    return out;
bad:
    free(stack);
    return NULL;
}

/* resolve_pair() with the lengths already known (render_pairs computes them once, not once per id) */
static char *resolve_with_lengths(uint32_t pair_index, dyn_arr_t *pair_arr, const size_t *L)
{
    const size_t n = L[pair_index];
    if (n == 0) /* undefined id, or one that refers to itself / a later id */
        return NULL;
    char *s = (char *)malloc(n + 1);
    if (!s)
        return NULL;
    char *end = expand_into(pair_index, pair_arr, s, n);
    if (!end)
    {
        free(s);
        return NULL;
    }
    *end = '\0';
    return s;
}

char *resolve_pair(uint32_t pair_index, dyn_arr_t *pair_arr, hash_table_t *memoization_table) /* bpe.c:23-92 */
{
    if (!pair_arr || pair_index > pair_arr->last_index)
        return NULL;
    char *memo = NULL;
    if (memoization_table && hash_table_search(memoization_table, &pair_index, &memo) && memo)
    {
        char *copy = (char *)malloc(strlen(memo) + 1);
        if (copy)
            strcpy(copy, memo);
        return copy;
    }
    size_t *L = expansion_lengths(pair_arr);
    if (!L)
        return NULL;
    char *s = resolve_with_lengths(pair_index, pair_arr, L);
    const size_t n = L[pair_index];
    free(L);
    if (!s)
        return NULL;
    if (memoization_table)
    {
        char *keep = (char *)malloc(n + 1);
        if (keep)
        {
            memcpy(keep, s, n + 1);
            if (!hash_table_insert(memoization_table, &pair_index, &keep))
                free(keep);
        }
    }
    return s;
}

char *decompress(uint32_t *encoding, size_t len, dyn_arr_t *pair_arr) /* bpe.c:341-394 */
{
    /* the expansion of the whole stream runs on the GPU (bpe_cuda_decode); the string it returns is
     * NUL-terminated like the reference's (bpe.c:390) */
    if (!encoding || !pair_arr || pair_arr->last_index < 255)
        return NULL;
    const size_t n_merges = pair_arr->last_index - 255;
    bpe_pair_t *merges = (bpe_pair_t *)malloc((n_merges ? n_merges : 1) * sizeof *merges);
    if (!merges)
        return NULL;
    for (size_t r = 0; r < n_merges; r++)
    {
        pair_t p;
        if (!dyn_arr_get(pair_arr, 256 + r, &p))
        {
            free(merges);
            return NULL;
        }
        merges[r].a = p.a;
        merges[r].b = p.b;
    }
    uint8_t *bytes = NULL;
    size_t n_bytes = 0;
    const int rc = bpe_cuda_decode(encoding, len, merges, n_merges, &bytes, &n_bytes, NULL);
    free(merges);
    if (rc)
    {
        fprintf(stderr, "decompress: %s\n", bpe_cuda_last_error());
        return NULL;
    }
    return (char *)bytes;
}

void render_pairs(dyn_arr_t *pair_arr) /* bpe.c:94-128 */
{
    if (!pair_arr)
        return;
    size_t *L = expansion_lengths(pair_arr);
    if (!L)
        return;
    for (size_t index = 256; index <= pair_arr->last_index; index++)
    {
        char *str = resolve_with_lengths((uint32_t)index, pair_arr, L);
        if (!str)
            break;
        fprintf(stdout, "%zu => %s\n", index, str);
        free(str);
    }
    free(L);
}

/* ---- merge-table file: little-endian {u32 a; u32 b} records from id 256, no header (bpe.c:243-339) */
bool dump_pairs(const char *path, dyn_arr_t *pair_arr)
{
    if (!path || !pair_arr)
    {
        fprintf(stderr, "Invalid arguments to dump_pairs\n");
        return false;
    }
    FILE *f = fopen(path, "wb");
    if (!f)
    {
        perror("fopen");
        return false;
    }
    /* every merge is written; the reference stops one short and overflows a uint16_t counter (bpe.c:258) */
    for (size_t index = 256; index <= pair_arr->last_index; index++)
    {
        pair_t p;
        if (!dyn_arr_get(pair_arr, index, &p) || fwrite(&p, sizeof p, 1, f) != 1)
        {
            fclose(f);
            return false;
        }
    }
    return fclose(f) == 0;
}

dyn_arr_t *read_pairs(const char *path)
{
    if (!path)
        return NULL;
    FILE *f = fopen(path, "rb");
    if (!f)
    {
        perror("fopen");
        return NULL;
    }
    dyn_arr_t *arr = new_vocabulary();
    pair_t p;
    size_t index = 256;
    while (arr && fread(&p, sizeof p, 1, f) == 1)
    {
        /* a merge may only combine ids that exist already (bpe.c:752-758 creates them in order); a file that says
         * otherwise would send every expansion into a cycle */
        if (p.a >= index || p.b >= index)
        {
            fprintf(stderr, "read_pairs: record %zu refers to id %u, which does not exist yet\n", index - 256,
                    p.a >= index ? p.a : p.b);
            dyn_arr_free(arr);
            arr = NULL;
            break;
        }
        if (!dyn_arr_set(arr, index++, &p))
        {
            dyn_arr_free(arr);
            arr = NULL;
        }
    }
    fclose(f);
    return arr;
}

/* ---- Graphviz view of the vocabulary (bpe.c:198-241): host only, needs `dot` on PATH ------------ */
void print_graph(dyn_arr_t *pair_arr, const char *png_name, bool add_ascii)
{
    if (!pair_arr || !png_name)
        return;
    FILE *f = fopen("temp_graph.dot", "w");
    if (!f)
    {
        perror("fopen");
        return;
    }
    fprintf(f, "digraph vocabulary {\n");
    for (size_t index = add_ascii ? 0 : 256; index <= pair_arr->last_index; index++)
    {
        pair_t p;
        if (!dyn_arr_get(pair_arr, index, &p))
            continue;
        if (index < 256)
            fprintf(f, "  %zu [label=\"%zu\"];\n", index, index);
        else
            fprintf(f, "  %zu -> %u;\n  %zu -> %u;\n", index, p.a, index, p.b);
    }
    fprintf(f, "}\n");
    fclose(f);
    char cmd[1024];
    snprintf(cmd, sizeof cmd, "dot -Tpng temp_graph.dot -o %s && rm -f temp_graph.dot", png_name);
    if (system(cmd) != 0)
        fprintf(stderr, "print_graph: `dot` failed or is not installed; temp_graph.dot kept\n");
}
