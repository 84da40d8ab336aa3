/* A caller with the call sequence of the reference's main.c:3-25, plus the round trip the reference
 * never wires up (decompress) and the additive knobs.  Usage: example_main <file> [max_merges] [gpus] */
#include "bpe/inc/bpe.h"

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        fprintf(stderr, "Usage: %s <file_path> [max_merges] [gpus]\n", argv[0]);
        return EXIT_FAILURE;
    }
    uint32_t *text;
    size_t text_len;
    dyn_arr_t *pair_arr = argc > 2 ? compress_n(argv[1], &text, &text_len, (size_t)strtoull(argv[2], NULL, 10), argc > 3 ? atoi(argv[3]) : 1)
                                   : compress(argv[1], &text, &text_len);
    if (!pair_arr)
        return EXIT_FAILURE;
    print_text(text, (int)text_len);
    char *back = decompress(text, text_len, pair_arr);
    char *orig = get_file(argv[1]);
    const int ok = back && orig && strcmp(back, orig) == 0;
    fprintf(stderr, "%zu merges, %zu tokens, round trip %s\n", pair_arr->last_index - 255, text_len, ok ? "ok" : "MISMATCH");
    free(back);
    free(orig);
    free(text);
    dyn_arr_free(pair_arr);
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}
