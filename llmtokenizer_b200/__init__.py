"""B200-native BPE merge-loop engine behind neofytr/LLMTokenizer's C API."""
from .bpe import (BpeCudaError, Context, compress, decode, dump_pairs, encode, encode_file, get_file, nccl_unique_id, print_text,
                  train_file,
                  read_pairs, train)

__all__ = ["BpeCudaError", "Context", "compress", "decode", "dump_pairs", "encode", "encode_file", "train_file", "get_file", "nccl_unique_id", "print_text",
           "read_pairs", "train"]
