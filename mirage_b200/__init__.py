"""mirage_b200 -- B200-native (sm_100a) implementation of the MIRAGE MultiViT hot path.

The public surface mirrors the reference (j-morano/MIRAGE):

    mirage_b200.mirage_hf.MIRAGEWrapper          <->  hf/mirage_hf.py:MIRAGEWrapper
    mirage_b200.mirage_wrapper.MIRAGEWrapper     <->  mirage_wrapper.py:MIRAGEWrapper (+ miragecls_factory)
    mirage_b200.model.{MIRAGEModel, MIRAGELight, model_factory}
    mirage_b200.input_adapters / output_adapters / criterion / utils

All compute goes through libmirage_b200.so (hand-written CUDA behind a C ABI, include/mirage_b200.h);
there is no CPU or library fallback.
"""
__version__ = "0.1.0"
