"""vkdt_b200 — B200-native raw development engine with vkdt's module contract.

The product is the C-ABI library libvkdt_b200.so (include/vkdt_b200.h), built in tree by `make -C vkdt_b200`
(or __graft_entry__.build()).  This package is the thin Python host-side mirror used by tests and bench.py:
ctypes bindings only — there is no Python or CPU compute path, importing `vkdt_b200.api` fails loudly when
the CUDA library has not been built.
"""
__all__ = ["api", "synth"]
