"""x265-mod-by-patman_b200 -- B200 (sm_100a) implementation of x265's analysis primitives.

The product is the C-ABI shared library  lib/libx265b200.so  (include/x265b200.h) built from
csrc/*.cu, plus the EncoderPrimitives table fillers lib/libx265b200_glue_<depth>.so.  This Python
package is plumbing only: ctypes bindings used by tests/, bench.py and __graft_entry__.py, with
PyTorch supplying device memory, streams and torch.distributed.

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
opening a context without a usable sm_100 device raises.
"""
from .binding import (Context, HostAPI, LIB_PATH, OP_SAD, OP_SATD, OP_SA8D, OP_SSE_PP, TR_DCT, TR_DST,  # noqa: F401
                      TR_LOWPASS, ME_DIA, ME_HEX, ME_UMH, ME_STAR, ME_SEA, ME_FULL, IP_KINDS, load_library, declared_symbols, build_library, GLUE_HEADER, glue_path,
                      Plane, FrameJob, PassResult, TmePU, TmeResult, TME_MAX_CAND, expand_levels, PASS_CMP, PASS_COEF, PASS_LEVELS, TU_INTER, TU_INTRA_LUMA)
