"""B200-native implementation of the Helping-Hands video-side forward path.

Layout (only what the hot path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (libhh_b200.so, header: include/hh_b200.h)
  _lib.py          ctypes binding;   ops.py  tensor-level operator wrappers
  model/LaviLa.py, model/tfm_decoder.py, model/metric.py, utils/box_ops.py
                   host-side mirrors of the reference modules of the same names (same classes, signatures,
                   state_dict keys) whose forwards call the C ABI
  parallel.py      data-parallel helpers: packed NCCL all-gather of embeddings, sharded similarity matrix
"""
__version__ = "0.1.0"
