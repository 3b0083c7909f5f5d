"""`lsi` -- B200-native drop-in for the hot path of google/layered-scene-inference.

Same module paths and call signatures as the reference package (`lsi.geometry.{ldi,sampling,projection}`,
`lsi.nnutils.{helpers,nets}`, `lsi.loss.loss`), but the functions run eagerly on CUDA `torch.Tensor`s and
dispatch through ctypes to the C-ABI library `lsi/_lib/liblsi_b200.so` (hand-written sm_100a CUDA; see
include/lsi_b200.h).  There is no CPU path: calling a compute function with CPU tensors, or without the
built library, raises RuntimeError.
"""
__version__ = '0.1.0'
