"""Stub so that `from tensorflow.contrib import slim` in lsi/nnutils/nets.py imports; the CNN is NOT
executed through the shim (slim semantics are restated in oracle/lsi_oracle_nets.py instead)."""
