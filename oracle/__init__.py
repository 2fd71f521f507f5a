"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the parity oracle for the BCn encode path:
  * `_ref/libref_oracle.so`  -- the UNMODIFIED reference encoder compiled from /root/reference/src (oracle/Makefile)
  * `_ref/librestate.so`     -- plain-C restatements (oracle/restate_*.c), each function citing reference file:line
  * `_ref/libbcdec.so`       -- spec block decoders (oracle/bcdec.c) for PSNR; the reference ships no decoder
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package (gfx_imagecompress_b200) never does.
"""
from .ref import RefOracle, Restated, Decoders, have_ref, have_restated, build  # noqa: F401
