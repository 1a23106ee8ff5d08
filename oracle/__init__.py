"""Test oracle for amcontrast3d_b200 — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this package.  The product path (amcontrast3d_b200/) never does; it fails loudly when its CUDA
library is missing instead of falling back to anything in here.

  ops_oracle.c / ops_oracle.py   CPU restatement of the reference's CUDA operator kernels
  loss_oracle.py                 torch restatement of the reference's Python loss / refinement
  ref_kernels.py + Makefile      the reference's own CUDA kernels compiled unmodified from
                                 /root/reference into oracle/_ref/ (GPU-side pin + GPU baseline)
"""
