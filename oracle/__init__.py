"""CPU oracle for the MSDeformAttn hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  ir_ads_b200/ never does (tests/test_boundary.py greps for it).

  msda_c      ctypes wrapper over the plain-C restatement (oracle/msda_oracle.c)
  msda_torch  torch restatement of the reference's grid_sample formulation (the CPU path the
              reference itself runs when no GPU is present)
  ref_cuda    ctypes wrapper over the reference's own CUDA kernels built for sm_100a
              (oracle/_ref/libmsda_refcuda.so, GPU only)
"""
