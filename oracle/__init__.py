"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's message-passing hot path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product
(graphcast-lite_b200/gcl_b200) never does and fails loudly when its CUDA library is missing.

  pyg_shim/      restated torch_geometric 2.5.3 symbols (PARITY UNPINNED: no upstream vectors)
  trimesh_shim/  restated trimesh 4.4.0 closest_point (PARITY UNPINNED for exact ties)
  graphs.py      restated graph construction (PINNED: bit-exact vs the unmodified reference code)
  model.py       restated encode-process-decode glue (PINNED: bit-exact vs unmodified models.py
                 running on the same shims)
  make_golden.py runs the UNMODIFIED reference (/root/reference, build container only) on the shims
                 and writes tests/golden/
"""
