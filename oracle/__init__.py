"""oracle/ — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  Nothing under ``opticalflowscivis_b200/``
imports it: the product path is CUDA-only and raises when ``libofsv.so`` is missing.

Contents
  ops_ref.py    torch fp32 restatement of the L1 operators (warp 2D/3D, correlation-81,
                upsample2d_flow_as, WarpingLayer_no_div); device-agnostic so that the GPU
                tests can also run it on ``cuda`` as the "reference CUDA-eager" variant.
  ifnet_ref.py  torch fp32 restatement of IFBlock / IFNet / Model.inference (2D and 3D)
                with the reference's ``state_dict`` key names.
  ofsv_oracle.c plain-C restatement of the L1 operators' arithmetic,
                built by ``oracle/Makefile`` into ``oracle/libofsv_oracle.so``.
  c_oracle.py   ctypes/numpy binding of that library.

Parity pinning: the reference ships no golden vectors for this path.
The restatements were pinned against the reference ITSELF, imported unmodified from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``; that script
asserts ref == oracle on every case and writes the small fixtures in ``tests/golden/*.npz``
that travel to the GPU box.  The un-vendored native ``correlation_cuda`` extension
(NVIDIA flownet2-pytorch correlation_package, no pinned version) cannot be run anywhere:
against it parity is "unpinned"; the correlation oracle is pinned to the in-tree twin
``UPFlow/utils/pytorch_correlation.py`` (Corr_pyTorch) instead.
"""
