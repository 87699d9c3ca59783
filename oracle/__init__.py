"""CPU oracle for the UniBEV uniform-BEV-encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``unibev_b200/`` may import this
package; it is used by ``tests/``, by ``__graft_entry__.smoke()`` as the checker
and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs as the timed
CPU restatement of the reference.

Parity status
-------------
* The reference's own files (``projects/UniBEV/unibev_plugin/models/modules``)
  are PINNED: ``tests/golden/make_golden.py`` imports them unmodified from
  ``/root/reference`` (behind import stubs for the absent third-party packages)
  and freezes their outputs into ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
  holds this oracle to those vectors.
* The arithmetic that lives in the un-vendored dependency mmcv-full==1.3.17
  (``multi_scale_deformable_attn_pytorch``, ``MultiScaleDeformableAttention``,
  ``FFN``, ``BaseTransformerLayer``) is restated in ``oracle/mmcv_semantics.py``
  from its published algorithm -- for that part parity is UNPINNED by reference
  tests (the reference ships none); it is cross-checked against the independent
  copy of the same algorithm in ``transformers`` and against a scalar C
  restatement of the mmcv CUDA kernel's per-corner semantics (``oracle/msda_core.c``).
"""
