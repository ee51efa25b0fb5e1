"""CPU oracle for the WT-PSE shape-regularization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the reported
CPU baseline.  The product path (``wt-pse-code_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Parity status
-------------
* Track R (whitening / MMD / KD-MSE / label thresholds -- what the reference's
  ``shape_networks.py`` and ``algorithms.py`` entry points really compute):
  **pinned**.  ``oracle/make_golden.py`` runs the unmodified reference from
  ``/root/reference`` (through ``oracle/ref_shim.py``) and stores its outputs under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every restatement here
  against those vectors.
* Track W (the DWT the task brief describes): the reference contains no wavelet
  code, so ``oracle/wavelet_np.py`` is a self-authored specification --
  **parity unpinned** (see DESIGN.md).
"""
