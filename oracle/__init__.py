"""CPU oracle for the FPL+ DSBN 3D U-Net hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch-fp32 / numpy)
restatement of the reference's arithmetic for the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``fpl-plus_b200``) never does and fails loudly when its
CUDA library is missing.

Parity pinning: the reference ships no tests.  Its only golden artefacts
(``dataset/weight/cyc121_vst1s-gan.npy`` + ``config_dual/data_vs/
train_vs_t1s_wi+wp.csv``) pin the FPL sort order, the ``<50 => 1`` sentinel and
the image-weight map; they are extracted to ``tests/golden/fpl_image_weights.json``
and checked in ``tests/test_oracle_golden.py``.  Everything else (logits, loss,
gradients, BN statistics, window stitching, MC-dropout statistics) is pinned by
running the reference's own importable modules in the build container with
fixed inputs (``oracle/gen_golden.py``) and committing the outputs under
``tests/golden/``.
"""
