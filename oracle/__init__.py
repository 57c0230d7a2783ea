"""CPU oracle for the varKoder image hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``varkoder_b200``) never does: it fails loudly when its
CUDA library is missing instead of falling back to anything in here.

Contents
--------
``dsk_oracle.c`` / :mod:`oracle.dsk`   restatement of the native half (FASTQ framing, sub-sampling rule,
                                        dsk canonical counts, dsk2ascii text) -- PARITY UNPINNED at the
                                        dsk/reformat.sh boundary (binaries and sources absent, the
                                        reference's tests hold no expected counts).
:mod:`oracle.dsk_numpy`                 a SECOND, independently written restatement of the same rules (flat
                                        window enumeration + bincount in numpy instead of a rolling window in
                                        C); the two must agree on every test input and on the BASELINE-size
                                        sample (``tests/test_oracle.py``, ``tests/test_gpu_parity.py``).
:mod:`oracle.image`                     restatement of the Python half (ladder, pixel tables, scatter,
                                        rank scaling) -- PINNED against the imported, unmodified
                                        reference (``tests/golden/*.npz`` made by ``oracle/make_golden.py``)
                                        and against the reference's shipped ``docs/*.png`` properties.
:mod:`oracle.ref_shim`                  imports the unmodified reference from /root/reference (only in
                                        the build container; never on the GPU box).
"""
