"""CPU oracle for the embed -> attack -> extract hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``image-in-speech-watermarking_b200/``) never imports this package
and raises if ``libwmk.so`` is missing.

The oracle is a plain PyTorch-CPU / numpy restatement of the reference's
algorithm.  Every function cites the reference file:line it follows.

Parity pinning: the reference ships no golden vectors or tests (SURVEY.md
section 4).  The restatement is pinned against outputs of the *unmodified
reference itself*, imported in the build container through
``oracle/shims.py`` and executed by ``oracle/make_golden.py``; the resulting
small fixtures are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py`` (CPU) and the ``-m gpu`` parity tests.
Third-party attack code that is absent from the reference tree (librosa
resample, libsndfile PCM_U8) is restated from its published algorithm and
is marked "parity unpinned" where it is defined.
"""
