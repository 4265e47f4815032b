"""ORACLE - test infrastructure only.

CPU restatements of the reference's spectral hot path used as the parity
checker: ``ref_port`` (scalar Python, mirrors the reference's cost) and
``c_oracle`` (batched C FFT, bit-identical, for large batches / large N).
Nothing under ``apda-fft_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` do.

The reference is pure Python, so there is no ``oracle/_ref`` build: the port is
pinned against the live reference (importable in the build container) by
``tests/golden/make_golden.py`` and the fixtures it commits.
"""
