"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's genotyping hot path, used as the parity checker by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
Nothing under ``tredparse_b200/`` imports this package; the product path fails loudly when its CUDA
library is missing instead of falling back to anything here.

Modules
-------
``sw``                 ctypes access to ``_build/libsw_oracle.so`` (our plain-C restatement, sw_oracle.c)
                       and, when present, ``_ref/libssw_ref.so`` (the reference's own ssw.c, unmodified).
``evidence_oracle``    restatement of tredparse/bam_parser.py (templates, read selection,
                       classification, tallies, paired-end distances, depth).
``likelihood_oracle``  dense, line-faithful restatement of tredparse/models.py.
``genotype_oracle``    the tred.run loop body (one (sample, locus) problem -> JSON-shaped dict).
"""
