"""ORACLE -- test infrastructure only.

CPU restatement of the reference's algorithm for the segmentation hot path. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import it.
The product package ``xview2_b200`` never does (tests/test_no_oracle_in_product.py enforces this).
"""
