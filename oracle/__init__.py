"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/ref_ops.py header).  Parity unpinned against
MinkowskiEngine itself; pinned on the reference's in-tree restatement and known-answer tests."""
