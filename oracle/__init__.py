"""oracle/ — CPU restatement of SPEC v0.  TEST INFRASTRUCTURE, NOT PRODUCT.

PARITY UNPINNED against upstream: /root/reference holds only README.md (lines 1-2: the
project name and the title "Semantic-Aware Multi-Modal Reinforcement Learning for
Autonomous Driving").  There is no upstream code, test, golden vector or fixture to pin
this oracle to, so it restates SPEC v0 (multimodal_sc_b200/spec.py) and is
pinned instead by
  * the published Philox4x32-10 known-answer vectors,
  * closed-form GAE and hand-computed PPO-loss cases,
  * invariants (unit power, empirical SNR, Rician moments, fp64 gradcheck)   -- all in tests/test_oracle.py,
  * frozen golden vectors it emitted once (tests/golden/oracle_v0.pt, scripts/make_golden.py; round-1 additions in
    tests/golden/oracle_v1.pt, scripts/make_golden_v1.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product package `multimodal_sc_b200` never does.
"""
