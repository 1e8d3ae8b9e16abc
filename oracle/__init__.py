"""CPU oracle for the HoliRobPose inference forward -- TEST INFRASTRUCTURE, not product code.

A restatement of the reference algorithm (fp32, PyTorch-CPU ops for conv/linear exactly as the reference dispatches
them, numpy for the kinematics) used as the checker in tests/, `__graft_entry__.smoke()` and as the timed CPU baseline
in `bench.py` (`cpu_baseline` leg and `--impl reference`). Nothing under holistic-robot-pose-estimation-study_b200/
imports it, and the product path never falls back to it.

Pinning: the reference ships no tests, golden vectors or fixtures for this path, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container: oracle/refrun/harness.py imports the
unmodified /root/reference code, oracle/refrun/make_golden.py dumps its outputs on seeded inputs into tests/golden/,
and tests/test_oracle_golden.py asserts the port reproduces them. In-tree known answers (limb lengths const.py:108-124,
rot6d(I) full_net.py:205) are asserted on top.
"""
