"""CPU oracle for the salt-identification U-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``open-solution-salt-identification_b200/salt_b200``).  It is used by

* ``tests/``                      - as the checker the CUDA engine is compared with,
* ``__graft_entry__.smoke()``     - same, one tiny case,
* ``bench.py``                    - only for the ``cpu_baseline`` leg and ``--impl reference``.

The oracle is a plain-PyTorch (CPU, fp32) *restatement* of the reference algorithm
(`/root/reference/common_blocks/...`), written functionally over a ``state_dict``
instead of as ``nn.Module`` classes.  Every function cites the reference
file:line it follows.

Pinning: the reference ships no tests, golden vectors or fixtures for this path
 => by the reference's own tests parity is UNPINNED.
The restatement is instead pinned against the *unmodified reference modules
executed in the build container* (``oracle/make_golden.py`` imports them from
``/root/reference`` through ``oracle/ref_shims.py`` and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` replays those vectors.
"""
