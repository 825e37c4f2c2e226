"""CPU oracle for the MD_RDM depth-map fusion path.

TEST INFRASTRUCTURE ONLY.  Nothing under `md_rdm_b200/` imports this package.
The only legitimate importers are `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`, where it is the checker
or the timed CPU baseline -- never the product path.

Two restatements of the reference algorithm (az16/MD_RDM, `network/RDM_Net.py`
and `network/computations.py`), both in torch-on-CPU because every arithmetic
step of the reference is a torch ATen call:

* `oracle.fusion_ref`   -- vectorised restatement (seconds at batch 16).
* `oracle.literal`      -- loop-for-loop restatement of the two stages whose cost
                           in the reference is Python loops (pair build, Lloyd);
                           it is what `bench.py --impl reference` times.

Parity pinning: the reference ships no tests or golden vectors,
so both restatements are pinned against the *reference code itself*, imported
unmodified from /root/reference in the build container by
`tools/make_golden.py`; the resulting vectors live in `tests/golden/` and are
re-checked by `tests/test_oracle_golden.py` on every run.
"""
