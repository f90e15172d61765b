"""CPU oracle for the torch_rw hot path -- TEST INFRASTRUCTURE ONLY.

`oracle.orc` wraps oracle/trw_oracle.c (a plain-C restatement of the reference's csrc/cpu
implementation, glibc rand() stream included); `oracle.ref` loads the unmodified reference
extension from oracle/_ref/ when it has been built (oracle/build_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (torch_random_walk_b200/) never does.
"""
from . import orc  # noqa: F401
