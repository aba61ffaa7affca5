"""CPU oracle for the BSFG Gibbs sweep -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package imports this.  Allowed users: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.  Parity is pinned to the compiled reference, oracle/_ref (see bsfg_oracle.py header)."""
