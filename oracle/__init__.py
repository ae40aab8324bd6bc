"""ORACLE — test infrastructure only. CPU restatements of the reference's hot path used as the parity
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under shapeclipper_b200/ imports this package."""
