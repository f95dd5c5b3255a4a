"""CPU oracle of the occupancy hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may import
anything from this package; nothing under biolith_b200/ does (tests/test_abi.py checks).  Modules:
occupancy (numpy: enumerated + closed forms), c_oracle (C/OpenMP port, liboccu_oracle.so), nuts (numpy NUTS).
"""
