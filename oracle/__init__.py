"""CPU oracle package (TEST INFRASTRUCTURE ONLY — see oracle/c/ark_oracle.c and oracle/py/bls12_381.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
