"""TEST INFRASTRUCTURE ONLY: CPU oracle for the ggdmc DE-MCMC / LBA hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this package.  Nothing under ggdmc_b200/ does.
"""
