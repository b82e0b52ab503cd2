"""CPU restatement of the reference algorithm -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs)."""
