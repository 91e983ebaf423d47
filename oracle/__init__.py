"""TEST INFRASTRUCTURE ONLY: CPU oracles for the parity tests (never imported by
mpifdtd_b200/)."""
