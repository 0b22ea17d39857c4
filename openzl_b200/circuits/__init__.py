"""Workload generators: circuits expressed as R1CS for the Groth16 prover path."""
from .r1cs import R1CS
from .poseidon import PoseidonChain, PoseidonParams

__all__ = ["R1CS", "PoseidonChain", "PoseidonParams"]
