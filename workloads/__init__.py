"""Benchmark / test workloads (synthetic problem families and the reference's fixture shapes).
Measurement and test infrastructure: not part of the product package piplib_b200/."""
