"""Quantity DAG, types and estimation operators (mirror of ``mlmc/quantity`` in the reference)."""
