"""Baselines that are timed NEXT TO the product (bench.py); never imported by fplplus_b200."""
