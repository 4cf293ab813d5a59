"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the VAEformer encode -> quantize -> entropy-code -> decode path
(plus the recipe that compiles the reference's own native coder into oracle/_ref). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import from here; nothing under
cra5_b200/ does.
"""
